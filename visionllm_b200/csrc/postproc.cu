// Fused Grounding-DINO post-processing: the step right after the hot path, which defines the
// externally visible indices.  Replaces the chain of torch ops in
//   visionllmv2/eval/eval_det.py:18-56   post_process_det_gdino      (sigmoid -> top-k over Q*K -> //, % -> box gather/scale)
//   visionllmv2/eval/eval_det.py:59-104  post_process_instseg_gdino  (+ mask: 4x bilinear -> crop -> bilinear to the original
//                                                                      size -> sigmoid() > 0.5)
// by two kernels:
//   det_topk_kernel:   one CTA per image.  Radix select (4 x 8-bit passes over the monotone bit pattern of the fp32
//                      probability) finds the k-th largest sigmoid(logit); elements above it plus the lowest-index
//                      ties fill the k slots; a bitonic sort orders them by (probability desc, flat index asc);
//                      `idx // K`, `idx % K` (int64), cxcywh -> xyxy and the (w, h, w, h) scale are applied in place.
//   mask_chain_kernel: one thread per 4 output pixels of a selected query's mask: the two chained
//                      F.interpolate(mode='bilinear', align_corners=False) calls are evaluated analytically (ATen's
//                      upsample_bilinear2d arithmetic restated: scale = in / out, src = scale * (dst + 0.5) - 0.5 clamped
//                      at 0, 2 x 2 taps, row-then-column order), the crop is a clamp of the intermediate coordinate,
//                      and only the final `sigmoid(x) > 0.5` byte is written -- 16 L2-resident taps per output pixel
//                      instead of two full-resolution fp32 intermediates in HBM (~2.5 GB -> ~0.13 GB per 1024^2 image).
// Index contract: probabilities are compared as fp32 (same as torch.topk on the sigmoid output); when probabilities tie,
// the lower flat index comes first (torch leaves the order of ties unspecified).
#include "common.cuh"
#include <float.h>
#include <limits.h>

namespace {

__device__ __forceinline__ float sigmoid_f32(float x) { return 1.0f / (1.0f + expf(-x)); }   // ATen: 1 / (1 + exp(-x))

constexpr int TOPK_THREADS = 1024;
constexpr int TOPK_MAX = 1024;

// logits [B, Q, ld] fp32, first K columns of each row are scored.  Outputs per image: scores [k] fp32,
// topk_indexes [k] i64 (flat index q*K + c), box_idx [k] i64, labels [k] i64, boxes [k, 4] fp32 (xyxy, scaled).
__global__ void __launch_bounds__(TOPK_THREADS)
det_topk_kernel(const float* __restrict__ logits, const float* __restrict__ pred_boxes, const float* __restrict__ sizes_hw,
                int Q, int K, int ld, int k, float* __restrict__ scores, int64_t* __restrict__ topk_indexes,
                int64_t* __restrict__ box_idx, int64_t* __restrict__ labels, float* __restrict__ boxes) {
  __shared__ unsigned s_hist[256];
  __shared__ unsigned s_prefix, s_need, s_cnt_gt, s_cnt_eq;
  __shared__ unsigned s_key[TOPK_MAX];
  __shared__ int s_idx[TOPK_MAX];
  __shared__ unsigned s_wcnt[TOPK_THREADS / 32];
  const int b = blockIdx.x, tid = threadIdx.x;
  const long long n = (long long)Q * K;
  const float* lg = logits + (size_t)b * Q * ld;
  auto key_of = [&](long long i) -> unsigned {
    const int q = (int)(i / K), c = (int)(i - (long long)q * K);
    return __float_as_uint(sigmoid_f32(lg[(size_t)q * ld + c]));       // prob >= 0: uint order == float order
  };
  // ---- radix select: the k-th largest key -------------------------------------------------------
  if (tid == 0) { s_prefix = 0u; s_need = (unsigned)k; }
  __syncthreads();
  for (int pass = 0; pass < 4; ++pass) {
    const int shift = 24 - 8 * pass;
    if (tid < 256) s_hist[tid] = 0u;
    __syncthreads();
    const unsigned prefix = s_prefix;
    const unsigned mask_hi = pass == 0 ? 0u : (0xffffffffu << (shift + 8));
    for (long long base = 0; base < n; base += TOPK_THREADS) {       // warp-uniform trip count (ballot / match below)
      const long long i = base + tid;
      unsigned key = 0u;
      bool part = false;
      if (i < n) { key = key_of(i); part = (key & mask_hi) == prefix; }
      // probabilities share a handful of exponent bytes: aggregate equal digits inside the warp before the shared atomic
      const unsigned act = __ballot_sync(0xffffffffu, part);
      if (part) {
        const unsigned digit = (key >> shift) & 255u;
        const unsigned peers = __match_any_sync(act, digit);
        if ((peers & ((1u << (tid & 31)) - 1u)) == 0u) atomicAdd(&s_hist[digit], (unsigned)__popc(peers));
      }
    }
    __syncthreads();
    if (tid == 0) {
      unsigned need = s_need, acc = 0u;
      int d = 255;
      for (; d > 0; --d) {                                  // digits from large to small
        if (acc + s_hist[d] >= need) break;
        acc += s_hist[d];
      }
      s_need = need - acc;                                  // how many of digit d (under the prefix) are still wanted
      s_prefix = prefix | ((unsigned)d << shift);
    }
    __syncthreads();
  }
  const unsigned kth = s_prefix;                            // the k-th largest key; s_need = ties to take
  const unsigned ties_wanted = s_need;
  // ---- collect: every key > kth, and the `ties_wanted` lowest-index keys == kth -------------------
  if (tid == 0) { s_cnt_gt = 0u; s_cnt_eq = 0u; }
  __syncthreads();
  const unsigned n_gt = (unsigned)k - ties_wanted;
  // ties must be taken in index order: walk the array in order, chunk by chunk, with a block-wide exclusive count
  for (long long base = 0; base < n; base += TOPK_THREADS) {
    const long long i = base + tid;
    unsigned key = 0u;
    bool gt = false, eq = false;
    if (i < n) { key = key_of(i); gt = key > kth; eq = key == kth; }
    if (gt) { const unsigned slot = atomicAdd(&s_cnt_gt, 1u); s_key[slot] = key; s_idx[slot] = (int)i; }
    // ordered slot for ties: ballot within the warp + ordered scan over warps through shared memory
    const unsigned bal = __ballot_sync(0xffffffffu, eq);
    if ((tid & 31) == 0) s_wcnt[tid >> 5] = __popc(bal);
    __syncthreads();
    if (eq) {
      unsigned before = s_cnt_eq;
      for (int w = 0; w < (tid >> 5); ++w) before += s_wcnt[w];
      before += __popc(bal & ((1u << (tid & 31)) - 1u));
      if (before < ties_wanted) { s_key[n_gt + before] = key; s_idx[n_gt + before] = (int)i; }
    }
    __syncthreads();
    if (tid == 0) {
      unsigned tot = 0u;
      for (int w = 0; w < TOPK_THREADS / 32; ++w) tot += s_wcnt[w];
      s_cnt_eq += tot;
    }
    __syncthreads();
  }
  // ---- bitonic sort of the k selected entries: probability descending, flat index ascending -------
  int P2 = 1;
  while (P2 < k) P2 <<= 1;
  for (int i = k + tid; i < P2; i += TOPK_THREADS) { s_key[i] = 0u; s_idx[i] = INT_MAX; }   // pads sort last
  __syncthreads();
  auto before = [&](unsigned ka, int ia, unsigned kb, int ib) { return ka > kb || (ka == kb && ia < ib); };
  for (int size = 2; size <= P2; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = tid; t < P2 / 2; t += TOPK_THREADS) {
        const int lo = 2 * t - (t & (stride - 1)), hi = lo + stride;
        const bool up = ((lo & size) == 0);
        const unsigned ka = s_key[lo], kb = s_key[hi];
        const int ia = s_idx[lo], ib = s_idx[hi];
        const bool swap = up ? before(kb, ib, ka, ia) : before(ka, ia, kb, ib);
        if (swap) { s_key[lo] = kb; s_key[hi] = ka; s_idx[lo] = ib; s_idx[hi] = ia; }
      }
      __syncthreads();
    }
  }
  // ---- epilogue: //, %, box gather, cxcywh -> xyxy, scale ---------------------------------------------
  const float img_h = sizes_hw[2 * b], img_w = sizes_hw[2 * b + 1];
  for (int j = tid; j < k; j += TOPK_THREADS) {
    const int flat = s_idx[j];
    const int q = flat / K, c = flat - q * K;
    const size_t o = (size_t)b * k + j;
    scores[o] = __uint_as_float(s_key[j]);
    topk_indexes[o] = flat;
    box_idx[o] = q;
    labels[o] = c;
    const float4 bx = *reinterpret_cast<const float4*>(pred_boxes + ((size_t)b * Q + q) * 4);
    const float hw_ = __fmul_rn(0.5f, bx.z), hh_ = __fmul_rn(0.5f, bx.w);
    float4 r;
    r.x = __fmul_rn(__fsub_rn(bx.x, hw_), img_w);
    r.y = __fmul_rn(__fsub_rn(bx.y, hh_), img_h);
    r.z = __fmul_rn(__fadd_rn(bx.x, hw_), img_w);
    r.w = __fmul_rn(__fadd_rn(bx.y, hh_), img_h);
    *reinterpret_cast<float4*>(boxes + o * 4) = r;
  }
}

// ATen upsample_bilinear2d (align_corners = False, scales unset): source coordinate of output index `dst`
__device__ __forceinline__ float src_index(float scale, int dst) {
  const float s = scale * ((float)dst + 0.5f) - 0.5f;
  return s < 0.f ? 0.f : s;
}

struct Taps {
  int i0, i1;
  float l0, l1;
};
__device__ __forceinline__ Taps taps_of(float scale, int dst, int in_size) {
  Taps t;
  const float r = src_index(scale, dst);
  t.i0 = (int)r;
  t.i1 = t.i0 + ((t.i0 < in_size - 1) ? 1 : 0);
  t.l1 = r - (float)t.i0;
  t.l0 = 1.0f - t.l1;
  return t;
}

// value of the FIRST interpolation (src [H, W] -> [H*stride, W*stride]) at integer output (y, x)
__device__ __forceinline__ float up1(const float* __restrict__ src, int H, int W, float inv_h, float inv_w, int y, int x) {
  const Taps ty = taps_of(inv_h, y, H), tx = taps_of(inv_w, x, W);
  const float* r0 = src + (size_t)ty.i0 * W;
  const float* r1 = src + (size_t)ty.i1 * W;
  return ty.l0 * (tx.l0 * __ldg(r0 + tx.i0) + tx.l1 * __ldg(r0 + tx.i1)) +
         ty.l1 * (tx.l0 * __ldg(r1 + tx.i0) + tx.l1 * __ldg(r1 + tx.i1));
}

// masks [Q, H, W] fp32 of ONE image, box_idx [k] i64 -> out [k, oh, ow] uint8 (bool)
__global__ void __launch_bounds__(256)
mask_chain_kernel(const float* __restrict__ masks, const int64_t* __restrict__ box_idx, int H, int W, int stride, int crop_h,
                  int crop_w, int oh, int ow, float scale_h, float scale_w, unsigned char* __restrict__ out) {
  const int det = blockIdx.z;
  const float* src = masks + (size_t)box_idx[det] * H * W;
  const float inv_h = (float)H / (float)(H * stride), inv_w = (float)W / (float)(W * stride);   // ATen: (float)in / out
  const int y = blockIdx.y * blockDim.y + threadIdx.y;
  const int x0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (y >= oh || x0 >= ow) return;
  const Taps ty = taps_of(scale_h, y, crop_h);
  unsigned char res[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int x = x0 + j;
    res[j] = 0;
    if (x < ow) {
      const Taps tx = taps_of(scale_w, x, crop_w);
      const float v00 = up1(src, H, W, inv_h, inv_w, ty.i0, tx.i0), v01 = up1(src, H, W, inv_h, inv_w, ty.i0, tx.i1);
      const float v10 = up1(src, H, W, inv_h, inv_w, ty.i1, tx.i0), v11 = up1(src, H, W, inv_h, inv_w, ty.i1, tx.i1);
      const float v = ty.l0 * (tx.l0 * v00 + tx.l1 * v01) + ty.l1 * (tx.l0 * v10 + tx.l1 * v11);
      res[j] = sigmoid_f32(v) > 0.5f ? 1 : 0;
    }
  }
  unsigned char* o = out + ((size_t)det * oh + y) * ow + x0;
  if (x0 + 3 < ow && ((reinterpret_cast<uintptr_t>(o) & 3) == 0)) {
    *reinterpret_cast<uchar4*>(o) = make_uchar4(res[0], res[1], res[2], res[3]);
  } else {
    for (int j = 0; j < 4 && x0 + j < ow; ++j) o[j] = res[j];
  }
}

}  // namespace

extern "C" {

int vllm_det_postprocess_f32(const float* logits, const float* pred_boxes, const float* sizes_hw, int batch, int num_queries,
                             int num_classes, int logits_ld, int topk, float* scores, int64_t* topk_indexes,
                             int64_t* box_idx, int64_t* labels, float* boxes, void* stream) {
  if (batch < 0 || num_queries <= 0 || num_classes <= 0 || logits_ld < num_classes || topk <= 0) return VLLM_EINVAL;
  if (batch == 0) return VLLM_OK;
  if (!logits || !pred_boxes || !sizes_hw || !scores || !topk_indexes || !box_idx || !labels || !boxes) return VLLM_EINVAL;
  const long long n = (long long)num_queries * num_classes;
  if (topk > n || topk > TOPK_MAX || n > INT_MAX) return VLLM_EUNSUPPORTED;   // caller clamps k = min(topk, Q*K) like the reference
  if (!vllm_aligned(pred_boxes, 16) || !vllm_aligned(boxes, 16)) return VLLM_EALIGN;
  det_topk_kernel<<<batch, TOPK_THREADS, 0, (cudaStream_t)stream>>>(logits, pred_boxes, sizes_hw, num_queries, num_classes,
                                                                   logits_ld, topk, scores, topk_indexes, box_idx, labels,
                                                                   boxes);
  VLLM_CHECK_LAUNCH();
  return VLLM_OK;
}

int vllm_mask_postprocess_f32(const float* masks, const int64_t* box_idx, int num_det, int mask_h, int mask_w, int mask_stride,
                              int crop_h, int crop_w, int out_h, int out_w, unsigned char* out, void* stream) {
  if (num_det < 0 || mask_h <= 0 || mask_w <= 0 || mask_stride <= 0 || crop_h <= 0 || crop_w <= 0 || out_h <= 0 || out_w <= 0)
    return VLLM_EINVAL;
  if (num_det == 0) return VLLM_OK;
  if (!masks || !box_idx || !out) return VLLM_EINVAL;
  if (num_det > 65535) return VLLM_EUNSUPPORTED;
  // the crop cannot exceed the upsampled mask (python slicing clamps, eval_det.py:95)
  if (crop_h > mask_h * mask_stride) crop_h = mask_h * mask_stride;
  if (crop_w > mask_w * mask_stride) crop_w = mask_w * mask_stride;
  const float scale_h = (float)crop_h / (float)out_h, scale_w = (float)crop_w / (float)out_w;
  dim3 block(64, 4);
  dim3 grid((unsigned)((out_w + 255) / 256), (unsigned)((out_h + 3) / 4), (unsigned)num_det);
  if (grid.y > 65535) return VLLM_EUNSUPPORTED;
  mask_chain_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(masks, box_idx, mask_h, mask_w, mask_stride, crop_h, crop_w, out_h,
                                                             out_w, scale_h, scale_w, out);
  VLLM_CHECK_LAUNCH();
  return VLLM_OK;
}

}  // extern "C"
