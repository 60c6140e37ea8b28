// Vectorised HBM-bound row kernels of the forward hot path (bf16 in/out, fp32 math).
//
//  * RMSNorm  -- replaces apex FusedRMSNorm `cuApplyRMSNorm` (apex/csrc/layer_norm_cuda_kernel.cu:353-437,
//    CPU semantics apex/normalization/fused_layer_norm.py:16-29) == InternRMSNorm
//    (internvit/modeling_intern_vit.py:33-44) == InternLM2RMSNorm (internlm2/modeling_internlm2.py:114-128)
//    == HF LlamaRMSNorm: fp32 sum of squares, x * rsqrt(mean + eps) rounded to the input dtype, THEN times weight.
//    Strided rows so the InternViT q/k norm over the flattened 3200-d slices of the packed qkv tensor
//    (modeling_intern_vit.py:149-153) runs in place with no `torch.stack([q, k, v])` copy.
//  * LayerNorm (CLIP / GDINO nn.LayerNorm): fp32 mean/var, affine.
//  * RoPE, rotate-half form (HF Llama apply_rotary_pos_emb; internlm2/modeling_internlm2.py:218-232):
//    q' = q*cos + rotate_half(q)*sin with cos/sin rounded to bf16 like the reference's cached tables.
//
// One CTA per row; 16-byte loads; the row stays in registers between the reduction and the scaling pass
// (1 read + 1 write of the activation -- the roofline for these ops).
#include "common.cuh"

namespace {

constexpr int NT = 256;          // threads per row CTA
constexpr int MAX_VEC = 8;       // up to 8 x (8 bf16) per thread = 16384 columns

__device__ __forceinline__ float block_sum(float v, float* sh) {
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) sh[w] = v;
  __syncthreads();
  float t = (l < NT / 32) ? sh[l] : 0.f;
  t = warp_sum(t);
  __syncthreads();
  return t;
}

struct Row8 { float v[8]; };
__device__ __forceinline__ Row8 ld8(const __nv_bfloat16* p) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
  Row8 r;
#pragma unroll
  for (int i = 0; i < 4; ++i) { const float2 f = __bfloat1622float2(h[i]); r.v[2 * i] = f.x; r.v[2 * i + 1] = f.y; }
  return r;
}
__device__ __forceinline__ void st8(__nv_bfloat16* p, const float* v) {
  uint4 u; __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = u;
}

// mode 0: RMSNorm (weight only); mode 1: LayerNorm (weight + bias)
template <int MODE>
__global__ void __launch_bounds__(NT)
norm_rows_kernel(const __nv_bfloat16* __restrict__ x, long long ldx, const __nv_bfloat16* __restrict__ w,
                 const __nv_bfloat16* __restrict__ b, __nv_bfloat16* __restrict__ y, long long ldy, int cols,
                 float eps) {
  __shared__ float sh[NT / 32];
  const long long row = blockIdx.x;
  const __nv_bfloat16* xr = x + row * ldx;
  __nv_bfloat16* yr = y + row * ldy;
  const int nvec = cols / 8;
  Row8 reg[MAX_VEC];
  float s = 0.f, s2 = 0.f;
#pragma unroll
  for (int i = 0; i < MAX_VEC; ++i) {
    const int v = threadIdx.x + i * NT;
    if (v < nvec) {
      reg[i] = ld8(xr + v * 8);
#pragma unroll
      for (int j = 0; j < 8; ++j) { s += reg[i].v[j]; s2 += reg[i].v[j] * reg[i].v[j]; }
    }
  }
  float mean = 0.f, inv;
  if (MODE == 1) {
    mean = block_sum(s, sh) / cols;
    // two-pass variance on the register-resident row
    float d2 = 0.f;
#pragma unroll
    for (int i = 0; i < MAX_VEC; ++i) {
      const int v = threadIdx.x + i * NT;
      if (v < nvec) {
#pragma unroll
        for (int j = 0; j < 8; ++j) { const float d = reg[i].v[j] - mean; d2 += d * d; }
      }
    }
    inv = rsqrtf(block_sum(d2, sh) / cols + eps);
  } else {
    inv = rsqrtf(block_sum(s2, sh) / cols + eps);
  }
#pragma unroll
  for (int i = 0; i < MAX_VEC; ++i) {
    const int v = threadIdx.x + i * NT;
    if (v < nvec) {
      const Row8 wv = ld8(w + v * 8);
      float o[8];
      if (MODE == 1) {
        const Row8 bv = ld8(b + v * 8);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = (reg[i].v[j] - mean) * inv * wv.v[j] + bv.v[j];
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          // reference: (x * rsqrt).to(input_dtype) first, then weight * that (both bf16 roundings kept)
          const float n = __bfloat162float(__float2bfloat16(reg[i].v[j] * inv));
          o[j] = wv.v[j] * n;
        }
      }
      st8(yr + v * 8, o);
    }
  }
}

// qk [T, heads, 128]-style rows inside a packed tensor: element (t, h, d) at base + t*ld + h*hd + d.
// cos/sin [T, hd] bf16 (already gathered per position).  In place.
__global__ void __launch_bounds__(128)
rope_kernel(__nv_bfloat16* __restrict__ x, long long ld, const __nv_bfloat16* __restrict__ cs,
            const __nv_bfloat16* __restrict__ sn, int heads, int hd, long long tokens) {
  // one warp per (token, head); lane handles hd/64 pairs (d, d + hd/2)
  const long long wid = (long long)blockIdx.x * 4 + (threadIdx.x >> 5);
  if (wid >= tokens * heads) return;
  const long long t = wid / heads;
  const int h = (int)(wid % heads);
  __nv_bfloat16* p = x + t * ld + (long long)h * hd;
  const __nv_bfloat16* c = cs + t * hd;
  const __nv_bfloat16* s = sn + t * hd;
  const int half = hd / 2;
  for (int d = (threadIdx.x & 31) * 2; d < half; d += 64) {
    const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(p + d);
    const __nv_bfloat162 bq = *reinterpret_cast<const __nv_bfloat162*>(p + d + half);
    const float2 af = __bfloat1622float2(a), bf = __bfloat1622float2(bq);
    const float2 c0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(c + d));
    const float2 c1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(c + d + half));
    const float2 s0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(s + d));
    const float2 s1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(s + d + half));
    // reference rounds each bf16 product, then the sum: (q*cos) + (rotate_half(q)*sin)
    auto r = [](float v) { return __bfloat162float(__float2bfloat16(v)); };
    float2 lo, hi;
    lo.x = r(af.x * c0.x) + r(-bf.x * s0.x); lo.y = r(af.y * c0.y) + r(-bf.y * s0.y);
    hi.x = r(bf.x * c1.x) + r(af.x * s1.x);  hi.y = r(bf.y * c1.y) + r(af.y * s1.y);
    *reinterpret_cast<__nv_bfloat162*>(p + d) = __floats2bfloat162_rn(lo.x, lo.y);
    *reinterpret_cast<__nv_bfloat162*>(p + d + half) = __floats2bfloat162_rn(hi.x, hi.y);
  }
}

int check_rows(const void* x, long long ldx, const void* y, long long ldy, long long rows, int cols) {
  if (rows < 0 || cols <= 0) return VLLM_EINVAL;
  if (rows == 0) return 1000;
  if (!x || !y) return VLLM_EINVAL;
  if (cols % 8 || cols > 8 * NT * MAX_VEC) return VLLM_EUNSUPPORTED;
  if (ldx % 8 || ldy % 8 || !vllm_aligned(x, 16) || !vllm_aligned(y, 16)) return VLLM_EALIGN;
  if (rows > 2147483647LL) return VLLM_EUNSUPPORTED;
  return VLLM_OK;
}

}  // namespace

extern "C" {

int vllm_rmsnorm_bf16(const void* x, long long ldx, const void* weight, void* y, long long ldy, long long rows,
                      int cols, float eps, void* stream) {
  int rc = check_rows(x, ldx, y, ldy, rows, cols);
  if (rc == 1000) return VLLM_OK;
  if (rc) return rc;
  if (!weight || !vllm_aligned(weight, 16)) return VLLM_EINVAL;
  norm_rows_kernel<0><<<(unsigned)rows, NT, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)x, ldx, (const __nv_bfloat16*)weight, nullptr, (__nv_bfloat16*)y, ldy, cols, eps);
  VLLM_CHECK_LAUNCH();
  return VLLM_OK;
}

int vllm_layernorm_bf16(const void* x, long long ldx, const void* weight, const void* bias, void* y, long long ldy,
                        long long rows, int cols, float eps, void* stream) {
  int rc = check_rows(x, ldx, y, ldy, rows, cols);
  if (rc == 1000) return VLLM_OK;
  if (rc) return rc;
  if (!weight || !bias || !vllm_aligned(weight, 16) || !vllm_aligned(bias, 16)) return VLLM_EINVAL;
  norm_rows_kernel<1><<<(unsigned)rows, NT, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)x, ldx, (const __nv_bfloat16*)weight, (const __nv_bfloat16*)bias, (__nv_bfloat16*)y, ldy,
      cols, eps);
  VLLM_CHECK_LAUNCH();
  return VLLM_OK;
}

int vllm_rope_bf16(void* x, long long ld, const void* cos, const void* sin, long long tokens, int heads,
                   int head_dim, void* stream) {
  if (tokens < 0 || heads <= 0 || head_dim <= 0) return VLLM_EINVAL;
  if (tokens == 0) return VLLM_OK;
  if (!x || !cos || !sin) return VLLM_EINVAL;
  if (head_dim % 4 || ld % 2) return VLLM_EUNSUPPORTED;
  if (!vllm_aligned(x, 4) || !vllm_aligned(cos, 4) || !vllm_aligned(sin, 4)) return VLLM_EALIGN;
  const long long warps = tokens * heads;
  const long long blocks = (warps + 3) / 4;
  if (blocks > 2147483647LL) return VLLM_EUNSUPPORTED;
  rope_kernel<<<(unsigned)blocks, 128, 0, (cudaStream_t)stream>>>((__nv_bfloat16*)x, ld, (const __nv_bfloat16*)cos,
                                                                 (const __nv_bfloat16*)sin, heads, head_dim, tokens);
  VLLM_CHECK_LAUNCH();
  return VLLM_OK;
}

}  // extern "C"
