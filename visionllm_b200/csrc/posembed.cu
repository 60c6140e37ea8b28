// Sine position embeddings of the Grounding-DINO stage as ONE kernel each (the reference: ~25 tiny elementwise launches per
// call -- slice, mul, div, sin, cos, stack, flatten, cat, cast, add):
//   * GroundingDinoSinePositionEmbedding (grounding_dino/modeling_ov_grounding_dino_mask_dn.py:529-564) of a feature level,
//     cast to the model dtype and added to the level embedding (:2420-2424): features (y_embed, x_embed);
//   * get_proposal_pos_embed / gen_sineembed_for_position of the decoder (:1755-1790): features (y, x, w, h) * 2 pi.
// out[r, f * nd + d] = (d even ? sinf : cosf)((feat_f[r] * pre) / dim_t[d]), fp32 IEEE arithmetic exactly like the torch
// elementwise kernels (no contraction possible: one multiply, one divide, one libdevice call); dim_t
// (= temperature ** (2 * (d // 2) / nd)) is computed once by torch and passed in, so the `pow` is the reference's own.
// bf16 output: round to bf16 (the `.to(dtype)`), then optionally add a bf16 row vector in fp32 and round again (the bf16
// tensor add of `pos + level_embed`).
#include "common.cuh"

namespace {

struct SineArgs {
  const float* feat[4];
  long long stride;        // element stride between consecutive rows of every feature
  int nfeat, nd;
  float pre;               // multiplier applied to the feature first (2 pi for the decoder form); 0 = none
  const float* dim_t;      // [nd]
  long long rows;
  void* out; long long ldo; int out_bf16;
  long long rpb, obs;      // output row r lives at (r / rpb) * obs + (r % rpb) * ldo elements (a level slab of a [B, S, C] buffer)
  const __nv_bfloat16* add_row;   // optional [nfeat * nd]
};

__global__ void __launch_bounds__(256)
sine_embed_kernel(const __grid_constant__ SineArgs a) {
  const int chunks = a.nfeat * a.nd / 8;                       // 8 outputs per thread
  const long long total = a.rows * chunks;
  for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < total; v += (long long)gridDim.x * blockDim.x) {
    const long long r = v / chunks;
    const int c = (int)(v - r * chunks);
    const int f = (c * 8) / a.nd, d0 = (c * 8) - f * a.nd;
    float x = a.feat[f][r * a.stride];
    if (a.pre != 0.f) x = __fmul_rn(x, a.pre);
    float o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float e = __fdiv_rn(x, a.dim_t[d0 + j]);
      o[j] = (j & 1) ? cosf(e) : sinf(e);                      // d0 is a multiple of 8: parity of d = parity of j
    }
    const long long ob = r / a.rpb;
    const long long off = ob * a.obs + (r - ob * a.rpb) * a.ldo + c * 8;
    if (a.out_bf16) {
      __nv_bfloat16 h[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        h[j] = __float2bfloat16_rn(o[j]);
        if (a.add_row) h[j] = __float2bfloat16_rn(__fadd_rn(__bfloat162float(h[j]), __bfloat162float(a.add_row[c * 8 + j])));
      }
      *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(a.out) + off) = *reinterpret_cast<const uint4*>(h);
    } else {
      float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(a.out) + off);
      dst[0] = make_float4(o[0], o[1], o[2], o[3]);
      dst[1] = make_float4(o[4], o[5], o[6], o[7]);
    }
  }
}

}  // namespace

extern "C" {

int vllm_sine_embed_f32(const float* f0, const float* f1, const float* f2, const float* f3, long long feat_stride, int nfeat,
                        float pre_scale, const float* dim_t, int nd, long long rows, void* out, long long ldo, int out_bf16,
                        long long rows_per_batch, long long out_batch_stride, const void* add_row_bf16, void* stream) {
  if (rows_per_batch < 0 || (rows_per_batch > 0 && out_batch_stride < rows_per_batch * ldo)) return VLLM_EINVAL;
  if (rows < 0 || nfeat < 1 || nfeat > 4 || nd <= 0 || nd % 8 || ldo < (long long)nfeat * nd || feat_stride <= 0) return VLLM_EINVAL;
  if (rows == 0) return VLLM_OK;
  const float* fs[4] = {f0, f1, f2, f3};
  for (int i = 0; i < nfeat; ++i) if (!fs[i]) return VLLM_EINVAL;
  if (!dim_t || !out || (add_row_bf16 && !out_bf16)) return VLLM_EINVAL;
  if (ldo % (out_bf16 ? 8 : 4) || out_batch_stride % (out_bf16 ? 8 : 4) || !vllm_aligned(out, 16)) return VLLM_EALIGN;
  SineArgs a;
  for (int i = 0; i < 4; ++i) a.feat[i] = fs[i];
  a.stride = feat_stride; a.nfeat = nfeat; a.nd = nd; a.pre = pre_scale; a.dim_t = dim_t; a.rows = rows;
  a.out = out; a.ldo = ldo; a.out_bf16 = out_bf16;
  a.rpb = rows_per_batch > 0 ? rows_per_batch : rows; a.obs = rows_per_batch > 0 ? out_batch_stride : 0;
  a.add_row = (const __nv_bfloat16*)add_row_bf16;
  long long blocks = (rows * (nfeat * nd / 8) + 255) / 256;
  const long long cap = (long long)vllm_num_sms() * 16;
  if (blocks > cap) blocks = cap;
  sine_embed_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(a);
  VLLM_CHECK_LAUNCH();
  return VLLM_OK;
}

}  // extern "C"
