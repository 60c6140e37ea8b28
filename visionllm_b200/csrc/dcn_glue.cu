// Elementwise companions of the DCNv3 module (visionllmv2/model/ops_dcnv3/modules/dcnv3.py:318-349) -- the torch glue the
// reference runs as a dozen elementwise launches, as two passes:
//
//  * dcn_prep_kernel: the packed projection  om[row, :] = [ offset (G*K*2) | mask logits (G*K) | centre-scale logit (G) ]
//    (fp32, one GEMM, internimage.py) -> contiguous offset [rows, G*K*2], mask = softmax over the K taps of each group
//    (`F.softmax(mask.reshape(N,H,W,G,-1), -1)`, :325-326) [rows, G*K], and sigmoid(centre-scale) [rows, G]
//    (CenterFeatureScaleModule, :77-90).  One thread per (pixel, group).
//  * dcn_blend_kernel: x = core * (1 - s) + x_proj * s with s broadcast over the group's channels (:343-348), then the
//    cast to the module dtype for output_proj -- core and x_proj fp32 in, bf16 out, 8 channels per thread.
#include "common.cuh"

namespace {

template <int K>
__global__ void __launch_bounds__(256)
dcn_prep_kernel(const float* __restrict__ om, long long ld, float* __restrict__ offset, float* __restrict__ mask,
                float* __restrict__ scale, long long rows, int G) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * G) return;
  const long long row = idx / G;
  const int g = (int)(idx % G);
  const float* r = om + row * ld;
  float* op = offset + row * (G * K * 2) + g * K * 2;
#pragma unroll
  for (int j = 0; j < K * 2; ++j) op[j] = r[g * K * 2 + j];
  float m[K];
  float mx = -3.4e38f;
#pragma unroll
  for (int j = 0; j < K; ++j) { m[j] = r[G * K * 2 + g * K + j]; mx = fmaxf(mx, m[j]); }
  float sum = 0.f;
#pragma unroll
  for (int j = 0; j < K; ++j) { m[j] = expf(m[j] - mx); sum += m[j]; }
  float* mp = mask + row * (G * K) + g * K;
#pragma unroll
  for (int j = 0; j < K; ++j) mp[j] = m[j] / sum;
  if (scale) scale[row * G + g] = 1.f / (1.f + expf(-r[G * K * 3 + g]));
}

__global__ void __launch_bounds__(256)
dcn_blend_kernel(const float* __restrict__ core, const float* __restrict__ xproj, const float* __restrict__ scale,
                 __nv_bfloat16* __restrict__ out, long long rows, int C, int gc) {
  const int cv = C / 8;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * cv) return;
  const long long row = idx / cv;
  const int c0 = (int)(idx % cv) * 8;
  const float4* cp = reinterpret_cast<const float4*>(core + row * C + c0);
  float v[8];
  { const float4 a = cp[0], b = cp[1]; v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w; }
  if (scale) {
    const float4* xp = reinterpret_cast<const float4*>(xproj + row * C + c0);
    float x[8];
    { const float4 a = xp[0], b = xp[1]; x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w; }
    const int G = C / gc;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float s = scale[row * G + (c0 + j) / gc];
      v[j] = v[j] * (1.f - s) + x[j] * s;
    }
  }
  uint4 pk; __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
  *reinterpret_cast<uint4*>(out + row * C + c0) = pk;
}

}  // namespace

extern "C" {

int vllm_dcnv3_prep_f32(const void* packed, long long ld, void* offset, void* mask, void* scale, long long rows,
                        int group, int taps, void* stream) {
  if (rows < 0 || group <= 0 || taps <= 0) return VLLM_EINVAL;
  if (rows == 0) return VLLM_OK;
  if (!packed || !offset || !mask) return VLLM_EINVAL;
  if (ld < (long long)group * taps * 3 + (scale ? group : 0)) return VLLM_EINVAL;
  const long long n = rows * group;
  const long long blocks = (n + 255) / 256;
  if (blocks > 2147483647LL) return VLLM_EUNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  if (taps == 9)
    dcn_prep_kernel<9><<<(unsigned)blocks, 256, 0, st>>>((const float*)packed, ld, (float*)offset, (float*)mask,
                                                         (float*)scale, rows, group);
  else if (taps == 25)
    dcn_prep_kernel<25><<<(unsigned)blocks, 256, 0, st>>>((const float*)packed, ld, (float*)offset, (float*)mask,
                                                          (float*)scale, rows, group);
  else
    return VLLM_EUNSUPPORTED;
  VLLM_CHECK_LAUNCH();
  return VLLM_OK;
}

int vllm_dcnv3_blend_bf16(const void* core, const void* xproj, const void* scale, void* out, long long rows,
                          int channels, int group_channels, void* stream) {
  if (rows < 0 || channels <= 0 || group_channels <= 0 || channels % group_channels) return VLLM_EINVAL;
  if (rows == 0) return VLLM_OK;
  if (!core || !out || (scale && !xproj)) return VLLM_EINVAL;
  if (channels % 8) return VLLM_EUNSUPPORTED;
  if (!vllm_aligned(core, 16) || !vllm_aligned(out, 16) || (xproj && !vllm_aligned(xproj, 16))) return VLLM_EALIGN;
  const long long n = rows * (channels / 8);
  const long long blocks = (n + 255) / 256;
  if (blocks > 2147483647LL) return VLLM_EUNSUPPORTED;
  dcn_blend_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const float*)core, (const float*)xproj,
                                                                        (const float*)scale, (__nv_bfloat16*)out, rows,
                                                                        channels, group_channels);
  VLLM_CHECK_LAUNCH();
  return VLLM_OK;
}

}  // extern "C"
