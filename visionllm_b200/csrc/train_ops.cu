// Row kernels of the training-side path of the LLM decoder (BASELINE cfg 5 "fwd+bwd step"): the backward of the
// HBM-bound forward ops of fused_ops.cu / the GEMM epilogues, and the loss.  bf16 activations / gradients, fp32 math.
//
//   rmsnorm_bwd_kernel     autograd of HF LlamaRMSNorm == InternLM2RMSNorm (internlm2/modeling_internlm2.py:114-128) ==
//                          apex rms_backward_affine (apex/csrc/layer_norm_cuda_kernel.cu cuComputeGradInput / GradGammaBeta):
//                          n = x * rsqrt(mean(x^2) + eps); y = w * n;  dn = dy * w;
//                          dx = rstd * (dn - n * mean(dn * n));  dw += sum_rows dy * n
//   swiglu_fwd / bwd       h = silu(g) * u on the interleaved (gate, up) columns the gate|up GEMM produces
//                          (LlamaMLP / InternLM2MLP: down(act(gate(x)) * up(x))); dg = dh * u * silu'(g), du = dh * silu(g)
//   softmax_causal_kernel  P = softmax(scale * S) over keys <= query (fp32 softmax like HF eager attention), zeros above
//                          the diagonal; attn_ds_kernel: dS = scale * P * (dP - sum_k dP * P)  (softmax backward), zeros above
//   ce_loss_kernel         visionllmv2/model/modeling_visionllmv2.py:741-757: CrossEntropyLoss (mean over labels != -100) of
//                          fp32 logits rows vs int64 labels; writes the loss sum and dlogits = (softmax - onehot) / n_valid.
#include "common.cuh"
#include <math.h>
#include <type_traits>

namespace {

__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) { const float2 t = __bfloat1622float2(h[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 u; __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  return u;
}

template <int NT>
__device__ __forceinline__ float block_sum(float v, float* sh) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int i = 0; i < NT / 32; ++i) t += sh[i];
  return t;
}
template <int NT>
__device__ __forceinline__ float block_max(float v, float* sh) {
  v = warp_max(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = -INFINITY;
#pragma unroll
  for (int i = 0; i < NT / 32; ++i) t = fmaxf(t, sh[i]);
  return t;
}

// One CTA (256 threads) per row; VPT 16-byte vectors per thread stay in registers.  dw: fp32 [cols], atomically
// accumulated (the caller zeroes it), one atomic per column per CTA-owned row group.
template <int VPT>
__global__ void __launch_bounds__(256)
rmsnorm_bwd_kernel(const __nv_bfloat16* __restrict__ x, long long ldx, const __nv_bfloat16* __restrict__ w,
                   const __nv_bfloat16* __restrict__ dy, long long ldy, __nv_bfloat16* __restrict__ dx, long long lddx,
                   float* __restrict__ dw, long long rows, int cols, float eps, int rows_per_cta,
                   float* __restrict__ partials = nullptr) {
  __shared__ float sh[8];
  const int nvec = cols / 8;
  float dwacc[VPT][8];
#pragma unroll
  for (int i = 0; i < VPT; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) dwacc[i][j] = 0.f;
  const long long r0 = (long long)blockIdx.x * rows_per_cta;
  for (long long row = r0; row < r0 + rows_per_cta && row < rows; ++row) {
    uint4 xr[VPT], gr[VPT];
    float s2 = 0.f;
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
      const int v = threadIdx.x + i * 256;
      xr[i] = gr[i] = make_uint4(0u, 0u, 0u, 0u);
      if (v < nvec) {
        xr[i] = *(reinterpret_cast<const uint4*>(x + row * ldx) + v);
        gr[i] = *(reinterpret_cast<const uint4*>(dy + row * ldy) + v);
        float f[8]; unpack8(xr[i], f);
#pragma unroll
        for (int j = 0; j < 8; ++j) s2 += f[j] * f[j];
      }
    }
    const float rstd = rsqrtf(block_sum<256>(s2, sh) / cols + eps);
    float dot = 0.f;                                        // sum dn * n
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
      const int v = threadIdx.x + i * 256;
      if (v < nvec) {
        float f[8], g[8], wv[8];
        unpack8(xr[i], f); unpack8(gr[i], g); unpack8(__ldg(reinterpret_cast<const uint4*>(w) + v), wv);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float n = f[j] * rstd;
          dot += g[j] * wv[j] * n;
          // dw uses the forward's bf16-rounded normalised value (y = w * bf16(n)), like autograd through the cast
          dwacc[i][j] += g[j] * __bfloat162float(__float2bfloat16(n));
        }
      }
    }
    const float mdot = block_sum<256>(dot, sh) / cols;
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
      const int v = threadIdx.x + i * 256;
      if (v < nvec) {
        float f[8], g[8], wv[8], o[8];
        unpack8(xr[i], f); unpack8(gr[i], g); unpack8(__ldg(reinterpret_cast<const uint4*>(w) + v), wv);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = rstd * (g[j] * wv[j] - f[j] * rstd * mdot);
        *(reinterpret_cast<uint4*>(dx + row * lddx) + v) = pack8(o);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < VPT; ++i) {
    const int v = threadIdx.x + i * 256;
    if (v < nvec) {
      if (partials) {                                      // one coalesced row of partials per CTA, summed by the second kernel
        float* pr = partials + (size_t)blockIdx.x * cols + v * 8;
        *reinterpret_cast<float4*>(pr) = make_float4(dwacc[i][0], dwacc[i][1], dwacc[i][2], dwacc[i][3]);
        *reinterpret_cast<float4*>(pr + 4) = make_float4(dwacc[i][4], dwacc[i][5], dwacc[i][6], dwacc[i][7]);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) atomicAdd(dw + v * 8 + j, dwacc[i][j]);
      }
    }
  }
}

// dw[c] = sum over the n per-CTA partial rows in a fixed order (deterministic, unlike the atomic form): a CTA owns 32 columns,
// its 8 warps take every 8th row (128-byte coalesced reads), shared-memory tree at the end
__global__ void __launch_bounds__(256)
colsum_partials_kernel(const float* __restrict__ partials, float* __restrict__ dw, int n, int cols) {
  __shared__ float sh[8][32];
  const int lane = threadIdx.x & 31, grp = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lane;
  float a = 0.f;
  if (c < cols)
    for (int r = grp; r < n; r += 8) a += partials[(size_t)r * cols + c];
  sh[grp][lane] = a;
  __syncthreads();
  if (grp == 0 && c < cols) {
    float t = 0.f;
#pragma unroll
    for (int g = 0; g < 8; ++g) t += sh[g][lane];
    dw[c] = t;
  }
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + __expf(-x)); }

// gu [rows, 2I] interleaved (g0, u0, g1, u1, ...) -> h [rows, I]; 8 outputs per thread
__global__ void __launch_bounds__(256)
swiglu_fwd_kernel(const __nv_bfloat16* __restrict__ gu, long long ldgu, __nv_bfloat16* __restrict__ h, long long ldh,
                  long long rows, int inter) {
  const int vec_per_row = inter / 8;
  const long long total = rows * vec_per_row;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / vec_per_row;
    const int v = (int)(i - row * vec_per_row);
    const uint4* src = reinterpret_cast<const uint4*>(gu + row * ldgu) + 2 * v;
    float a[8], b[8], o[8];
    unpack8(src[0], a); unpack8(src[1], b);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      o[j] = a[2 * j] * sigmoidf_(a[2 * j]) * a[2 * j + 1];
      o[4 + j] = b[2 * j] * sigmoidf_(b[2 * j]) * b[2 * j + 1];
    }
    *(reinterpret_cast<uint4*>(h + row * ldh) + v) = pack8(o);
  }
}

__global__ void __launch_bounds__(256)
swiglu_bwd_kernel(const __nv_bfloat16* __restrict__ gu, long long ldgu, const __nv_bfloat16* __restrict__ dh, long long lddh,
                  __nv_bfloat16* __restrict__ dgu, long long lddgu, long long rows, int inter) {
  const int vec_per_row = inter / 8;
  const long long total = rows * vec_per_row;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / vec_per_row;
    const int v = (int)(i - row * vec_per_row);
    const uint4* src = reinterpret_cast<const uint4*>(gu + row * ldgu) + 2 * v;
    float a[8], b[8], d[8], oa[8], ob[8];
    unpack8(src[0], a); unpack8(src[1], b);
    unpack8(*(reinterpret_cast<const uint4*>(dh + row * lddh) + v), d);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      {
        const float g = a[2 * j], u = a[2 * j + 1], s = sigmoidf_(g);
        oa[2 * j] = d[j] * u * s * (1.f + g * (1.f - s));
        oa[2 * j + 1] = d[j] * g * s;
      }
      {
        const float g = b[2 * j], u = b[2 * j + 1], s = sigmoidf_(g);
        ob[2 * j] = d[4 + j] * u * s * (1.f + g * (1.f - s));
        ob[2 * j + 1] = d[4 + j] * g * s;
      }
    }
    uint4* dst = reinterpret_cast<uint4*>(dgu + row * lddgu) + 2 * v;
    dst[0] = pack8(oa); dst[1] = pack8(ob);
  }
}

// S [n_mat, T, T] bf16 (row pitch lds, matrix pitch T*lds) -> P in place: softmax over keys j <= i of scale * S, zeros for j > i.
// One CTA per row; T <= 256 * 8 * VPT.
template <int VPT>
__global__ void __launch_bounds__(256)
softmax_causal_kernel(__nv_bfloat16* __restrict__ s, long long lds, int T, float scale) {
  __shared__ float sh[8];
  const long long row_g = blockIdx.x;                        // matrix * T + i
  const int i = (int)(row_g % T);
  __nv_bfloat16* sr = s + row_g * lds;
  const int nvec = T / 8;
  float v[VPT][8];
  float mx = -INFINITY;
#pragma unroll
  for (int k = 0; k < VPT; ++k) {
    const int vi = threadIdx.x + k * 256;
#pragma unroll
    for (int j = 0; j < 8; ++j) v[k][j] = -INFINITY;
    if (vi < nvec && vi * 8 <= i) {
      float f[8]; unpack8(*(reinterpret_cast<const uint4*>(sr) + vi), f);
#pragma unroll
      for (int j = 0; j < 8; ++j) if (vi * 8 + j <= i) { v[k][j] = f[j] * scale; mx = fmaxf(mx, v[k][j]); }
    }
  }
  mx = block_max<256>(mx, sh);
  float sum = 0.f;
#pragma unroll
  for (int k = 0; k < VPT; ++k)
#pragma unroll
    for (int j = 0; j < 8; ++j) { v[k][j] = __expf(v[k][j] - mx); sum += v[k][j]; }     // exp(-inf) = 0 above the diagonal
  const float inv = 1.f / block_sum<256>(sum, sh);
  // zeros are only needed inside the 256-aligned diagonal block: the causal batched GEMMs (gemm.cu, causal 2 / 3) never read
  // a 64-column k-block that lies wholly above a tile's diagonal block, and tiles are at most 256 rows
  const int wvec = min(nvec, ((i >> 8) + 1) * 32);
#pragma unroll
  for (int k = 0; k < VPT; ++k) {
    const int vi = threadIdx.x + k * 256;
    if (vi < wvec) {
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = v[k][j] * inv;
      *(reinterpret_cast<uint4*>(sr) + vi) = pack8(o);
    }
  }
}

// dP [.., T, T] -> dS in place: dS = scale * P * (dP - sum_j dP_j P_j); zeros above the diagonal (P is zero there)
template <int VPT>
__global__ void __launch_bounds__(256)
attn_ds_kernel(const __nv_bfloat16* __restrict__ p, __nv_bfloat16* __restrict__ dp, long long ld, int T, float scale) {
  __shared__ float sh[8];
  const long long row_g = blockIdx.x;
  const int i = (int)(row_g % T);
  const __nv_bfloat16* pr = p + row_g * ld;
  __nv_bfloat16* dr = dp + row_g * ld;
  const int nvec = T / 8;
  float pv[VPT][8], dv[VPT][8];
  float dot = 0.f;
#pragma unroll
  for (int k = 0; k < VPT; ++k) {
    const int vi = threadIdx.x + k * 256;
#pragma unroll
    for (int j = 0; j < 8; ++j) { pv[k][j] = 0.f; dv[k][j] = 0.f; }
    if (vi < nvec && vi * 8 <= i) {
      unpack8(*(reinterpret_cast<const uint4*>(pr) + vi), pv[k]);
      unpack8(*(reinterpret_cast<const uint4*>(dr) + vi), dv[k]);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (vi * 8 + j > i) { pv[k][j] = 0.f; dv[k][j] = 0.f; }       // dP above the diagonal may be uncomputed garbage
        dot += pv[k][j] * dv[k][j];
      }
    }
  }
  dot = block_sum<256>(dot, sh);
  const int wvec = min(nvec, ((i >> 8) + 1) * 32);            // see softmax_causal_kernel
#pragma unroll
  for (int k = 0; k < VPT; ++k) {
    const int vi = threadIdx.x + k * 256;
    if (vi < wvec) {
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = scale * pv[k][j] * (dv[k][j] - dot);
      *(reinterpret_cast<uint4*>(dr) + vi) = pack8(o);
    }
  }
}

// [B, T, 3, H, D] (the packed q | k | v projection rows) <-> [3, B, H, T, D] (every (batch, head) matrix of q, k, v stacked
// along rows for the block-diagonal batched GEMMs of the attention backward), 16-byte vectors; `to_stacked` = 0 goes back
// (the packed gradient).  NP = 3 for qkv, 1 for a plain [B, T, H, D] tensor (dO).
__global__ void __launch_bounds__(256)
head_stack_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, int B, int T, int NP, int H, int DV,
                  int to_stacked, long long n_vec) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += (long long)gridDim.x * blockDim.x) {
    // i indexes the STACKED tensor [NP, B, H, T, DV] (coalesced on the stacked side, 256-byte runs on the packed side)
    const int d = (int)(i % DV);
    long long r = i / DV;
    const int t = (int)(r % T); r /= T;
    const int h = (int)(r % H); r /= H;
    const int b = (int)(r % B);
    const int part = (int)(r / B);
    const long long packed = ((((long long)b * T + t) * NP + part) * H + h) * DV + d;
    if (to_stacked) dst[i] = src[packed];
    else dst[packed] = src[i];
  }
}

// logits fp32 [rows, V] (pitch ld), labels int64 [rows] (-100 = ignore), n_valid: device int64 scalar.
// loss_sum += -log softmax(logits)[label] (fp32 atomic); dlogits bf16 [rows, V] (pitch ldd) = (softmax - onehot) / n_valid,
// zero rows for ignored labels.  One CTA (512 threads) per row, three passes over the row (max, sum, write).
__global__ void __launch_bounds__(512)
ce_loss_kernel(const float* __restrict__ logits, long long ld, const int64_t* __restrict__ labels,
               const int64_t* __restrict__ n_valid, int V, float* __restrict__ loss_sum, __nv_bfloat16* __restrict__ dlogits,
               long long ldd) {
  __shared__ float sh[16];
  const long long row = blockIdx.x;
  const float* lr = logits + row * ld;
  const long long label = labels[row];
  __nv_bfloat16* dr = dlogits ? dlogits + row * ldd : nullptr;
  if (label < 0 || label >= V) {                             // IGNORE_INDEX (-100): no loss, zero gradient
    if (dr) for (int c = threadIdx.x; c < V; c += 512) dr[c] = __float2bfloat16(0.f);
    return;
  }
  float mx = -INFINITY;
  for (int c = threadIdx.x; c < V; c += 512) mx = fmaxf(mx, lr[c]);
  mx = block_max<512>(mx, sh);
  float sum = 0.f;
  for (int c = threadIdx.x; c < V; c += 512) sum += expf(lr[c] - mx);
  sum = block_sum<512>(sum, sh);
  const float lse = mx + logf(sum);
  if (threadIdx.x == 0) atomicAdd(loss_sum, lse - lr[label]);
  if (dr) {
    const float invn = 1.f / (float)(*n_valid);
    for (int c = threadIdx.x; c < V; c += 512) {
      float g = expf(lr[c] - lse);
      if (c == label) g -= 1.f;
      dr[c] = __float2bfloat16(g * invn);
    }
  }
}

}  // namespace

extern "C" {

int vllm_rmsnorm_bwd_bf16(const void* x, long long ldx, const void* weight, const void* dy, long long ldy, void* dx,
                          long long lddx, float* dweight, long long rows, int cols, float eps, void* stream) {
  if (rows < 0 || cols <= 0 || cols % 8) return VLLM_EINVAL;
  if (rows == 0) return VLLM_OK;
  if (!x || !weight || !dy || !dx || !dweight) return VLLM_EINVAL;
  if (ldx % 8 || ldy % 8 || lddx % 8 || !vllm_aligned(x, 16) || !vllm_aligned(dy, 16) || !vllm_aligned(dx, 16)) return VLLM_EALIGN;
  const int nvec = cols / 8;
  const int rows_per_cta = (int)((rows + (long long)vllm_num_sms() * 8 - 1) / ((long long)vllm_num_sms() * 8));
  const long long blocks = (rows + rows_per_cta - 1) / rows_per_cta;
  cudaStream_t st = (cudaStream_t)stream;
  auto go = [&](auto vpt) -> int {
    constexpr int VPT = decltype(vpt)::value;
    rmsnorm_bwd_kernel<VPT><<<(unsigned)blocks, 256, 0, st>>>((const __nv_bfloat16*)x, ldx, (const __nv_bfloat16*)weight,
                                                             (const __nv_bfloat16*)dy, ldy, (__nv_bfloat16*)dx, lddx, dweight,
                                                             rows, cols, eps, rows_per_cta);
    VLLM_CHECK_LAUNCH();
    return VLLM_OK;
  };
  if (nvec <= 256) return go(std::integral_constant<int, 1>{});
  if (nvec <= 512) return go(std::integral_constant<int, 2>{});
  if (nvec <= 1024) return go(std::integral_constant<int, 4>{});
  return VLLM_EUNSUPPORTED;
}

static int rmsnorm_bwd_ctas(long long rows) {
  const int rows_per_cta = (int)((rows + (long long)vllm_num_sms() * 8 - 1) / ((long long)vllm_num_sms() * 8));
  return rows_per_cta > 0 ? (int)((rows + rows_per_cta - 1) / rows_per_cta) : 0;
}

/* rows of the `partials` workspace vllm_rmsnorm_bwd_ws_bf16 needs for `rows` activation rows */
int vllm_rmsnorm_bwd_partials(long long rows) { return rows > 0 ? rmsnorm_bwd_ctas(rows) : 0; }

int vllm_rmsnorm_bwd_ws_bf16(const void* x, long long ldx, const void* weight, const void* dy, long long ldy, void* dx,
                             long long lddx, float* dweight, float* partials, int n_partials, long long rows, int cols,
                             float eps, void* stream) {
  if (rows < 0 || cols <= 0 || cols % 8) return VLLM_EINVAL;
  if (rows == 0) return VLLM_OK;
  if (!x || !weight || !dy || !dx || !dweight || !partials) return VLLM_EINVAL;
  if (ldx % 8 || ldy % 8 || lddx % 8 || !vllm_aligned(x, 16) || !vllm_aligned(dy, 16) || !vllm_aligned(dx, 16) ||
      !vllm_aligned(partials, 16))
    return VLLM_EALIGN;
  const int nvec = cols / 8;
  const int rows_per_cta = (int)((rows + (long long)vllm_num_sms() * 8 - 1) / ((long long)vllm_num_sms() * 8));
  const int blocks = rmsnorm_bwd_ctas(rows);
  if (n_partials < blocks) return VLLM_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  auto go = [&](auto vpt) -> int {
    constexpr int VPT = decltype(vpt)::value;
    rmsnorm_bwd_kernel<VPT><<<(unsigned)blocks, 256, 0, st>>>((const __nv_bfloat16*)x, ldx, (const __nv_bfloat16*)weight,
                                                             (const __nv_bfloat16*)dy, ldy, (__nv_bfloat16*)dx, lddx, dweight,
                                                             rows, cols, eps, rows_per_cta, partials);
    VLLM_CHECK_LAUNCH();
    colsum_partials_kernel<<<(unsigned)((cols + 31) / 32), 256, 0, st>>>(partials, dweight, blocks, cols);
    VLLM_CHECK_LAUNCH();
    return VLLM_OK;
  };
  if (nvec <= 256) return go(std::integral_constant<int, 1>{});
  if (nvec <= 512) return go(std::integral_constant<int, 2>{});
  if (nvec <= 1024) return go(std::integral_constant<int, 4>{});
  return VLLM_EUNSUPPORTED;
}

int vllm_head_stack_bf16(const void* src, void* dst, int batch, int tokens, int parts, int heads, int head_dim,
                         int to_stacked, void* stream) {
  if (batch < 0 || tokens < 0 || parts <= 0 || heads <= 0 || head_dim <= 0) return VLLM_EINVAL;
  const long long n_vec = (long long)batch * tokens * parts * heads * (head_dim / 8);
  if (n_vec == 0) return VLLM_OK;
  if (!src || !dst) return VLLM_EINVAL;
  if (head_dim % 8) return VLLM_EUNSUPPORTED;
  if (!vllm_aligned(src, 16) || !vllm_aligned(dst, 16)) return VLLM_EALIGN;
  long long blocks = (n_vec + 255) / 256;
  const long long cap = (long long)vllm_num_sms() * 16;
  if (blocks > cap) blocks = cap;
  head_stack_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const uint4*)src, (uint4*)dst, batch, tokens, parts,
                                                                       heads, head_dim / 8, to_stacked ? 1 : 0, n_vec);
  VLLM_CHECK_LAUNCH();
  return VLLM_OK;
}

int vllm_swiglu_fwd_bf16(const void* gate_up, long long ldgu, void* h, long long ldh, long long rows, int inter, void* stream) {
  if (rows < 0 || inter <= 0 || inter % 8) return VLLM_EINVAL;
  if (rows == 0) return VLLM_OK;
  if (!gate_up || !h) return VLLM_EINVAL;
  if (ldgu % 8 || ldh % 8 || !vllm_aligned(gate_up, 16) || !vllm_aligned(h, 16)) return VLLM_EALIGN;
  long long blocks = (rows * (inter / 8) + 255) / 256;
  const long long cap = (long long)vllm_num_sms() * 16;
  if (blocks > cap) blocks = cap;
  swiglu_fwd_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)gate_up, ldgu, (__nv_bfloat16*)h, ldh,
                                                                        rows, inter);
  VLLM_CHECK_LAUNCH();
  return VLLM_OK;
}

int vllm_swiglu_bwd_bf16(const void* gate_up, long long ldgu, const void* dh, long long lddh, void* dgate_up, long long lddgu,
                         long long rows, int inter, void* stream) {
  if (rows < 0 || inter <= 0 || inter % 8) return VLLM_EINVAL;
  if (rows == 0) return VLLM_OK;
  if (!gate_up || !dh || !dgate_up) return VLLM_EINVAL;
  if (ldgu % 8 || lddh % 8 || lddgu % 8 || !vllm_aligned(gate_up, 16) || !vllm_aligned(dh, 16) || !vllm_aligned(dgate_up, 16))
    return VLLM_EALIGN;
  long long blocks = (rows * (inter / 8) + 255) / 256;
  const long long cap = (long long)vllm_num_sms() * 16;
  if (blocks > cap) blocks = cap;
  swiglu_bwd_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)gate_up, ldgu,
                                                                        (const __nv_bfloat16*)dh, lddh, (__nv_bfloat16*)dgate_up,
                                                                        lddgu, rows, inter);
  VLLM_CHECK_LAUNCH();
  return VLLM_OK;
}

int vllm_softmax_causal_bf16(void* s, long long ld, long long n_mat, int T, float scale, void* stream) {
  if (n_mat < 0 || T <= 0 || T % 8 || ld < T || ld % 8) return VLLM_EINVAL;
  if (n_mat == 0) return VLLM_OK;
  if (!s) return VLLM_EINVAL;
  if (!vllm_aligned(s, 16)) return VLLM_EALIGN;
  const long long rows = n_mat * T;
  if (rows > 2147483647LL) return VLLM_EUNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  const int nvec = T / 8;
  if (nvec <= 256) softmax_causal_kernel<1><<<(unsigned)rows, 256, 0, st>>>((__nv_bfloat16*)s, ld, T, scale);
  else if (nvec <= 512) softmax_causal_kernel<2><<<(unsigned)rows, 256, 0, st>>>((__nv_bfloat16*)s, ld, T, scale);
  else if (nvec <= 1024) softmax_causal_kernel<4><<<(unsigned)rows, 256, 0, st>>>((__nv_bfloat16*)s, ld, T, scale);
  else return VLLM_EUNSUPPORTED;
  VLLM_CHECK_LAUNCH();
  return VLLM_OK;
}

int vllm_attn_ds_bf16(const void* p, void* dp, long long ld, long long n_mat, int T, float scale, void* stream) {
  if (n_mat < 0 || T <= 0 || T % 8 || ld < T || ld % 8) return VLLM_EINVAL;
  if (n_mat == 0) return VLLM_OK;
  if (!p || !dp) return VLLM_EINVAL;
  if (!vllm_aligned(p, 16) || !vllm_aligned(dp, 16)) return VLLM_EALIGN;
  const long long rows = n_mat * T;
  if (rows > 2147483647LL) return VLLM_EUNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  const int nvec = T / 8;
  if (nvec <= 256) attn_ds_kernel<1><<<(unsigned)rows, 256, 0, st>>>((const __nv_bfloat16*)p, (__nv_bfloat16*)dp, ld, T, scale);
  else if (nvec <= 512) attn_ds_kernel<2><<<(unsigned)rows, 256, 0, st>>>((const __nv_bfloat16*)p, (__nv_bfloat16*)dp, ld, T, scale);
  else if (nvec <= 1024) attn_ds_kernel<4><<<(unsigned)rows, 256, 0, st>>>((const __nv_bfloat16*)p, (__nv_bfloat16*)dp, ld, T, scale);
  else return VLLM_EUNSUPPORTED;
  VLLM_CHECK_LAUNCH();
  return VLLM_OK;
}

int vllm_ce_loss_f32(const float* logits, long long ld, const int64_t* labels, const int64_t* n_valid, long long rows, int vocab,
                     float* loss_sum, void* dlogits, long long ldd, void* stream) {
  if (rows < 0 || vocab <= 0 || ld < vocab) return VLLM_EINVAL;
  if (rows == 0) return VLLM_OK;
  if (!logits || !labels || !loss_sum || (dlogits && (!n_valid || ldd < vocab))) return VLLM_EINVAL;
  if (rows > 2147483647LL) return VLLM_EUNSUPPORTED;
  ce_loss_kernel<<<(unsigned)rows, 512, 0, (cudaStream_t)stream>>>(logits, ld, labels, n_valid, vocab, loss_sum,
                                                                   (__nv_bfloat16*)dlogits, ldd);
  VLLM_CHECK_LAUNCH();
  return VLLM_OK;
}

}  // extern "C"
