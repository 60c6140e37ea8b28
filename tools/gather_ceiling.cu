// Microbenchmark: how many 128-byte lines per clock can one B200 SM pull through L1 / shared memory with the MSDA
// access pattern (a warp instruction = 32 lanes x 16 B covering FOUR independent 128-byte rows)?  This is the bound
// DESIGN.md 6.2 argues from: the fp32 gather needs 64 such rows per (query, head), whatever the kernel around it does.
//
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o build/gather_ceiling tools/gather_ceiling.cu
//   build/gather_ceiling            -> one JSON line
//
// Variants (all: 512 threads per CTA, 2 CTAs per SM, no math besides the accumulate):
//   l1_lines    LDG.128, random rows inside a 32 KB window per CTA (L1-resident after the first touch)
//   l1_lines_u8 the same, 8 independent loads in flight per lane
//   smem_lines  LDS.128 from a 32 KB shared-memory window, same row pattern (conflict-free: a row = all 32 banks)
//   smem_half   LDS.128, 64-byte rows (bf16 D=32): 8 rows per instruction, horizontally adjacent pairs -> 4 wavefronts
//   l2_lines    LDG.128, random rows inside a 16 MB tensor (~ the value tensor of one image; L2-resident, L1-missing)
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned hash32(unsigned x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x;
}

template <int UNROLL>
__global__ void __launch_bounds__(512, 2) l1_lines_kernel(const float4* __restrict__ buf, size_t window_rows,
                                                         size_t total_rows, int iters, float* sink, int spread) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned grp = lane >> 3, cq = lane & 7;          // 4 rows per instruction, 8 x 16 B per row
  const size_t base = spread ? 0 : ((size_t)blockIdx.x * window_rows) % (total_rows - window_rows);
  const size_t span = spread ? total_rows : window_rows;
  float4 acc = make_float4(0, 0, 0, 0);
  unsigned s = hash32(blockIdx.x * 977u + warp * 131u + grp);
  for (int i = 0; i < iters; i += UNROLL) {
    float4 v[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      s = s * 1664525u + 1013904223u;
      const size_t row = base + (size_t)((s >> 9) & (unsigned)(span - 1));     // span is a power of two
      v[u] = __ldg(buf + row * 8 + cq);
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) { acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w; }
  }
  if (acc.x + acc.y + acc.z + acc.w == 123.456f) sink[0] = acc.x;
}

template <int ROWB>   // row bytes: 128 (fp32 D=32) or 64 (bf16 D=32)
__global__ void __launch_bounds__(512, 2) smem_lines_kernel(const float4* __restrict__ buf, int iters, float* sink) {
  extern __shared__ float4 sm[];                           // 32 KB window
  constexpr int ROWS = 32768 / ROWB;
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) sm[i] = buf[(size_t)blockIdx.x * 2048 + i];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int LPR = ROWB / 16;                           // lanes per row
  const unsigned grp = lane / LPR, cq = lane % LPR;
  float4 acc = make_float4(0, 0, 0, 0);
  // 64-byte rows: lanes 2k, 2k+1 groups take a horizontally adjacent PAIR of rows (the two x-corners of a sample),
  // which always covers complementary bank halves -> conflict-free quarter-warps
  unsigned s = hash32(blockIdx.x * 977u + warp * 131u + (ROWB == 64 ? (grp >> 1) : grp));
  for (int i = 0; i < iters; i += 4) {
    float4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      s = s * 1664525u + 1013904223u;
      unsigned row = (s >> 9) & (unsigned)(ROWS - 1);
      if (ROWB == 64) row = (row + (grp & 1)) & (unsigned)(ROWS - 1);
      v[u] = sm[row * LPR + cq];
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) { acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w; }
  }
  if (acc.x + acc.y + acc.z + acc.w == 123.456f) sink[0] = acc.x;
}

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("{\"error\": \"%s\"}\n", cudaGetErrorString(e)); return 1; } } while (0)

template <typename F> static float time_ms(F f, int reps) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(a);
  for (int r = 0; r < reps; ++r) f();
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  return ms / reps;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  const int sms = p.multiProcessorCount, ctas = sms * 2, iters = 4096;
  const size_t total_rows = (size_t)22282240 / 128 * 8;    // 8 images' worth of value rows (178 MB)
  float4* buf; float* sink;
  CK(cudaMalloc(&buf, total_rows * 128)); CK(cudaMalloc(&sink, 4));
  CK(cudaMemset(buf, 0, total_rows * 128));
  CK(cudaFuncSetAttribute(smem_lines_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768));
  CK(cudaFuncSetAttribute(smem_lines_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768));
  const double instr = (double)ctas * 16 * iters;          // warp-level load instructions per launch
  auto report = [&](const char* name, float ms, double rows_per_instr, double bytes_per_instr, bool last) {
    const double rows = instr * rows_per_instr;
    printf("\"%s\": {\"ms\": %.4f, \"rows_per_s\": %.4g, \"rows_per_clk_per_sm_at_max_clock\": %.3f, \"TBps\": %.3f}%s",
           name, ms, rows / (ms * 1e-3), rows / (ms * 1e-3) / sms / (clk_khz * 1e3), instr * bytes_per_instr / (ms * 1e-3) / 1e12,
           last ? "" : ", ");
  };
  printf("{\"device\": \"%s\", \"sms\": %d, \"max_clock_mhz\": %.0f, \"ctas\": %d, \"threads_per_cta\": 512, ", p.name, sms,
         clk_khz / 1e3, ctas);
  float ms;
  ms = time_ms([&] { l1_lines_kernel<4><<<ctas, 512>>>(buf, 256, total_rows, iters, sink, 0); }, 5);
  report("l1_lines_128B_rows_ldg128_u4", ms, 4, 512, false);
  ms = time_ms([&] { l1_lines_kernel<8><<<ctas, 512>>>(buf, 256, total_rows, iters, sink, 0); }, 5);
  report("l1_lines_128B_rows_ldg128_u8", ms, 4, 512, false);
  ms = time_ms([&] { smem_lines_kernel<128><<<ctas, 512, 32768>>>(buf, iters, sink); }, 5);
  report("smem_128B_rows_lds128", ms, 4, 512, false);
  ms = time_ms([&] { smem_lines_kernel<64><<<ctas, 512, 32768>>>(buf, iters, sink); }, 5);
  report("smem_64B_rows_lds128", ms, 8, 512, false);
  ms = time_ms([&] { l1_lines_kernel<8><<<ctas, 512>>>(buf, 256, 131072, iters / 4, sink, 1); }, 5);
  { const double save = instr; (void)save; }
  printf("\"l2_lines_128B_rows_ldg128_u8_16MB\": {\"ms\": %.4f, \"rows_per_s\": %.4g, \"TBps\": %.3f}", ms,
         instr / 4 * 4 / (ms * 1e-3), instr / 4 * 512 / (ms * 1e-3) / 1e12);
  CK(cudaGetLastError());
  printf("}\n");
  return 0;
}
