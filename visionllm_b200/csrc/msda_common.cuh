// Shared by msda.cu and msda_win.cu: the sampling-index arithmetic of the reference kernel (bit-exact contract,
// SURVEY.md 8a-a17), the query-tiling descriptor and the streaming loads of sampling_loc / attn_weight.
#pragma once
#include "common.cuh"
#include <limits.h>

#define MSDA_MAX_LEVELS 8

struct MsdaTiling {
  int mode;                          // 0: tiles of consecutive queries; 1: 2-D pixel patches
  int n_tiles;                       // tiles per (batch, head)
  int tile_start[MSDA_MAX_LEVELS + 1];
  int H[MSDA_MAX_LEVELS], W[MSDA_MAX_LEVELS];
  int q_start[MSDA_MAX_LEVELS];
  int tiles_w[MSDA_MAX_LEVELS];
};

// ---------------------------------------------------------------------------
// Index arithmetic shared by every path (reference .cuh:238-241 and :22-29).
// ---------------------------------------------------------------------------
template <typename T> struct MsdaGeom {
  int h_low, w_low;
  T lh, lw;
  int mask;  // bit0: sample in range; bits1..4: corner (ll, lh, hl, hh) in bounds
};

__device__ __forceinline__ float msda_mul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float msda_sub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float msda_add(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ double msda_mul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double msda_sub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double msda_add(double a, double b) { return __dadd_rn(a, b); }

template <typename T>
__device__ __forceinline__ MsdaGeom<T> msda_geom(T loc_w, T loc_h, int H, int W) {
  MsdaGeom<T> g;
  const T h_im = msda_sub(msda_mul(loc_h, (T)H), (T)0.5);
  const T w_im = msda_sub(msda_mul(loc_w, (T)W), (T)0.5);
  g.mask = 0; g.h_low = 0; g.w_low = 0; g.lh = 0; g.lw = 0;
  if (h_im > (T)-1 && w_im > (T)-1 && h_im < (T)H && w_im < (T)W) {
    // The reference calls floorf() for every scalar type (kernel.cuh:22-23).
    const int h_low = (int)floorf((float)h_im);
    const int w_low = (int)floorf((float)w_im);
    g.h_low = h_low; g.w_low = w_low;
    g.lh = msda_sub(h_im, (T)h_low);
    g.lw = msda_sub(w_im, (T)w_low);
    const int h_high = h_low + 1, w_high = w_low + 1;
    int m = 1;
    if (h_low >= 0 && w_low >= 0) m |= 2;
    if (h_low >= 0 && w_high <= W - 1) m |= 4;
    if (h_high <= H - 1 && w_low >= 0) m |= 8;
    if (h_high <= H - 1 && w_high <= W - 1) m |= 16;
    g.mask = m;
  }
  return g;
}

__device__ __forceinline__ float2 ld_stream_f2(const float* p) {
  float2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0, %1}, [%2];" : "=f"(r.x), "=f"(r.y) : "l"(p));
  return r;
}
__device__ __forceinline__ float ld_stream_f1(const float* p) {
  float r;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
  return r;
}

