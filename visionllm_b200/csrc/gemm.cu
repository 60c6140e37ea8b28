// bf16 GEMM with fused epilogue on the sm_100a tensor cores (tcgen05 + TMEM + TMA).
//
//   C[M, N] = epilogue( A[M, K] . B[N, K]^T )      A, B bf16 K-major (B is an nn.Linear weight [out, in])
//   epilogue: (+bias[N]) -> act -> (*colscale[N]) -> (+residual[M, N]) -> bf16 | fp32
//   act = SWIGLU pairs columns (2j, 2j+1) = (gate_j, up_j) and writes N/2 columns.
//
// This one kernel carries every dense projection on the hot path:
//   InternViT  qkv / proj(+bias, *ls1, +x) / fc1(+bias, GELU) / fc2(+bias, *ls2, +x)
//              (internvit/modeling_intern_vit.py:112,124,172-173,206-208), patch-embed as im2col GEMM (:73-85)
//   vl_bridge  Linear+GELU+Linear (modeling_visionllmv2.py:162-184)
//   LLM        q/k/v/o, gate|up (SwiGLU), down(+residual)  (HF LlamaDecoderLayer; internlm2/modeling_internlm2.py:235-360)
//   GDINO      value/offset/weight/output projections, FFN (ReLU)  (grounding_dino/...mask_dn.py:674-677,1116-1117)
//
// Structure (persistent, warp-specialised, one CTA or one CTA pair per SM):
//   warp 0   TMA producer: A tile 128x64 and B tile (256|128)x64 bf16, 128B swizzle, STAGES-deep mbarrier ring
//   warp 1   MMA issuer : one elected lane issues tcgen05.mma kind::f16, UMMA 128x256x16 (cta_group::1) or
//            256x256x16 over a CTA pair (cta_group::2); fp32 accumulators in TMEM, 2 accumulator stages
//            (2 x 256 columns) so the epilogue of tile i overlaps the main loop of tile i+1
//   warp 2   TMEM allocator
//   warps 4-11 epilogue (two warpgroups, one per 128-column half of the tile): tcgen05.ld 32 lanes x 32 columns ->
//            registers -> fused math -> 16-byte global stores
// Tiles are visited in groups of GROUP_M row-blocks so concurrently running CTAs share B (weights) in L2.
#include "common.cuh"
#include "tc_common.cuh"

namespace {

constexpr int BM = 128, BN = 256, BK = 64, ACC = 2, THREADS = 384, EPI_WARPS = 8;
constexpr int A_BYTES = BM * BK * 2;
constexpr int EPI_PITCH = 128 + 16;                 // 64 bf16 columns + 16 B pad: conflict-free 16-byte accesses
constexpr int EPI_STAGE_BYTES = 32 * EPI_PITCH;     // 32 rows per epilogue warp

enum { ACT_NONE = 0, ACT_GELU = 1, ACT_RELU = 2, ACT_SILU = 3, ACT_SWIGLU = 4, ACT_QUICKGELU = 5 };

struct GemmArgs {
  int M, N, K;
  void* C; int ldc;
  const __nv_bfloat16* bias; const __nv_bfloat16* colscale; const __nv_bfloat16* residual; int ldr;
  int act; int out_f32;
  const unsigned char* row_keep;   // optional [M]: rows with 0 are written as exact zeros (value.masked_fill of the MSDA module)
  int tiles_m, tiles_n;  // tiles_m counts 128*CG-row blocks
  int group_m;           // row-blocks per rasterisation group (see pick_group_m)
  // Implicit convolution over a zero-padded channels-last image (vllm_conv_rows_bf16): the K axis is a_segs segments
  // of a_seg_kb k-blocks; segment s reads A rows shifted by s * a_seg_rows (one image row of the padded map), and a
  // row of the A tensor map spans kw consecutive pixels (row pitch = C elements, row length = kw*C: rows overlap).
  int a_seg_kb, a_seg_rows;
  // Fused reduce-scatter push (vllm_gemm_bf16_scatter, tensor-parallel o_proj): row block d = row / sc_rows of C goes
  // to sc_dst[d] (a peer GPU's receive slot, row pitch ldc, local row = row - d * sc_rows); every epilogue warp bumps
  // sc_flag[d] once per tile after its stores (release at system scope).  sc_rows == 0: plain GEMM.
  int sc_rows;
  void* sc_dst[8];
  uint32_t* sc_flag[8];
  // MN-major operands (vllm_gemm_bf16_tn: the backward GEMMs -- dgrad reads the nn.Linear weight [N_red, K_out] as
  // B[j, k] = W[k, j], wgrad reads grad_output / activations [tokens, features] with the token axis as K): the operand
  // is a row-major [K, MN] matrix; a 64-k x 64-mn TMA box is one 128B-swizzled MN-major UMMA atom row, chunks of 64 mn
  // are 8 KB apart in the stage (LBO), 8 k-rows 1 KB apart (SBO), a k-step of 16 advances 2 KB.
  int a_mn, b_mn;
  // Block-diagonal batching (vllm_gemm_bf16_batched: the attention-backward GEMMs over all (batch, head) matrices of a
  // layer in one launch): every operand is a stack of `bt_rows`-row matrices along its row axis; the output row block
  // m0 belongs to matrix m0 / bt_rows and only meets that matrix's B rows / K range.  causal: 1 = skip output tiles
  // strictly above the diagonal (S = Q K^T, dP = dO V^T), 2 = the K range starts at the tile's first row (dV = P^T dO,
  // dK = dS^T Q: P, dS are zero below), 3 = the K range ends at the tile's last row (dQ = dS K).
  int bt_rows, causal;
};

// per-tile K range / operand offsets of the batched mode (identity when bt_rows == 0)
struct TilePlan { int skip, kb0, kb1, b_row_off, k_off; };
__device__ __forceinline__ TilePlan plan_tile(const GemmArgs& g, int m0, int n0, int rows_per_tile, int num_kb) {
  TilePlan p{0, 0, num_kb, 0, 0};
  if (g.bt_rows) {
    const int bh = m0 / g.bt_rows, ml = m0 - bh * g.bt_rows;
    p.b_row_off = bh * g.N;                             // K-major B: a stack of [N, K] matrices along the rows
    p.k_off = bh * g.K;                                 // MN-major operands: stacks of [K, M|N] matrices along the K rows
    if (g.causal == 1 && n0 >= ml + rows_per_tile) p.skip = 1;
    if (g.causal == 2) p.kb0 = ml / BK;
    if (g.causal == 3) { const int e = (ml + rows_per_tile + BK - 1) / BK; p.kb1 = e < num_kb ? e : num_kb; }
  }
  return p;
}

template <int CG> struct Cfg {
  static constexpr int B_ROWS = BN / CG;               // B rows held by one CTA
  static constexpr int B_BYTES = B_ROWS * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = CG == 1 ? 3 : 5;
  static constexpr int SMEM = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/ + 4 * BN * 4 /*bias, scale: double-buffered*/ +
                              EPI_WARPS * EPI_STAGE_BYTES /*per-warp store staging*/;
};

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float silu(float x) { return x / (1.f + __expf(-x)); }

__device__ __forceinline__ void tile_coords(int t, int tiles_m, int tiles_n, int group_m, int& tm, int& tn) {
  const int per_group = group_m * tiles_n;
  const int group = t / per_group;
  const int first = group * group_m;
  const int gsz = min(group_m, tiles_m - first);
  const int in = t - group * per_group;
  tm = first + in % gsz;
  tn = in / gsz;
}

template <int CG>
__global__ void __launch_bounds__(THREADS, 1)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                         const GemmArgs g) {
  using C_ = Cfg<CG>;
  constexpr int STAGES = C_::STAGES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (tc::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - tc::smem_u32(smem_raw));
  const uint32_t bar_base = smem_base + STAGES * C_::STAGE_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8 * s; };
  auto empty_bar = [&](int s) { return bar_base + 8 * (STAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8 * (2 * STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8 * (2 * STAGES + ACC + a); };
  const uint32_t tmem_slot = bar_base + 8 * (2 * STAGES + 2 * ACC);
  uint32_t* tmem_slot_gen = reinterpret_cast<uint32_t*>(smem_gen + STAGES * C_::STAGE_BYTES + 8 * (2 * STAGES + 2 * ACC));
  float* s_bs = reinterpret_cast<float*>(smem_gen + STAGES * C_::STAGE_BYTES + 256);   // [2][bias BN | scale BN]
  uint8_t* s_epi = reinterpret_cast<uint8_t*>(s_bs + 4 * BN);   // EPI_WARPS x 32 rows x EPI_PITCH bytes

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = CG == 2 ? tc::cluster_ctarank() : 0;
  const bool leader = rank == 0;
  const int cluster_id = blockIdx.x / CG, n_clusters = gridDim.x / CG;
  const int n_tiles = g.tiles_m * g.tiles_n;
  const int num_kb = (g.K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    tc::tma_prefetch_desc(&tmap_a);
    tc::tma_prefetch_desc(&tmap_b);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) { tc::mbar_init(full_bar(s), 1); tc::mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < ACC; ++a) { tc::mbar_init(tfull_bar(a), 1); tc::mbar_init(tempty_bar(a), EPI_WARPS * CG); }
    tc::mbar_fence_init();
  }
  if (warp == 2) tc::tmem_alloc<CG>(tmem_slot, ACC * BN);
  tc::tc_fence_before();
  if constexpr (CG == 2) tc::cluster_sync(); else __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (tc::elect_one()) {
      int stage = 0; uint32_t phase = 0;
      for (int t = cluster_id; t < n_tiles; t += n_clusters) {
        int tm, tn; tile_coords(t, g.tiles_m, g.tiles_n, g.group_m, tm, tn);
        const TilePlan tp = plan_tile(g, tm * CG * BM, tn * BN, CG * BM, num_kb);
        if (tp.skip) continue;
        const int bt_m0 = g.bt_rows ? (tm * CG * BM) / g.bt_rows * g.bt_rows : 0;
        // K-major A: global stacked row; MN-major A: column inside its matrix (the stack runs along the K rows)
        const int row_a = (tm * CG + (int)rank) * BM - (g.a_mn ? bt_m0 : 0);
        const int row_b = tn * BN + (int)rank * C_::B_ROWS + (g.b_mn ? 0 : tp.b_row_off);
        const int ka_off = g.a_mn ? tp.k_off : 0, kb_off = g.b_mn ? tp.k_off : 0;
        for (int kb = tp.kb0; kb < tp.kb1; ++kb) {
          tc::mbar_wait(empty_bar(stage), phase ^ 1);
          const uint32_t sa = smem_base + stage * C_::STAGE_BYTES, sb = sa + A_BYTES;
          int ka = kb * BK + ka_off, ra = row_a;
          if (g.a_seg_kb) {
            const int seg = kb / g.a_seg_kb;
            ka = (kb - seg * g.a_seg_kb) * BK;
            ra = row_a + seg * g.a_seg_rows;
          }
          if constexpr (CG == 1) {
            tc::mbar_arrive_expect_tx(full_bar(stage), C_::STAGE_BYTES);
            if (g.a_mn) {
              for (int h = 0; h < BM / 64; ++h) tc::tma_load_2d(sa + h * 8192, &tmap_a, full_bar(stage), ra + 64 * h, ka);
            } else {
              tc::tma_load_2d(sa, &tmap_a, full_bar(stage), ka, ra);
            }
            if (g.b_mn) {
              for (int h = 0; h < C_::B_ROWS / 64; ++h)
                tc::tma_load_2d(sb + h * 8192, &tmap_b, full_bar(stage), row_b + 64 * h, kb * BK + kb_off);
            } else {
              tc::tma_load_2d(sb, &tmap_b, full_bar(stage), kb * BK, row_b);
            }
          } else {
            if (leader) tc::mbar_arrive_expect_tx(full_bar(stage), 2 * C_::STAGE_BYTES);
            if (g.a_mn) {
              for (int h = 0; h < BM / 64; ++h) tc::tma_load_2d_cg2(sa + h * 8192, &tmap_a, full_bar(stage), ra + 64 * h, ka);
            } else {
              tc::tma_load_2d_cg2(sa, &tmap_a, full_bar(stage), ka, ra);
            }
            if (g.b_mn) {
              for (int h = 0; h < C_::B_ROWS / 64; ++h)
                tc::tma_load_2d_cg2(sb + h * 8192, &tmap_b, full_bar(stage), row_b + 64 * h, kb * BK + kb_off);
            } else {
              tc::tma_load_2d_cg2(sb, &tmap_b, full_bar(stage), kb * BK, row_b);
            }
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA) =====================
    if (leader) {
      const uint32_t idesc = tc::umma_idesc_bf16_f32(BM * CG, BN) | (g.a_mn ? (1u << 15) : 0u) | (g.b_mn ? (1u << 16) : 0u);
      int stage = 0; uint32_t phase = 0; int acc = 0; uint32_t acc_phase = 0;
      for (int t = cluster_id; t < n_tiles; t += n_clusters) {
        int tm_, tn_; tile_coords(t, g.tiles_m, g.tiles_n, g.group_m, tm_, tn_);
        const TilePlan tp = plan_tile(g, tm_ * CG * BM, tn_ * BN, CG * BM, num_kb);
        if (tp.skip) continue;
        tc::mbar_wait(tempty_bar(acc), acc_phase ^ 1);
        tc::tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = tp.kb0; kb < tp.kb1; ++kb) {
          tc::mbar_wait(full_bar(stage), phase);
          tc::tc_fence_after();
          if (tc::elect_one()) {
            const uint32_t sa = smem_base + stage * C_::STAGE_BYTES, sb = sa + A_BYTES;
            // K-major: a k-step of 16 bf16 advances the start address by 32 B inside the 128 B swizzle row;
            // MN-major: by 16 k-rows x 128 B = 2 KB (descriptor address units are 16 B)
            const uint64_t adesc = g.a_mn ? tc::umma_desc_mnmajor_sw128(sa) : tc::umma_desc_kmajor_sw128(sa);
            const uint64_t bdesc = g.b_mn ? tc::umma_desc_mnmajor_sw128(sb) : tc::umma_desc_kmajor_sw128(sb);
            const uint64_t astep = g.a_mn ? 128 : 2, bstep = g.b_mn ? 128 : 2;
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              tc::umma_f16<CG>(d_tmem, adesc + astep * k, bdesc + bstep * k, idesc, (kb != tp.kb0) || (k != 0));
            tc::umma_commit<CG>(empty_bar(stage));
            if (kb == tp.kb1 - 1) tc::umma_commit<CG>(tfull_bar(acc));
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if (++acc == ACC) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    // 8 epilogue warps: warp w reads TMEM lanes 32*(w%4).. (hardware rule); warps 4-7 take columns 0-127 of the
    // tile, warps 8-11 columns 128-255, so a tile drains twice as fast (matters when K is short: GDINO K = 256)
    const int quarter = warp & 3;
    const int col_half = (warp - 4) >> 2;
    const int et = threadIdx.x - 128;  // 0..255
    int acc = 0; uint32_t acc_phase = 0;
    const bool swiglu = g.act == ACT_SWIGLU;
    // bias / column scale of a tile live in shared memory (broadcast reads in the epilogue math); the NEXT tile's slice is
    // fetched into registers at the top of a tile and parked in the other buffer at its end, so the global-load latency
    // (~1 us, as long as a whole K = 256 main loop) never sits between two tiles (static_assert: one element per thread)
    static_assert(BN == 256, "one bias / scale element per epilogue thread");
    int staged_tile = -1, buf = 0;
    auto fetch_bs = [&](int tn_, float& b_, float& s_) {
      const int col = tn_ * BN + et;
      b_ = (g.bias && col < g.N) ? __bfloat162float(g.bias[col]) : 0.f;
      s_ = (g.colscale && col < g.N) ? __bfloat162float(g.colscale[col]) : 1.f;
    };
    for (int t = cluster_id; t < n_tiles; t += n_clusters) {
      int tm, tn; tile_coords(t, g.tiles_m, g.tiles_n, g.group_m, tm, tn);
      if (plan_tile(g, tm * CG * BM, tn * BN, CG * BM, num_kb).skip) continue;
      const int n0 = tn * BN;
      const int row = (tm * CG + (int)rank) * BM + quarter * 32 + lane;
      if (staged_tile != t) {                              // first tile (or the predicted successor was skipped)
        float b0, s0; fetch_bs(tn, b0, s0);
        asm volatile("bar.sync 1, 256;" ::: "memory");
        s_bs[buf * 2 * BN + et] = b0; s_bs[buf * 2 * BN + BN + et] = s0;
        asm volatile("bar.sync 1, 256;" ::: "memory");
      }
      float* s_bias = s_bs + buf * 2 * BN;
      float* s_scale = s_bias + BN;
      const int t_next = t + n_clusters;
      float nb = 0.f, ns = 1.f;
      if (t_next < n_tiles) { int tm2, tn2; tile_coords(t_next, g.tiles_m, g.tiles_n, g.group_m, tm2, tn2); fetch_bs(tn2, nb, ns); }
      tc::mbar_wait(tfull_bar(acc), acc_phase);
      tc::tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * BN;
      const bool row_ok = row < g.M;
      const bool zero_row = g.row_keep && row_ok && !g.row_keep[row];
      // the 32-column step of the general path: direct row-per-thread stores (fp32 output, SwiGLU, ragged N)
      auto direct32 = [&](int c) {
        const int col0 = n0 + c;
        if (col0 >= g.N) return;  // uniform
        uint32_t r[32];
        tc::tmem_ld_32x32(taddr + c, r);
        tc::tmem_ld_wait();
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) + s_bias[c + j];
        if (g.act == ACT_GELU) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = gelu_erf(v[j]);
        } else if (g.act == ACT_RELU) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
        } else if (g.act == ACT_SILU) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = silu(v[j]);
        } else if (g.act == ACT_QUICKGELU) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = v[j] / (1.f + __expf(-1.702f * v[j]));
        }
        if (swiglu) {
          // columns (2j, 2j+1) = (gate, up) -> 16 outputs at column col0/2
          if (row_ok) {
            const int oc0 = col0 >> 1;
            const int n_out = g.N >> 1;
            float o[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) o[j] = silu(v[2 * j]) * v[2 * j + 1];
            if (g.residual) {
              const __nv_bfloat16* rp = g.residual + (size_t)row * g.ldr + oc0;
#pragma unroll
              for (int j = 0; j < 16; ++j) if (oc0 + j < n_out) o[j] += __bfloat162float(rp[j]);
            }
            __nv_bfloat16* cp = reinterpret_cast<__nv_bfloat16*>(g.C) + (size_t)row * g.ldc + oc0;
            if (oc0 + 16 <= n_out) {
#pragma unroll
              for (int q = 0; q < 2; ++q) {
                uint4 pk;
                __nv_bfloat162 p0 = __floats2bfloat162_rn(o[8 * q + 0], o[8 * q + 1]);
                __nv_bfloat162 p1 = __floats2bfloat162_rn(o[8 * q + 2], o[8 * q + 3]);
                __nv_bfloat162 p2 = __floats2bfloat162_rn(o[8 * q + 4], o[8 * q + 5]);
                __nv_bfloat162 p3 = __floats2bfloat162_rn(o[8 * q + 6], o[8 * q + 7]);
                pk.x = *reinterpret_cast<uint32_t*>(&p0); pk.y = *reinterpret_cast<uint32_t*>(&p1);
                pk.z = *reinterpret_cast<uint32_t*>(&p2); pk.w = *reinterpret_cast<uint32_t*>(&p3);
                *reinterpret_cast<uint4*>(cp + 8 * q) = pk;
              }
            } else {
              for (int j = 0; j < 16; ++j) if (oc0 + j < n_out) cp[j] = __float2bfloat16(o[j]);
            }
          }
          return;
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] *= s_scale[c + j];
        if (row_ok) {
          const bool full = col0 + 32 <= g.N;
          if (g.residual) {
            const __nv_bfloat16* rp = g.residual + (size_t)row * g.ldr + col0;
            if (full) {
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const uint4 rr = *reinterpret_cast<const uint4*>(rp + 8 * q);
                const __nv_bfloat162* r2 = reinterpret_cast<const __nv_bfloat162*>(&rr);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const float2 f = __bfloat1622float2(r2[j]);
                  v[8 * q + 2 * j] += f.x; v[8 * q + 2 * j + 1] += f.y;
                }
              }
            } else {
              for (int j = 0; j < 32; ++j) if (col0 + j < g.N) v[j] += __bfloat162float(rp[j]);
            }
          }
          if (zero_row) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = 0.f;
          }
          if (g.out_f32) {
            float* cp = reinterpret_cast<float*>(g.C) + (size_t)row * g.ldc + col0;
            if (full) {
#pragma unroll
              for (int q = 0; q < 8; ++q)
                *reinterpret_cast<float4*>(cp + 4 * q) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
            } else {
              for (int j = 0; j < 32; ++j) if (col0 + j < g.N) cp[j] = v[j];
            }
          } else {
            __nv_bfloat16* cp = reinterpret_cast<__nv_bfloat16*>(g.C) + (size_t)row * g.ldc + col0;
            if (full) {
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                uint4 pk;
                __nv_bfloat162 p0 = __floats2bfloat162_rn(v[8 * q + 0], v[8 * q + 1]);
                __nv_bfloat162 p1 = __floats2bfloat162_rn(v[8 * q + 2], v[8 * q + 3]);
                __nv_bfloat162 p2 = __floats2bfloat162_rn(v[8 * q + 4], v[8 * q + 5]);
                __nv_bfloat162 p3 = __floats2bfloat162_rn(v[8 * q + 6], v[8 * q + 7]);
                pk.x = *reinterpret_cast<uint32_t*>(&p0); pk.y = *reinterpret_cast<uint32_t*>(&p1);
                pk.z = *reinterpret_cast<uint32_t*>(&p2); pk.w = *reinterpret_cast<uint32_t*>(&p3);
                *reinterpret_cast<uint4*>(cp + 8 * q) = pk;
              }
            } else {
              for (int j = 0; j < 32; ++j) if (col0 + j < g.N) cp[j] = __float2bfloat16(v[j]);
            }
          }
        }
            };
      uint8_t* my_stage = s_epi + (warp - 4) * EPI_STAGE_BYTES;
#pragma unroll 1
      for (int c = col_half * (BN / 2); c < (col_half + 1) * (BN / 2); c += 64) {
        const int col0 = n0 + c;
        if (col0 >= g.N) break;  // uniform
        if (swiglu || g.out_f32 || col0 + 64 > g.N) { direct32(c); direct32(c + 32); continue; }
        // ---- bf16 fast path: stage 32 rows x 64 columns per warp in smem, then store 128-byte row runs ----
        auto half = [&](const uint32_t (&r)[32], const int h) {
          float v[32];
          if (g.bias) {                      // 16-byte broadcast reads: 8 LDS.128 instead of 32 LDS.32
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const float4 b4 = *reinterpret_cast<const float4*>(&s_bias[c + 32 * h + 4 * q]);
              v[4 * q] = __uint_as_float(r[4 * q]) + b4.x; v[4 * q + 1] = __uint_as_float(r[4 * q + 1]) + b4.y;
              v[4 * q + 2] = __uint_as_float(r[4 * q + 2]) + b4.z; v[4 * q + 3] = __uint_as_float(r[4 * q + 3]) + b4.w;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
          }
          if (g.act == ACT_GELU) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = gelu_erf(v[j]);
          } else if (g.act == ACT_RELU) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
          } else if (g.act == ACT_SILU) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = silu(v[j]);
          } else if (g.act == ACT_QUICKGELU) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = v[j] / (1.f + __expf(-1.702f * v[j]));
          }
          if (g.colscale) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const float4 s4 = *reinterpret_cast<const float4*>(&s_scale[c + 32 * h + 4 * q]);
              v[4 * q] *= s4.x; v[4 * q + 1] *= s4.y; v[4 * q + 2] *= s4.z; v[4 * q + 3] *= s4.w;
            }
          }
          if (g.residual && row_ok) {
            const __nv_bfloat16* rp = g.residual + (size_t)row * g.ldr + col0 + 32 * h;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const uint4 rr = *reinterpret_cast<const uint4*>(rp + 8 * q);
              const __nv_bfloat162* r2 = reinterpret_cast<const __nv_bfloat162*>(&rr);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float2 f = __bfloat1622float2(r2[j]);
                v[8 * q + 2 * j] += f.x; v[8 * q + 2 * j + 1] += f.y;
              }
            }
          }
          if (zero_row) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = 0.f;
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            uint4 pk;
            __nv_bfloat162 p0 = __floats2bfloat162_rn(v[8 * q + 0], v[8 * q + 1]);
            __nv_bfloat162 p1 = __floats2bfloat162_rn(v[8 * q + 2], v[8 * q + 3]);
            __nv_bfloat162 p2 = __floats2bfloat162_rn(v[8 * q + 4], v[8 * q + 5]);
            __nv_bfloat162 p3 = __floats2bfloat162_rn(v[8 * q + 6], v[8 * q + 7]);
            pk.x = *reinterpret_cast<uint32_t*>(&p0); pk.y = *reinterpret_cast<uint32_t*>(&p1);
            pk.z = *reinterpret_cast<uint32_t*>(&p2); pk.w = *reinterpret_cast<uint32_t*>(&p3);
            *reinterpret_cast<uint4*>(my_stage + lane * EPI_PITCH + h * 64 + q * 16) = pk;
          }
                };
        {
          uint32_t r0[32], r1[32];
          tc::tmem_ld_32x32(taddr + c, r0);
          tc::tmem_ld_wait();
          tc::tmem_ld_32x32(taddr + c + 32, r1);     // in flight while the first half is processed
          half(r0, 0);
          tc::tmem_ld_wait();
          half(r1, 1);
        }
        __syncwarp();
        {
          const int row_base = (tm * CG + (int)rank) * BM + quarter * 32;
          __nv_bfloat16* cbase = reinterpret_cast<__nv_bfloat16*>(g.C) + col0;
          int row_local = row_base;
          if (g.sc_rows && row_base < g.M) {         // peer push: the 32 rows of a warp share one destination
            const int d = row_base / g.sc_rows;
            cbase = reinterpret_cast<__nv_bfloat16*>(g.sc_dst[d]) + col0;
            row_local = row_base - d * g.sc_rows;
          }
#pragma unroll
          for (int it = 0; it < 8; ++it) {           // each instruction: 4 rows x 128 contiguous bytes
            const int rr = it * 4 + (lane >> 3), piece = lane & 7;
            const uint4 val = *reinterpret_cast<const uint4*>(my_stage + rr * EPI_PITCH + piece * 16);
            if (row_base + rr < g.M)
              *reinterpret_cast<uint4*>(cbase + (size_t)(row_local + rr) * g.ldc + piece * 8) = val;
          }
        }
        __syncwarp();
      }
      // accumulator stage drained: hand it back to the MMA issuer
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (CG == 1) tc::mbar_arrive(tempty_bar(acc));
        else tc::mbar_arrive_cluster(tempty_bar(acc), 0);
      }
      if (g.sc_rows) {                               // tile pushed: publish it to the owner of these rows (after the
        const int row_base = (tm * CG + (int)rank) * BM + quarter * 32;   // TMEM hand-back: the fence waits for NVLink acks)
        __threadfence_system();
        __syncwarp();
        if (lane == 0 && row_base < g.M)
          asm volatile("red.release.sys.global.add.u32 [%0], %1;" ::"l"(g.sc_flag[row_base / g.sc_rows]), "r"(1u) : "memory");
      }
      if (++acc == ACC) { acc = 0; acc_phase ^= 1; }
      if (t_next < n_tiles) {                              // park the successor's bias / scale (nobody reads that buffer now)
        s_bs[(buf ^ 1) * 2 * BN + et] = nb; s_bs[(buf ^ 1) * 2 * BN + BN + et] = ns;
        asm volatile("bar.sync 1, 256;" ::: "memory");
        staged_tile = t_next; buf ^= 1;
      }
    }
  }

  // ===================== teardown =====================
  tc::tc_fence_before();
  if constexpr (CG == 2) tc::cluster_sync(); else __syncthreads();
  if (warp == 2) tc::tmem_dealloc<CG>(tmem_base, ACC * BN);
}

// Rasterisation: a group of `group_m` row-blocks sweeps all column-blocks before the next group starts, so a wave
// of 74 CTA pairs covers ~8 x 9 tiles (near-square = minimal A+B bytes per wave).  Measured (profiles/
// r1_gemm_raster_experiment.md): DRAM traffic and ncu durations move by < 3 % between group sizes 1..8 at the
// ViT shapes and get worse below 8 for the wide LLM gate|up GEMM, so 8 stays; the knob remains for sweeps.
int g_group_m_override = 0;
// SM budgets (0 = every SM): a persistent GEMM CTA owns its SM (168 registers x 384 threads = the whole register file), so a
// link-bound kernel of another stream -- the peer pushes of the tensor-parallel exchange -- can only overlap a GEMM on SMs
// the GEMM's grid leaves free.  g_sm_limit caps every launch, g_scatter_sm_limit the scatter GEMM (vllm_gemm_bf16_scatter),
// whose epilogue stores ride NVLink while the other micro-batch's GEMM runs beside it (visionllm_b200/tp.py).
int g_sm_limit = 0, g_scatter_sm_limit = 0;
int pick_group_m(int, int, long long) { return g_group_m_override > 0 ? g_group_m_override : 8; }

// 0: auto (cta_group::2 CTA pairs, 256x256 tiles; cta_group::1 when K <= 512, where the cross-CTA hand-offs of
// a pair cost more than the halved B traffic saves: 617 vs 494 TFLOP/s at K = 256); 1 / 2: force
int g_gemm_variant = 0;

template <int CG>
int launch_gemm(const void* A, int lda, const void* B, int ldb, GemmArgs g, cudaStream_t st, long long a_rows = -1,
                int a_cols = -1) {
  using C_ = Cfg<CG>;
  CUtensorMap ta, tb;
  const uint64_t nb = g.bt_rows ? (uint64_t)(g.M / g.bt_rows) : 1;     // matrices in the stack (batched mode)
  const uint64_t m_local = g.bt_rows ? (uint64_t)g.bt_rows : (uint64_t)g.M;
  int rc = g.a_mn ? vllm_make_tmap_bf16(&ta, A, nb * (uint64_t)g.K, m_local, (uint64_t)lda, 64)       // [K, M] rows, 64 x 64 boxes
                  : vllm_make_tmap_bf16(&ta, A, (uint64_t)(a_rows < 0 ? g.M : a_rows),
                                        (uint64_t)(a_cols < 0 ? g.K : a_cols), (uint64_t)lda, BM);
  if (rc) return rc;
  rc = g.b_mn ? vllm_make_tmap_bf16(&tb, B, nb * (uint64_t)g.K, (uint64_t)g.N, (uint64_t)ldb, 64)
              : vllm_make_tmap_bf16(&tb, B, nb * (uint64_t)g.N, (uint64_t)g.K, (uint64_t)ldb, C_::B_ROWS);
  if (rc) return rc;
  g.tiles_m = (g.M + BM * CG - 1) / (BM * CG);
  g.tiles_n = (g.N + BN - 1) / BN;
  const int n_tiles = g.tiles_m * g.tiles_n;
  int sms = vllm_num_sms();
  const int limit = g.sc_rows ? (g_scatter_sm_limit > 0 ? g_scatter_sm_limit : g_sm_limit) : g_sm_limit;
  if (limit > 0 && limit < sms) sms = limit < CG ? CG : limit;
  int clusters = sms / CG;
  if (clusters > n_tiles) clusters = n_tiles;
  g.group_m = pick_group_m(sms / CG, g.tiles_n, (long long)g.N * g.K * 2);
  static bool attr_set[3] = {false, false, false};
  if (!attr_set[CG]) {
    cudaError_t e = cudaFuncSetAttribute(gemm_bf16_tcgen05_kernel<CG>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         C_::SMEM);
    if (e != cudaSuccess) return (int)e;
    attr_set[CG] = true;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(clusters * CG);
  cfg.blockDim = dim3(THREADS);
  cfg.dynamicSmemBytes = C_::SMEM;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, gemm_bf16_tcgen05_kernel<CG>, ta, tb, g);
  if (e != cudaSuccess) return (int)e;
  return VLLM_OK;
}

}  // namespace

extern "C" {

int vllm_gemm_set_variant(int v) { g_gemm_variant = (v == 1 || v == 2) ? v : 0; return VLLM_OK; }
int vllm_gemm_set_group_m(int gm) { g_group_m_override = gm; return VLLM_OK; }
int vllm_gemm_set_sm_limit(int all_gemms, int scatter_gemm) {
  if (all_gemms < 0 || scatter_gemm < 0) return VLLM_EINVAL;
  g_sm_limit = all_gemms; g_scatter_sm_limit = scatter_gemm;
  return VLLM_OK;
}

static int gemm_bf16_common(const void* A, int lda, const void* B, int ldb, void* C, int ldc, int M, int N, int K,
                            const void* bias, const void* colscale, const void* residual, int ldr, int act, int out_f32,
                            const unsigned char* row_keep, void* stream);

int vllm_gemm_bf16(const void* A, int lda, const void* B, int ldb, void* C, int ldc, int M, int N, int K,
                   const void* bias, const void* colscale, const void* residual, int ldr, int act, int out_f32,
                   void* stream) {
  return gemm_bf16_common(A, lda, B, ldb, C, ldc, M, N, K, bias, colscale, residual, ldr, act, out_f32, nullptr, stream);
}

int vllm_gemm_bf16_rowmask(const void* A, int lda, const void* B, int ldb, void* C, int ldc, int M, int N, int K,
                           const void* bias, const void* colscale, const void* residual, int ldr, int act, int out_f32,
                           const unsigned char* row_keep, void* stream) {
  if (!row_keep || act == ACT_SWIGLU) return VLLM_EINVAL;
  return gemm_bf16_common(A, lda, B, ldb, C, ldc, M, N, K, bias, colscale, residual, ldr, act, out_f32, row_keep, stream);
}

static int gemm_bf16_common(const void* A, int lda, const void* B, int ldb, void* C, int ldc, int M, int N, int K,
                            const void* bias, const void* colscale, const void* residual, int ldr, int act, int out_f32,
                            const unsigned char* row_keep, void* stream) {
  if (M < 0 || N <= 0 || K <= 0 || lda < K || ldb < K) return VLLM_EINVAL;
  if (M == 0) return VLLM_OK;
  if (!A || !B || !C) return VLLM_EINVAL;
  if (act < 0 || act > ACT_QUICKGELU) return VLLM_EINVAL;
  const int n_out = act == ACT_SWIGLU ? N / 2 : N;
  if (act == ACT_SWIGLU && (N % 2 || out_f32 || colscale)) return VLLM_EUNSUPPORTED;
  if (ldc < n_out || (residual && ldr < n_out)) return VLLM_EINVAL;
  // TMA: 16-byte aligned bases and row pitches; vector epilogue: 16-byte aligned rows
  if (!vllm_aligned(A, 16) || !vllm_aligned(B, 16) || (lda % 8) || (ldb % 8)) return VLLM_EALIGN;
  const int celt = out_f32 ? 4 : 2;
  if (!vllm_aligned(C, 16) || ((size_t)ldc * celt) % 16 || (residual && (!vllm_aligned(residual, 16) || ldr % 8)))
    return VLLM_EALIGN;
  GemmArgs g{};
  g.M = M; g.N = N; g.K = K; g.C = C; g.ldc = ldc;
  g.bias = (const __nv_bfloat16*)bias; g.colscale = (const __nv_bfloat16*)colscale;
  g.residual = (const __nv_bfloat16*)residual; g.ldr = ldr; g.act = act; g.out_f32 = out_f32;
  g.row_keep = row_keep;
  cudaStream_t st = (cudaStream_t)stream;
  const int cg = g_gemm_variant ? g_gemm_variant : (K <= 512 ? 1 : 2);
  if (cg == 2) return launch_gemm<2>(A, lda, B, ldb, g, st);
  return launch_gemm<1>(A, lda, B, ldb, g, st);
}

int vllm_gemm_bf16_tn(const void* A, int lda, int a_mn_major, const void* B, int ldb, int b_mn_major, void* C, int ldc, int M,
                      int N, int K, int out_f32, void* stream) {
  // C[M, N] = sum_k A(m, k) * B(n, k) with either operand stored K-major ([rows = M|N, cols = K], like vllm_gemm_bf16) or
  // MN-major ([rows = K, cols = M|N], pitch lda / ldb): the backward GEMMs of a Linear y = x W^T without transposed copies --
  //   dgrad  dx[T, in]  = dy[T, out] (K-major A, K = out) x W[out, in] as MN-major B
  //   wgrad  dW[out, in] = dy[T, out] as MN-major A (K = T) x x[T, in] as MN-major B.
  if (M < 0 || N <= 0 || K <= 0) return VLLM_EINVAL;
  if (M == 0) return VLLM_OK;
  if (!A || !B || !C) return VLLM_EINVAL;
  if (lda < (a_mn_major ? M : K) || ldb < (b_mn_major ? N : K) || ldc < N) return VLLM_EINVAL;
  if (!vllm_aligned(A, 16) || !vllm_aligned(B, 16) || (lda % 8) || (ldb % 8)) return VLLM_EALIGN;
  if (!vllm_aligned(C, 16) || ((size_t)ldc * (out_f32 ? 4 : 2)) % 16) return VLLM_EALIGN;
  GemmArgs g{};
  g.M = M; g.N = N; g.K = K; g.C = C; g.ldc = ldc; g.out_f32 = out_f32;
  g.a_mn = a_mn_major ? 1 : 0; g.b_mn = b_mn_major ? 1 : 0;
  cudaStream_t st = (cudaStream_t)stream;
  const int cg = g_gemm_variant ? g_gemm_variant : (K <= 512 ? 1 : 2);
  if (cg == 2) return launch_gemm<2>(A, lda, B, ldb, g, st);
  return launch_gemm<1>(A, lda, B, ldb, g, st);
}

int vllm_gemm_bf16_batched(const void* A, int lda, int a_mn_major, const void* B, int ldb, int b_mn_major, void* C, int ldc,
                           int n_batch, int M, int N, int K, int causal, int out_f32, void* stream) {
  // n_batch independent products C_b[M, N] = A_b . B_b^T in ONE launch: every operand / the output is a stack of its
  // n_batch matrices along the row axis (K-major A: [n_batch*M, K]; MN-major A: [n_batch*K, M]; same for B; C: [n_batch*M, N]).
  // causal (square attention matrices, M == the sequence length): see GemmArgs.  The attention-backward GEMMs of a layer.
  if (n_batch < 0 || M <= 0 || N <= 0 || K <= 0 || causal < 0 || causal > 3) return VLLM_EINVAL;
  if (n_batch == 0) return VLLM_OK;
  if (!A || !B || !C) return VLLM_EINVAL;
  if (M % 256 || ((a_mn_major || b_mn_major) && K % BK)) return VLLM_EUNSUPPORTED;     // tiles must not straddle matrices
  if ((long long)n_batch * M > 2147483647LL || (long long)n_batch * K > 2147483647LL || (long long)n_batch * N > 2147483647LL)
    return VLLM_EUNSUPPORTED;
  if (lda < (a_mn_major ? M : K) || ldb < (b_mn_major ? N : K) || ldc < N) return VLLM_EINVAL;
  if (!vllm_aligned(A, 16) || !vllm_aligned(B, 16) || (lda % 8) || (ldb % 8)) return VLLM_EALIGN;
  if (!vllm_aligned(C, 16) || ((size_t)ldc * (out_f32 ? 4 : 2)) % 16) return VLLM_EALIGN;
  GemmArgs g{};
  g.M = n_batch * M; g.N = N; g.K = K; g.C = C; g.ldc = ldc; g.out_f32 = out_f32;
  g.a_mn = a_mn_major ? 1 : 0; g.b_mn = b_mn_major ? 1 : 0;
  g.bt_rows = M; g.causal = causal;
  cudaStream_t st = (cudaStream_t)stream;
  const int cg = g_gemm_variant ? g_gemm_variant : (K <= 512 ? 1 : 2);
  if (cg == 2) return launch_gemm<2>(A, lda, B, ldb, g, st);
  return launch_gemm<1>(A, lda, B, ldb, g, st);
}

int vllm_gemm_bf16_scatter(const void* A, int lda, const void* B, int ldb, void* const* dst, void* const* flags,
                           int n_dst, int rows_per_dst, int ldc, int N, int K, void* stream) {
  // C = A . B^T (plain bf16, no epilogue math), M = n_dst * rows_per_dst; row block d is stored to dst[d] and
  // counted on flags[d] (8 arrivals per 128 x 256 tile: (rows_per_dst / 128) * ceil(N / 256) * 8 per source and pass).
  if (n_dst <= 0 || n_dst > 8 || rows_per_dst <= 0 || N <= 0 || K <= 0 || lda < K || ldb < K || ldc < N) return VLLM_EINVAL;
  if (!A || !B || !dst || !flags) return VLLM_EINVAL;
  if (rows_per_dst % BM || N % 64) return VLLM_EUNSUPPORTED;    // a warp's 32 rows share a destination; bf16 fast path
  if ((long long)n_dst * rows_per_dst > 2147483647LL) return VLLM_EUNSUPPORTED;
  if (!vllm_aligned(A, 16) || !vllm_aligned(B, 16) || (lda % 8) || (ldb % 8) || (ldc % 8)) return VLLM_EALIGN;
  GemmArgs g{};
  g.M = n_dst * rows_per_dst; g.N = N; g.K = K; g.ldc = ldc; g.sc_rows = rows_per_dst;
  for (int d = 0; d < n_dst; ++d) {
    if (!dst[d] || !flags[d]) return VLLM_EINVAL;
    if (!vllm_aligned(dst[d], 16)) return VLLM_EALIGN;
    g.sc_dst[d] = dst[d]; g.sc_flag[d] = (uint32_t*)flags[d];
  }
  g.C = dst[0];
  cudaStream_t st = (cudaStream_t)stream;
  const int cg = g_gemm_variant ? g_gemm_variant : (K <= 512 ? 1 : 2);
  if (cg == 2) return launch_gemm<2>(A, lda, B, ldb, g, st);
  return launch_gemm<1>(A, lda, B, ldb, g, st);
}

int vllm_conv_rows_bf16(const void* xpad, long long pad_pixels, int channels, int padded_width, int kernel_h,
                        int kernel_w, const void* weight, int ldw, void* out, int ldo, int out_channels,
                        const void* bias, int act, void* stream) {
  // out[i, :] = epi(sum_{dy,dx} xpad[i + dy*padded_width + dx, :] . weight[:, (dy*kw + dx)*C : +C]^T) for every flat
  // pixel i of the padded map; rows whose window crosses an image edge are don't-care and sliced away by the caller.
  if (pad_pixels < 0 || channels <= 0 || padded_width <= 0 || kernel_h <= 0 || kernel_w <= 0 || out_channels <= 0)
    return VLLM_EINVAL;
  if (pad_pixels == 0) return VLLM_OK;
  if (!xpad || !weight || !out) return VLLM_EINVAL;
  if (act < 0 || act > ACT_QUICKGELU || act == ACT_SWIGLU) return VLLM_EINVAL;
  const long long seg_k = (long long)kernel_w * channels;
  // a segment must be whole k-blocks so that A's and B's K coordinates stay aligned; 16-byte row pitches
  if (seg_k % BK || channels % 8 || ldw % 8 || ldo % 8 || ldw < seg_k * kernel_h || ldo < out_channels) return VLLM_EUNSUPPORTED;
  if (pad_pixels > 2147483647LL - 4096 || seg_k * kernel_h > 2147483647LL) return VLLM_EUNSUPPORTED;
  if (!vllm_aligned(xpad, 16) || !vllm_aligned(weight, 16) || !vllm_aligned(out, 16)) return VLLM_EALIGN;
  const long long a_rows = pad_pixels - (kernel_w - 1);      // last rows whose kw-pixel span stays inside the buffer
  if (a_rows <= 0) return VLLM_EINVAL;
  GemmArgs g{};
  g.M = (int)pad_pixels; g.N = out_channels; g.K = (int)(seg_k * kernel_h); g.C = out; g.ldc = ldo;
  g.bias = (const __nv_bfloat16*)bias; g.act = act;
  g.a_seg_kb = (int)(seg_k / BK); g.a_seg_rows = padded_width;
  cudaStream_t st = (cudaStream_t)stream;
  const int cg = g_gemm_variant ? g_gemm_variant : (g.K <= 512 ? 1 : 2);
  if (cg == 2) return launch_gemm<2>(xpad, channels, weight, ldw, g, st, a_rows, (int)seg_k);
  return launch_gemm<1>(xpad, channels, weight, ldw, g, st, a_rows, (int)seg_k);
}

}  // extern "C"
