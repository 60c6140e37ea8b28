// Fused attention on the sm_100a tensor cores (tcgen05 + TMEM + TMA), head_dim 128.
//
// Same operator as csrc/attention.cu (softmax(Q K^T * scale [+ causal / key-length mask]) V; replaces
// flash_attn_varlen_qkvpacked_func, internvit/flash_attention.py:51-54, and the HF-Llama / InternLM2 attention,
// internlm2/modeling_internlm2.py:362-546); this is the Blackwell-native dataflow:
//
//   CTA = 256 query rows of one (batch, head): two 128-row Q tiles "ping-pong" so the tensor pipe works on one
//   tile while the other tile's softmax runs.
//   warp 0      TMA producer: Q0,Q1 once; K_j, V_j tiles of 128 keys, 2-stage ring (3-D tensor maps
//               [batch, token, column] so rows past the sequence end are zero-filled, never the next image)
//   warp 1      MMA issuer (one lane):  S_i = Q_i K_j^T   (SS form, both operands K-major in 128B-swizzled smem)
//                                       O_i = P_i V_j     (TS form: A = P_i read from TMEM, B = V_j MN-major smem)
//   warp 2      TMEM allocator (512 columns: S0 | S1 | O0 | O1; P_i aliases the first 64 columns of S_i)
//   warps 4-7   softmax warpgroup of Q tile 0, warps 8-11 of Q tile 1: ONE THREAD PER QUERY ROW (TMEM lane = row),
//               so row max / row sum need no shuffles: tcgen05.ld S (once, 128 registers) -> max -> exp2 -> bf16 P
//               -> tcgen05.st.  O_i accumulates IN TMEM across KV tiles (PV MMA with accumulate); the online-softmax
//               rescale of O is lazy: a row keeps a stale reference max until the true max has grown by more
//               than 2^8, and only then (warp-uniform vote) O_i is loaded, scaled and stored back.
//   Scores and probabilities never leave the SM; HBM traffic = Q, K, V read + O written.
#include "common.cuh"
#include "tc_common.cuh"

namespace {

constexpr int D = 128, BQ = 128, BKV = 128, KV_STAGES = 2;
constexpr int TILE_BYTES = 128 * 128 * 2;          // one 128x128 bf16 tile = two 64-column swizzled halves
constexpr int HALF_BYTES = TILE_BYTES / 2;
constexpr int THREADS = 384;
constexpr int SMEM = 2 * TILE_BYTES + 2 * KV_STAGES * TILE_BYTES + 1024 + 256;

struct AttnTcArgs {
  __nv_bfloat16* o;
  long long o_bs, o_ts;
  const int* seqlens;
  int Tq, Tk, heads, kv_heads, causal;
  float scale_log2;
};

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(lbo_bytes >> 4) << 16;
  d |= (uint64_t)(sbo_bytes >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// D[tmem] (+)= A[tmem] * B[smem desc]
__device__ __forceinline__ void umma_ts_f16(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st_32x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// 2^x on the FMA/ALU pipes (Cody-Waite split + degree-3 minimax, max rel. error 1.0e-4).  Tried for half of the
// exponentials to relieve the MUFU unit (FA4's trick): measured SLOWER here (598 vs 678 TFLOP/s at the ViT shape,
// profiles/r1_attn_bench.json history) -- the softmax warps are issue/latency-bound, not MUFU-bound, at two warps
// per scheduler -- so it is kept for reference but not used.
__device__ __forceinline__ float ex2_poly(float x) {
  x = fmaxf(x, -126.f);
  const float xr = x + 12582912.f;                 // 1.5 * 2^23: low mantissa bits now hold round(x)
  const float f = x - (xr - 12582912.f);           // f in [-0.5, 0.5]
  float p = fmaf(0.055008938f, f, 0.24221096f);
  p = fmaf(p, f, 0.69328293f);
  p = fmaf(p, f, 1.0f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(xr) << 23));
}
template <int N> __device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N> __device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }

__global__ void __launch_bounds__(THREADS, 1)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                   const __grid_constant__ CUtensorMap tm_v, const AttnTcArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (tc::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - tc::smem_u32(smem_raw));
  const uint32_t sQ = base;                                  // Q0 | Q1
  const uint32_t sK = sQ + 2 * TILE_BYTES;                   // K stages
  const uint32_t sV = sK + KV_STAGES * TILE_BYTES;           // V stages
  const uint32_t bar = sV + KV_STAGES * TILE_BYTES;
  // barrier map (8 bytes each)
  const uint32_t q_full = bar;
  auto k_full = [&](int s) { return bar + 8 * (1 + s); };
  auto v_full = [&](int s) { return bar + 8 * (3 + s); };
  auto k_empty = [&](int s) { return bar + 8 * (5 + s); };
  auto v_empty = [&](int s) { return bar + 8 * (7 + s); };
  auto s_full = [&](int i) { return bar + 8 * (9 + i); };
  auto p_ready = [&](int i) { return bar + 8 * (11 + i); };
  auto o_full = [&](int i) { return bar + 8 * (13 + i); };
  auto o_free = [&](int i) { return bar + 8 * (15 + i); };
  const uint32_t tmem_slot = bar + 8 * 17;
  uint32_t* tmem_slot_gen = reinterpret_cast<uint32_t*>(gen + (tmem_slot - base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 2 * BQ, head = blockIdx.y, b = blockIdx.z;
  const int kvh = head / (a.heads / a.kv_heads);
  const int len = a.seqlens ? min(a.seqlens[b], a.Tk) : a.Tk;
  const int coff = a.Tk - a.Tq;
  int k_end = len;
  if (a.causal) k_end = min(k_end, q0 + 2 * BQ + coff);
  const int n_tiles = k_end > 0 ? (k_end + BKV - 1) / BKV : 0;

  if (warp == 0 && lane == 0) {
    tc::tma_prefetch_desc(&tm_q); tc::tma_prefetch_desc(&tm_k); tc::tma_prefetch_desc(&tm_v);
  }
  if (warp == 1 && lane == 0) {
    tc::mbar_init(q_full, 1);
    for (int s = 0; s < KV_STAGES; ++s) {
      tc::mbar_init(k_full(s), 1); tc::mbar_init(v_full(s), 1); tc::mbar_init(k_empty(s), 1); tc::mbar_init(v_empty(s), 1);
    }
    for (int i = 0; i < 2; ++i) {
      tc::mbar_init(s_full(i), 1); tc::mbar_init(p_ready(i), 128); tc::mbar_init(o_full(i), 1); tc::mbar_init(o_free(i), 128);  /* unused since O accumulates in TMEM */
    }
    tc::mbar_fence_init();
  }
  if (warp == 2) tc::tmem_alloc<1>(tmem_slot, 512);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = *tmem_slot_gen;

  if (warp < 4) {
    reg_dec<56>();
    if (warp == 0) {
      // ===================== TMA producer =====================
      if (tc::elect_one() && n_tiles > 0) {
        tc::mbar_arrive_expect_tx(q_full, 2 * TILE_BYTES);
        for (int i = 0; i < 2; ++i)
          for (int h = 0; h < 2; ++h)
            tma_load_3d(sQ + i * TILE_BYTES + h * HALF_BYTES, &tm_q, q_full, head * D + h * 64, q0 + i * BQ, b);
        for (int j = 0; j < n_tiles; ++j) {
          const int s = j % KV_STAGES;
          const uint32_t ph = ((j / KV_STAGES) & 1) ^ 1;
          tc::mbar_wait(k_empty(s), ph);
          tc::mbar_arrive_expect_tx(k_full(s), TILE_BYTES);
          for (int h = 0; h < 2; ++h)
            tma_load_3d(sK + s * TILE_BYTES + h * HALF_BYTES, &tm_k, k_full(s), kvh * D + h * 64, j * BKV, b);
          tc::mbar_wait(v_empty(s), ph);
          tc::mbar_arrive_expect_tx(v_full(s), TILE_BYTES);
          for (int h = 0; h < 2; ++h)
            tma_load_3d(sV + s * TILE_BYTES + h * HALF_BYTES, &tm_v, v_full(s), kvh * D + h * 64, j * BKV, b);
        }
      }
      __syncwarp();
    } else if (warp == 1) {
      // ===================== MMA issuer =====================
      constexpr uint32_t idesc_s = tc::umma_idesc_bf16_f32(BQ, BKV);
      constexpr uint32_t idesc_pv = tc::umma_idesc_bf16_f32(BQ, D) | (1u << 16);   // B (= V) is MN-major
      auto issue_s = [&](int i, int j) {   // S_i = Q_i K_j^T
        const int s = j % KV_STAGES;
        const uint32_t qa = sQ + i * TILE_BYTES, ka = sK + s * TILE_BYTES;
#pragma unroll
        for (int kk = 0; kk < D / 16; ++kk) {
          const uint32_t off = (kk >> 2) * HALF_BYTES + (kk & 3) * 32;
          tc::umma_f16<1>(tmem + i * 128, umma_desc_sw128(qa + off, 16, 1024), umma_desc_sw128(ka + off, 16, 1024),
                          idesc_s, kk != 0);
        }
        tc::umma_commit<1>(s_full(i));
      };
      auto issue_pv = [&](int i, int j) {  // O_i (+)= P_i V_j, accumulator lives in TMEM across tiles
        const int s = j % KV_STAGES;
        const uint32_t va = sV + s * TILE_BYTES;
#pragma unroll
        for (int kk = 0; kk < BKV / 16; ++kk)
          umma_ts_f16(tmem + 256 + i * 128, tmem + i * 128 + kk * 8,
                      umma_desc_sw128(va + kk * 2048, HALF_BYTES, 1024), idesc_pv, (j | kk) != 0);
        tc::umma_commit<1>(o_full(i));
      };
      if (n_tiles > 0) {
        tc::mbar_wait(q_full, 0);
        tc::mbar_wait(k_full(0), 0);
        tc::tc_fence_after();
        if (tc::elect_one()) { issue_s(0, 0); issue_s(1, 0); tc::umma_commit<1>(k_empty(0)); }
        __syncwarp();
        for (int j = 0; j < n_tiles; ++j) {
          const int s = j % KV_STAGES;
          const uint32_t kv_ph = (j / KV_STAGES) & 1, jp = j & 1;
          const bool more = j + 1 < n_tiles;
          tc::mbar_wait(v_full(s), kv_ph);
          tc::mbar_wait(p_ready(0), jp);
          tc::tc_fence_after();
          if (tc::elect_one()) issue_pv(0, j);
          __syncwarp();
          if (more) {
            tc::mbar_wait(k_full((j + 1) % KV_STAGES), ((j + 1) / KV_STAGES) & 1);
            tc::tc_fence_after();
            if (tc::elect_one()) issue_s(0, j + 1);
            __syncwarp();
          }
          tc::mbar_wait(p_ready(1), jp);
          tc::tc_fence_after();
          if (tc::elect_one()) {
            issue_pv(1, j);
            tc::umma_commit<1>(v_empty(s));
            if (more) { issue_s(1, j + 1); tc::umma_commit<1>(k_empty((j + 1) % KV_STAGES)); }
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ===================== softmax warpgroups: one thread per query row =====================
    reg_inc<200>();
    const int i = (warp - 4) >> 2;                    // Q tile of this warpgroup
    const int quarter = warp & 3;
    const int row = q0 + i * BQ + quarter * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    const uint32_t tS = tmem + lane_addr + i * 128, tO = tmem + lane_addr + 256 + i * 128;
    constexpr float RESCALE_THRESHOLD = 8.f;          // log2 domain: tolerate a 2^8 stale reference max
    float m = -INFINITY, l = 0.f;
    for (int j = 0; j < n_tiles; ++j) {
      const uint32_t jp = j & 1;
      tc::mbar_wait(s_full(i), jp);
      tc::tc_fence_after();
      uint32_t r[BKV];
#pragma unroll
      for (int c = 0; c < BKV; c += 32) tc::tmem_ld_32x32(tS + c, *reinterpret_cast<uint32_t(*)[32]>(&r[c]));
      tc::tmem_ld_wait();
      const int n0 = j * BKV;
      const bool need_mask = (n0 + BKV > len) || (a.causal && (n0 + BKV - 1 > q0 + i * BQ + coff));
      const int lim = (a.causal ? min(len, row + coff + 1) : len) - n0;   // tile-local keys < lim are visible
      if (need_mask) {
#pragma unroll
        for (int k = 0; k < BKV; ++k) if (k >= lim) r[k] = 0xff800000u;   // -inf
      }
      float mx = -INFINITY;
#pragma unroll
      for (int k = 0; k < BKV; k += 2) mx = fmaxf(mx, fmaxf(__uint_as_float(r[k]), __uint_as_float(r[k + 1])));
      const float m_tile = mx * a.scale_log2;
      // lazy online-softmax rescale (first tile: just adopt the max, O is overwritten by the first PV MMA)
      const bool grow = m_tile > m + RESCALE_THRESHOLD;            // also true while m == -inf and the tile has a key
      if (j == 0) {
        m = m_tile;
      } else if (__any_sync(0xffffffffu, grow)) {
        const float m_new = grow ? m_tile : m;
        const float corr = (m == -INFINITY) ? 0.f : ex2(m - m_new);   // rows that do not grow: corr == 1
        l *= corr;
        m = m_new;
        tc::mbar_wait(o_full(i), jp ^ 1);                              // PV of tile j-1 has landed in O_i
        tc::tc_fence_after();
#pragma unroll
        for (int c = 0; c < D; c += 32) {
          uint32_t o[32];
          tc::tmem_ld_32x32(tO + c, o);
          tc::tmem_ld_wait();
#pragma unroll
          for (int k = 0; k < 32; ++k) o[k] = __float_as_uint(__uint_as_float(o[k]) * corr);
          asm volatile(
              "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
              "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
              "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(tO + c),
              "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3]), "r"(o[4]), "r"(o[5]), "r"(o[6]), "r"(o[7]), "r"(o[8]),
              "r"(o[9]), "r"(o[10]), "r"(o[11]), "r"(o[12]), "r"(o[13]), "r"(o[14]), "r"(o[15]), "r"(o[16]),
              "r"(o[17]), "r"(o[18]), "r"(o[19]), "r"(o[20]), "r"(o[21]), "r"(o[22]), "r"(o[23]), "r"(o[24]),
              "r"(o[25]), "r"(o[26]), "r"(o[27]), "r"(o[28]), "r"(o[29]), "r"(o[30]), "r"(o[31])
              : "memory");
        }
      }
      const float m_use = (m == -INFINITY) ? 0.f : m;
      // p = exp2(s*scale - m): bf16 P written over the first 64 columns of S
      float rs0 = 0.f, rs1 = 0.f, rs2 = 0.f, rs3 = 0.f;       // independent chains for the row sum
#pragma unroll
      for (int c = 0; c < BKV; c += 32) {
        uint32_t pk[16];
#pragma unroll
        for (int k = 0; k < 32; k += 4) {
          const float p0 = ex2(fmaf(__uint_as_float(r[c + k]), a.scale_log2, -m_use));
          const float p1 = ex2(fmaf(__uint_as_float(r[c + k + 1]), a.scale_log2, -m_use));
          const float p2 = ex2(fmaf(__uint_as_float(r[c + k + 2]), a.scale_log2, -m_use));
          const float p3 = ex2(fmaf(__uint_as_float(r[c + k + 3]), a.scale_log2, -m_use));
          rs0 += p0; rs1 += p1; rs2 += p2; rs3 += p3;
          __nv_bfloat162 h0 = __floats2bfloat162_rn(p0, p1), h1 = __floats2bfloat162_rn(p2, p3);
          pk[k >> 1] = *reinterpret_cast<uint32_t*>(&h0);
          pk[(k >> 1) + 1] = *reinterpret_cast<uint32_t*>(&h1);
        }
        tmem_st_32x16(tS + (c >> 1), pk);
      }
      tmem_st_wait();
      tc::tc_fence_before();
      tc::mbar_arrive(p_ready(i));
      l += (rs0 + rs1) + (rs2 + rs3);
    }
    float acc[D];
    if (n_tiles > 0) {
      tc::mbar_wait(o_full(i), (n_tiles - 1) & 1);
      tc::tc_fence_after();
#pragma unroll
      for (int c = 0; c < D; c += 32) tc::tmem_ld_32x32(tO + c, *reinterpret_cast<uint32_t(*)[32]>(&acc[c]));
      tc::tmem_ld_wait();
    } else {
#pragma unroll
      for (int d = 0; d < D; ++d) acc[d] = 0.f;
    }
    if (row < a.Tq) {
      const float inv = l > 0.f ? 1.f / l : 0.f;
      __nv_bfloat16* op = a.o + b * a.o_bs + (long long)row * a.o_ts + head * D;
#pragma unroll
      for (int c = 0; c < D; c += 8) {
        uint4 u;
        __nv_bfloat162 h0 = __floats2bfloat162_rn(acc[c] * inv, acc[c + 1] * inv);
        __nv_bfloat162 h1 = __floats2bfloat162_rn(acc[c + 2] * inv, acc[c + 3] * inv);
        __nv_bfloat162 h2 = __floats2bfloat162_rn(acc[c + 4] * inv, acc[c + 5] * inv);
        __nv_bfloat162 h3 = __floats2bfloat162_rn(acc[c + 6] * inv, acc[c + 7] * inv);
        u.x = *reinterpret_cast<uint32_t*>(&h0); u.y = *reinterpret_cast<uint32_t*>(&h1);
        u.z = *reinterpret_cast<uint32_t*>(&h2); u.w = *reinterpret_cast<uint32_t*>(&h3);
        *reinterpret_cast<uint4*>(op + c) = u;
      }
    }
  }

  tc::tc_fence_before();
  __syncthreads();
  if (warp == 2) tc::tmem_dealloc<1>(tmem, 512);
}

int make_tmap_3d(CUtensorMap* m, const void* base, uint64_t cols, uint64_t tokens, uint64_t batch, uint64_t token_pitch,
                 uint64_t batch_pitch) {
  PFN_cuTensorMapEncodeTiled_v12000 enc = vllm_tma_encoder();
  if (!enc) return -100;
  cuuint64_t dims[3] = {cols, tokens, batch};
  cuuint64_t strides[2] = {token_pitch * 2, batch_pitch * 2};
  cuuint32_t box[3] = {64, 128, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  if (batch == 1) strides[1] = token_pitch * 2 * (tokens > 0 ? tokens : 1);   // any valid multiple of 16
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -101;
}

}  // namespace

// Called by vllm_attention_bf16 (attention.cu) for head_dim == 128.  Returns VLLM_EUNSUPPORTED when the tensor
// maps cannot describe the views (caller then uses the mma.sync kernel).
int vllm_attention_tc_d128(const void* q, const void* k, const void* v, void* o, int batch, int Tq, int Tk, int heads,
                           int kv_heads, long long q_bs, long long q_ts, long long k_bs, long long k_ts, long long v_bs,
                           long long v_ts, long long o_bs, long long o_ts, const int* seqlens, int causal, float scale,
                           cudaStream_t st) {
  CUtensorMap tq, tk, tv;
  if (make_tmap_3d(&tq, q, (uint64_t)heads * D, Tq, batch, q_ts, q_bs)) return VLLM_EUNSUPPORTED;
  if (make_tmap_3d(&tk, k, (uint64_t)kv_heads * D, Tk, batch, k_ts, k_bs)) return VLLM_EUNSUPPORTED;
  if (make_tmap_3d(&tv, v, (uint64_t)kv_heads * D, Tk, batch, v_ts, v_bs)) return VLLM_EUNSUPPORTED;
  static bool set = false;
  if (!set) {
    cudaError_t e = cudaFuncSetAttribute(attn_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    if (e != cudaSuccess) return (int)e;
    set = true;
  }
  AttnTcArgs a;
  a.o = (__nv_bfloat16*)o; a.o_bs = o_bs; a.o_ts = o_ts; a.seqlens = seqlens; a.Tq = Tq; a.Tk = Tk;
  a.heads = heads; a.kv_heads = kv_heads; a.causal = causal; a.scale_log2 = scale * 1.4426950408889634f;
  dim3 grid((Tq + 2 * BQ - 1) / (2 * BQ), heads, batch);
  attn_fwd_tc_kernel<<<grid, THREADS, SMEM, st>>>(tq, tk, tv, a);
  VLLM_CHECK_LAUNCH();
  return VLLM_OK;
}
