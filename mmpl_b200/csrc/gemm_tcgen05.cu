// Persistent warp-specialised tcgen05 GEMM for sm_100a:
//   out[M,N] = epilogue( A[M,K] . W[N,K]^T + bias[N] )      (bf16 in, fp32 accumulate in TMEM)
// This is every nn.Linear on the CausalWanModel hot path (reference: wan/modules/causal_model.py:
// 111-113 q/k/v, :230 o, :267-269 ffn; wan/modules/model.py:172-193 cross-attention), with the
// element-wise work that follows each Linear in the reference folded into the epilogue at the
// reference's own bf16 rounding points:
//   EPI_BIAS           y = bf16(acc + b)
//   EPI_BIAS_GELU      bf16(gelu_tanh(y))                         (ffn.1, causal_model.py:268)
//   EPI_BIAS_SILU      bf16(silu(y))                              (time_embedding.1)
//   EPI_BIAS_RES       bf16(res + y)                              (cross-attn residual, :314)
//   EPI_BIAS_GATE_RES  bf16(res + bf16(y * gate[frame(row)]))     (:310, :322)
//
// Structure: warp 0 = TMA producer, warp 1 = MMA issuer (+ TMEM owner), warps 2..9 = epilogue (warps 2-5 take the
// left half of the tile's columns, warps 6-9 the right half: two warps per scheduler hide each other's latencies).
// A/W tiles are [128|BN rows] x [64 k] bf16 boxes in 128B-swizzled smem; the 128 x BN fp32
// accumulator is double-buffered in TMEM so the epilogue of tile i overlaps the main loop of tile i+1.
#include <cstdlib>

#include "host_util.h"
#include "mmpl_b200.h"
#include "ptx.cuh"

namespace mmpl {

constexpr int kBM = 128;
constexpr int kBK = 64;
constexpr int kGemmThreads = 320;  // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue (two per TMEM lane quadrant)
constexpr int kEpiWarps = 8;

struct GemmParams {
  int M, N, K;
  __nv_bfloat16* out;
  int64_t ldo;
  const __nv_bfloat16* bias;  // may be null
  const __nv_bfloat16* res;   // residual, may alias out
  int64_t ldr;
  const __nv_bfloat16* gate;  // [frames][gate_stride] vectors of N
  int64_t gate_stride;
  int rows_per_frame;
  int tiles_m, tiles_n;
  // stream-K part of the pair kernel's schedule (see PairSched): tiles [0, tiles_dp) are whole-tile work items, the
  // k blocks of the remaining tiles are dealt out evenly
  int tiles_dp;
  int group_m;         // tile order of the pair kernel: groups of group_m tile rows, row-fastest inside a group
  float* sk_part;      // [pairs][2 CTAs][128 rows][256 columns] fp32 partial accumulators
  uint32_t* sk_flags;  // [pairs][2 CTAs]: 1 = the partial of this CTA's head segment is in sk_part
};

// Tap-GEMM convolution (CONV instantiations of the single-CTA kernel only, see conv3d_cl): the K axis is
// taps x channel blocks; k block kb reads channel block kb % conv_kb_per_tap of the A rows shifted by
// conv_tap_off[kb / conv_kb_per_tap]; only rows that are interior positions of the padded [.., conv_hp, conv_wp] grid
// are stored. A separate (derived) parameter block so that the Linear kernels' parameter layout, and with it their
// generated code, is exactly what was measured.
struct ConvGemmParams : GemmParams {
  int conv_kb_per_tap;
  int conv_hp, conv_wp;
  int conv_tap_off[27];
};
template <bool CONV>
struct ParamsOf { using type = GemmParams; };
template <>
struct ParamsOf<true> { using type = ConvGemmParams; };

template <int BN>
struct GemmCfg {
  static constexpr int kStages = (BN == 256) ? 4 : (BN == 192 ? 5 : (BN == 128 ? 6 : 8));
  static constexpr int kABytes = kBM * kBK * 2;
  static constexpr int kBBytes = BN * kBK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kTmemCols = (2 * BN <= 32) ? 32 : (2 * BN <= 64 ? 64 : (2 * BN <= 128 ? 128 : (2 * BN <= 256 ? 256 : 512)));
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align*/ + 256 /*barriers*/;
};

// gelu_tanh(x) = 0.5 x (1 + tanh(u)) = x * sigmoid(2u),  u = sqrt(2/pi) (x + 0.044715 x^3):
// one ex2 and one rcp on the MUFU and five FMA-pipe operations per element.
__device__ __forceinline__ float gelu_tanh_f(float x) {
  const float kK = -2.0f * 1.4426950408889634f * 0.7978845608028654f;  // -2 log2(e) sqrt(2/pi)
  const float w = fmaf(x * x, kK * 0.044715f, kK);
  const float e = fast_exp2(x * w);  // exp(-2u)
  return __fdividef(x, 1.0f + e);
}
__device__ __forceinline__ float silu_f(float x) { return __fdividef(x, 1.0f + __expf(-x)); }
__device__ __forceinline__ void round2f(float& a, float& b) {  // both to bf16 and back: one packed conversion + two integer ops
  const uint32_t p = pack_bf16x2(a, b);
  a = bf16_lo(p);
  b = bf16_hi(p);
}

// Epilogue of one accumulator row: `taddr` addresses this warp's 32 TMEM lanes at the tile's first column; the
// thread owns output row `row`, columns [col_base, col_base + BN). Rounds to bf16 where the reference does.
// `part` (stream-K owner segments only): `nparts` fp32 partial accumulator rows of this thread's row, `part_stride`
// floats apart, written by other CTAs; they are added to the TMEM accumulator before the epilogue arithmetic.
template <int BN, int EPI, bool CONV = false>
__device__ __forceinline__ void epilogue_tile(const typename ParamsOf<CONV>::type& p, uint32_t taddr, int row, int col_base, int c_begin,
                                              int c_end, const float* part = nullptr, int nparts = 0,
                                              int64_t part_stride = 0) {
  constexpr bool kHasRes = (EPI == MMPL_EPI_BIAS_RES || EPI == MMPL_EPI_BIAS_GATE_RES);
  bool row_ok = row < p.M;
  bool halo_row = false;
  if constexpr (CONV) {
    // rows are positions of the zero-haloed grid: interior rows get the convolution, halo rows are written as zeros, so
    // the output is a valid padded input of the next layer whatever the buffer held before (no fill pass anywhere)
    const int wp = row % p.conv_wp;
    const int hp = (row / p.conv_wp) % p.conv_hp;
    const bool interior = wp >= 1 && wp < p.conv_wp - 1 && hp >= 1 && hp < p.conv_hp - 1;
    halo_row = row_ok && !interior;
    row_ok = row_ok && interior;
  }
  const __nv_bfloat16* gate_row = nullptr;
  if (EPI == MMPL_EPI_BIAS_GATE_RES && row_ok)
    gate_row = p.gate + static_cast<int64_t>(row / p.rows_per_frame) * p.gate_stride;
  // N is a multiple of 8: 8-column (16-byte) groups; two groups are stored together as one 32-byte sector
  // (st.global.v8) when the row pitch allows, so no sector is written partially.
  const bool wide = (p.ldo % 16 == 0) && ((reinterpret_cast<uintptr_t>(p.out) & 31) == 0);
  const bool wide_res = kHasRes && (p.ldr % 16 == 0) && ((reinterpret_cast<uintptr_t>(p.res) & 31) == 0);
  // Residual of one 32-column chunk of this thread's row (64 bytes); loaded one chunk AHEAD of its use, so that the
  // global-load latency overlaps the tcgen05.ld and the arithmetic of the previous chunk instead of adding to every chunk.
  auto load_res = [&](int c, uint32_t (&r)[16]) {
#pragma unroll
    for (int i = 0; i < 16; ++i) r[i] = 0;
    const int col0 = col_base + c * 32;
    if (!kHasRes || !row_ok || col0 >= p.N) return;
    const __nv_bfloat16* src = p.res + static_cast<int64_t>(row) * p.ldr + col0;
#pragma unroll
    for (int g2 = 0; g2 < 2; ++g2) {
      if (col0 + g2 * 16 + 16 <= p.N && wide_res) {
        asm volatile("ld.global.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                     : "=r"(r[g2 * 8]), "=r"(r[g2 * 8 + 1]), "=r"(r[g2 * 8 + 2]), "=r"(r[g2 * 8 + 3]), "=r"(r[g2 * 8 + 4]),
                       "=r"(r[g2 * 8 + 5]), "=r"(r[g2 * 8 + 6]), "=r"(r[g2 * 8 + 7])
                     : "l"(src + g2 * 16)
                     : "memory");
      } else {
        if (col0 + g2 * 16 < p.N) {
          const uint4 t = *reinterpret_cast<const uint4*>(src + g2 * 16);
          r[g2 * 8] = t.x; r[g2 * 8 + 1] = t.y; r[g2 * 8 + 2] = t.z; r[g2 * 8 + 3] = t.w;
        }
        if (col0 + g2 * 16 + 8 < p.N) {
          const uint4 t = *reinterpret_cast<const uint4*>(src + g2 * 16 + 8);
          r[g2 * 8 + 4] = t.x; r[g2 * 8 + 5] = t.y; r[g2 * 8 + 6] = t.z; r[g2 * 8 + 7] = t.w;
        }
      }
    }
  };
  uint32_t res_next[16];
  load_res(c_begin, res_next);
#pragma unroll 1
  for (int c = c_begin; c < c_end; ++c) {
    uint32_t acc[32];
    tmem_ld_32x32(taddr + c * 32, acc);
    uint32_t rres[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) rres[i] = res_next[i];
    if (kHasRes && c + 1 < c_end) load_res(c + 1, res_next);
    tmem_ld_wait();
    const int col0 = col_base + c * 32;
    if (row_ok) {
      for (int i = 0; i < nparts; ++i) {
        const float4* src = reinterpret_cast<const float4*>(part + i * part_stride + c * 32);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 v = __ldcg(src + j);  // L2: written by another SM during this launch
          acc[4 * j] = __float_as_uint(__uint_as_float(acc[4 * j]) + v.x);
          acc[4 * j + 1] = __float_as_uint(__uint_as_float(acc[4 * j + 1]) + v.y);
          acc[4 * j + 2] = __float_as_uint(__uint_as_float(acc[4 * j + 2]) + v.z);
          acc[4 * j + 3] = __float_as_uint(__uint_as_float(acc[4 * j + 3]) + v.w);
        }
      }
    }
    if (row_ok && col0 < p.N) {
#pragma unroll
      for (int g2 = 0; g2 < 2; ++g2) {
        uint32_t ow[8];
        bool have[2] = {false, false};
#pragma unroll
        for (int gg = 0; gg < 2; ++gg) {
          const int g = g2 * 2 + gg;
          const int col = col0 + g * 8;
          if (col >= p.N) continue;
          have[gg] = true;
          uint4 bv = make_uint4(0, 0, 0, 0);
          if (p.bias) bv = __ldg(reinterpret_cast<const uint4*>(p.bias + col));
          uint4 gv = make_uint4(0, 0, 0, 0);
          if (EPI == MMPL_EPI_BIAS_GATE_RES)
            gv = __ldg(reinterpret_cast<const uint4*>(gate_row + col));
          const uint32_t* bw = reinterpret_cast<const uint32_t*>(&bv);
          const uint32_t* rw = &rres[g * 4];
          const uint32_t* gw = reinterpret_cast<const uint32_t*>(&gv);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            // roundings in pairs: cvt.rn.bf16x2.f32 issues at 2 clk per warp instruction, the scalar conversion at 8
            // (profiles/r01_microbench_pipes.txt), and an unfused Linear needs none before the pack below
            float y0 = __uint_as_float(acc[g * 8 + 2 * j]) + bf16_lo(bw[j]);
            float y1 = __uint_as_float(acc[g * 8 + 2 * j + 1]) + bf16_hi(bw[j]);
            if (EPI != MMPL_EPI_BIAS) round2f(y0, y1);
            if (EPI == MMPL_EPI_BIAS_GELU) {
              y0 = gelu_tanh_f(y0);
              y1 = gelu_tanh_f(y1);
            } else if (EPI == MMPL_EPI_BIAS_SILU) {
              y0 = silu_f(y0);
              y1 = silu_f(y1);
            } else if (EPI == MMPL_EPI_BIAS_RES) {
              y0 = bf16_lo(rw[j]) + y0;
              y1 = bf16_hi(rw[j]) + y1;
            } else if (EPI == MMPL_EPI_BIAS_GATE_RES) {
              y0 *= bf16_lo(gw[j]);
              y1 *= bf16_hi(gw[j]);
              round2f(y0, y1);
              y0 += bf16_lo(rw[j]);
              y1 += bf16_hi(rw[j]);
            }
            ow[gg * 4 + j] = pack_bf16x2(y0, y1);
          }
        }
        __nv_bfloat16* dst = p.out + static_cast<int64_t>(row) * p.ldo + col0 + g2 * 16;
        if (have[0] && have[1] && wide) {
          st_global_v8(dst, ow);
        } else {
          if (have[0]) *reinterpret_cast<uint4*>(dst) = make_uint4(ow[0], ow[1], ow[2], ow[3]);
          if (have[1]) *reinterpret_cast<uint4*>(dst + 8) = make_uint4(ow[4], ow[5], ow[6], ow[7]);
        }
      }
    } else if (CONV && halo_row && col0 < p.N) {
      __nv_bfloat16* dst = p.out + static_cast<int64_t>(row) * p.ldo + col0;
#pragma unroll
      for (int g = 0; g < 4; ++g)
        if (col0 + g * 8 < p.N) *reinterpret_cast<uint4*>(dst + g * 8) = make_uint4(0, 0, 0, 0);
    }
  }
}

template <int BN, int EPI, bool CONV = false>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                 const typename ParamsOf<CONV>::type p) {
  using Cfg = GemmCfg<BN>;
  constexpr int ST = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + ST * Cfg::kABytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + ST * Cfg::kStageBytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + ST;
  uint64_t* tmem_full = bars + 2 * ST;
  uint64_t* tmem_empty = bars + 2 * ST + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * ST + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_kb = (p.K + kBK - 1) / kBK;
  const int num_tiles = p.tiles_m * p.tiles_n;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int i = 0; i < ST; ++i) {
        mbar_init(&full_bar[i], 1);
        mbar_init(&empty_bar[i], 1);
      }
      for (int i = 0; i < 2; ++i) {
        mbar_init(&tmem_full[i], 1);
        mbar_init(&tmem_empty[i], kEpiWarps);
      }
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, Cfg::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();  // PDL: the prologue above overlapped the previous kernel; its outputs (A, residual) are valid from here on
  pdl_launch_dependents();

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    // (whole warp, converged; one elected lane issues: see the "_elect" wrappers in ptx.cuh)
    int s = 0;
    uint32_t ph = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      const int m_blk = t % p.tiles_m;
      const int n_blk = t / p.tiles_m;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&empty_bar[s], ph ^ 1);
        mbar_arrive_expect_tx_elect(&full_bar[s], Cfg::kStageBytes);
        if constexpr (CONV) {
          // tap-GEMM: the same 128 grid positions shifted by this tap's offset (negative / past-the-end rows and
          // channels beyond Cin are zero-filled by TMA); W is packed [Cout][taps][Cin rounded up to 64]
          const int tap = kb / p.conv_kb_per_tap;
          const int cb = kb - tap * p.conv_kb_per_tap;
          tma_load_2d_elect(smem_a + s * Cfg::kABytes, &map_a, &full_bar[s], cb * kBK, m_blk * kBM + p.conv_tap_off[tap],
                            kEvictNormal);
        } else {
          tma_load_2d_elect(smem_a + s * Cfg::kABytes, &map_a, &full_bar[s], kb * kBK, m_blk * kBM, kEvictNormal);
        }
        tma_load_2d_elect(smem_b + s * Cfg::kBBytes, &map_b, &full_bar[s], kb * kBK, n_blk * BN, kEvictLast);
        if (++s == ST) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // -------------------------------------------------------------- MMA issuer
    constexpr uint32_t idesc = make_idesc_bf16(kBM, BN, 0, 0);
    int s = 0;
    uint32_t ph = 0;
    int it = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
      const int as = it & 1;
      const uint32_t aph = (it >> 1) & 1;
      mbar_wait(&tmem_empty[as], aph ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + as * BN;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        // descriptor = constant high word + start address (>> 4) in the low word
        constexpr uint64_t kDesc0 = make_smem_desc_sw128_const(16, 1024);
        const uint32_t a_lo = desc_lo(kDesc0) + ((smem_u32(smem_a) + s * Cfg::kABytes) >> 4);
        const uint32_t b_lo = desc_lo(kDesc0) + ((smem_u32(smem_b) + s * Cfg::kBBytes) >> 4);
#pragma unroll
        for (int k = 0; k < kBK / 16; ++k) {
          // +32 bytes (= 2 in the >>4 address field) per 16-element K step inside the swizzle atom
          umma_ss_elect(tmem_d, a_lo + 2 * k, desc_hi(kDesc0), b_lo + 2 * k, desc_hi(kDesc0), idesc, (kb | k) != 0 ? 1u : 0u);
        }
        tc_commit_elect(&empty_bar[s]);
        if (kb == num_kb - 1) tc_commit_elect(&tmem_full[as]);
        if (++s == ST) { s = 0; ph ^= 1; }
      }
    }
  } else {
    // ---------------------------------------------------------------- epilogue
    const int lane_base = (warp & 3) * 32;  // TMEM lanes this warp may access
    int it = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
      const int m_blk = t % p.tiles_m;
      const int n_blk = t / p.tiles_m;
      const int as = it & 1;
      const uint32_t aph = (it >> 1) & 1;
      mbar_wait(&tmem_full[as], aph);
      tc_fence_after();
      const int row = m_blk * kBM + lane_base + lane;
      const int half = (warp - 2) >> 2;
      if constexpr ((BN / 32) % 2 == 0) {
        constexpr int kChunks = BN / 32 / 2;  // 32-column chunks per warp
        epilogue_tile<BN, EPI, CONV>(p, tmem_base + (static_cast<uint32_t>(lane_base) << 16) + as * BN, row, n_blk * BN,
                                     half * kChunks, (half + 1) * kChunks);
      } else {
        // odd chunk count (BN = 96: three): the left-half warps take the first two, the right-half warps the last
        constexpr int kChunksLo = (BN / 32 + 1) / 2;
        epilogue_tile<BN, EPI, CONV>(p, tmem_base + (static_cast<uint32_t>(lane_base) << 16) + as * BN, row, n_blk * BN,
                                     half ? kChunksLo : 0, half ? BN / 32 : kChunksLo);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[as]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// ------------------------------------------------------------------------------------------------
// Shared-A variant of the kernel above for the short-K projections (cfg2's o / cross-q / cross-o, M = 4680, N = K = 1536).
// Their 128 x 128 tiles fit the machine exactly (444 = 3 x 148) but load 32 KB per 64-wide k block for 256 clk of MMA,
// 128 B/clk/SM, above what L2 delivers to one SM (measured 341 clk per k block); the cta_group::2 tiles halve the bytes
// but quantise M into 4 waves. Here a cluster of two CTAs computes two 128 x BN tiles that are neighbours along N, i.e.
// share their A rows: each CTA loads its own W tile and HALF of the A tile (64 rows), which TMA multicasts into the
// shared memory of both. 8 + 16 KB per CTA and k block at BN = 128 (96 B/clk/SM, 222 cluster tiles = 3 waves on 74
// clusters), 8 + 24 KB at BN = 192 (85 B/clk/SM, 148 cluster tiles = 2 waves). Each CTA runs its own cta_group::1 MMAs
// on its own TMEM; only the shared-memory ring is coupled:
//   full[s]  (per CTA, 1 arrival + bytes): the CTA's own producer arms it with the bytes that land in ITS shared memory
//            (both A halves + its W tile); the peer's multicast credits its half to the barrier at the same offset here.
//   empty[s] (per CTA, 2 arrivals): a slot is rewritten by both producers, so it is free only once the MMAs of BOTH CTAs
//            have read it: each MMA warp commits to empty[s] of both CTAs (multicast commit).
// A peer's bytes may be credited before the local producer has armed the phase: the phase cannot complete without the
// producer's own arrival, and they cannot arrive a phase early because the peer's producer waits for this CTA's MMA
// commit of the previous use of the slot, which follows this CTA's full[s] wait.
__device__ __forceinline__ void tma_load_2d_mcast_elect(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int32_t c0,
                                                        int32_t c1, uint16_t cta_mask, uint64_t policy) {
  asm volatile(
      "{\n\t.reg .pred e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "@e cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster.L2::cache_hint"
      " [%0], [%1, {%4, %5}], [%2], %3, %6;\n\t}"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "h"(cta_mask), "r"(c0), "r"(c1),
      "l"(policy)
      : "memory");
}
// Commit of this CTA's tcgen05 ops, arriving on the barrier at this offset in every CTA of `cta_mask`.
__device__ __forceinline__ void tc_commit_mcast_elect(uint64_t* bar, uint16_t cta_mask) {
  asm volatile(
      "{\n\t.reg .pred e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n\t}"
      ::"r"(smem_u32(bar)), "h"(cta_mask)
      : "memory");
}

template <int BN, int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kGemmThreads, 1)
gemm_bf16_shared_a_kernel(const __grid_constant__ CUtensorMap map_a /* box: 64 rows */, const __grid_constant__ CUtensorMap map_b,
                          const GemmParams p) {
  using Cfg = GemmCfg<BN>;
  constexpr int ST = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + ST * Cfg::kABytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + ST * Cfg::kStageBytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + ST;
  uint64_t* tmem_full = bars + 2 * ST;
  uint64_t* tmem_empty = bars + 2 * ST + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * ST + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int cta = static_cast<int>(cluster_ctarank());
  const int cluster = blockIdx.x >> 1;
  const int num_clusters = gridDim.x >> 1;
  const int num_kb = (p.K + kBK - 1) / kBK;
  const int num_tiles = p.tiles_m * (p.tiles_n >> 1);  // cluster tiles of 128 x 2 BN; tiles_n is even (host)

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int i = 0; i < ST; ++i) {
        mbar_init(&full_bar[i], 1);
        mbar_init(&empty_bar[i], 2);  // the MMA warps of both CTAs
      }
      for (int i = 0; i < 2; ++i) {
        mbar_init(&tmem_full[i], 1);
        mbar_init(&tmem_empty[i], kEpiWarps);
      }
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, Cfg::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  cluster_sync();  // the peer's barriers are initialised before anything of this CTA can reach them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();  // PDL (see gemm_bf16_kernel)
  pdl_launch_dependents();

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    int s = 0;
    uint32_t ph = 0;
    for (int t = cluster; t < num_tiles; t += num_clusters) {
      const int m_blk = t % p.tiles_m;
      const int n_blk = 2 * (t / p.tiles_m) + cta;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&empty_bar[s], ph ^ 1);
        mbar_arrive_expect_tx_elect(&full_bar[s], Cfg::kStageBytes);
        // rows [64 cta, 64 cta + 64) of the A tile, into both CTAs at the same offset
        tma_load_2d_mcast_elect(smem_a + s * Cfg::kABytes + cta * (Cfg::kABytes / 2), &map_a, &full_bar[s], kb * kBK,
                                m_blk * kBM + cta * (kBM / 2), 0x3, kEvictNormal);
        tma_load_2d_elect(smem_b + s * Cfg::kBBytes, &map_b, &full_bar[s], kb * kBK, n_blk * BN, kEvictLast);
        if (++s == ST) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // -------------------------------------------------------------- MMA issuer
    constexpr uint32_t idesc = make_idesc_bf16(kBM, BN, 0, 0);
    int s = 0;
    uint32_t ph = 0;
    int it = 0;
    for (int t = cluster; t < num_tiles; t += num_clusters, ++it) {
      const int as = it & 1;
      const uint32_t aph = (it >> 1) & 1;
      mbar_wait(&tmem_empty[as], aph ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + as * BN;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        constexpr uint64_t kDesc0 = make_smem_desc_sw128_const(16, 1024);
        const uint32_t a_lo = desc_lo(kDesc0) + ((smem_u32(smem_a) + s * Cfg::kABytes) >> 4);
        const uint32_t b_lo = desc_lo(kDesc0) + ((smem_u32(smem_b) + s * Cfg::kBBytes) >> 4);
#pragma unroll
        for (int k = 0; k < kBK / 16; ++k) {
          umma_ss_elect(tmem_d, a_lo + 2 * k, desc_hi(kDesc0), b_lo + 2 * k, desc_hi(kDesc0), idesc, (kb | k) != 0 ? 1u : 0u);
        }
        tc_commit_mcast_elect(&empty_bar[s], 0x3);
        if (kb == num_kb - 1) tc_commit_elect(&tmem_full[as]);
        if (++s == ST) { s = 0; ph ^= 1; }
      }
    }
  } else {
    // ---------------------------------------------------------------- epilogue
    const int lane_base = (warp & 3) * 32;
    int it = 0;
    for (int t = cluster; t < num_tiles; t += num_clusters, ++it) {
      const int m_blk = t % p.tiles_m;
      const int n_blk = 2 * (t / p.tiles_m) + cta;
      const int as = it & 1;
      const uint32_t aph = (it >> 1) & 1;
      mbar_wait(&tmem_full[as], aph);
      tc_fence_after();
      const int row = m_blk * kBM + lane_base + lane;
      const int half = (warp - 2) >> 2;
      constexpr int kChunks = BN / 32 / 2;
      static_assert((BN / 32) % 2 == 0, "shared-A tile width");
      epilogue_tile<BN, EPI>(p, tmem_base + (static_cast<uint32_t>(lane_base) << 16) + as * BN, row, n_blk * BN,
                             half * kChunks, (half + 1) * kChunks);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[as]);
    }
  }

  tc_fence_before();
  cluster_sync();  // no CTA leaves while the peer's multicast loads or commits can still reach its shared memory
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// Work list of one CTA pair in the cta_group::2 kernel. Whole tiles alone leave the last wave partly empty (cfg2:
// M = 4680, N = 1536 is 19 x 6 = 114 tiles of 256 x 256 on 74 pairs, 1.54 waves that take 2), so the schedule is
// "stream-K" for the tail: tiles [0, tiles_dp) (a multiple of the pair count) are dealt out whole, round-robin; the
// k blocks of the other tiles form one sequence of (tile, k block) units cut into equal contiguous ranges, one per
// pair. A range is a head segment (the last k blocks of a tile), whole tiles, and a tail segment (the first k
// blocks of a tile). The pair that holds a tile's FIRST k blocks owns the tile: it runs that segment last, by
// which time the pairs holding the rest of the tile -- it is the head segment of their ranges, the first thing
// they ran -- have left their fp32 partial accumulators in sk_part and raised sk_flags; the owner adds them to its
// own accumulator in a fixed order (deterministic) and runs the epilogue. Stream-K segments run before the whole
// tiles, so the hand-over happens mid-kernel behind the main loops.
struct PairSched {
  int num_kb, tiles_dp, num_pairs;
  int sk_pos, sk_hi, dp_t;
  __device__ __forceinline__ static int range_lo(int pair, int num_pairs, int units) {
    return static_cast<int>(static_cast<long long>(pair) * units / num_pairs);
  }
  __device__ __forceinline__ void init(const GemmParams& p, int pair, int npairs, int nkb) {
    num_kb = nkb;
    tiles_dp = p.tiles_dp;
    num_pairs = npairs;
    const int units = (p.tiles_m * p.tiles_n - p.tiles_dp) * nkb;
    sk_pos = range_lo(pair, npairs, units);
    sk_hi = range_lo(pair + 1, npairs, units);
    dp_t = pair;
  }
  // next segment: tile index and its k-block range [kb0, kb1)
  __device__ __forceinline__ bool next(int& tile, int& kb0, int& kb1) {
    if (sk_pos < sk_hi) {
      const int t = sk_pos / num_kb;
      const int end = min(sk_hi, (t + 1) * num_kb);
      kb0 = sk_pos - t * num_kb;
      kb1 = end - t * num_kb;
      tile = tiles_dp + t;
      sk_pos = end;
      return true;
    }
    if (dp_t < tiles_dp) {
      tile = dp_t;
      kb0 = 0;
      kb1 = num_kb;
      dp_t += num_pairs;
      return true;
    }
    return false;
  }
};

// Tile index -> (tile row, tile column). Tiles that run at the same time (consecutive indices) cover group_m tile rows x
// (pairs / group_m) tile columns, so one wave re-reads group_m A panels and a few W panels from L2 instead of all
// of A (at the Wan-14B shapes A alone is 112 MB: row-fastest order over all 43 tile rows streams it from HBM in
// every wave).
__device__ __forceinline__ void tile_coords(const GemmParams& p, int t, int& m_blk, int& n_blk) {
  const int per_group = p.group_m * p.tiles_n;
  const int g = t / per_group;
  const int r = t - g * per_group;
  const int m0 = g * p.group_m;
  const int gm = min(p.group_m, p.tiles_m - m0);
  n_blk = r / gm;
  m_blk = m0 + (r - n_blk * gm);
}

// Stream-K head segment: this thread's accumulator row (columns [c_begin*32, c_end*32)) -> fp32 workspace row.
__device__ __forceinline__ void store_partial_row(uint32_t taddr, float* dst, bool row_ok, int c_begin, int c_end) {
#pragma unroll 1
  for (int c = c_begin; c < c_end; ++c) {
    uint32_t acc[32];
    tmem_ld_32x32(taddr + c * 32, acc);
    tmem_ld_wait();
    if (row_ok) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t w[8] = {acc[8 * j], acc[8 * j + 1], acc[8 * j + 2], acc[8 * j + 3],
                               acc[8 * j + 4], acc[8 * j + 5], acc[8 * j + 6], acc[8 * j + 7]};
        st_global_v8(dst + c * 32 + 8 * j, w);
      }
    }
  }
}
__device__ __forceinline__ void epi_group_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }  // warps 2..9

// ------------------------------------------------------------------------------------------------
// cta_group::2 variant: a pair of CTAs (one cluster of 2, same TPC) computes a 256 x 256 tile with
// tcgen05.mma.cta_group::2 (M = 256, N = 256). Each CTA loads 128 rows of A and 128 rows (= N/2) of W
// per 64-wide k block (32 KB per stage instead of 48 KB) and holds 128 accumulator rows x 256 columns in
// its own TMEM, double-buffered. Per-SM L2->SMEM traffic drops from 96 to 64 B/clk at full MMA rate,
// which is what bounds the single-CTA 128x256 kernel (ncu: tensor pipe ~48 % active).
// Barriers: the leader's full[s] collects the TMA bytes of both CTAs (expect_tx = 64 KB); the leader's
// MMA thread releases smem slots and publishes accumulators with multicast commits to both CTAs; the
// epilogue warps of both CTAs arrive on the leader's tmem_empty.
constexpr int kPairStages = 6;
constexpr int kPairStageBytes = 2 * kBM * kBK * 2;  // 16 KB A + 16 KB W per CTA
constexpr int kPairSmemBytes = kPairStages * kPairStageBytes + 1024 + 256;
constexpr int kPairBN = 256;

// BN = tile width of the pair: 256, or 224 where that fills the machine better. cfg2's N = 1536 projections are
// 19 x 6 = 114 tiles of 256 x 256 on 74 pairs (2 waves, the second 54 % full) but 19 x 7 = 133 tiles of 256 x 224
// (2 waves of tiles that are 12.5 % cheaper); cuBLAS picks the same width for this shape (nvjet 256x224, ncu
// comparison in profiles/). The smem slots and the TMEM accumulator spacing stay those of BN = 256.
template <int BN, int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kGemmThreads, 1)
gemm_bf16_pair_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                      const GemmParams p) {
  constexpr int ST = kPairStages;
  static_assert(BN % 32 == 0 && BN <= kPairBN && (BN / 2) % 8 == 0, "pair tile width");
  constexpr uint32_t kTxBytes = 2 * (kBM * kBK * 2 + (BN / 2) * kBK * 2);  // A + W bytes of both CTAs per stage
  constexpr int kChunksLo = (BN / 32 + 1) / 2;  // 32-column chunks of the left-half epilogue warps; the rest go right
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + ST * (kPairStageBytes / 2);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + ST * kPairStageBytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + ST;
  uint64_t* tmem_full = bars + 2 * ST;
  uint64_t* tmem_empty = bars + 2 * ST + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * ST + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t cta = cluster_ctarank();
  const bool leader = cta == 0;
  const int pair = blockIdx.x >> 1;
  const int num_pairs = gridDim.x >> 1;
  const int num_kb = (p.K + kBK - 1) / kBK;
  const int num_tiles = p.tiles_m * p.tiles_n;  // tiles of 256 x 256

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int i = 0; i < ST; ++i) {
        mbar_init(&full_bar[i], 1);
        mbar_init(&empty_bar[i], 1);
      }
      for (int i = 0; i < 2; ++i) {
        mbar_init(&tmem_full[i], 1);
        mbar_init(&tmem_empty[i], 2 * kEpiWarps);  // epilogue warps of both CTAs
      }
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc_2cta(tmem_slot, 512);
    tmem_relinquish_2cta();
  }
  tc_fence_before();
  cluster_sync();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();  // PDL (see above)
  pdl_launch_dependents();

  if (warp == 0) {
    // ------------------------------------------------ TMA producer (both CTAs, own halves)
    int s = 0;
    uint32_t ph = 0;
    PairSched sched;
    sched.init(p, pair, num_pairs, num_kb);
    int t, kb0, kb1;
    while (sched.next(t, kb0, kb1)) {
      int m_blk, n_blk;
      tile_coords(p, t, m_blk, n_blk);
      const int row_a = m_blk * 256 + static_cast<int>(cta) * kBM;
      const int row_b = n_blk * BN + static_cast<int>(cta) * (BN / 2);
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&empty_bar[s], ph ^ 1);
        if (leader) mbar_arrive_expect_tx_elect(&full_bar[s], kTxBytes);
        tma_load_2d_2cta_elect(smem_a + s * (kPairStageBytes / 2), &map_a, &full_bar[s], kb * kBK, row_a, kEvictNormal);
        tma_load_2d_2cta_elect(smem_b + s * (kPairStageBytes / 2), &map_b, &full_bar[s], kb * kBK, row_b, kEvictLast);
        if (++s == ST) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------ MMA issuer (leader CTA only)
    if (leader) {
      constexpr uint32_t idesc = make_idesc_bf16(256, BN, 0, 0);  // N = BN columns, half from each CTA's W rows
      int s = 0;
      uint32_t ph = 0;
      int it = 0;
      PairSched sched;
      sched.init(p, pair, num_pairs, num_kb);
      int t, kb0, kb1;
      for (; sched.next(t, kb0, kb1); ++it) {
        const int as = it & 1;
        const uint32_t aph = (it >> 1) & 1;
        mbar_wait(&tmem_empty[as], aph ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * kPairBN;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          constexpr uint64_t kDesc0 = make_smem_desc_sw128_const(16, 1024);
          const uint32_t a_lo = desc_lo(kDesc0) + ((smem_u32(smem_a) + s * (kPairStageBytes / 2)) >> 4);
          const uint32_t b_lo = desc_lo(kDesc0) + ((smem_u32(smem_b) + s * (kPairStageBytes / 2)) >> 4);
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k)
            umma_ss_2cta_elect(tmem_d, a_lo + 2 * k, desc_hi(kDesc0), b_lo + 2 * k, desc_hi(kDesc0), idesc,
                               ((kb - kb0) | k) != 0 ? 1u : 0u);
          tc_commit_2cta_elect(&empty_bar[s], 0x3);
          if (kb == kb1 - 1) tc_commit_2cta_elect(&tmem_full[as], 0x3);
          if (++s == ST) { s = 0; ph ^= 1; }
        }
      }
    }
  } else {
    // ------------------------------------------------ epilogue (both CTAs, own 128 rows)
    const int lane_base = (warp & 3) * 32;
    int it = 0;
    PairSched sched;
    sched.init(p, pair, num_pairs, num_kb);
    const int sk_units = (num_tiles - p.tiles_dp) * num_kb;
    constexpr int64_t kSlotFloats = 128 * BN;  // one CTA's partial accumulator
    int t, kb0, kb1;
    for (; sched.next(t, kb0, kb1); ++it) {
      int m_blk, n_blk;
      tile_coords(p, t, m_blk, n_blk);
      const int as = it & 1;
      const uint32_t aph = (it >> 1) & 1;
      mbar_wait(&tmem_full[as], aph);
      tc_fence_after();
      const int row = m_blk * 256 + static_cast<int>(cta) * kBM + lane_base + lane;
      const int half = (warp - 2) >> 2;
      const int c_begin = half * kChunksLo;
      const int c_end = c_begin + ((BN / 32) % 2 == 0 ? kChunksLo : (half ? BN / 32 - kChunksLo : kChunksLo));
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(lane_base) << 16) + as * kPairBN;
      if (kb0 > 0) {
        // head segment of this pair's stream-K range: leave the partial accumulator for the tile's owner
        float* dst = p.sk_part + (static_cast<int64_t>(pair) * 2 + cta) * kSlotFloats +
                     static_cast<int64_t>(lane_base + lane) * BN;
        store_partial_row(taddr, dst, row < p.M, c_begin, c_end);
        __threadfence();
        epi_group_sync();
        if (warp == 2 && lane == 0)
          asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p.sk_flags + pair * 2 + cta), "r"(1u) : "memory");
      } else if (kb1 < num_kb) {
        // tail segment: this pair owns the tile. The rest of its k blocks are the head segments of the next pairs.
        const int tile_end = (t - p.tiles_dp + 1) * num_kb;
        int q_last = pair;
        for (int covered = sched.sk_hi; covered < tile_end; ) {
          ++q_last;
          covered = min(tile_end, PairSched::range_lo(q_last + 1, num_pairs, sk_units));
        }
        if (warp == 2 && lane == 0) {
          for (int q = pair + 1; q <= q_last; ++q) {
            uint32_t* flag = p.sk_flags + q * 2 + cta;
            uint32_t v;
            do {
              asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
            } while (v == 0);
            asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(flag), "r"(0u) : "memory");  // re-arm for the next launch
          }
        }
        epi_group_sync();
        const float* part = p.sk_part + (static_cast<int64_t>(pair + 1) * 2 + cta) * kSlotFloats +
                            static_cast<int64_t>(lane_base + lane) * BN;
        epilogue_tile<BN, EPI>(p, taddr, row, n_blk * BN, c_begin, c_end, part, q_last - pair, 2 * kSlotFloats);
      } else {
        epilogue_tile<BN, EPI>(p, taddr, row, n_blk * BN, c_begin, c_end);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(&tmem_empty[as], 0);
    }
  }

  tc_fence_before();
  cluster_sync();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2cta(tmem_base, 512);
  }
}

// Stream-K workspace: one fp32 partial-accumulator slot and one flag per CTA (grown on demand; one per process, used
// by launches on one stream at a time, like the attention workspace). Flags are zero between launches: every raised
// flag is consumed and cleared by exactly one owner inside the same launch.
static float* g_sk_part = nullptr;
static uint32_t* g_sk_flags = nullptr;
static int g_sk_pairs = 0;
// 0 never (default), -1 automatic, 1 whenever the tile count is not a multiple of the pair count; MMPL_GEMM_STREAMK sets it.
// Off by default: measured on B200 (profiles/README.md, round-1 stream-K A/B) the tail schedule loses 9-11 us per
// launch at the cfg2 shapes. The GEMMs run power-capped at ~1.4-1.65 GHz, so a half-empty last wave is not idle time
// to recover -- the pairs that are still busy clock higher -- while the owner's fix-up epilogue is exposed.
static int g_streamk_mode = getenv("MMPL_GEMM_STREAMK") ? atoi(getenv("MMPL_GEMM_STREAMK")) : 0;
void gemm_set_streamk(int mode) { g_streamk_mode = mode; }
static int ensure_streamk_workspace(int pairs) {
  if (pairs <= g_sk_pairs) return MMPL_OK;
  if (g_sk_part) MMPL_CUDA(cudaFree(g_sk_part));
  if (g_sk_flags) MMPL_CUDA(cudaFree(g_sk_flags));
  g_sk_part = nullptr;
  g_sk_flags = nullptr;
  g_sk_pairs = 0;
  MMPL_CUDA(cudaMalloc(&g_sk_part, static_cast<size_t>(pairs) * 2 * 128 * kPairBN * sizeof(float)));
  MMPL_CUDA(cudaMalloc(&g_sk_flags, static_cast<size_t>(pairs) * 2 * sizeof(uint32_t)));
  bump_workspace_generation();
  MMPL_CUDA(cudaMemset(g_sk_flags, 0, static_cast<size_t>(pairs) * 2 * sizeof(uint32_t)));
  g_sk_pairs = pairs;
  return MMPL_OK;
}

template <int BN, int EPI>
static int launch_gemm_pair(const CUtensorMap* ma, const CUtensorMap* mb, GemmParams p, cudaStream_t stream) {
  auto kern = gemm_bf16_pair_kernel<BN, EPI>;
  static bool attr_set = false;
  if (!attr_set) {
    MMPL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kPairSmemBytes));
    attr_set = true;
  }
  p.tiles_m = (p.M + 255) / 256;
  p.tiles_n = (p.N + BN - 1) / BN;
  const int tiles = p.tiles_m * p.tiles_n;
  const int max_pairs = sm_count() / 2;
  int pairs = tiles < max_pairs ? tiles : max_pairs;
  // Schedule (PairSched): whole tiles while they fill every pair; stream-K over the k blocks of the rest when whole
  // tiles would leave more than 4 % of the last wave empty.
  p.tiles_dp = tiles;
  {
    static const int env_gm = getenv("MMPL_GEMM_GROUP_M") ? atoi(getenv("MMPL_GEMM_GROUP_M")) : 0;
    const bool a_fits_l2 = static_cast<int64_t>(p.M) * p.K * 2 <= (32ll << 20);
    // A larger than L2's comfortable share: if a whole row of tiles fits in one wave, run whole rows together (each A
    // panel is then read from HBM once: cfg2's ffn.2 with its 84 MB A read 157 MB with groups of 8 rows, 111 MB is the
    // minimum); otherwise groups of 8 tile rows.
    const int rows_per_wave = max_pairs / p.tiles_n;
    p.group_m = env_gm > 0 ? env_gm : (a_fits_l2 ? p.tiles_m : (rows_per_wave >= 2 ? rows_per_wave : 8));
    if (p.group_m > p.tiles_m) p.group_m = p.tiles_m;
  }
  const int num_kb = (p.K + kBK - 1) / kBK;
  const int waves = (tiles + max_pairs - 1) / max_pairs;
  const double fill = static_cast<double>(tiles) / (static_cast<double>(waves) * max_pairs);
  const bool sk = g_streamk_mode > 0 || (g_streamk_mode < 0 && fill < 0.96 && 2 * tiles >= max_pairs && num_kb >= 8);
  if (sk && tiles % max_pairs != 0 && static_cast<int64_t>(tiles) * num_kb >= 4 * max_pairs) {
    pairs = max_pairs;
    const int full = tiles / pairs;
    p.tiles_dp = full >= 1 ? (full - 1) * pairs : 0;
    const int st = ensure_streamk_workspace(pairs);
    if (st != MMPL_OK) return st;
    p.sk_part = g_sk_part;
    p.sk_flags = g_sk_flags;
  }
  MMPL_CUDA_LAUNCH(launch_kernel(kern, 2 * pairs, kGemmThreads, kPairSmemBytes, stream, *ma, *mb, p));
  MMPL_CUDA(cudaGetLastError());
  return MMPL_OK;
}

template <int BN>
static int dispatch_epi_pair(int epi, const CUtensorMap* ma, const CUtensorMap* mb, const GemmParams& p,
                             cudaStream_t stream) {
  switch (epi) {
    case MMPL_EPI_BIAS: return launch_gemm_pair<BN, MMPL_EPI_BIAS>(ma, mb, p, stream);
    case MMPL_EPI_BIAS_GELU: return launch_gemm_pair<BN, MMPL_EPI_BIAS_GELU>(ma, mb, p, stream);
    case MMPL_EPI_BIAS_SILU: return launch_gemm_pair<BN, MMPL_EPI_BIAS_SILU>(ma, mb, p, stream);
    case MMPL_EPI_BIAS_RES: return launch_gemm_pair<BN, MMPL_EPI_BIAS_RES>(ma, mb, p, stream);
    case MMPL_EPI_BIAS_GATE_RES: return launch_gemm_pair<BN, MMPL_EPI_BIAS_GATE_RES>(ma, mb, p, stream);
    default: set_error("gemm: unknown epilogue %d", epi); return MMPL_ERR_ARG;
  }
}

// Width of the pair tile: the candidate with the lowest (waves x tile width), i.e. the shortest main loop per pair.
static int pick_pair_bn(int M, int N) {
  const int pairs = (sm_count() > 0 ? sm_count() : 148) / 2;
  const int tiles_m = (M + 255) / 256;
  int best_bn = 256;
  long long best = -1;
  for (int bn : {256, 224}) {
    const int tiles = tiles_m * ((N + bn - 1) / bn);
    const long long cost = static_cast<long long>((tiles + pairs - 1) / pairs) * bn;
    if (best < 0 || cost < best) { best = cost; best_bn = bn; }
  }
  return best_bn;
}

template <int BN, int EPI>
static int launch_gemm(const CUtensorMap* ma, const CUtensorMap* mb, GemmParams p, cudaStream_t stream) {
  using Cfg = GemmCfg<BN>;
  auto kern = gemm_bf16_kernel<BN, EPI>;
  static bool attr_set = false;
  if (!attr_set) {
    MMPL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    attr_set = true;
  }
  p.tiles_m = (p.M + kBM - 1) / kBM;
  p.tiles_n = (p.N + BN - 1) / BN;
  const int tiles = p.tiles_m * p.tiles_n;
  const int grid = tiles < sm_count() ? tiles : sm_count();
  MMPL_CUDA_LAUNCH(launch_kernel(kern, grid, kGemmThreads, Cfg::kSmemBytes, stream, *ma, *mb, p));
  MMPL_CUDA(cudaGetLastError());
  return MMPL_OK;
}

template <int BN>
static int dispatch_epi(int epi, const CUtensorMap* ma, const CUtensorMap* mb, const GemmParams& p,
                        cudaStream_t stream) {
  switch (epi) {
    case MMPL_EPI_BIAS: return launch_gemm<BN, MMPL_EPI_BIAS>(ma, mb, p, stream);
    case MMPL_EPI_BIAS_GELU: return launch_gemm<BN, MMPL_EPI_BIAS_GELU>(ma, mb, p, stream);
    case MMPL_EPI_BIAS_SILU: return launch_gemm<BN, MMPL_EPI_BIAS_SILU>(ma, mb, p, stream);
    case MMPL_EPI_BIAS_RES: return launch_gemm<BN, MMPL_EPI_BIAS_RES>(ma, mb, p, stream);
    case MMPL_EPI_BIAS_GATE_RES: return launch_gemm<BN, MMPL_EPI_BIAS_GATE_RES>(ma, mb, p, stream);
    default: set_error("gemm: unknown epilogue %d", epi); return MMPL_ERR_ARG;
  }
}

// Shared-A clusters (gemm_bf16_shared_a_kernel): two 128 x BN tiles per cluster, N must hold an even number of tiles.
template <int BN, int EPI>
static int launch_gemm_shared_a(const CUtensorMap* ma, const CUtensorMap* mb, GemmParams p, cudaStream_t stream) {
  using Cfg = GemmCfg<BN>;
  auto kern = gemm_bf16_shared_a_kernel<BN, EPI>;
  static bool attr_set = false;
  if (!attr_set) {
    MMPL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    attr_set = true;
  }
  p.tiles_m = (p.M + kBM - 1) / kBM;
  p.tiles_n = (p.N + BN - 1) / BN;
  const int tiles = p.tiles_m * (p.tiles_n / 2);
  const int max_clusters = sm_count() / 2;
  const int clusters = tiles < max_clusters ? tiles : max_clusters;
  MMPL_CUDA_LAUNCH(launch_kernel(kern, 2 * clusters, kGemmThreads, Cfg::kSmemBytes, stream, *ma, *mb, p));
  MMPL_CUDA(cudaGetLastError());
  return MMPL_OK;
}

template <int BN>
static int dispatch_epi_shared_a(int epi, const CUtensorMap* ma, const CUtensorMap* mb, const GemmParams& p,
                                 cudaStream_t stream) {
  switch (epi) {
    case MMPL_EPI_BIAS: return launch_gemm_shared_a<BN, MMPL_EPI_BIAS>(ma, mb, p, stream);
    case MMPL_EPI_BIAS_GELU: return launch_gemm_shared_a<BN, MMPL_EPI_BIAS_GELU>(ma, mb, p, stream);
    case MMPL_EPI_BIAS_SILU: return launch_gemm_shared_a<BN, MMPL_EPI_BIAS_SILU>(ma, mb, p, stream);
    case MMPL_EPI_BIAS_RES: return launch_gemm_shared_a<BN, MMPL_EPI_BIAS_RES>(ma, mb, p, stream);
    case MMPL_EPI_BIAS_GATE_RES: return launch_gemm_shared_a<BN, MMPL_EPI_BIAS_GATE_RES>(ma, mb, p, stream);
    default: set_error("gemm: unknown epilogue %d", epi); return MMPL_ERR_ARG;
  }
}

// Kernel for the short-K shapes that gemm_bf16() takes off the pair kernel (see there): 128 until the B200 A/B of the
// candidates says otherwise.
constexpr int kShortKDefault = 128;

// Tile width: 256 when it keeps the machine full, otherwise narrower tiles for more CTAs.
static int pick_bn(int M, int N) {
  if (N % 128 != 0 || N < 128) return 64;
  const int tiles_m = (M + kBM - 1) / kBM;
  if (N % 256 == 0) {
    const int t256 = tiles_m * (N / 256);
    const int sms = sm_count() > 0 ? sm_count() : 148;
    // wave efficiency of each choice: tiles / (ceil(tiles / sms) * sms)
    const int t128 = tiles_m * (N / 128);
    const double e256 = double(t256) / (double((t256 + sms - 1) / sms) * sms);
    const double e128 = double(t128) / (double((t128 + sms - 1) / sms) * sms);
    return (e256 + 0.08 >= e128) ? 256 : 128;
  }
  return 128;
}

int gemm_bf16(const void* a, int64_t lda, const void* w, int64_t ldw, const void* bias, void* out,
              int64_t ldo, int M, int N, int K, int epilogue, const void* residual, int64_t ldr,
              const void* gate, int64_t gate_stride, int rows_per_frame, int force_bn,
              cudaStream_t stream) {
  MMPL_CHECK(device_is_sm100(), MMPL_ERR_ARCH, "gemm: requires an sm_100 device");
  MMPL_CHECK(M > 0 && N > 0 && K > 0, MMPL_ERR_SHAPE, "gemm: bad shape M=%d N=%d K=%d", M, N, K);
  MMPL_CHECK(N % 8 == 0 && K % 8 == 0 && ldo % 8 == 0, MMPL_ERR_SHAPE,
             "gemm: N, K and ldo must be multiples of 8 (N=%d K=%d ldo=%lld)", N, K, (long long)ldo);
  if (epilogue == MMPL_EPI_BIAS_RES || epilogue == MMPL_EPI_BIAS_GATE_RES)
    MMPL_CHECK(residual != nullptr && ldr % 8 == 0, MMPL_ERR_ARG, "gemm: residual epilogue needs residual with ldr %% 8 == 0");
  if (epilogue == MMPL_EPI_BIAS_GATE_RES)
    MMPL_CHECK(gate != nullptr && rows_per_frame > 0 && gate_stride % 8 == 0, MMPL_ERR_ARG,
               "gemm: gate epilogue needs gate, rows_per_frame > 0 and gate_stride %% 8 == 0");
  // tile_n 512 / 448 select the cta_group::2 kernel (256 x 256 / 256 x 224 tile per CTA pair); auto-selected for large problems
  bool use_pair = force_bn == 512 || force_bn == 448 || (force_bn == 0 && N % 256 == 0 && M >= 512);
  if (use_pair && force_bn == 0 && K <= 2048) {
    // Short-K problems whose 256 x 256 tiles leave the last wave mostly empty while 128 x 128 tiles fill theirs:
    // cfg2's o / cross-attention projections (M = 4680, N = K = 1536) are 114 pair tiles on 74 pairs (1.54 waves) but
    // 444 = 3 x 148 single-CTA tiles; measured 20.3 us against 22.3 us. With a long K the pair kernel still wins.
    const int sms = sm_count() > 0 ? sm_count() : 148;
    const int tp = ((M + 255) / 256) * (N / 256), ts = ((M + kBM - 1) / kBM) * (N / 128);
    const double fill_pair = double(tp) / (double((tp + sms / 2 - 1) / (sms / 2)) * (sms / 2));
    const double fill_single = double(ts) / (double((ts + sms - 1) / sms) * sms);
    if (fill_pair < 0.8 && fill_single > 0.95) {
      use_pair = false;
      // which short-K kernel: 128 = single-CTA 128 x 128 tiles; 192 = single-CTA 128 x 192; 1128 / 1192 = shared-A
      // clusters of two 128 x 128 / 128 x 192 tiles (gemm_bf16_shared_a_kernel). MMPL_GEMM_SHORT_K overrides (A/B).
      static const int env_sk = getenv("MMPL_GEMM_SHORT_K") ? atoi(getenv("MMPL_GEMM_SHORT_K")) : 0;
      const int want = env_sk ? env_sk : kShortKDefault;
      force_bn = 128;
      if (want == 192 && N % 192 == 0) force_bn = 192;
      if (want == 1128) force_bn = 1128;  // N % 256 == 0 here: an even number of 128-wide tiles
      if (want == 1192 && N % 384 == 0) force_bn = 1192;
    }
  }
  if (force_bn == 1128 || force_bn == 1192) {
    const int bn = force_bn - 1000;
    MMPL_CHECK(((N + bn - 1) / bn) % 2 == 0, MMPL_ERR_ARG, "gemm: shared-A clusters need an even number of %d-wide tiles (N=%d)", bn, N);
    const CUtensorMap* ma = get_tensor_map_bf16(a, M, K, lda, kBM / 2);  // each CTA loads (and multicasts) half of the A tile
    const CUtensorMap* mb = get_tensor_map_bf16(w, N, K, ldw, bn);
    if (!ma || !mb) return MMPL_ERR_CUDA;
    GemmParams pp{};
    pp.M = M; pp.N = N; pp.K = K;
    pp.out = static_cast<__nv_bfloat16*>(out);
    pp.ldo = ldo;
    pp.bias = static_cast<const __nv_bfloat16*>(bias);
    pp.res = static_cast<const __nv_bfloat16*>(residual);
    pp.ldr = ldr;
    pp.gate = static_cast<const __nv_bfloat16*>(gate);
    pp.gate_stride = gate_stride;
    pp.rows_per_frame = rows_per_frame > 0 ? rows_per_frame : 1;
    return bn == 192 ? dispatch_epi_shared_a<192>(epilogue, ma, mb, pp, stream) : dispatch_epi_shared_a<128>(epilogue, ma, mb, pp, stream);
  }
  if (use_pair) {
    MMPL_CHECK(N % 8 == 0, MMPL_ERR_SHAPE, "gemm: N must be a multiple of 8");
    static const int env_bn = getenv("MMPL_GEMM_PAIR_BN") ? atoi(getenv("MMPL_GEMM_PAIR_BN")) : 0;
    const int pbn = force_bn == 448 ? 224 : (force_bn == 512 ? 256 : (env_bn == 224 || env_bn == 256 ? env_bn : pick_pair_bn(M, N)));
    const CUtensorMap* pa = get_tensor_map_bf16(a, M, K, lda, kBM);
    const CUtensorMap* pb = get_tensor_map_bf16(w, N, K, ldw, pbn / 2);  // each CTA loads half of the tile's W rows
    if (!pa || !pb) return MMPL_ERR_CUDA;
    GemmParams pp{};
    pp.M = M; pp.N = N; pp.K = K;
    pp.out = static_cast<__nv_bfloat16*>(out);
    pp.ldo = ldo;
    pp.bias = static_cast<const __nv_bfloat16*>(bias);
    pp.res = static_cast<const __nv_bfloat16*>(residual);
    pp.ldr = ldr;
    pp.gate = static_cast<const __nv_bfloat16*>(gate);
    pp.gate_stride = gate_stride;
    pp.rows_per_frame = rows_per_frame > 0 ? rows_per_frame : 1;
    return pbn == 224 ? dispatch_epi_pair<224>(epilogue, pa, pb, pp, stream) : dispatch_epi_pair<256>(epilogue, pa, pb, pp, stream);
  }
  const int bn = force_bn ? force_bn : pick_bn(M, N);
  MMPL_CHECK(bn == 64 || bn == 128 || bn == 192 || bn == 256, MMPL_ERR_ARG, "gemm: tile width %d not supported", bn);

  const CUtensorMap* ma = get_tensor_map_bf16(a, M, K, lda, kBM);
  const CUtensorMap* mb = get_tensor_map_bf16(w, N, K, ldw, bn);
  if (!ma || !mb) return MMPL_ERR_CUDA;

  GemmParams p{};
  p.M = M; p.N = N; p.K = K;
  p.out = static_cast<__nv_bfloat16*>(out);
  p.ldo = ldo;
  p.bias = static_cast<const __nv_bfloat16*>(bias);
  p.res = static_cast<const __nv_bfloat16*>(residual);
  p.ldr = ldr;
  p.gate = static_cast<const __nv_bfloat16*>(gate);
  p.gate_stride = gate_stride;
  p.rows_per_frame = rows_per_frame > 0 ? rows_per_frame : 1;
  if (bn == 256) return dispatch_epi<256>(epilogue, ma, mb, p, stream);
  if (bn == 192) return dispatch_epi<192>(epilogue, ma, mb, p, stream);
  if (bn == 128) return dispatch_epi<128>(epilogue, ma, mb, p, stream);
  return dispatch_epi<64>(epilogue, ma, mb, p, stream);
}

// ------------------------------------------------------------------------------------------------------------------
// Causal 3-D convolution as a tap-GEMM (wan/modules/vae.py:16-36 CausalConv3d, stride 1; also its per-frame Conv2d
// 3x3 and the 1x1 convolutions), channels-last:
//   in  [history + T][H + 2][W + 2][Cin]    bf16, one-pixel halo of zeros; `history` (0..KT-1) carried frames in front,
//                                            the KT-1-history frames before them are zeros that are never stored: their
//                                            rows have negative coordinates, which TMA fills with zeros
//   out [T][H + 2][W + 2][Cout]              bf16; interior positions = the convolution, halo positions = 0
// out(t, h, w, :) = bias + sum_taps in(t + dt - (KT-1), h + dh - KH/2, w + dw - KW/2, :) . w[:, tap, :]  (+ residual)
// Every grid position (halo included) is one GEMM row; a tap is a row shift of the same A matrix, so the A operand
// is read in place by TMA: no im2col buffer. The halo rows cost (H+2)(W+2)/(HW) - 1 extra MMA work (0.7 % at 480x832).
// K layout. KW = 1: one K span of Cin per tap. KW = 3: the three dw taps of one (dt, dh) are three CONSECUTIVE rows of
// the channels-last grid, i.e. one contiguous span of 3*Cin elements starting at row r - 1; the A tensor map describes
// that directly (row pitch Cin, row length 3*Cin: rows overlap), so K per (dt, dh) is 3*Cin rounded up to 64 instead of
// three times Cin rounded up to 64: 320 instead of 384 at the VAE's 96-channel level, 64 instead of 192 for its 3- and
// 16-channel ends.
//   w   KW = 1: [Cout][KT*KH][Cin64]; KW = 3: [Cout][KT*KH][pad64(3*Cin)] with (dw, c) order inside a span; bf16, zero padded
template <int BN, int EPI>
static int launch_conv(const CUtensorMap* ma, const CUtensorMap* mb, ConvGemmParams p, cudaStream_t stream) {
  using Cfg = GemmCfg<BN>;
  auto kern = gemm_bf16_kernel<BN, EPI, true>;
  static bool attr_set = false;
  if (!attr_set) {
    MMPL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    attr_set = true;
  }
  p.tiles_m = (p.M + kBM - 1) / kBM;
  p.tiles_n = (p.N + BN - 1) / BN;
  const int tiles = p.tiles_m * p.tiles_n;
  const int grid = tiles < sm_count() ? tiles : sm_count();
  MMPL_CUDA_LAUNCH(launch_kernel(kern, grid, kGemmThreads, Cfg::kSmemBytes, stream, *ma, *mb, p));
  MMPL_CUDA(cudaGetLastError());
  return MMPL_OK;
}

int conv3d_cl(const void* in, const void* w_packed, const void* bias, void* out, const void* residual, int T, int H,
              int W, int Cin, int Cout, int KT, int KH, int KW, int history, cudaStream_t stream) {
  MMPL_CHECK(device_is_sm100(), MMPL_ERR_ARCH, "conv3d: requires an sm_100 device");
  MMPL_CHECK(in && w_packed && out, MMPL_ERR_ARG, "conv3d: null argument");
  MMPL_CHECK(T > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0, MMPL_ERR_SHAPE, "conv3d: bad shape");
  MMPL_CHECK(Cin % 8 == 0 && Cout % 8 == 0, MMPL_ERR_SHAPE,
             "conv3d: channel counts must be multiples of 8 (pad the layout): Cin=%d Cout=%d", Cin, Cout);
  MMPL_CHECK((KT == 1 || KT == 3) && (KH == 1 || KH == 3) && KH == KW, MMPL_ERR_SHAPE,
             "conv3d: kernel %dx%dx%d not supported (1 or 3 per axis, square)", KT, KH, KW);
  MMPL_CHECK(history >= 0 && history <= KT - 1, MMPL_ERR_ARG, "conv3d: history=%d frames, the kernel takes 0..%d", history, KT - 1);
  const int Hp = H + 2, Wp = W + 2;
  const int64_t rows_out = static_cast<int64_t>(T) * Hp * Wp;
  const int64_t rows_in = static_cast<int64_t>(T + history) * Hp * Wp;
  MMPL_CHECK(rows_in < (int64_t(1) << 31) - 65536, MMPL_ERR_SHAPE, "conv3d: %lld grid positions exceed the 32-bit row index",
             (long long)rows_in);
  const bool rowpack = KW == 3;
  const int spans = rowpack ? KT * KH : KT * KH * KW;     // K spans ("taps" of the kernel's k loop)
  const int span = rowpack ? 3 * Cin : Cin;
  const int kb_per_span = (span + kBK - 1) / kBK;

  ConvGemmParams p{};
  p.M = static_cast<int>(rows_out);
  p.N = Cout;
  p.K = spans * kb_per_span * kBK;
  p.out = static_cast<__nv_bfloat16*>(out);
  p.ldo = Cout;
  p.bias = static_cast<const __nv_bfloat16*>(bias);
  p.res = static_cast<const __nv_bfloat16*>(residual);
  p.ldr = Cout;
  p.rows_per_frame = 1;
  p.conv_kb_per_tap = kb_per_span;
  p.conv_hp = Hp;
  p.conv_wp = Wp;
  int i = 0;
  for (int dt = 0; dt < KT; ++dt)
    for (int dh = 0; dh < KH; ++dh) {
      const int frame_row = ((dt - (KT - 1) + history) * Hp + (dh - KH / 2)) * Wp;  // first row of the tap's image row
      if (rowpack) {
        p.conv_tap_off[i++] = frame_row - 1;
      } else {
        for (int dw = 0; dw < KW; ++dw) p.conv_tap_off[i++] = frame_row + (dw - KW / 2);
      }
    }

  // tile width: 96 for the VAE's 96- and 192-channel levels (UMMA N = 96; a 128- / 256-wide tile would compute a quarter
  // of its columns for nothing), otherwise the narrowest of 64 / 128 / 256 that holds Cout or divides it
  int bn = Cout <= 64 ? 64 : ((Cout == 96 || Cout == 192) ? 96 : (Cout <= 128 ? 128 : ((Cout <= 256 || Cout % 256 == 0) ? 256 : 128)));
  static const bool no_bn96 = getenv("MMPL_CONV_BN96") != nullptr && atoi(getenv("MMPL_CONV_BN96")) == 0;  // A/B switch
  if (no_bn96 && bn == 96) bn = Cout <= 128 ? 128 : 256;
  // KW = 3: rows of 3*Cin elements that start Cin apart (overlapping); the last two grid positions are left out of the
  // map so that no row of it reaches past the buffer (they are halo positions, and TMA zero-fills them as operands)
  const CUtensorMap* ma = rowpack ? get_tensor_map_bf16(in, static_cast<uint64_t>(rows_in - 2), 3 * Cin, Cin, kBM)
                                  : get_tensor_map_bf16(in, static_cast<uint64_t>(rows_in), Cin, Cin, kBM);
  const CUtensorMap* mb = get_tensor_map_bf16(w_packed, Cout, static_cast<uint64_t>(p.K), static_cast<uint64_t>(p.K), bn);
  if (!ma || !mb) return MMPL_ERR_CUDA;
  const bool res = residual != nullptr;
  if (bn == 96) return res ? launch_conv<96, MMPL_EPI_BIAS_RES>(ma, mb, p, stream) : launch_conv<96, MMPL_EPI_BIAS>(ma, mb, p, stream);
  if (bn == 256) return res ? launch_conv<256, MMPL_EPI_BIAS_RES>(ma, mb, p, stream) : launch_conv<256, MMPL_EPI_BIAS>(ma, mb, p, stream);
  if (bn == 128) return res ? launch_conv<128, MMPL_EPI_BIAS_RES>(ma, mb, p, stream) : launch_conv<128, MMPL_EPI_BIAS>(ma, mb, p, stream);
  return res ? launch_conv<64, MMPL_EPI_BIAS_RES>(ma, mb, p, stream) : launch_conv<64, MMPL_EPI_BIAS>(ma, mb, p, stream);
}

}  // namespace mmpl
