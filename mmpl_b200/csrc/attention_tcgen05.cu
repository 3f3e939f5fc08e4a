// tcgen05/TMEM flash attention for sm_100a, head_dim 128, bf16, no mask:
//   out[q, h, :] = softmax(Q[q,h,:] . K[kv,h,:]^T / sqrt(128)) . V[kv,h,:]
// Replaces the flash_attn_varlen_func call behind wan/modules/attention.py:32-136 for the two hot
// call sites: self-attention over the KV cache (wan/modules/causal_model.py:220-224, KV rows
// [max(0,local_end-32760), local_end) read in place from the cache tensor) and cross-attention over
// the 512 cached text keys (wan/modules/model.py:189). The KV operand is a list of up to 8 row
// segments, so the same kernel serves the frame-slot visibility gather of CausalFPSWanModel
// (wan/modules/causal_fps_model.py:219-264) without materialising cache[:, all_indices]; a segment
// may also come from a second (K,V) pair, which is the "cat(gathered cache, new K/V)" of the last
// MMPL stage (:254-264).
//
// Work unit = one head x two 128-row query tiles that ping-pong on the tensor pipe, against T KV tiles
// of 128 rows. The kernel is persistent; to balance the SMs every unit's KV range is cut into `split`
// equal chunks ("pieces", split chosen by a cost model on the host) and CTA b runs pieces b, b+G, ...
// Pieces are ordered (head, chunk, query pair), so the CTAs that run side by side read the same K/V tiles
// and share them through L2. With split == 1 a piece is a whole unit and writes the normalised bf16
// output directly; otherwise it writes its un-normalised fp32 O and (max, sum) to a workspace and
// attn_combine_kernel merges the chunks. (cfg2: 228 units on 148 SMs = 1.54 waves of work; a
// one-CTA-per-unit grid takes 2 full waves.)
//   warp 0      TMA producer (Q per piece; K and V tiles through 2-stage rings)
//   warp 1      MMA issuer:  S_q = Q_q K^T (SS),  O_q += P_q V (A = P from TMEM, B = V MN-major)
//   warps 4-7   softmax for query tile 0 (thread = row), warps 8-11 for query tile 1
// TMEM (512 columns): S0 [0,128) S1 [128,256) O0 [256,384) O1 [384,512); bf16 P_q overwrites the
// first 64 columns of S_q. O is rescaled lazily (only when the running max grows by > 2^8), by the
// softmax warps themselves.
#include <cstdlib>
#include <type_traits>

#include "host_util.h"
#include "mmpl_b200.h"
#include "ptx.cuh"

#ifndef MMPL_ATTN_NS
#define MMPL_ATTN_NS prod
#define MMPL_ATTN_PROD 1
#endif
#ifndef MMPL_ATTN_HALVES
#define MMPL_ATTN_HALVES 1
#endif
// 0 (namespace prod): S of a KV tile is handed to the softmax warps whole. 1 (namespace half, compiled by
// attention_tcgen05_half.cu): S is produced, exponentiated and consumed in two independent 64-column halves, which
// shortens the S -> P -> P.V -> next S round trip. Measured on B200 (profiles/r01_microbench_attention_halftile.txt):
// the half-tile pipeline is 22 % faster for the 4-tile cross-attention (23.6 vs 30.2 us) and 0-6 % slower for the
// long self-attention ranges (twice the barriers, fences and N=64 MMAs per tile), and inside the cfg2 step the
// cross-attention gain does not show (972.4 vs 974.1 ms per step): opt-in through the dispatcher
// (attention_dispatch.cu: MMPL_ATTN_HALF / MMPL_ATTN_HALF_TILES), the whole-tile build is the default everywhere.
#ifndef MMPL_ATTN_SPLIT_S
#define MMPL_ATTN_SPLIT_S 0
#endif

#ifndef MMPL_ATTN_TIMING
#define MMPL_ATTN_TIMING 0
#endif
#if MMPL_ATTN_TIMING
// Development build only (tools/attn_timing.py): per-phase clock64() totals of CTA 0.
//   [qt*4 + 0..3] softmax warp of query tile qt: waiting for S, computing P, storing P + arrive, tiles
//   [8 + qt] MMA warp waiting for P_qt   [10] waiting for V   [11] waiting for K   [12] MMA warp total   [13] tiles
__device__ long long g_attn_dbg[32];
#define TSTAMP(var) const long long var = clock64()
#define TACC(slot, val) dbg[slot] += (val)
#else
#define TSTAMP(var)
#define TACC(slot, val)
#endif

namespace mmpl {
namespace MMPL_ATTN_NS {

// Warpgroup 0 = control (warp 0 TMA producer, warp 1 MMA issuer, 2 idle); warpgroups 1, 2 = softmax of query tile 0 / 1.
// setmaxnreg moves registers from the control warpgroup to the softmax warpgroups; the sum must stay within the launch
// allocation (threads x registers ptxas reports): exceeding it deadlocks setmaxnreg.inc.
// 12 warps, one thread per query row: 128 x 96 + 256 x 200 = 63488 <= 384 x 168.
constexpr int kAttnThreads = 384;
constexpr int kSoftmaxWarpsPerTile = 4;
#ifndef MMPL_ATTN_CTRL_REGS
#define MMPL_ATTN_CTRL_REGS 96
#define MMPL_ATTN_SOFTMAX_REGS 200
#endif
constexpr int kFirstSoftmaxWarp = 4;
constexpr int kQTile = 128;
constexpr int kKVTile = 128;
constexpr int kHD = 128;
constexpr int kBoxBytes = 128 * 64 * 2;  // one [128 rows][64 bf16] swizzled box = 16 KB
constexpr int kKVStages = 2;
constexpr int kAttnSmem = 4 * kBoxBytes /*Q*/ + kKVStages * 2 * kBoxBytes /*K*/ +
                          kKVStages * 2 * kBoxBytes /*V*/ + 1024 + 256 /*barriers*/;
constexpr int kMaxSeg = 8;
constexpr int kUnitRows = 2 * kQTile;
#ifndef MMPL_ATTN_POLY_NUM
#define MMPL_ATTN_POLY_NUM 4
#endif
constexpr int kPolyNum = MMPL_ATTN_POLY_NUM;  // pairs per kPolyDen whose exp2 runs on the FMA/ALU pipes
constexpr int kPolyDen = 16;
// Which of every 16 element pairs take the polynomial path: alternating with the MUFU pairs in program order, so that
// the FMA-pipe work of a polynomial pair issues in the shadow of the 8-clk MUFU instructions next to it (ptxas keeps
// close to source order; with the polynomial pairs bunched at the front of each group the two pipes took turns).
__host__ __device__ constexpr bool pair_is_poly(int i) {
  return ((i % kPolyDen) & 1) == 0 && (i % kPolyDen) < 2 * kPolyNum;
}

struct AttnParams {
  int Lq, H;
  int QP;         // query-tile pairs per head
  int T;          // KV tiles per unit
  int split;      // KV chunks per unit (= workspace slots per unit)
  int n_pieces;   // U * split
  __nv_bfloat16* out;
  int64_t ldo;
  float scale_log2;
  float* part_o;   // [U][P][256][128] un-normalised O of partial pieces
  float* part_ml;  // [U][P][256][2]   (running max in the scaled log2 domain, running sum)
  int nseg;
  int seg_start[kMaxSeg];
  int seg_rows[kMaxSeg];
  int seg_src[kMaxSeg];
  // range schedule (PieceIter): heads are taken in groups of `hg`; 0 = uniform-split schedule
  int hg;
  int G;          // CTAs of flash_attn_kernel (needed to locate the pieces of a unit when merging)
  // hybrid schedule (hg == H): units [0, u_base) run whole, CTA c takes units c, c+G, ... (u_base = rounds * G); only
  // the remaining units [u_base, U) are cut into ranges
  int u_base;
  // merge of partial pieces inside flash_attn_kernel: per (unit, query tile) arrival counters (zero between launches),
  // nullptr = leave the merge to attn_combine_kernel
  int* merge_cnt;
};

// A piece = KV tiles [t0, t0 + n) of one unit. `whole`: the piece is the entire unit and writes the normalised bf16
// output; otherwise it leaves un-normalised fp32 O and (max, sum) in workspace slot `slot` for attn_combine_kernel.
struct Piece { int head, q_row0, t0, n, slot; bool whole; };

// Two schedules, chosen by the host (flash_attn_impl):
//  * uniform split (hg == 0): every unit's KV range is cut into `split` equal chunks; pieces are ordered (head, chunk,
//    query pair) and CTA b runs pieces b, b+G, ... Used for short KV ranges (cross-attention: 4 tiles per unit).
//  * ranges (hg > 0): the heads are taken in groups of hg (sized so that a group's K/V stays L2-resident). Inside a
//    group the (unit, KV tile) pairs, unit-major, form one sequence that is cut into G equal contiguous ranges, CTA c
//    takes range c of every group in turn. All CTAs get the same number of KV tiles (228 units on 148 SMs is 1.54
//    units each, which whole units or equal chunks can only approximate), and a CTA starts 2-3 pieces per group
//    instead of one per chunk. A unit that straddles a range boundary is merged by attn_combine_kernel; the slot
//    of a partial piece is (group, CTA, piece starts its CTA's range ? 0 : 1), so 2*G slots per group suffice.
//  * hybrid (hg == H, u_base > 0): U units on G CTAs is U/G = r + f units each. The first r*G units run whole, one per
//    CTA and round, neighbouring CTAs on the same head streaming its K/V in lockstep from tile 0 (the uniform
//    schedule's L2 sharing, no partials); only the last f*G units go through the range schedule, so every CTA still
//    gets the same number of KV tiles, the partials are few (2 per CTA) and merged in the kernel while L2-hot, and the
//    K/V the free-running CTAs of that phase touch is f*G/QP heads' worth instead of all of it (cfg2: 228 units =
//    148 whole + 80 ranged over 5 heads).
struct RangeGroup {
  int heads, units;     // heads and units in this group (units: only those the range schedule covers)
  int first_unit;       // global index of the group's first ranged unit
  int W;                // KV-tile units in this group = units * T  (host guarantees W * (G + 1) < 2^31)
  __host__ __device__ __forceinline__ void set(const AttnParams& p, int g) {
    heads = p.hg < p.H - g * p.hg ? p.hg : p.H - g * p.hg;
    first_unit = g * p.hg * p.QP + p.u_base;  // u_base > 0 only with a single group
    units = heads * p.QP - p.u_base;
    W = units * p.T;
  }
  __host__ __device__ __forceinline__ int lo(int c, int G) const { return c * W / G; }
  // the CTA whose range contains position x: largest c with lo(c) <= x
  __host__ __device__ __forceinline__ int cta_of(int x, int G) const { return ((x + 1) * G + W - 1) / W - 1; }
};

struct PieceIter {
  int G, c;
  // uniform split
  int pi;
  // ranges
  int g, ngroups;
  int pos, lo, hi;
  int wu;  // hybrid: next whole unit of this CTA
  RangeGroup grp;
  __host__ __device__ __forceinline__ void init(const AttnParams& p, int cta, int grid) {
    G = grid;
    c = cta;
    pi = cta;
    wu = cta;
    g = -1;
    ngroups = p.hg > 0 ? (p.H + p.hg - 1) / p.hg : 0;
    pos = hi = lo = 0;
  }
  __host__ __device__ __forceinline__ bool next(const AttnParams& p, Piece& pc) {
    if (p.hg == 0) {
      if (pi >= p.n_pieces) return false;
      const int per_head = p.split * p.QP;
      pc.head = pi / per_head;
      const int r = pi - pc.head * per_head;
      const int chunk = r / p.QP;
      const int qp = r - chunk * p.QP;
      pc.q_row0 = qp * (2 * kQTile);
      pc.t0 = chunk * p.T / p.split;
      pc.n = (chunk + 1) * p.T / p.split - pc.t0;
      pc.slot = (pc.head * p.QP + qp) * p.split + chunk;
      pc.whole = p.split == 1;
      pi += G;
      return true;
    }
    if (wu < p.u_base) {
      pc.head = wu / p.QP;
      pc.q_row0 = (wu - pc.head * p.QP) * (2 * kQTile);
      pc.t0 = 0;
      pc.n = p.T;
      pc.whole = true;
      pc.slot = 0;
      wu += G;
      return true;
    }
    while (pos >= hi) {
      if (++g >= ngroups) return false;
      grp.set(p, g);
      lo = pos = grp.lo(c, G);
      hi = grp.lo(c + 1, G);
    }
    const int ul = pos / p.T;
    const int u = grp.first_unit + ul;
    pc.head = u / p.QP;
    pc.q_row0 = (u - pc.head * p.QP) * (2 * kQTile);
    pc.t0 = pos - ul * p.T;
    const int end = hi < (ul + 1) * p.T ? hi : (ul + 1) * p.T;
    pc.n = end - pos;
    pc.whole = pc.n == p.T;
    pc.slot = (g * G + c) * 2 + (pos == lo ? 0 : 1);
    pos = end;
    return true;
  }
};

// The pieces a unit was cut into: uniform split -> slots u*split + i; range schedule -> one per CTA whose range
// overlaps the unit (same arithmetic as PieceIter; a CTA whose range is empty holds none: slot() = -1).
struct UnitPieces {
  int np, u, c_first, g, start;
  RangeGroup grp;
  __host__ __device__ __forceinline__ void init(const AttnParams& p, int unit) {
    u = unit;
    c_first = g = start = 0;
    if (p.hg == 0) {
      np = p.split;
    } else {
      g = (u / p.QP) / p.hg;
      grp.set(p, g);
      start = (u - grp.first_unit) * p.T;  // callers never pass a whole unit of the hybrid schedule (u < u_base)
      c_first = grp.cta_of(start, p.G);
      np = grp.cta_of(start + p.T - 1, p.G) - c_first + 1;
    }
  }
  __host__ __device__ __forceinline__ int slot(const AttnParams& p, int i) const {
    if (p.hg == 0) return u * p.split + i;
    const int c = c_first + i;
    const int lo = grp.lo(c, p.G);
    if (grp.lo(c + 1, p.G) <= lo) return -1;
    return (g * p.G + c) * 2 + (lo >= start ? 0 : 1);  // does the piece open its CTA's range?
  }
  // pieces that really exist (the arrival count that makes a piece the last one)
  __host__ __device__ __forceinline__ int count(const AttnParams& p) const {
    if (p.hg == 0) return np;
    int n = 0;
    for (int i = 0; i < np; ++i) n += slot(p, i) >= 0 ? 1 : 0;
    return n;
  }
};

// Select over the (at most 8) segment parameters without indexing the kernel-parameter arrays dynamically
// (a dynamic index would make the compiler copy them to local memory).
__device__ __forceinline__ int seg_field(const int (&a)[kMaxSeg], int i) {
  int r = a[0];
#pragma unroll
  for (int k = 1; k < kMaxSeg; ++k) r = (i == k) ? a[k] : r;
  return r;
}

// Walks the KV tiles of a unit in order: source, first row and number of valid rows of the current tile.
struct TileIter {
  int seg, src, row, rows_left;
  __device__ __forceinline__ void load_seg(const AttnParams& p) {
    src = seg_field(p.seg_src, seg);
    row = seg_field(p.seg_start, seg);
    rows_left = seg_field(p.seg_rows, seg);
  }
  __device__ __forceinline__ void init(const AttnParams& p, int j) {  // position on tile j of the unit
    seg = 0;
    load_seg(p);
    while (seg < p.nseg - 1) {
      const int nt = (rows_left + kKVTile - 1) / kKVTile;
      if (j < nt) break;
      j -= nt;
      ++seg;
      load_seg(p);
    }
    row += j * kKVTile;
    rows_left -= j * kKVTile;
  }
  __device__ __forceinline__ int valid() const { return min(kKVTile, rows_left); }
  __device__ __forceinline__ void next(const AttnParams& p) {
    row += kKVTile;
    rows_left -= kKVTile;
    if (rows_left <= 0 && seg < p.nseg - 1) {
      ++seg;
      load_seg(p);
    }
  }
};

__global__ void __launch_bounds__(kAttnThreads, 1)
flash_attn_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k0,
                  const __grid_constant__ CUtensorMap map_v0, const __grid_constant__ CUtensorMap map_k1,
                  const __grid_constant__ CUtensorMap map_v1, const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_q = smem;                                   // [qtile][hd half] boxes
  uint8_t* smem_k = smem_q + 4 * kBoxBytes;                 // [stage][hd half]
  uint8_t* smem_v = smem_k + kKVStages * 2 * kBoxBytes;     // [stage][hd half]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_v + kKVStages * 2 * kBoxBytes);
  uint64_t* q_full = bars;                 // 1
  uint64_t* q_empty = bars + 1;            // 1
  uint64_t* k_full = bars + 2;             // 2
  uint64_t* k_empty = bars + 4;            // 2
  uint64_t* v_full = bars + 6;             // 2
  uint64_t* v_empty = bars + 8;            // 2
  uint64_t* s_full = bars + 10;            // 4: [query tile][half of the KV tile] (whole-tile variant: [query tile])
  uint64_t* p_full = bars + 14;            // 4: [query tile][half of the KV tile]
  uint64_t* o_full = bars + 18;            // 1
  uint64_t* o_empty = bars + 19;           // 2
  uint64_t* pv_last = bars + 21;           // 2 (half-tile pipeline): first P.V half of a piece's last KV tile has completed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 23);
  int* merge_ticket = reinterpret_cast<int*>(bars + 24);  // [query tile]: arrival ticket of the current partial piece

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int G = gridDim.x;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_q);
    tma_prefetch_desc(&map_k0);
    tma_prefetch_desc(&map_v0);
  }
  if (warp == 1) {
    if (lane == 0) {
      mbar_init(q_full, 1);
      mbar_init(q_empty, 1);
      for (int i = 0; i < kKVStages; ++i) {
        mbar_init(&k_full[i], 1);
        mbar_init(&k_empty[i], 1);
        mbar_init(&v_full[i], 1);
        mbar_init(&v_empty[i], 1);
      }
      for (int i = 0; i < 2; ++i) {
        mbar_init(&s_full[i], 1);
        mbar_init(&s_full[2 + i], 1);
        mbar_init(&p_full[2 * i], kSoftmaxWarpsPerTile);
        mbar_init(&p_full[2 * i + 1], kSoftmaxWarpsPerTile);
        mbar_init(&o_empty[i], kSoftmaxWarpsPerTile);
        mbar_init(&pv_last[i], 1);
      }
      mbar_init(o_full, 1);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();  // PDL: everything above overlapped the tail of the previous kernel; Q/K/V are valid from here on
  pdl_launch_dependents();

  if (warp < kFirstSoftmaxWarp) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(MMPL_ATTN_CTRL_REGS));
    if (warp == 0) {
      // -------------------------------------------------------- TMA producer
      // (whole warp, converged; one elected lane issues each TMA: see the "_elect" wrappers in ptx.cuh)
      int g = 0, piece = 0;
      PieceIter pit;
      pit.init(p, blockIdx.x, G);
      Piece pc;
      for (; pit.next(p, pc); ++piece) {
        const int t0 = pc.t0, n = pc.n, head = pc.head, q_row0 = pc.q_row0;
        mbar_wait(q_empty, (piece & 1) ^ 1);
        mbar_arrive_expect_tx_elect(q_full, 4 * kBoxBytes);
        for (int qt = 0; qt < 2; ++qt)
          for (int hh = 0; hh < 2; ++hh)
            tma_load_2d_elect(smem_q + (qt * 2 + hh) * kBoxBytes, &map_q, q_full, head * kHD + hh * 64,
                              q_row0 + qt * kQTile, kEvictFirst);
        TileIter it;
        it.init(p, t0);
        for (int jj = 0; jj < n; ++jj, ++g, it.next(p)) {
          const CUtensorMap* mk = it.src ? &map_k1 : &map_k0;
          const CUtensorMap* mv = it.src ? &map_v1 : &map_v0;
          const int st = g & 1;
          const uint32_t ph = (g >> 1) & 1;
          mbar_wait(&k_empty[st], ph ^ 1);
          mbar_arrive_expect_tx_elect(&k_full[st], 2 * kBoxBytes);
          for (int hh = 0; hh < 2; ++hh)
            tma_load_2d_elect(smem_k + (st * 2 + hh) * kBoxBytes, mk, &k_full[st], head * kHD + hh * 64, it.row, kEvictLast);
          mbar_wait(&v_empty[st], ph ^ 1);
          mbar_arrive_expect_tx_elect(&v_full[st], 2 * kBoxBytes);
          for (int hh = 0; hh < 2; ++hh)
            tma_load_2d_elect(smem_v + (st * 2 + hh) * kBoxBytes, mv, &v_full[st], head * kHD + hh * 64, it.row, kEvictLast);
        }
      }
    } else if (warp == 1) {
      // ---------------------------------------------------------- MMA issuer
      constexpr uint32_t idesc_qk = make_idesc_bf16(128, 128, 0, 0);
      constexpr uint32_t idesc_pv = make_idesc_bf16(128, 128, 0, 1);
      const uint32_t q_addr = smem_u32(smem_q);
      const uint32_t k_addr = smem_u32(smem_k);
      const uint32_t v_addr = smem_u32(smem_v);

      // S_q = Q_q . K(stage)^T : 8 k-steps of 16 over head_dim (two 64-wide swizzled halves)
      // (every lane runs the issue code, one elected lane issues each tcgen05 instruction: ptx.cuh "_elect" wrappers;
      //  descriptors = constant high word + (smem address >> 4) in the low word)
      constexpr uint64_t kDescK = make_smem_desc_sw128_const(16, 1024);          // K-major Q / K tiles
      constexpr uint64_t kDescV = make_smem_desc_sw128_const(kBoxBytes, 1024);   // MN-major V tile
      auto issue_qk = [&](int qt, int st) {
        const uint32_t a0 = desc_lo(kDescK) + ((q_addr + qt * 2 * kBoxBytes) >> 4);
        const uint32_t b0 = desc_lo(kDescK) + ((k_addr + st * 2 * kBoxBytes) >> 4);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const uint32_t off = (k >> 2) * (kBoxBytes >> 4) + 2 * (k & 3);
          umma_ss_elect(tmem_base + qt * 128, a0 + off, desc_hi(kDescK), b0 + off, desc_hi(kDescK), idesc_qk, k != 0 ? 1u : 0u);
        }
      };
      // O_q += P_q . V(stage) : 8 k-steps of 16 kv rows; V is MN-major (hd contiguous), the two
      // 64-wide hd halves are 16 KB apart (LBO), 8-row kv groups 1 KB apart (SBO).
      auto issue_pv = [&](int qt, int st, int half, bool first) {
        const uint32_t b0 = desc_lo(kDescV) + ((v_addr + st * 2 * kBoxBytes) >> 4);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const int k = half * 4 + kk;
          umma_ts_elect(tmem_base + 256 + qt * 128, tmem_base + qt * 128 + 8 * k, b0 + k * (2048 >> 4), desc_hi(kDescV),
                        idesc_pv, (first && k == 0) ? 0u : 1u);
        }
      };

#if MMPL_ATTN_TIMING
      long long dbg[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
      const long long tm_begin = clock64();
#endif
#if MMPL_ATTN_SPLIT_S
      // Half-tile pipeline. S_q of a KV tile lives in two independent 64-column halves (KV rows 0-63 / 64-127), each
      // with its own S-ready and P-ready barrier; bf16 P of a half overwrites the first 32 columns of that half.
      // Per query tile and half the issue order is  P_h(j).V_h(j)  then  Q.K_h(j+1)^T : the next S half is on its way
      // as soon as this P half has been consumed, while the softmax warps are still busy with the other half, so
      // neither the softmax warps nor the tensor pipe wait for a whole-tile round trip (the whole-tile variant measured
      // 1190 of 2700 clk per tile with the softmax warps waiting for S). The commit on s_full[q][h] that follows
      // P.V_h(j) and Q.K_h(j+1) also tells the softmax warps that this P.V has completed, which they need to know
      // before they rescale O; on the last tile of a piece, where no Q.K follows, pv_last[q] does that for half 0.
      constexpr uint32_t idesc_qk_h = make_idesc_bf16(128, 64, 0, 0);
      auto issue_qk_half = [&](int qt, int st, int h) {
        const uint32_t a0 = desc_lo(kDescK) + ((q_addr + qt * 2 * kBoxBytes) >> 4);
        const uint32_t b0 = desc_lo(kDescK) + ((k_addr + st * 2 * kBoxBytes + h * 64 * 128) >> 4);  // KV rows h*64 ..
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const uint32_t off = (k >> 2) * (kBoxBytes >> 4) + 2 * (k & 3);
          umma_ss_elect(tmem_base + qt * 128 + h * 64, a0 + off, desc_hi(kDescK), b0 + off, desc_hi(kDescK), idesc_qk_h,
                        k != 0 ? 1u : 0u);
        }
      };
      auto issue_pv_half = [&](int qt, int st, int h, bool first) {
        const uint32_t b0 = desc_lo(kDescV) + ((v_addr + st * 2 * kBoxBytes) >> 4);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const int k = h * 4 + kk;
          umma_ts_elect(tmem_base + 256 + qt * 128, tmem_base + qt * 128 + h * 64 + 8 * kk, b0 + k * (2048 >> 4),
                        desc_hi(kDescV), idesc_pv, (first && kk == 0) ? 0u : 1u);
        }
      };
      int g = 0, piece = 0;
      PieceIter pit;
      pit.init(p, blockIdx.x, G);
      Piece pc;
      for (; pit.next(p, pc); ++piece) {
        const int n = pc.n;
        mbar_wait(q_full, piece & 1);
        mbar_wait(&k_full[g & 1], (g >> 1) & 1);
        tc_fence_after();
        for (int qt = 0; qt < 2; ++qt)
          for (int h = 0; h < 2; ++h) {
            issue_qk_half(qt, g & 1, h);
            tc_commit_elect(&s_full[2 * qt + h]);
          }
        tc_commit_elect(&k_empty[g & 1]);
        if (n == 1) tc_commit_elect(q_empty);
        for (int jj = 0; jj < n; ++jj) {
          const int gj = g + jj;
          const int st = gj & 1;
          const uint32_t ph = (gj >> 1) & 1;
          const int stn = (gj + 1) & 1;
          const uint32_t phn = ((gj + 1) >> 1) & 1;
          const bool more = jj + 1 < n;
          mbar_wait(&v_full[st], ph);
          for (int qt = 0; qt < 2; ++qt) {
            for (int h = 0; h < 2; ++h) {
              mbar_wait(&p_full[2 * qt + h], gj & 1);
              if (jj == 0 && h == 0) mbar_wait(&o_empty[qt], (piece & 1) ^ 1);  // previous piece's O has been read out
              if (qt == 0 && h == 0 && more) mbar_wait(&k_full[stn], phn);
              tc_fence_after();
              issue_pv_half(qt, st, h, jj == 0 && h == 0);
              if (more) {
                issue_qk_half(qt, stn, h);
                tc_commit_elect(&s_full[2 * qt + h]);
              } else if (h == 0) {
                tc_commit_elect(&pv_last[qt]);
              }
            }
          }
          tc_commit_elect(&v_empty[st]);
          if (more) {
            tc_commit_elect(&k_empty[stn]);
            if (jj + 2 == n) tc_commit_elect(q_empty);  // last QK of the piece issued: Q smem may be reloaded
          }
        }
        tc_commit_elect(o_full);
        g += n;
      }
#else
      int g = 0, piece = 0;
      PieceIter pit;
      pit.init(p, blockIdx.x, G);
      Piece pc;
      for (; pit.next(p, pc); ++piece) {
        const int n = pc.n;
        mbar_wait(q_full, piece & 1);
        mbar_wait(&k_full[g & 1], (g >> 1) & 1);
        tc_fence_after();
        issue_qk(0, g & 1);
        tc_commit_elect(&s_full[0]);
        issue_qk(1, g & 1);
        tc_commit_elect(&s_full[1]);
        tc_commit_elect(&k_empty[g & 1]);
        if (n == 1) tc_commit_elect(q_empty);
        for (int jj = 0; jj < n; ++jj) {
          const int gj = g + jj;
          const int st = gj & 1;
          const uint32_t ph = (gj >> 1) & 1;
          const int stn = (gj + 1) & 1;
          const uint32_t phn = ((gj + 1) >> 1) & 1;
          TSTAMP(tv0);
          mbar_wait(&v_full[st], ph);
          TSTAMP(tv1);
          TACC(10, tv1 - tv0); TACC(13, 1);
          for (int qt = 0; qt < 2; ++qt) {
            // P arrives in two halves (KV rows 0-63, 64-127) so that P.V starts while the second half is still
            // being exponentiated
#if MMPL_ATTN_HALVES
            mbar_wait(&p_full[2 * qt], gj & 1);
            if (jj == 0) mbar_wait(&o_empty[qt], (piece & 1) ^ 1);  // previous piece's O has been read out
            tc_fence_after();
            issue_pv(qt, st, 0, jj == 0);
            mbar_wait(&p_full[2 * qt + 1], gj & 1);
#else
            TSTAMP(tp0);
            mbar_wait(&p_full[2 * qt + 1], gj & 1);
            TSTAMP(tp1);
            TACC(8 + qt, tp1 - tp0);
            if (jj == 0) mbar_wait(&o_empty[qt], (piece & 1) ^ 1);
            tc_fence_after();
            issue_pv(qt, st, 0, jj == 0);
#endif
            TSTAMP(tk0);
            if (qt == 0 && jj + 1 < n) mbar_wait(&k_full[stn], phn);
            TSTAMP(tk1);
            TACC(11, tk1 - tk0);
            tc_fence_after();
            issue_pv(qt, st, 1, false);
            if (qt == 1) tc_commit_elect(&v_empty[st]);
            if (jj + 1 < n) {
              issue_qk(qt, stn);
              tc_commit_elect(&s_full[qt]);
              if (qt == 1) {
                tc_commit_elect(&k_empty[stn]);
                if (jj + 2 == n) tc_commit_elect(q_empty);  // last QK of the piece issued: Q smem may be reloaded
              }
            }
          }
        }
        tc_commit_elect(o_full);
        g += n;
      }
#endif  // MMPL_ATTN_SPLIT_S
#if MMPL_ATTN_TIMING
      dbg[12] = clock64() - tm_begin;
      if (blockIdx.x == 0 && lane == 0)
        for (int i = 8; i < 14; ++i) atomicAdd(reinterpret_cast<unsigned long long*>(&g_attn_dbg[i]), (unsigned long long)dbg[i]);
#endif
    }
  } else {
    // ------------------------------------------------------------- softmax
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(MMPL_ATTN_SOFTMAX_REGS));
    const int qt = (warp - kFirstSoftmaxWarp) >> 2;
    const int lane_base = (warp & 3) * 32;
    const uint32_t t_lane = static_cast<uint32_t>(lane_base) << 16;
    const uint32_t t_s = tmem_base + t_lane + qt * 128;
    const uint32_t t_o = tmem_base + t_lane + 256 + qt * 128;
    const int row_in_unit = qt * kQTile + lane_base + lane;
#if MMPL_ATTN_TIMING
    long long dbg[4] = {0, 0, 0, 0};
#endif
    int g = 0, piece = 0;
    int sbase = 0;  // half-tile pipeline: phases of s_full[qt][*] consumed by earlier pieces
    PieceIter pit;
    pit.init(p, blockIdx.x, G);
    Piece pc;
    for (; pit.next(p, pc); ++piece) {
      const int t0 = pc.t0, n = pc.n, head = pc.head;
      const int q_row = pc.q_row0 + row_in_unit;
#if MMPL_ATTN_SPLIT_S
      float m_run = -INFINITY;  // running max, in the scaled log2 domain
      float l_run = 0.f;
      TileIter it;
      it.init(p, t0);
      for (int jj = 0; jj < n; ++jj, it.next(p)) {
        const int valid = it.valid();
#pragma unroll 1
        for (int h = 0; h < 2; ++h) {
          const int valid_h = valid - h * 64;  // valid KV rows in this half (<= 0: none; >= 64: all)
          TSTAMP(ts0);
          mbar_wait(&s_full[2 * qt + h], (sbase + jj) & 1);
          tc_fence_after();
          TSTAMP(ts1);
          uint32_t sv[64];
          tmem_ld_32x32(t_s + h * 64, *reinterpret_cast<uint32_t(*)[32]>(&sv[0]));
          tmem_ld_32x32(t_s + h * 64 + 32, *reinterpret_cast<uint32_t(*)[32]>(&sv[32]));
          tmem_ld_wait();
          if (valid_h < 64) {
#pragma unroll
            for (int c = 0; c < 64; ++c)
              if (c >= valid_h) sv[c] = 0xFF800000u;  // -inf
          }
          float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};  // four independent chains of 3-input max
#pragma unroll
          for (int c = 0; c < 32; ++c)
            mx4[c & 3] = fmaxf(fmaxf(mx4[c & 3], __uint_as_float(sv[2 * c])), __uint_as_float(sv[2 * c + 1]));
          const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
          const float m_new = fmaxf(m_run, mx * p.scale_log2);
          const bool grow = m_new > m_run + 8.0f;
          if (__any_sync(0xffffffffu, grow)) {
            const float alpha = fast_exp2(m_run - m_new);
            l_run *= alpha;
            m_run = m_new;
            if (jj > 0 || h > 0) {
              // O holds the P.V of every half before this one; the last of them may still be executing (or not even
              // issued). The commit that follows it is the next phase of the *other* half's S barrier: after P.V_b(jj-1)
              // comes S_b(jj) (h == 0), after P.V_a(jj) comes S_a(jj+1) (h == 1) or, on the last tile, pv_last. No
              // later P.V can start before this warp group has delivered the P of this half.
              if (h == 1 && jj + 1 == n) mbar_wait(&pv_last[qt], piece & 1);
              else mbar_wait(&s_full[2 * qt + (h ^ 1)], (sbase + jj + h) & 1);
              tc_fence_after();
#pragma unroll 1
              for (int c = 0; c < 4; ++c) {
                uint32_t o[32];
                tmem_ld_32x32(t_o + c * 32, o);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
                tmem_st_32x32(t_o + c * 32, o);
              }
            }
          }
          uint64_t sum2 = pack_f32x2(0.f, 0.f);
          const uint64_t sc2 = pack_f32x2(p.scale_log2, p.scale_log2);
          const uint64_t nm2 = pack_f32x2(-m_run, -m_run);
          uint32_t pk[32];
#pragma unroll
          for (int c = 0; c < 32; ++c) {
            const uint64_t x2 = fma_f32x2(pack_f32x2(__uint_as_float(sv[2 * c]), __uint_as_float(sv[2 * c + 1])), sc2, nm2);
            float p0, p1;
            if (pair_is_poly(c)) {
              exp2_poly_x2(x2, p0, p1);
            } else {
              float x0, x1;
              unpack_f32x2(x2, x0, x1);
              p0 = fast_exp2(x0);
              p1 = fast_exp2(x1);
            }
            sum2 = add_f32x2(sum2, pack_f32x2(p0, p1));
            pk[c] = pack_bf16x2(p0, p1);
          }
          {
            float s_lo, s_hi;
            unpack_f32x2(sum2, s_lo, s_hi);
            l_run += s_lo + s_hi;
          }
          TSTAMP(ts2);
          tmem_st_32x32(t_s + h * 64, pk);
          tmem_st_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&p_full[2 * qt + h]);
          TSTAMP(ts3);
          TACC(0, ts1 - ts0); TACC(1, ts2 - ts1); TACC(2, ts3 - ts2); TACC(3, 1);
        }
      }
      sbase += n;
#else
      float m_run = -INFINITY;  // running max, in the scaled log2 domain
      float l_run = 0.f;
      TileIter it;
      it.init(p, t0);
      for (int jj = 0; jj < n; ++jj, it.next(p)) {
        const int valid = it.valid();
        TSTAMP(ts0);
        mbar_wait(&s_full[qt], (g + jj) & 1);
        tc_fence_after();
        TSTAMP(ts1);
        uint32_t sv[128];
#pragma unroll
        for (int c = 0; c < 4; ++c) tmem_ld_32x32(t_s + c * 32, *reinterpret_cast<uint32_t(*)[32]>(&sv[c * 32]));
        tmem_ld_wait();
        if (valid < kKVTile) {
#pragma unroll
          for (int c = 0; c < 128; ++c)
            if (c >= valid) sv[c] = 0xFF800000u;  // -inf
        }
        float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};  // four independent chains of 3-input max
#pragma unroll
        for (int c = 0; c < 64; ++c)
          mx4[c & 3] = fmaxf(fmaxf(mx4[c & 3], __uint_as_float(sv[2 * c])), __uint_as_float(sv[2 * c + 1]));
        const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
        const float m_new = fmaxf(m_run, mx * p.scale_log2);
        const bool grow = m_new > m_run + 8.0f;
        if (__any_sync(0xffffffffu, grow)) {
          const float alpha = fast_exp2(m_run - m_new);
          l_run *= alpha;
          m_run = m_new;
          if (jj > 0) {
#pragma unroll 1
            for (int c = 0; c < 4; ++c) {
              uint32_t o[32];
              tmem_ld_32x32(t_o + c * 32, o);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
              tmem_st_32x32(t_o + c * 32, o);
            }
          }
        }
        // P = 2^(S*scale - m): packed fp32x2 FMA for the scaling and the row sum; kPolyNum of every kPolyDen pairs
        // are exponentiated on the FMA/ALU pipes (exp2_poly_x2) instead of the MUFU, which is otherwise the pipe
        // that paces the softmax (16 ex2/clk/SM against 128 rows x 128 columns per tile).
        uint64_t sum2 = pack_f32x2(0.f, 0.f);
        const uint64_t sc2 = pack_f32x2(p.scale_log2, p.scale_log2);
        const uint64_t nm2 = pack_f32x2(-m_run, -m_run);
        uint32_t pk[64];
#pragma unroll
        for (int c = 0; c < 64; ++c) {
          const uint64_t x2 = fma_f32x2(pack_f32x2(__uint_as_float(sv[2 * c]), __uint_as_float(sv[2 * c + 1])), sc2, nm2);
          float p0, p1;
          if (pair_is_poly(c)) {
            exp2_poly_x2(x2, p0, p1);
          } else {
            float x0, x1;
            unpack_f32x2(x2, x0, x1);
            p0 = fast_exp2(x0);
            p1 = fast_exp2(x1);
          }
          sum2 = add_f32x2(sum2, pack_f32x2(p0, p1));
          pk[c] = pack_bf16x2(p0, p1);
        }
        float sum;
        {
          float s_lo, s_hi;
          unpack_f32x2(sum2, s_lo, s_hi);
          sum = s_lo + s_hi;
        }
        TSTAMP(ts2);
        tmem_st_32x32(t_s, *reinterpret_cast<uint32_t(*)[32]>(&pk[0]));
#if MMPL_ATTN_HALVES
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[2 * qt]);
#endif
        tmem_st_32x32(t_s + 32, *reinterpret_cast<uint32_t(*)[32]>(&pk[32]));
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[2 * qt + 1]);
        l_run += sum;
        TSTAMP(ts3);
        TACC(0, ts1 - ts0); TACC(1, ts2 - ts1); TACC(2, ts3 - ts2); TACC(3, 1);
      }
#endif  // MMPL_ATTN_SPLIT_S
      // piece epilogue
      mbar_wait(o_full, piece & 1);
      tc_fence_after();
      if (pc.whole) {
        // whole unit: O / l -> bf16 -> out[q_row, head*128 : head*128+128]
        const float inv_l = 1.0f / l_run;
        __nv_bfloat16* orow = p.out + static_cast<int64_t>(q_row) * p.ldo + head * kHD;
        const bool wide_out = (p.ldo % 16 == 0) && ((reinterpret_cast<uintptr_t>(p.out) & 31) == 0);
        // the TMEM loads are issued two at a time before a wait (the S registers are dead here): two TMEM round trips
        // instead of four on the critical path between two pieces
#pragma unroll 1
        for (int c2 = 0; c2 < 2; ++c2) {
          uint32_t o[64];
          tmem_ld_32x32(t_o + c2 * 64, *reinterpret_cast<uint32_t(*)[32]>(&o[0]));
          tmem_ld_32x32(t_o + c2 * 64 + 32, *reinterpret_cast<uint32_t(*)[32]>(&o[32]));
          tmem_ld_wait();
          if (q_row < p.Lq) {
#pragma unroll
            for (int gq = 0; gq < 4; ++gq) {
              uint32_t w[8];
#pragma unroll
              for (int i = 0; i < 8; ++i)
                w[i] = pack_bf16x2(__uint_as_float(o[gq * 16 + 2 * i]) * inv_l, __uint_as_float(o[gq * 16 + 2 * i + 1]) * inv_l);
              if (wide_out) {
                st_global_v8(orow + c2 * 64 + gq * 16, w);
              } else {
                *reinterpret_cast<uint4*>(orow + c2 * 64 + gq * 16) = make_uint4(w[0], w[1], w[2], w[3]);
                *reinterpret_cast<uint4*>(orow + c2 * 64 + gq * 16 + 8) = make_uint4(w[4], w[5], w[6], w[7]);
              }
            }
          }
        }
      } else {
        // partial unit: un-normalised fp32 O and (m, l) to the workspace slot of this piece
        const int64_t base = static_cast<int64_t>(pc.slot) * kUnitRows + row_in_unit;
        float* po = p.part_o + base * kHD;
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          uint32_t o[32];
          tmem_ld_32x32(t_o + c * 32, o);
          tmem_ld_wait();
#pragma unroll
          for (int gq = 0; gq < 4; ++gq) {
            const uint32_t w[8] = {o[gq * 8], o[gq * 8 + 1], o[gq * 8 + 2], o[gq * 8 + 3],
                                   o[gq * 8 + 4], o[gq * 8 + 5], o[gq * 8 + 6], o[gq * 8 + 7]};
            st_global_v8(po + c * 32 + gq * 8, w);  // workspace rows are 512 B, cudaMalloc-aligned
          }
        }
        *reinterpret_cast<float2*>(p.part_ml + base * 2) = make_float2(m_run, l_run);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&o_empty[qt]);
      if (!pc.whole && p.merge_cnt != nullptr) {
        // Merge inside the kernel: the piece of a unit that finishes last adds up all of them (its own included, read
        // back from L2 like the others) and writes the bf16 output, so the partials never make the round trip through
        // HBM that a separate merge kernel pays (ncu: 88 MB read back, 35 us per launch at L_kv >= 18720).
        // Per query tile: this warpgroup's 128 rows are fenced and counted independently of the other one.
        const int u = head * p.QP + pc.q_row0 / kUnitRows;
        UnitPieces up;
        up.init(p, u);
        __threadfence();
        auto wg_sync = [&]() {  // the four softmax warps of this query tile
          if (qt == 0) asm volatile("bar.sync 1, 128;" ::: "memory");
          else asm volatile("bar.sync 2, 128;" ::: "memory");
        };
        wg_sync();
        if ((warp & 3) == 0 && lane == 0) merge_ticket[qt] = atomicAdd(p.merge_cnt + 2 * u + qt, 1);
        wg_sync();
        if (merge_ticket[qt] == up.count(p) - 1) {
          __threadfence();
          if ((warp & 3) == 0 && lane == 0) p.merge_cnt[2 * u + qt] = 0;  // re-armed for the next launch
          // Lane = row for the (max, sum) statistics; the O rows are then read one row per warp instruction (512 B,
          // coalesced; lane = 4 columns), 32 rows in flight per piece, with the row's weight broadcast by shuffle.
          float M = -INFINITY;
          for (int i = 0; i < up.np; ++i) {
            const int sl = up.slot(p, i);
            if (sl >= 0) M = fmaxf(M, __ldcg(p.part_ml + (static_cast<int64_t>(sl) * kUnitRows + row_in_unit) * 2));
          }
          float L = 0.f;
          for (int i = 0; i < up.np; ++i) {
            const int sl = up.slot(p, i);
            if (sl < 0) continue;
            const float2 ml = __ldcg(reinterpret_cast<const float2*>(p.part_ml + (static_cast<int64_t>(sl) * kUnitRows + row_in_unit) * 2));
            L += exp2f(ml.x - M) * ml.y;
          }
          const float inv = 1.0f / L;
          const int warp_row0 = qt * kQTile + lane_base;  // first row (within the unit) of this warp
          const int rows_ok = p.Lq - (pc.q_row0 + warp_row0);
#pragma unroll 1
          for (int hr = 0; hr < 32; hr += 16) {  // 16 rows at a time: 64 accumulator + 64 load registers
            float4 acc[16];
#pragma unroll
            for (int r = 0; r < 16; ++r) acc[r] = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int i = 0; i < up.np; ++i) {
              const int sl = up.slot(p, i);
              if (sl < 0) continue;
              const int64_t r0 = static_cast<int64_t>(sl) * kUnitRows + warp_row0;
              const float w_own = exp2f(__ldcg(p.part_ml + (r0 + lane) * 2) - M) * inv;
              const float4* src = reinterpret_cast<const float4*>(p.part_o + (r0 + hr) * kHD) + lane;
              float4 v[16];
#pragma unroll
              for (int r = 0; r < 16; ++r) v[r] = __ldcg(src + r * (kHD / 4));
#pragma unroll
              for (int r = 0; r < 16; ++r) {
                const float w = __shfl_sync(0xffffffffu, w_own, hr + r);
                acc[r].x += w * v[r].x; acc[r].y += w * v[r].y; acc[r].z += w * v[r].z; acc[r].w += w * v[r].w;
              }
            }
            __nv_bfloat16* obase = p.out + static_cast<int64_t>(pc.q_row0 + warp_row0 + hr) * p.ldo + head * kHD + lane * 4;
#pragma unroll
            for (int r = 0; r < 16; ++r)
              if (hr + r < rows_ok)
                *reinterpret_cast<uint2*>(obase + static_cast<int64_t>(r) * p.ldo) =
                    make_uint2(pack_bf16x2(acc[r].x, acc[r].y), pack_bf16x2(acc[r].z, acc[r].w));
          }
        }
      }
      g += n;
    }
#if MMPL_ATTN_TIMING
    if (blockIdx.x == 0 && lane == 0 && (warp & 3) == 0)
      for (int i = 0; i < 4; ++i) atomicAdd(reinterpret_cast<unsigned long long*>(&g_attn_dbg[qt * 4 + i]), (unsigned long long)dbg[i]);
#endif
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// Merges the pieces of every unit that was split across CTAs:
//   out = sum_i 2^(m_i - M) O_i / sum_i 2^(m_i - M) l_i,  M = max_i m_i.   One warp per query row.
// The pieces of unit u: uniform split -> slots u*split + i; range schedule -> one per CTA whose range of the unit's
// head group overlaps the unit (same arithmetic as PieceIter), none if a single CTA ran the whole unit.
__global__ void __launch_bounds__(256)
attn_combine_kernel(const AttnParams p) {
  pdl_wait();
  pdl_launch_dependents();
  const int u = blockIdx.x >> 5;
  const int row_in_unit = (blockIdx.x & 31) * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  const int head = u / p.QP;
  const int q_row = (u - head * p.QP) * kUnitRows + row_in_unit;
  if (u < p.u_base) return;          // hybrid schedule: ran whole
  UnitPieces up;
  up.init(p, u);
  const int np = up.np;
  if (p.hg != 0 && np == 1) return;  // written directly by the CTA that ran the whole unit
  if (q_row >= p.Lq) return;
  auto slot_of = [&](int i) -> int { return up.slot(p, i); };
  // pieces are taken four at a time with all their loads issued before the first use: the kernel is a single pass over
  // ~90 MB of partials that have mostly left L2, i.e. bound by how many loads are in flight
  float M = -INFINITY;
  for (int i = 0; i < np; ++i) {
    const int sl = slot_of(i);
    if (sl >= 0) M = fmaxf(M, p.part_ml[(static_cast<int64_t>(sl) * kUnitRows + row_in_unit) * 2]);
  }
  float L = 0.f;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int i0 = 0; i0 < np; i0 += 4) {
    float2 ml[4];
    float4 o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int sl = i0 + k < np ? slot_of(i0 + k) : -1;
      ml[k] = make_float2(-INFINITY, 0.f);
      o[k] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (sl >= 0) {
        const int64_t r = static_cast<int64_t>(sl) * kUnitRows + row_in_unit;
        ml[k] = *reinterpret_cast<const float2*>(p.part_ml + r * 2);
        o[k] = *reinterpret_cast<const float4*>(p.part_o + r * kHD + lane * 4);
      }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float w = exp2f(ml[k].x - M);  // 0 for the slots that are not there
      L += w * ml[k].y;
      acc.x += w * o[k].x; acc.y += w * o[k].y; acc.z += w * o[k].z; acc.w += w * o[k].w;
    }
  }
  const float inv = 1.0f / L;
  uint2 w2 = make_uint2(pack_bf16x2(acc.x * inv, acc.y * inv), pack_bf16x2(acc.z * inv, acc.w * inv));
  *reinterpret_cast<uint2*>(p.out + static_cast<int64_t>(q_row) * p.ldo + head * kHD + lane * 4) = w2;
}

// Work partition of one launch (host): fills the schedule fields of `p` (QP, T, split, n_pieces, hg, u_base) for Lq, H
// already set, and returns the grid size G and the number of workspace slots for partial pieces. `sms` = CTAs of the
// persistent grid (the SM count; smaller in tests). Pure host arithmetic: also reachable without a GPU through
// attn_plan_pieces() / mmpl_attn_plan(), which the CPU test-suite uses to check that every (unit, KV tile) pair is
// covered exactly once and that the merge bookkeeping (UnitPieces) agrees with the pieces the CTAs run.
static void plan_schedule(AttnParams& p, int T, int sms, int force_split, int& G, size_t& slots) {
  const int Lq = p.Lq, H = p.H;
  // work partition: U units of T KV tiles over G persistent CTAs, by one of the two schedules of PieceIter.
  // Cost model (in KV-tile times per SM; measured on B200 with tools/bench_kernels.py): a piece costs ~6 tile times on
  // top of its tiles (Q load, pipeline fill, epilogue); merging costs 2 x 128 KB of traffic per partial piece ~ 0.029
  // tile times each.
  p.QP = (Lq + kUnitRows - 1) / kUnitRows;
  p.T = T;
  const int U = p.QP * H;
  const double piece_fixed = 6.0, merge_per_piece = 0.0291, merge_fixed = 6.0;  // merge_fixed: the combine launch itself
  int best_split = 1;
  double best = 1e30;
  for (int sp = 1; sp <= 8 && sp <= T; ++sp) {
    if (sp > 1 && T / sp < 8) break;
    const int rounds = (U * sp + sms - 1) / sms;
    const double cost = rounds * (static_cast<double>(T) / sp + piece_fixed) + (sp > 1 ? merge_fixed + merge_per_piece * U * sp : 0.0);
    if (cost < best - 1e-9) { best = cost; best_split = sp; }
  }
  // Range schedule. Its CTAs sit at different KV positions of the same head, so (unlike the uniform schedule, whose
  // CTAs stream the same K/V tiles in lockstep and are served by L2 together) it pays only while the K/V of all heads
  // stays L2-resident: measured on B200 (profiles/README.md) 6 % faster than the best uniform split for L_kv = 9360
  // and 14040 at cfg2 (58 / 86 MB of K/V), slower from 115 MB up, erratic with several head groups. Hence: one group
  // of all heads, only below the L2 budget, and only when the cost model prefers it.
  int hg = 0, u_base = 0;
  {
    static const int l2_mb = getenv("MMPL_ATTN_L2_MB") ? atoi(getenv("MMPL_ATTN_L2_MB")) : 90;
    static const int mode_env = getenv("MMPL_ATTN_RANGES") ? atoi(getenv("MMPL_ATTN_RANGES")) : -1;  // 0 never, 1 whenever possible, -1 cost model
    static const int hybrid_env = getenv("MMPL_ATTN_HYBRID") ? atoi(getenv("MMPL_ATTN_HYBRID")) : -1;  // same
    const double head_mb = static_cast<double>(T) * kKVTile * 512.0 / (1 << 20);
    const double kv_mb = H * head_mb;
    const double pieces_per_cta = static_cast<double>(U) / sms + 1.0;
    const double cost_ranges = static_cast<double>(U) * T / sms + piece_fixed * pieces_per_cta + merge_fixed + merge_per_piece * 2.0 * sms;
    const bool fits32 = static_cast<long long>(U) * T * (sms + 1) < (1ll << 31);  // PieceIter's range arithmetic is 32-bit
    const bool possible = T >= 16 && static_cast<long long>(U) * T >= 8ll * sms && kv_mb <= l2_mb && fits32;
    if (possible && (mode_env == 1 || (mode_env < 0 && cost_ranges < best))) { hg = H; best = cost_ranges; }
    // Hybrid schedule: floor(U/G) rounds of whole units in lockstep, the remaining U mod G units as ranges (merged in
    // the kernel). Same tile count per CTA as the pure range schedule, but only the heads of the ranged units are read
    // by free-running CTAs, so it stays inside L2 where the pure range schedule does not (cfg2 at L_kv = 32760: 5 of
    // 12 heads, 84 MB), and it has no merge kernel, which the uniform split pays for (37 us per launch there).
    const int rounds = U / sms, rem = U - rounds * sms;
    if (rounds >= 1 && rem > 0) {
      const int heads2 = H - (rounds * sms) / p.QP;  // heads the ranged units belong to
      const double cost_hybrid = static_cast<double>(U) * T / sms + piece_fixed * (rounds + 2.0);
      const bool hybrid_possible = T >= 16 && static_cast<long long>(rem) * T >= 6ll * sms && heads2 * head_mb <= l2_mb && fits32;
      if (hybrid_possible && (hybrid_env == 1 || (hybrid_env < 0 && cost_hybrid < best))) {
        hg = H;
        u_base = rounds * sms;
        best = cost_hybrid;
      }
    }
  }
  if (force_split > 0 && force_split <= T) {
    best_split = force_split;
    hg = 0;
    u_base = 0;
  } else if (force_split <= -1000) {
    // test hook: hybrid schedule whenever there is at least one whole round and something left over
    const int rounds = U / sms, rem = U - rounds * sms;
    hg = 0;
    u_base = 0;
    if (rounds >= 1 && rem > 0 && static_cast<long long>(rem) * T >= 2ll * sms &&
        static_cast<long long>(U) * T * (sms + 1) < (1ll << 31)) {
      hg = H;
      u_base = rounds * sms;
    }
  } else if (force_split < 0 && T >= 2 && static_cast<long long>(U) * T >= 2ll * sms &&
             static_cast<long long>(U) * T * (sms + 1) < (1ll << 31)) {
    hg = -force_split < H ? -force_split : H;  // test hook: range schedule with this many heads per group
    u_base = 0;
  }
  static const bool nonpersistent = getenv("MMPL_ATTN_NONPERSISTENT") != nullptr;
  slots = 0;
  if (hg > 0) {
    p.hg = hg;
    p.u_base = u_base;
    p.split = 1;
    p.n_pieces = 0;
    G = sms;
    slots = static_cast<size_t>((H + hg - 1) / hg) * G * 2;
  } else {
    p.hg = 0;
    p.split = best_split;
    p.n_pieces = U * p.split;
    G = (p.n_pieces < sms || nonpersistent) ? p.n_pieces : sms;
    slots = p.split > 1 ? static_cast<size_t>(U) * p.split : 0;
  }
}

// Workspace for partial pieces (grown on demand; one per process, used by launches on one stream at a time) and the
// arrival counters of the in-kernel merge (zero between launches: the last piece of a unit clears its counter).
static float* g_part = nullptr;
static size_t g_part_bytes = 0;
static int* g_merge_cnt = nullptr;
static size_t g_merge_cnt_n = 0;
int flash_attn_impl(const void* q, int64_t ldq, int Lq, int H, const void* k0, const void* v0,
                    int64_t ldkv0, int rows0, const void* k1, const void* v1, int64_t ldkv1, int rows1,
                    int nseg, const int* seg_start, const int* seg_rows, const int* seg_src, void* out,
                    int64_t ldo, float softmax_scale, int force_split, int force_ctas, cudaStream_t stream) {
  MMPL_CHECK(device_is_sm100(), MMPL_ERR_ARCH, "flash_attn: requires an sm_100 device");
  MMPL_CHECK(Lq > 0 && H > 0, MMPL_ERR_SHAPE, "flash_attn: bad Lq=%d H=%d", Lq, H);
  MMPL_CHECK(nseg >= 1 && nseg <= kMaxSeg, MMPL_ERR_SHAPE, "flash_attn: nseg=%d out of [1,%d]", nseg, kMaxSeg);
  MMPL_CHECK(ldo % 8 == 0, MMPL_ERR_SHAPE, "flash_attn: ldo must be a multiple of 8");
  AttnParams p{};
  p.Lq = Lq;
  p.H = H;
  p.out = static_cast<__nv_bfloat16*>(out);
  p.ldo = ldo;
  p.scale_log2 = softmax_scale * 1.4426950408889634f;
  p.nseg = nseg;
  bool uses1 = false;
  int T = 0;
  for (int i = 0; i < nseg; ++i) {
    const int src = seg_src ? seg_src[i] : 0;
    const int lim = src ? rows1 : rows0;
    MMPL_CHECK(seg_rows[i] > 0 && seg_start[i] >= 0 && seg_start[i] + seg_rows[i] <= lim, MMPL_ERR_SHAPE,
               "flash_attn: segment %d [%d,+%d) outside its %d-row source", i, seg_start[i], seg_rows[i], lim);
    p.seg_start[i] = seg_start[i];
    p.seg_rows[i] = seg_rows[i];
    p.seg_src[i] = src;
    uses1 |= (src != 0);
    T += (seg_rows[i] + kKVTile - 1) / kKVTile;
  }
  MMPL_CHECK(!uses1 || (k1 && v1), MMPL_ERR_ARG, "flash_attn: segment refers to a missing second K/V source");
  const CUtensorMap* mq = get_tensor_map_bf16(q, Lq, static_cast<uint64_t>(H) * kHD, ldq, 128);
  const CUtensorMap* mk0 = get_tensor_map_bf16(k0, rows0, static_cast<uint64_t>(H) * kHD, ldkv0, 128);
  const CUtensorMap* mv0 = get_tensor_map_bf16(v0, rows0, static_cast<uint64_t>(H) * kHD, ldkv0, 128);
  const CUtensorMap* mk1 = uses1 ? get_tensor_map_bf16(k1, rows1, static_cast<uint64_t>(H) * kHD, ldkv1, 128) : mk0;
  const CUtensorMap* mv1 = uses1 ? get_tensor_map_bf16(v1, rows1, static_cast<uint64_t>(H) * kHD, ldkv1, 128) : mv0;
  if (!mq || !mk0 || !mv0 || !mk1 || !mv1) return MMPL_ERR_CUDA;

  int G = 0;
  size_t slots = 0;
  plan_schedule(p, T, (force_ctas > 0 && force_ctas < sm_count()) ? force_ctas : sm_count(), force_split, G, slots);
  p.G = G;
  const int U = p.QP * H;
  if (slots > 0) {
    const size_t need = slots * kUnitRows * (kHD + 2) * sizeof(float);
    if (need > g_part_bytes) {
      if (g_part) MMPL_CUDA(cudaFree(g_part));
      g_part = nullptr;
      g_part_bytes = 0;
      MMPL_CUDA(cudaMalloc(&g_part, need));
      g_part_bytes = need;
      bump_workspace_generation();
    }
    p.part_o = g_part;
    p.part_ml = g_part + slots * kUnitRows * kHD;
    // Who merges the partial pieces. Range schedule: the pieces of a unit run on neighbouring CTAs at the same time, the
    // last one to finish merges them inside the kernel while they are still in L2 (3.5 % faster than the merge kernel
    // at L_kv = 9360 / 14040). Uniform split: the chunks of a unit run rounds apart, the partials have left L2 by the
    // time the last chunk is done and the in-kernel merge stalls that CTA's softmax warps on HBM reads (10-25 % slower
    // at L_kv >= 18720): attn_combine_kernel. MMPL_ATTN_MERGE=inline|kernel forces one for both (tests, A/B).
    static const char* merge_env = getenv("MMPL_ATTN_MERGE");
    const bool merge_in_kernel = merge_env ? merge_env[0] == 'i' : p.hg > 0;
    if (merge_in_kernel) {
      const size_t need_cnt = static_cast<size_t>(U) * 2;
      if (need_cnt > g_merge_cnt_n) {
        if (g_merge_cnt) MMPL_CUDA(cudaFree(g_merge_cnt));
        g_merge_cnt = nullptr;
        g_merge_cnt_n = 0;
        MMPL_CUDA(cudaMalloc(&g_merge_cnt, need_cnt * sizeof(int)));
        MMPL_CUDA(cudaMemset(g_merge_cnt, 0, need_cnt * sizeof(int)));
        g_merge_cnt_n = need_cnt;
        bump_workspace_generation();
      }
      p.merge_cnt = g_merge_cnt;
    }
  }

  static bool attr_set = false;
  if (!attr_set) {
    MMPL_CUDA(cudaFuncSetAttribute(flash_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmem));
    attr_set = true;
  }
  MMPL_CUDA_LAUNCH(launch_kernel(flash_attn_kernel, G, kAttnThreads, kAttnSmem, stream, *mq, *mk0, *mv0, *mk1, *mv1, p));
  MMPL_CUDA(cudaGetLastError());
  if (slots > 0 && p.merge_cnt == nullptr) {
    MMPL_CUDA_LAUNCH(launch_kernel(attn_combine_kernel, U * 32, 256, 0, stream, p));
    MMPL_CUDA(cudaGetLastError());
  }
  return MMPL_OK;
}

#ifdef MMPL_ATTN_PROD
// Host-only: the pieces every CTA of a launch would run, in order (no GPU needed). rows of `pieces`:
// {cta, head, q_row0, t0, n, whole, slot, unit_pieces, merge_ok}: unit_pieces = how many pieces UnitPieces (the merge
// bookkeeping) says the piece's unit has, merge_ok = 1 if this piece's slot is among the slots UnitPieces enumerates
// for the unit (always 1 for whole pieces). Returns the number of pieces, or -1 if `max_pieces` is too small.
int attn_plan_pieces(int Lq, int H, int kv_tiles, int ctas, int force_split, int* sched, int* pieces, int max_pieces) {
  AttnParams p{};
  p.Lq = Lq;
  p.H = H;
  int G = 0;
  size_t slots = 0;
  plan_schedule(p, kv_tiles, ctas, force_split, G, slots);
  p.G = G;
  if (sched) {
    sched[0] = p.hg == 0 ? 0 : (p.u_base > 0 ? 2 : 1);  // 0 uniform split, 1 ranges, 2 hybrid
    sched[1] = p.split;
    sched[2] = p.hg;
    sched[3] = p.u_base;
    sched[4] = G;
    sched[5] = static_cast<int>(slots);
    sched[6] = p.QP;
  }
  int n = 0;
  for (int c = 0; c < G; ++c) {
    PieceIter it;
    it.init(p, c, G);
    Piece pc;
    while (it.next(p, pc)) {
      if (n >= max_pieces) return -1;
      int np = 1, ok = 1;
      if (!pc.whole) {
        const int u = pc.head * p.QP + pc.q_row0 / kUnitRows;
        UnitPieces up;
        up.init(p, u);
        np = up.count(p);
        ok = 0;
        for (int i = 0; i < up.np; ++i) ok |= (up.slot(p, i) == pc.slot) ? 1 : 0;
      }
      int* r = pieces + static_cast<size_t>(n) * 9;
      r[0] = c; r[1] = pc.head; r[2] = pc.q_row0; r[3] = pc.t0; r[4] = pc.n; r[5] = pc.whole ? 1 : 0; r[6] = pc.slot;
      r[7] = np; r[8] = ok;
      ++n;
    }
  }
  return n;
}
#endif

#if MMPL_ATTN_TIMING
extern "C" int mmpl_attn_debug_read(long long* out32, int reset) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out32, g_attn_dbg, sizeof(long long) * 32);
  if (reset) {
    long long z[32] = {0};
    cudaMemcpyToSymbol(g_attn_dbg, z, sizeof(z));
  }
  return 0;
}
#endif

}  // namespace MMPL_ATTN_NS
}  // namespace mmpl
