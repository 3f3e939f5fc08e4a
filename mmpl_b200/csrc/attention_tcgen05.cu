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
// One CTA = one head x two 128-row query tiles that ping-pong on the tensor pipe:
//   warp 0      TMA producer (Q once; K and V tiles through 2-stage rings)
//   warp 1      MMA issuer:  S_q = Q_q K^T (SS),  O_q += P_q V (A = P from TMEM, B = V MN-major)
//   warps 4-7   softmax for query tile 0 (thread = row), warps 8-11 for query tile 1
// TMEM (512 columns): S0 [0,128) S1 [128,256) O0 [256,384) O1 [384,512); bf16 P_q overwrites the
// first 64 columns of S_q. O is rescaled lazily (only when the running max grows by > 2^8), by the
// softmax warps themselves.
#include "host_util.h"
#include "mmpl_b200.h"
#include "ptx.cuh"

namespace mmpl {

constexpr int kAttnThreads = 384;
constexpr int kQTile = 128;
constexpr int kKVTile = 128;
constexpr int kHD = 128;
constexpr int kBoxBytes = 128 * 64 * 2;  // one [128 rows][64 bf16] swizzled box = 16 KB
constexpr int kKVStages = 2;
constexpr int kAttnSmem = 4 * kBoxBytes /*Q*/ + kKVStages * 2 * kBoxBytes /*K*/ +
                          kKVStages * 2 * kBoxBytes /*V*/ + 1024 + 256;
constexpr int kMaxSeg = 8;

struct AttnParams {
  int Lq;
  __nv_bfloat16* out;
  int64_t ldo;
  float scale_log2;
  int nseg;
  int seg_start[kMaxSeg];
  int seg_rows[kMaxSeg];
  int seg_src[kMaxSeg];
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
  uint64_t* k_full = bars + 1;             // 2
  uint64_t* k_empty = bars + 3;            // 2
  uint64_t* v_full = bars + 5;             // 2
  uint64_t* v_empty = bars + 7;            // 2
  uint64_t* s_full = bars + 9;             // 2 (per query tile)
  uint64_t* p_full = bars + 11;            // 2
  uint64_t* o_full = bars + 13;            // 1
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int head = blockIdx.y;
  const int q_row0 = blockIdx.x * (2 * kQTile);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_q);
    tma_prefetch_desc(&map_k0);
    tma_prefetch_desc(&map_v0);
  }
  if (warp == 1) {
    if (lane == 0) {
      mbar_init(q_full, 1);
      for (int i = 0; i < kKVStages; ++i) {
        mbar_init(&k_full[i], 1);
        mbar_init(&k_empty[i], 1);
        mbar_init(&v_full[i], 1);
        mbar_init(&v_empty[i], 1);
      }
      for (int i = 0; i < 2; ++i) {
        mbar_init(&s_full[i], 1);
        mbar_init(&p_full[i], 4);
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

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(56));
    if (warp == 0 && lane == 0) {
      // -------------------------------------------------------- TMA producer
      mbar_arrive_expect_tx(q_full, 4 * kBoxBytes);
      for (int qt = 0; qt < 2; ++qt)
        for (int hh = 0; hh < 2; ++hh)
          tma_load_2d(smem_q + (qt * 2 + hh) * kBoxBytes, &map_q, q_full, head * kHD + hh * 64,
                      q_row0 + qt * kQTile, kEvictFirst);
      int j = 0;
      for (int sg = 0; sg < p.nseg; ++sg) {
        const CUtensorMap* mk = p.seg_src[sg] ? &map_k1 : &map_k0;
        const CUtensorMap* mv = p.seg_src[sg] ? &map_v1 : &map_v0;
        const int nt = (p.seg_rows[sg] + kKVTile - 1) / kKVTile;
        for (int t = 0; t < nt; ++t, ++j) {
          const int st = j & 1;
          const uint32_t ph = (j >> 1) & 1;
          const int row = p.seg_start[sg] + t * kKVTile;
          mbar_wait(&k_empty[st], ph ^ 1);
          mbar_arrive_expect_tx(&k_full[st], 2 * kBoxBytes);
          for (int hh = 0; hh < 2; ++hh)
            tma_load_2d(smem_k + (st * 2 + hh) * kBoxBytes, mk, &k_full[st], head * kHD + hh * 64, row, kEvictLast);
          mbar_wait(&v_empty[st], ph ^ 1);
          mbar_arrive_expect_tx(&v_full[st], 2 * kBoxBytes);
          for (int hh = 0; hh < 2; ++hh)
            tma_load_2d(smem_v + (st * 2 + hh) * kBoxBytes, mv, &v_full[st], head * kHD + hh * 64, row, kEvictLast);
        }
      }
    } else if (warp == 1) {
      // ---------------------------------------------------------- MMA issuer
      int n_tiles = 0;
      for (int sg = 0; sg < p.nseg; ++sg) n_tiles += (p.seg_rows[sg] + kKVTile - 1) / kKVTile;
      constexpr uint32_t idesc_qk = make_idesc_bf16(128, 128, 0, 0);
      constexpr uint32_t idesc_pv = make_idesc_bf16(128, 128, 0, 1);
      const uint32_t q_addr = smem_u32(smem_q);
      const uint32_t k_addr = smem_u32(smem_k);
      const uint32_t v_addr = smem_u32(smem_v);

      // S_q = Q_q . K(stage)^T : 8 k-steps of 16 over head_dim (two 64-wide swizzled halves)
      auto issue_qk = [&](int qt, int st) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int hh = k >> 2;
          const uint64_t da = make_smem_desc_sw128(q_addr + (qt * 2 + hh) * kBoxBytes, 16, 1024) + 2 * (k & 3);
          const uint64_t db = make_smem_desc_sw128(k_addr + (st * 2 + hh) * kBoxBytes, 16, 1024) + 2 * (k & 3);
          umma_ss(tmem_base + qt * 128, da, db, idesc_qk, k != 0 ? 1u : 0u);
        }
      };
      // O_q += P_q . V(stage) : 8 k-steps of 16 kv rows; V is MN-major (hd contiguous), the two
      // 64-wide hd halves are 16 KB apart (LBO), 8-row kv groups 1 KB apart (SBO).
      auto issue_pv = [&](int qt, int st, bool first) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const uint64_t db = make_smem_desc_sw128(v_addr + st * 2 * kBoxBytes + k * 2048, kBoxBytes, 1024);
          umma_ts(tmem_base + 256 + qt * 128, tmem_base + qt * 128 + 8 * k, db, idesc_pv,
                  (first && k == 0) ? 0u : 1u);
        }
      };

      mbar_wait(q_full, 0);
      mbar_wait(&k_full[0], 0);
      tc_fence_after();
      if (lane == 0) {
        issue_qk(0, 0);
        tc_commit(&s_full[0]);
        issue_qk(1, 0);
        tc_commit(&s_full[1]);
        tc_commit(&k_empty[0]);
      }
      __syncwarp();
      for (int j = 0; j < n_tiles; ++j) {
        const int st = j & 1;
        const uint32_t ph = (j >> 1) & 1;
        const int stn = (j + 1) & 1;
        const uint32_t phn = ((j + 1) >> 1) & 1;
        mbar_wait(&v_full[st], ph);
        for (int qt = 0; qt < 2; ++qt) {
          mbar_wait(&p_full[qt], j & 1);
          if (qt == 0 && j + 1 < n_tiles) mbar_wait(&k_full[stn], phn);
          tc_fence_after();
          if (lane == 0) {
            issue_pv(qt, st, j == 0);
            if (qt == 1) tc_commit(&v_empty[st]);
            if (j + 1 < n_tiles) {
              issue_qk(qt, stn);
              tc_commit(&s_full[qt]);
              if (qt == 1) tc_commit(&k_empty[stn]);
            }
          }
          __syncwarp();
        }
      }
      if (lane == 0) tc_commit(o_full);
      __syncwarp();
    }
  } else {
    // ------------------------------------------------------------- softmax
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(224));
    const int qt = (warp - 4) >> 2;
    const int lane_base = (warp & 3) * 32;
    const uint32_t t_lane = static_cast<uint32_t>(lane_base) << 16;
    const uint32_t t_s = tmem_base + t_lane + qt * 128;
    const uint32_t t_o = tmem_base + t_lane + 256 + qt * 128;
    const int q_row = q_row0 + qt * kQTile + lane_base + lane;
    float m_run = -INFINITY;  // running max, in the scaled log2 domain
    float l_run = 0.f;
    int j = 0;
    for (int sg = 0; sg < p.nseg; ++sg) {
      const int nt = (p.seg_rows[sg] + kKVTile - 1) / kKVTile;
      for (int t = 0; t < nt; ++t, ++j) {
        const int valid = min(kKVTile, p.seg_rows[sg] - t * kKVTile);
        mbar_wait(&s_full[qt], j & 1);
        tc_fence_after();
        uint32_t s[128];
#pragma unroll
        for (int c = 0; c < 4; ++c) tmem_ld_32x32(t_s + c * 32, *reinterpret_cast<uint32_t(*)[32]>(&s[c * 32]));
        tmem_ld_wait();
        if (valid < kKVTile) {
#pragma unroll
          for (int c = 0; c < 128; ++c)
            if (c >= valid) s[c] = 0xFF800000u;  // -inf
        }
        float mx = __uint_as_float(s[0]);
#pragma unroll
        for (int c = 1; c < 128; ++c) mx = fmaxf(mx, __uint_as_float(s[c]));
        const float m_new = fmaxf(m_run, mx * p.scale_log2);
        const bool grow = m_new > m_run + 8.0f;
        if (__any_sync(0xffffffffu, grow)) {
          const float alpha = fast_exp2(m_run - m_new);  // 0 on the first tile (m_run = -inf)
          l_run *= alpha;
          m_run = m_new;
          if (j > 0) {
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
        float sum = 0.f;
        uint32_t pk[64];
#pragma unroll
        for (int c = 0; c < 64; ++c) {
          const float p0 = fast_exp2(fmaf(__uint_as_float(s[2 * c]), p.scale_log2, -m_run));
          const float p1 = fast_exp2(fmaf(__uint_as_float(s[2 * c + 1]), p.scale_log2, -m_run));
          sum += p0 + p1;
          pk[c] = pack_bf16x2(p0, p1);
        }
        l_run += sum;
        tmem_st_32x32(t_s, *reinterpret_cast<uint32_t(*)[32]>(&pk[0]));
        tmem_st_32x32(t_s + 32, *reinterpret_cast<uint32_t(*)[32]>(&pk[32]));
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[qt]);
      }
    }
    // epilogue: O / l -> bf16 -> out[q_row, head*128 : head*128+128]
    mbar_wait(o_full, 0);
    tc_fence_after();
    const float inv_l = 1.0f / l_run;
    __nv_bfloat16* orow = p.out + static_cast<int64_t>(q_row) * p.ldo + head * kHD;
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      uint32_t o[32];
      tmem_ld_32x32(t_o + c * 32, o);
      tmem_ld_wait();
      if (q_row < p.Lq) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint32_t w[4];
#pragma unroll
          for (int i = 0; i < 4; ++i)
            w[i] = pack_bf16x2(__uint_as_float(o[g * 8 + 2 * i]) * inv_l, __uint_as_float(o[g * 8 + 2 * i + 1]) * inv_l);
          *reinterpret_cast<uint4*>(orow + c * 32 + g * 8) = make_uint4(w[0], w[1], w[2], w[3]);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

int flash_attn_bf16(const void* q, int64_t ldq, int Lq, int H, const void* k0, const void* v0,
                    int64_t ldkv0, int rows0, const void* k1, const void* v1, int64_t ldkv1, int rows1,
                    int nseg, const int* seg_start, const int* seg_rows, const int* seg_src, void* out,
                    int64_t ldo, float softmax_scale, cudaStream_t stream) {
  MMPL_CHECK(device_is_sm100(), MMPL_ERR_ARCH, "flash_attn: requires an sm_100 device");
  MMPL_CHECK(Lq > 0 && H > 0, MMPL_ERR_SHAPE, "flash_attn: bad Lq=%d H=%d", Lq, H);
  MMPL_CHECK(nseg >= 1 && nseg <= kMaxSeg, MMPL_ERR_SHAPE, "flash_attn: nseg=%d out of [1,%d]", nseg, kMaxSeg);
  MMPL_CHECK(ldo % 8 == 0, MMPL_ERR_SHAPE, "flash_attn: ldo must be a multiple of 8");
  AttnParams p{};
  p.Lq = Lq;
  p.out = static_cast<__nv_bfloat16*>(out);
  p.ldo = ldo;
  p.scale_log2 = softmax_scale * 1.4426950408889634f;
  p.nseg = nseg;
  bool uses1 = false;
  for (int i = 0; i < nseg; ++i) {
    const int src = seg_src ? seg_src[i] : 0;
    const int lim = src ? rows1 : rows0;
    MMPL_CHECK(seg_rows[i] > 0 && seg_start[i] >= 0 && seg_start[i] + seg_rows[i] <= lim, MMPL_ERR_SHAPE,
               "flash_attn: segment %d [%d,+%d) outside its %d-row source", i, seg_start[i], seg_rows[i], lim);
    p.seg_start[i] = seg_start[i];
    p.seg_rows[i] = seg_rows[i];
    p.seg_src[i] = src;
    uses1 |= (src != 0);
  }
  MMPL_CHECK(!uses1 || (k1 && v1), MMPL_ERR_ARG, "flash_attn: segment refers to a missing second K/V source");
  const CUtensorMap* mq = get_tensor_map_bf16(q, Lq, static_cast<uint64_t>(H) * kHD, ldq, 128);
  const CUtensorMap* mk0 = get_tensor_map_bf16(k0, rows0, static_cast<uint64_t>(H) * kHD, ldkv0, 128);
  const CUtensorMap* mv0 = get_tensor_map_bf16(v0, rows0, static_cast<uint64_t>(H) * kHD, ldkv0, 128);
  const CUtensorMap* mk1 = uses1 ? get_tensor_map_bf16(k1, rows1, static_cast<uint64_t>(H) * kHD, ldkv1, 128) : mk0;
  const CUtensorMap* mv1 = uses1 ? get_tensor_map_bf16(v1, rows1, static_cast<uint64_t>(H) * kHD, ldkv1, 128) : mv0;
  if (!mq || !mk0 || !mv0 || !mk1 || !mv1) return MMPL_ERR_CUDA;

  static bool attr_set = false;
  if (!attr_set) {
    MMPL_CUDA(cudaFuncSetAttribute(flash_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmem));
    attr_set = true;
  }
  dim3 grid((Lq + 2 * kQTile - 1) / (2 * kQTile), H);
  flash_attn_kernel<<<grid, kAttnThreads, kAttnSmem, stream>>>(*mq, *mk0, *mv0, *mk1, *mv1, p);
  MMPL_CUDA(cudaGetLastError());
  return MMPL_OK;
}

}  // namespace mmpl
