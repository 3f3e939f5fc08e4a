// Public entry of the attention kernel: applies the overrides set through mmpl_attn_set_split() / mmpl_attn_set_ctas()
// and picks the kernel variant: namespace prod = whole-tile S hand-over (default for every call), namespace half =
// half-tile S pipeline (see MMPL_ATTN_SPLIT_S in attention_tcgen05.cu). The half-tile pipeline is faster in isolation
// where the S -> P round trip is exposed (cross-attention, 4 KV tiles per unit: 23.6 vs 30.2 us), slower on long KV
// ranges, and made no difference inside the cfg2 step (972.4 vs 974.1 ms), so it is opt-in: MMPL_ATTN_HALF=1 uses it
// for every call, MMPL_ATTN_HALF_TILES=n for calls with at most n KV tiles per unit.
#include <cstdlib>

#include "host_util.h"
#include "kernels.h"

namespace mmpl {

#define MMPL_DECLARE_ATTN_IMPL(ns)                                                                                          \
  namespace ns {                                                                                                            \
  int flash_attn_impl(const void* q, int64_t ldq, int Lq, int H, const void* k0, const void* v0, int64_t ldkv0, int rows0, \
                      const void* k1, const void* v1, int64_t ldkv1, int rows1, int nseg, const int* seg_start,            \
                      const int* seg_rows, const int* seg_src, void* out, int64_t ldo, float softmax_scale,                \
                      int force_split, int force_ctas, cudaStream_t stream);                                                \
  }
MMPL_DECLARE_ATTN_IMPL(prod)
MMPL_DECLARE_ATTN_IMPL(half)
namespace prod {
int attn_plan_pieces(int Lq, int H, int kv_tiles, int ctas, int force_split, int* sched, int* pieces, int max_pieces);
}
int flash_attn_plan(int Lq, int H, int kv_tiles, int ctas, int force_split, int* sched, int* pieces, int max_pieces) {
  return prod::attn_plan_pieces(Lq, H, kv_tiles, ctas, force_split, sched, pieces, max_pieces);
}

static int g_force_split = 0;
void flash_attn_force_split(int split) { g_force_split = split; }
static int g_force_ctas = 0;
void flash_attn_force_ctas(int ctas) { g_force_ctas = ctas; }


int flash_attn_bf16(const void* q, int64_t ldq, int Lq, int H, const void* k0, const void* v0, int64_t ldkv0, int rows0,
                    const void* k1, const void* v1, int64_t ldkv1, int rows1, int nseg, const int* seg_start,
                    const int* seg_rows, const int* seg_src, void* out, int64_t ldo, float softmax_scale,
                    cudaStream_t stream) {
  static const int forced = getenv("MMPL_ATTN_HALF") ? atoi(getenv("MMPL_ATTN_HALF")) : -1;
  static const int half_tiles = getenv("MMPL_ATTN_HALF_TILES") ? atoi(getenv("MMPL_ATTN_HALF_TILES")) : 0;
  bool use_half = forced == 1;
  if (forced < 0 && half_tiles > 0 && seg_rows != nullptr && nseg >= 1 && nseg <= 8) {
    long long tiles = 0;
    for (int i = 0; i < nseg; ++i) tiles += (seg_rows[i] + 127) / 128;
    use_half = tiles <= half_tiles;
  }
  auto impl = use_half ? half::flash_attn_impl : prod::flash_attn_impl;
  return impl(q, ldq, Lq, H, k0, v0, ldkv0, rows0, k1, v1, ldkv1, rows1, nseg, seg_start, seg_rows, seg_src, out, ldo,
              softmax_scale, g_force_split, g_force_ctas, stream);
}

}  // namespace mmpl
