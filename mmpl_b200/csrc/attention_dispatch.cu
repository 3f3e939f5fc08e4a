// Public entry of the attention kernel: applies the split override set through mmpl_attn_set_split().
#include "host_util.h"
#include "kernels.h"

namespace mmpl {

namespace prod {
int flash_attn_impl(const void* q, int64_t ldq, int Lq, int H, const void* k0, const void* v0, int64_t ldkv0, int rows0,
                    const void* k1, const void* v1, int64_t ldkv1, int rows1, int nseg, const int* seg_start,
                    const int* seg_rows, const int* seg_src, void* out, int64_t ldo, float softmax_scale, int force_split,
                    cudaStream_t stream);
}

static int g_force_split = 0;
void flash_attn_force_split(int split) { g_force_split = split; }

int flash_attn_bf16(const void* q, int64_t ldq, int Lq, int H, const void* k0, const void* v0, int64_t ldkv0, int rows0,
                    const void* k1, const void* v1, int64_t ldkv1, int rows1, int nseg, const int* seg_start,
                    const int* seg_rows, const int* seg_src, void* out, int64_t ldo, float softmax_scale,
                    cudaStream_t stream) {
  return prod::flash_attn_impl(q, ldq, Lq, H, k0, v0, ldkv0, rows0, k1, v1, ldkv1, rows1, nseg, seg_start, seg_rows, seg_src,
                               out, ldo, softmax_scale, g_force_split, stream);
}

}  // namespace mmpl
