// Internal C++ launchers behind the C ABI (include/mmpl_b200.h). Each returns an MMPL_* status.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "mmpl_b200.h"

namespace mmpl {

int gemm_bf16(const void* a, int64_t lda, const void* w, int64_t ldw, const void* bias, void* out,
              int64_t ldo, int M, int N, int K, int epilogue, const void* residual, int64_t ldr,
              const void* gate, int64_t gate_stride, int rows_per_frame, int force_bn, cudaStream_t stream);

// Causal 3-D convolution as a tap-GEMM over a zero-haloed channels-last grid (gemm_tcgen05.cu: conv3d_cl).
int conv3d_cl(const void* in, const void* w_packed, const void* bias, void* out, const void* residual, int T, int H,
              int W, int Cin, int Cout, int KT, int KH, int KW, int history, cudaStream_t stream);

// vae_pointwise.cu: HBM-bound kernels of the VAE segment connect on the zero-haloed channels-last grid
int vae_norm_act(const void* x, void* out, int64_t rows, int C, const void* gamma, int silu, cudaStream_t st);
int vae_upsample2x(const void* in, void* out, int frames, int H, int W, int C, cudaStream_t st);
int vae_pick_odd(const void* in, void* out, int frames, int Hin, int Win, int C, cudaStream_t st);
int softmax_rows(const void* s, int64_t lds, void* p, int64_t ldp, int rows, int L, float scale, cudaStream_t st);

void gemm_set_streamk(int mode);  // -1 automatic (default), 0 off, 1 forced

int flash_attn_bf16(const void* q, int64_t ldq, int Lq, int H, const void* k0, const void* v0,
                    int64_t ldkv0, int rows0, const void* k1, const void* v1, int64_t ldkv1, int rows1,
                    int nseg, const int* seg_start, const int* seg_rows, const int* seg_src, void* out,
                    int64_t ldo, float softmax_scale, cudaStream_t stream);

void flash_attn_force_split(int split);
void flash_attn_force_ctas(int ctas);
// host-only enumeration of a launch's work partition (see attention_tcgen05.cu: attn_plan_pieces)
int flash_attn_plan(int Lq, int H, int kv_tiles, int ctas, int force_split, int* sched, int* pieces, int max_pieces);

int ln_modulate(const void* x, int64_t ldx, void* out, int64_t ldo, int S, int D, float eps, const void* shift,
                const void* scale, int64_t mod_stride, int rows_per_frame, cudaStream_t st);
int ln_affine(const void* x, int64_t ldx, void* out, int64_t ldo, int S, int D, float eps, const void* weight,
              const void* bias, cudaStream_t st);
int rmsnorm(const void* x, int64_t ldx, void* out, int64_t ldo, int S, int D, const void* weight, float eps,
            cudaStream_t st);
int qk_norm_rope_kv(const void* q_in, const void* k_in, const void* v_in, int64_t ld_in, const void* wq,
                    const void* wk, const void* rope_table, void* q_out, int64_t ldq, void* k_dst, void* v_dst,
                    int64_t ldkv, int S, int D, int gh, int gw, int n_frames, const int* frame_pos,
                    const int* kv_row, float eps, cudaStream_t st);
int modulation_add(const void* mod, const void* src, int64_t src_fstride, int64_t src_jstride, void* out, int F,
                   int J, int D, cudaStream_t st);
// all blocks at once: out[l][f][6][D] = mods_dev[l][6][D] + src[f][6][D]; mods_dev = device table of L pointers
int modulation_add_layers(const void* const* mods_dev, const void* src, void* out, int L, int F, int D, cudaStream_t st);
int sinusoid_embedding(const double* t, void* out, int F, int dim, cudaStream_t st);
int skinny_linear(const void* x, int64_t ldx, const void* w, const void* b, void* out, int64_t ldo, int M, int N,
                  int K, int silu_in, int silu_out, cudaStream_t st);
int patchify(const void* x, int64_t stride_f, int64_t stride_c, void* a, int F, int C, int H, int W, cudaStream_t st);
int unpatchify_x0(const void* head, int64_t ldh, const void* xt, int64_t xt_stride_f, int64_t xt_stride_c,
                  const double* sigma, void* flow, void* x0, int F, int C, int H, int W, cudaStream_t st);
int add_noise(const void* x0, const void* noise, const float* sigma, void* out, int n_frames, int64_t per_frame,
              cudaStream_t st);

// sampler.cu: CFG combine + flow->x0 + UniPC corrector / predictor in one launch
int unipc_cfg_step(const void* flow_cond, const void* flow_uncond, const void* sample, const void* m_prev1, const void* m_prev2,
                   const void* last_sample, void* sample_next, void* x0_out, void* corrected_out, int64_t n,
                   const mmpl_unipc_coeffs* k, cudaStream_t st);

}  // namespace mmpl
