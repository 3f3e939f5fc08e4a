/* mmpl_b200 — C ABI of the B200-native chunk-wise causal denoising hot path.
 *
 * The reference (Tele-AI/MMPL) is pure Python/PyTorch and has no FFI layer: its boundary for this
 * path is three Python call signatures plus two dict layouts. This header is the
 * boundary a native binding of that path would use instead; mmpl_b200/_lib.py binds it with ctypes
 * and mmpl_b200/causal_model.py mirrors the reference classes on top of it.
 *
 * Conventions: every function returns 0 on success or a negative MMPL_ERR_* code, the message is
 * available from mmpl_last_error(); no exceptions cross the ABI. All data pointers are DEVICE
 * pointers to bf16 unless stated otherwise, arrays documented as "host" are read during the call.
 * `stream` is a cudaStream_t passed as void* (0 = default stream). The caller owns all tensors; a
 * context owns only its activation workspace. A context is not thread-safe; use one per GPU/process.
 * Requires an sm_100 (B200) device: there is no fallback path, calls fail with MMPL_ERR_ARCH.
 */
#ifndef MMPL_B200_H_
#define MMPL_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MMPL_ABI_VERSION 4

enum {
  MMPL_OK = 0,
  MMPL_ERR_SHAPE = -1, /* unsupported or inconsistent shape / stride                */
  MMPL_ERR_ARCH = -2,  /* current device is not sm_100                              */
  MMPL_ERR_CUDA = -3,  /* CUDA runtime / driver call failed (see mmpl_last_error()) */
  MMPL_ERR_ARG = -4,   /* missing or invalid argument                               */
  MMPL_ERR_STATE = -5  /* context not ready (e.g. a weight is not bound)            */
};

/* GEMM epilogues (the element-wise op that follows each nn.Linear in the reference block). */
enum {
  MMPL_EPI_BIAS = 0,          /* y = bf16(acc + bias)                                               */
  MMPL_EPI_BIAS_GELU = 1,     /* bf16(gelu_tanh(y))           wan/modules/causal_model.py:267-269    */
  MMPL_EPI_BIAS_SILU = 2,     /* bf16(silu(y))                causal_model.py:828 (time_embedding)   */
  MMPL_EPI_BIAS_RES = 3,      /* bf16(res + y)                causal_model.py:314 (cross-attn)       */
  MMPL_EPI_BIAS_GATE_RES = 4  /* bf16(res + bf16(y*gate_f))   causal_model.py:310,322                */
};

int mmpl_abi_version(void);
/* Identity of the sources this library was built from (mmpl_b200/_build.py: hash of every .cu / header and the
 * compiler flags). The loader refuses a library whose id differs from the tree it was started from. */
const char* mmpl_build_id(void);
const char* mmpl_last_error(void);

/* ---------------------------------------------------------------------------------------------
 * Per-kernel entry points (unit-testable pieces of CausalWanAttentionBlock.forward,
 * wan/modules/causal_model.py:274-326).
 * ------------------------------------------------------------------------------------------- */

/* out[M,N] = epilogue(A[M,K] . W[N,K]^T + bias[N]); replaces nn.Linear (+ fused follow-up op).
 * lda/ldw/ldo/ldr are row pitches in elements. residual may alias out. gate is [frames][gate_stride]
 * (row r uses frame r / rows_per_frame). tile_n: 0 = auto; 64/128/192/256 = single-CTA tiles of that width; 512/448 =
 * cta_group::2 pairs (256 x 256 / 256 x 224 per pair); 1128/1192 = clusters of two 128 x 128 / 128 x 192 tiles that
 * share their A tile by TMA multicast (needs an even number of tiles along N). */
int mmpl_gemm_bf16(const void* a, int64_t lda, const void* w, int64_t ldw, const void* bias, void* out,
                   int64_t ldo, int M, int N, int K, int epilogue, const void* residual, int64_t ldr,
                   const void* gate, int64_t gate_stride, int rows_per_frame, int tile_n, void* stream);

/* Causal 3-D convolution, stride 1, channels-last, as a tap-GEMM on the tensor cores: replaces CausalConv3d.forward
 * (wan/modules/vae.py:16-36: F.pad with all temporal padding in front + nn.Conv3d) and, with KT = 1, the per-frame
 * nn.Conv2d(3x3, padding 1) / 1x1 convolutions of the same file (:75-81, :233-235) -- the contraction that carries the
 * VAE "segment connect" either side of the anchor hand-off (Wan_fps_inference_parallel_4gpu_20s.py:191-205;
 * SURVEY.md 8(f) row 2). Layouts (bf16):
 *   in       [history + T][H + 2][W + 2][Cin]   one-pixel zero halo; `history` in 0..KT-1 = carried frames stored in front
 *                                               of the T frames (what `feat_cache` carries between chunks, vae.py:197-209);
 *                                               the remaining KT-1-history causal padding frames are implicit zeros
 *   w_packed KW = 1: [Cout][KT*KH][Cin64]       Cin64 = Cin rounded up to a multiple of 64, zero padded
 *            KW = 3: [Cout][KT*KH][K3]          K3 = 3*Cin rounded up to 64; inside a span the order is (dw, c): the three
 *                                               dw taps of an image row are read as ONE contiguous run of the grid
 *                                               (= weight.permute(0,2,3,4,1) of the reference parameter, rows padded)
 *   bias     [Cout] or NULL;  residual: NULL or a tensor laid out like `out` (ResidualBlock's `x + h`, vae.py:213)
 *   out      [T][H + 2][W + 2][Cout]            every position is written: interior = convolution, halo = 0
 * Cin and Cout must be multiples of 8 (pad channels in the layout); KT, KH = KW in {1, 3}. */
int mmpl_conv3d_cl(const void* in, const void* w_packed, const void* bias, void* out, const void* residual, int T, int H,
                   int W, int Cin, int Cout, int KT, int KH, int KW, int history, void* stream);

/* RMS_norm.forward (wan/modules/vae.py:39-55: F.normalize over channels * sqrt(C) * gamma), optionally followed by
 * nn.SiLU (the `RMS_norm, SiLU` pairs of ResidualBlock / the heads, vae.py:180-186,309-311,413-415), over `rows` rows of C
 * bf16 channels (any contiguous [rows, C] view of the haloed grid: zero rows stay zero). bf16 rounding after each of the
 * reference's operators. C % 8 == 0, C <= 1024. out may alias x. */
int mmpl_vae_norm_act(const void* x, void* out, int64_t rows, int C, const void* gamma, int silu, void* stream);

/* Upsample(scale_factor=(2,2), mode='nearest') of every frame (vae.py:58-64,75-78): in [frames][H+2][W+2][C] ->
 * out [frames][2H+2][2W+2][C]; every position of `out` is written (halo = 0). */
int mmpl_vae_upsample2x(const void* in, void* out, int frames, int H, int W, int C, void* stream);

/* The stride-2 pick that turns a stride-1 3x3 "same" convolution of the haloed grid into the reference's
 * nn.ZeroPad2d((0,1,0,1)) + nn.Conv2d(3, stride=(2,2)) (vae.py:85-88): out interior (i, j) = in interior (2i+1, 2j+1).
 * in [frames][Hin+2][Win+2][C] -> out [frames][Hin/2+2][Win/2+2][C]; every position of `out` is written (halo = 0). */
int mmpl_vae_pick_odd(const void* in, void* out, int frames, int Hin, int Win, int C, void* stream);

/* p[r, :] = softmax(scale * s[r, :]) in fp32, bf16 in and out (row pitches lds / ldp in elements): the softmax of the
 * single-head attention of AttentionBlock (vae.py:241-266), whose head dimension (384) is outside mmpl_flash_attn's. */
int mmpl_softmax_rows(const void* s, int64_t lds, void* p, int64_t ldp, int rows, int L, float scale, void* stream);

/* Non-causal softmax(Q K^T * scale) V, head_dim 128; replaces flash_attention()/attention()
 * (wan/modules/attention.py:32-185). q: [Lq, H, 128] with row pitch ldq, out likewise with ldo.
 * KV source 0 (k0,v0: [rows0, H, 128], pitch ldkv0) and optional source 1 are read in place through
 * `nseg` (<= 8) row segments (host arrays): segment i = rows [seg_start[i], seg_start[i]+seg_rows[i])
 * of source seg_src[i] (seg_src may be NULL = all source 0). */
int mmpl_flash_attn(const void* q, int64_t ldq, int Lq, int H, const void* k0, const void* v0,
                    int64_t ldkv0, int rows0, const void* k1, const void* v1, int64_t ldkv1, int rows1,
                    int nseg, const int* seg_start, const int* seg_rows, const int* seg_src, void* out,
                    int64_t ldo, float softmax_scale, void* stream);

/* Tuning / test hook for the attention work partition: split > 0 forces the uniform schedule with that many KV chunks
 * per unit; -1000 < split < 0 forces the range schedule with -split heads per group; split <= -1000 forces the hybrid
 * schedule (whole units in lockstep rounds, the remainder as ranges) wherever it applies; 0 = cost model (default). */
int mmpl_attn_set_split(int split);
/* Test hook: run the persistent attention grid with `ctas` CTAs instead of one per SM (0 = default), so that small
 * problems exercise the multi-round schedules. */
int mmpl_attn_set_ctas(int ctas);
/* Host-only (no GPU needed): the work partition mmpl_flash_attn would use for Lq query rows, H heads and kv_tiles
 * 128-row KV tiles per unit on a persistent grid of `ctas` CTAs (force_split as in mmpl_attn_set_split).
 * sched[7] = {schedule (0 uniform split, 1 ranges, 2 hybrid), split, heads per group, whole units, grid size, workspace
 * slots, query-tile pairs per head}; pieces[max_pieces][9] = {cta, head, q_row0, first KV tile, KV tiles, whole, slot,
 * pieces of the unit according to the merge bookkeeping, slot found by it}. Returns the number of pieces,
 * MMPL_ERR_SHAPE if max_pieces is too small, MMPL_ERR_ARG for invalid arguments. Used by the CPU tests to check that every (query tile pair,
 * KV tile) is covered exactly once by every schedule. */
int mmpl_attn_plan(int Lq, int H, int kv_tiles, int ctas, int force_split, int* sched, int* pieces, int max_pieces);
/* Tuning / test hook: stream-K tail schedule of the cta_group::2 GEMM: 0 = never (default; slower on B200 at the
 * cfg2 shapes, see gemm_tcgen05.cu), -1 = automatic (when whole 256x256 tiles would leave more than 4 % of the last
 * wave empty), 1 = whenever tiles % pairs != 0. */
int mmpl_gemm_set_streamk(int mode);

/* bf16(bf16(bf16(LayerNorm(x)) * bf16(1 + scale_f)) + shift_f); shift/scale are [frames][mod_stride]
 * (causal_model.py:305,318; WanLayerNorm model.py:89-99). D in {256,512,1536,5120}. */
int mmpl_ln_modulate(const void* x, int64_t ldx, void* out, int64_t ldo, int S, int D, float eps,
                     const void* shift, const void* scale, int64_t mod_stride, int rows_per_frame,
                     void* stream);
/* bf16(LayerNorm(x) * weight + bias)  (norm3, causal_model.py:314). */
int mmpl_ln_affine(const void* x, int64_t ldx, void* out, int64_t ldo, int S, int D, float eps,
                   const void* weight, const void* bias, void* stream);
/* WanRMSNorm over the full row (model.py:70-86). May run in place. */
int mmpl_rmsnorm(const void* x, int64_t ldx, void* out, int64_t ldo, int S, int D, const void* weight,
                 float eps, void* stream);

/* RMSNorm(q), RMSNorm(k), 3-D RoPE on both, roped q -> q_out, roped k and v -> KV-cache rows
 * (causal_model.py:111-113,193-217; causal_rope_apply :27-55; causal_fps_model.py:192-217).
 * rope_table: device double [1024][64][2] (cos,sin). Token r of frame f (r in [0, gh*gw)) goes to row
 * kv_row[f] + r of k_dst/v_dst and uses temporal position frame_pos[f] (host int arrays [n_frames]).
 * q_out may alias q_in. */
int mmpl_qk_norm_rope_kv(const void* q_in, const void* k_in, const void* v_in, int64_t ld_in,
                         const void* norm_q_weight, const void* norm_k_weight, const void* rope_table,
                         void* q_out, int64_t ldq, void* k_dst, void* v_dst, int64_t ldkv, int S, int D,
                         int gh, int gw, int n_frames, const int* frame_pos, const int* kv_row, float eps,
                         void* stream);

/* out[f][j][:] = bf16(mod[j][:] + src[f*src_fstride + j*src_jstride + :])  (causal_model.py:300,355) */
int mmpl_modulation_add(const void* mod, const void* src, int64_t src_fstride, int64_t src_jstride,
                        void* out, int F, int J, int D, void* stream);
/* sinusoidal_embedding_1d (model.py:15-25), float64 -> bf16. t: device double [F]. out: [F, dim]. */
int mmpl_sinusoid_embedding(const double* t, void* out, int F, int dim, void* stream);
/* Linear for M <= 32 rows with optional SiLU on the input and/or output (time MLP, causal_model.py:828-831). */
int mmpl_skinny_linear(const void* x, int64_t ldx, const void* w, const void* b, void* out, int64_t ldo,
                       int M, int N, int K, int silu_in, int silu_out, void* stream);
/* im2col of the (1,2,2) patch embedding (causal_model.py:812): x is [F][C][H][W] with the given
 * frame/channel strides; a is [F*(H/2)*(W/2), C*4]. */
int mmpl_patchify(const void* x, int64_t stride_f, int64_t stride_c, void* a, int F, int C, int H, int W,
                  void* stream);
/* unpatchify (causal_model.py:1094-1117) -> flow [F][C][H][W]; if x0 != NULL also
 * x0 = bf16(double(xt) - sigma_f * double(flow)) (utils/wan_wrapper.py:172-196). sigma: device double [F]. */
int mmpl_unpatchify_x0(const void* head, int64_t ldh, const void* xt, int64_t xt_stride_f,
                       int64_t xt_stride_c, const double* sigma, void* flow, void* x0, int F, int C, int H,
                       int W, void* stream);
/* Anchor hand-off between segments: replaces torch.save(latents_chunk{k}.pt) on the producer and the poll / torch.load /
 * os.remove loop on the consumer (pipeline/casual_fps_inference.py:380-383; Wan_fps_inference_parallel_4gpu_20s.py:183-189)
 * for a host that owns a raw NCCL communicator: ncclBroadcast of `bytes` bytes at `buf` from rank `root` of `nccl_comm`
 * (an ncclComm_t), enqueued on `stream` behind the anchor stage's last kernel. In place on every rank. t2v payload
 * [1,8,16,60,104] bf16 = 1.6 MB, i2v [1,3,16,60,104] = 0.6 MB. NCCL is looked up in the process at run time
 * (MMPL_ERR_STATE if it is not loaded). The Python host side uses torch.distributed's communicator instead
 * (mmpl_b200/segment_parallel.py: point-to-point isend / recv on the same stream). */
int mmpl_anchor_broadcast(void* nccl_comm, void* buf, int64_t bytes, int root, void* stream);

/* FlowMatchScheduler.add_noise (utils/scheduler.py:159-176): bf16((1-sigma_f)*x0 + sigma_f*noise), fp32.
 * sigma: device float [n_frames]; tensors are [n_frames][per_frame] contiguous. */
int mmpl_add_noise(const void* x0, const void* noise, const float* sigma, void* out, int n_frames,
                   int64_t per_frame, void* stream);

/* ---------------------------------------------------------------------------------------------
 * One denoising step of the MMPL hot loop between two backbone forwards:
 * CFG combine (pipeline/casual_fps_inference.py:366-374) + FlowUniPCMultistepScheduler.step
 * (wan/utils/fm_solvers_unipc.py:655-739: convert_model_output :318-321, multistep_uni_c_bh_update :486-626,
 * multistep_uni_p_bh_update :350-484) for solver_order 2 / "bh2" / predict_x0 / flow_prediction, as ONE launch.
 * The element-wise torch operators of the reference are executed in their order with a bf16 rounding after each, so
 * the result equals the operator sequence bit for bit. All scalars are host-computed per step:
 *   flow      = bf16(u + bf16(guidance * bf16(c - u)))                 (flow_uncond == NULL: flow = flow_cond)
 *   x0        = bf16(x - bf16(sigma * flow))
 *   corrected = corr_order == 0 ? x :
 *               bf16( bf16(bf16(corr_a*last) - bf16(corr_b*m1)) - bf16(corr_c * res) ),
 *               res = bf16(corr_rho1 * bf16(x0 - m1))  [+ bf16(corr_rho0 * D(m2 - m1, corr_rk)) when corr_order == 2]
 *   next      = bf16( bf16(bf16(pred_a*corrected) - bf16(pred_b*x0)) - bf16(pred_c * 0.5 * D(m1 - x0, pred_rk)) )
 *               (pred_order == 1: the last term is pred_c * 0)
 *   D(d, rk)  = true_division ? bf16(bf16(d) / rk) : bf16(rk * bf16(d))
 * torch evaluates `tensor / cpu_scalar` as a multiplication by the reciprocal on CUDA and as a division on the CPU, and
 * rounds a CPU scalar to bf16 before `cpu_scalar * tensor` on the CPU only; the host encodes either behaviour in the
 * coefficients (mmpl_b200/unipc.py) - the default is what the reference computes on a GPU.
 * m1 / m2 = x0 of the previous / second-previous step (model_outputs[-1], [-2]), last = the corrected sample of the
 * previous step (last_sample). Outputs may alias the input they replace (next = x, corrected = last, x0_out = m2).
 * All tensors: n contiguous bf16, 16-byte aligned. */
typedef struct {
  float guidance, sigma;
  int corr_order; /* 0 (first step), 1, 2 */
  float corr_a, corr_b, corr_c, corr_rk, corr_rho0, corr_rho1;
  int pred_order; /* 1, 2 */
  float pred_a, pred_b, pred_c, pred_rk;
  int true_division;
} mmpl_unipc_coeffs;

int mmpl_unipc_cfg_step(const void* flow_cond, const void* flow_uncond, const void* x, const void* m1, const void* m2,
                        const void* last, void* next, void* x0_out, void* corrected_out, int64_t n,
                        const mmpl_unipc_coeffs* coeffs /* host */, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Whole-forward entry point: CausalWanModel._forward_inference (causal_model.py:763-892) and the
 * KV-cache branch of CausalFPSWanModel (causal_fps_model.py:192-264) for batch size 1.
 * ------------------------------------------------------------------------------------------- */
typedef struct mmpl_ctx mmpl_ctx;

typedef struct {
  int dim, ffn_dim, num_heads, num_layers;
  int freq_dim, text_dim, text_len, in_dim, out_dim;
  float eps;
  int max_tokens; /* workspace capacity: largest n_frames * tokens-per-frame of any forward */
} mmpl_model_config;

int mmpl_ctx_create(const mmpl_model_config* cfg, mmpl_ctx** out);
void mmpl_ctx_destroy(mmpl_ctx* ctx);
/* Bind a parameter by its reference state-dict name (causal_model.py module tree), e.g.
 * "blocks.3.self_attn.o.weight", "head.modulation", "patch_embedding.weight"; the fused
 * [3*dim, dim] q/k/v projection is bound as "blocks.N.self_attn.qkv.{weight,bias}". The pointer must
 * stay valid while the context is used. */
int mmpl_bind_weight(mmpl_ctx* ctx, const char* name, const void* ptr, int64_t numel);
int mmpl_bind_rope_table(mmpl_ctx* ctx, const void* table /* device double [1024][64][2] */);
/* Number of kernels the context launched since the last call with reset != 0. */
int64_t mmpl_launch_count(mmpl_ctx* ctx, int reset);
/* Changes whenever the library reallocates one of its process-level device workspaces (attention partials, merge
 * counters, stream-K slots): a CUDA graph captured from mmpl_forward is valid only while this value is the one seen
 * at capture. */
int64_t mmpl_workspace_generation(void);
/* Adds n (may be negative) to the context's and the process's launch counters and returns the context's: a host that
 * captures a forward's launch sequence in a CUDA graph credits the captured count for every replay, so that the counters
 * keep meaning "kernels of this library that ran". */
int64_t mmpl_launch_credit(mmpl_ctx* ctx, int64_t n);
/* Kernels launched through any entry point of this library in this process (bench.py "gpu_launches"). */
int64_t mmpl_total_launches(int reset);

/* Device timing by kernel category for the roofline report: when a category's bit is set in `category_mask`,
 * mmpl_forward brackets each launch of that category with CUDA events on the launch stream.
 * mmpl_profile_read synchronises those events and returns, per category, the summed milliseconds, the summed
 * algorithmic work (FLOPs for categories 0-2, bytes for 3) and the launch count. */
enum { MMPL_PROF_SELF_ATTN = 0, MMPL_PROF_CROSS_ATTN = 1, MMPL_PROF_GEMM = 2, MMPL_PROF_POINTWISE = 3, MMPL_PROF_NCAT = 4 };
int mmpl_profile_enable(mmpl_ctx* ctx, int category_mask);
int mmpl_profile_mask(mmpl_ctx* ctx); /* the mask currently set */
int mmpl_profile_read(mmpl_ctx* ctx, double* ms, double* work, int64_t* launches, int reset);
/* The same spans split by call site inside CausalWanAttentionBlock.forward (causal_model.py:274-326): arrays of
 * MMPL_SITE_COUNT entries. Diagnostic only (bench.py "breakdown.sites", tools/). */
enum {
  MMPL_SITE_SELF_ATTN = 0, MMPL_SITE_CROSS_ATTN = 1, MMPL_SITE_GEMM_QKV = 2, MMPL_SITE_GEMM_O = 3, MMPL_SITE_GEMM_CQ = 4,
  MMPL_SITE_GEMM_CO = 5, MMPL_SITE_GEMM_FFN0 = 6, MMPL_SITE_GEMM_FFN2 = 7, MMPL_SITE_LN_MOD = 8, MMPL_SITE_LN_AFFINE = 9,
  MMPL_SITE_ROPE_KV = 10, MMPL_SITE_RMSNORM = 11, MMPL_SITE_MODADD = 12, MMPL_SITE_OTHER = 13, MMPL_SITE_COUNT = 14
};
int mmpl_profile_read_sites(mmpl_ctx* ctx, double* ms, double* work, int64_t* launches, int reset);

typedef struct {
  /* latent chunk [n_frames][in_dim][lat_h][lat_w] (strides in elements) */
  const void* latents;
  int64_t lat_stride_f, lat_stride_c;
  int n_frames, lat_h, lat_w;
  const double* timesteps; /* device [n_frames] */
  const void* context;     /* [text_len, text_dim]; required when cross_init == 0 */
  /* self-attention KV cache: host arrays [num_layers] of device pointers to [cache_rows, H, 128] */
  void* const* kv_k;
  void* const* kv_v;
  int64_t cache_rows;
  const int* frame_pos; /* host [n_frames]: temporal RoPE position of each frame                  */
  const int* kv_row;    /* host [n_frames]: cache row receiving each frame's first token           */
  int kv_to_tail;       /* 1: K/V of this call are NOT written to the cache but attended as an extra
                           trailing segment (last MMPL stage, causal_fps_model.py:254-264)          */
  int n_seg;            /* cache row segments attended after the write (host arrays, <= 7)         */
  const int* seg_start;
  const int* seg_rows;
  /* cross-attention cache: host arrays [num_layers] of device pointers to [text_len, H, 128] */
  void* const* cross_k;
  void* const* cross_v;
  int cross_init; /* 0: compute K/V from `context` into cross_k/cross_v first (model.py:174-180)  */
  /* outputs, [n_frames][out_dim][lat_h][lat_w] contiguous */
  void* flow;
  void* x0;            /* may be NULL */
  const double* sigma; /* device [n_frames], required when x0 != NULL */
} mmpl_forward_args;

int mmpl_forward(mmpl_ctx* ctx, const mmpl_forward_args* args, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MMPL_B200_H_ */
