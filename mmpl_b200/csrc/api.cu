// C ABI (include/mmpl_b200.h): thin extern "C" wrappers over the kernel launchers plus the whole-forward
// orchestration of CausalWanModel._forward_inference (wan/modules/causal_model.py:763-892) as a fixed
// sequence of launches on one stream. All integer bookkeeping (cache rows, RoPE positions, attended
// segments) is decided by the caller and passed in, so this file contains no cache policy.
#include <cuda_bf16.h>

#include <dlfcn.h>

#include <cstring>
#include <string>
#include <vector>

#include "host_util.h"
#include "kernels.h"
#include "mmpl_b200.h"

using namespace mmpl;

namespace {

struct LayerWeights {
  const void *qkv_w = nullptr, *qkv_b = nullptr, *norm_q = nullptr, *norm_k = nullptr, *o_w = nullptr, *o_b = nullptr;
  const void *norm3_w = nullptr, *norm3_b = nullptr;
  const void *cq_w = nullptr, *cq_b = nullptr, *ck_w = nullptr, *ck_b = nullptr, *cv_w = nullptr, *cv_b = nullptr;
  const void *co_w = nullptr, *co_b = nullptr, *cnorm_q = nullptr, *cnorm_k = nullptr;
  const void *ffn0_w = nullptr, *ffn0_b = nullptr, *ffn2_w = nullptr, *ffn2_b = nullptr;
  const void* modulation = nullptr;
};

struct GlobalWeights {
  const void *patch_w = nullptr, *patch_b = nullptr;
  const void *text0_w = nullptr, *text0_b = nullptr, *text2_w = nullptr, *text2_b = nullptr;
  const void *time0_w = nullptr, *time0_b = nullptr, *time2_w = nullptr, *time2_b = nullptr;
  const void *tproj_w = nullptr, *tproj_b = nullptr;
  const void *head_w = nullptr, *head_b = nullptr, *head_mod = nullptr;
};

struct FieldDesc {
  const char* name;
  size_t offset;
  int kind;  // expected numel: 0=D*D 1=D 2=3D*D 3=3D 4=Fd*D 5=Fd 6=6D 7=other(any)
};

#define LW(field) offsetof(LayerWeights, field)
const FieldDesc kLayerFields[] = {
    {"self_attn.qkv.weight", LW(qkv_w), 2},   {"self_attn.qkv.bias", LW(qkv_b), 3},
    {"self_attn.norm_q.weight", LW(norm_q), 1}, {"self_attn.norm_k.weight", LW(norm_k), 1},
    {"self_attn.o.weight", LW(o_w), 0},       {"self_attn.o.bias", LW(o_b), 1},
    {"norm3.weight", LW(norm3_w), 1},         {"norm3.bias", LW(norm3_b), 1},
    {"cross_attn.q.weight", LW(cq_w), 0},     {"cross_attn.q.bias", LW(cq_b), 1},
    {"cross_attn.k.weight", LW(ck_w), 0},     {"cross_attn.k.bias", LW(ck_b), 1},
    {"cross_attn.v.weight", LW(cv_w), 0},     {"cross_attn.v.bias", LW(cv_b), 1},
    {"cross_attn.o.weight", LW(co_w), 0},     {"cross_attn.o.bias", LW(co_b), 1},
    {"cross_attn.norm_q.weight", LW(cnorm_q), 1}, {"cross_attn.norm_k.weight", LW(cnorm_k), 1},
    {"ffn.0.weight", LW(ffn0_w), 4},          {"ffn.0.bias", LW(ffn0_b), 5},
    {"ffn.2.weight", LW(ffn2_w), 4},          {"ffn.2.bias", LW(ffn2_b), 1},
    {"modulation", LW(modulation), 6},
};
#undef LW
#define GW(field) offsetof(GlobalWeights, field)
const FieldDesc kGlobalFields[] = {
    {"patch_embedding.weight", GW(patch_w), 7}, {"patch_embedding.bias", GW(patch_b), 1},
    {"text_embedding.0.weight", GW(text0_w), 7}, {"text_embedding.0.bias", GW(text0_b), 1},
    {"text_embedding.2.weight", GW(text2_w), 0}, {"text_embedding.2.bias", GW(text2_b), 1},
    {"time_embedding.0.weight", GW(time0_w), 7}, {"time_embedding.0.bias", GW(time0_b), 1},
    {"time_embedding.2.weight", GW(time2_w), 0}, {"time_embedding.2.bias", GW(time2_b), 1},
    {"time_projection.1.weight", GW(tproj_w), 7}, {"time_projection.1.bias", GW(tproj_b), 6},
    {"head.head.weight", GW(head_w), 7},        {"head.head.bias", GW(head_b), 7},
    {"head.modulation", GW(head_mod), 7},
};
#undef GW

}  // namespace

struct mmpl_ctx {
  mmpl_model_config cfg;
  std::vector<LayerWeights> layers;
  GlobalWeights g;
  const void* rope_table = nullptr;
  int64_t launches = 0;
  // optional per-category device timing (bench.py roofline): events around every launch of a category
  int profile_mask = 0;
  std::vector<cudaEvent_t> ev_pool;
  size_t ev_used = 0;
  std::vector<std::pair<int, std::pair<size_t, size_t>>> ev_spans;  // (category, (start idx, stop idx))
  double prof_ms[MMPL_PROF_NCAT] = {0, 0, 0, 0};
  int64_t prof_launches[MMPL_PROF_NCAT] = {0, 0, 0, 0};
  double prof_work[MMPL_PROF_NCAT] = {0, 0, 0, 0};  // algorithmic FLOPs (cat 0-2) or bytes (cat 3)
  // the same spans by call site inside the block (MMPL_SITE_*): which Linear / norm the time belongs to
  int cur_site = MMPL_SITE_OTHER;
  double site_ms[MMPL_SITE_COUNT] = {};
  double site_work[MMPL_SITE_COUNT] = {};
  int64_t site_launches[MMPL_SITE_COUNT] = {};
  // workspace (device)
  char* ws = nullptr;
  size_t ws_bytes = 0;
  void *x = nullptr, *xm = nullptr, *qkv = nullptr, *attn = nullptr, *h = nullptr, *patch = nullptr, *hout = nullptr;
  void *sinus = nullptr, *t1 = nullptr, *e = nullptr, *e0 = nullptr, *emod = nullptr, *ehead = nullptr;
  void *ctx_h = nullptr, *ctx_e = nullptr, *tail_k = nullptr, *tail_v = nullptr;
  // device table of the blocks' `modulation` parameters (re-uploaded after a bind) for modulation_add_layers
  void* mod_table = nullptr;
  std::vector<const void*> mod_table_host;
  bool mod_table_dirty = true;
};

static int64_t g_total_launches = 0;
#define COUNTED(call)                         \
  do {                                        \
    const int _s = (call);                    \
    if (_s == MMPL_OK) ++g_total_launches;    \
    return _s;                                \
  } while (0)

extern "C" {

int mmpl_abi_version(void) { return MMPL_ABI_VERSION; }
#ifndef MMPL_BUILD_ID
#define MMPL_BUILD_ID "unknown"
#endif
const char* mmpl_build_id(void) { return MMPL_BUILD_ID; }

int mmpl_unipc_cfg_step(const void* flow_cond, const void* flow_uncond, const void* x, const void* m1, const void* m2,
                        const void* last, void* next, void* x0_out, void* corrected_out, int64_t n,
                        const mmpl_unipc_coeffs* coeffs, void* stream) {
  COUNTED(unipc_cfg_step(flow_cond, flow_uncond, x, m1, m2, last, next, x0_out, corrected_out, n, coeffs,
                         static_cast<cudaStream_t>(stream)));
}
const char* mmpl_last_error(void) { return last_error_buf(); }

int mmpl_gemm_bf16(const void* a, int64_t lda, const void* w, int64_t ldw, const void* bias, void* out,
                   int64_t ldo, int M, int N, int K, int epilogue, const void* residual, int64_t ldr,
                   const void* gate, int64_t gate_stride, int rows_per_frame, int tile_n, void* stream) {
  COUNTED(gemm_bf16(a, lda, w, ldw, bias, out, ldo, M, N, K, epilogue, residual, ldr, gate, gate_stride,
                   rows_per_frame, tile_n, static_cast<cudaStream_t>(stream)));
}

int mmpl_conv3d_cl(const void* in, const void* w_packed, const void* bias, void* out, const void* residual, int T, int H,
                   int W, int Cin, int Cout, int KT, int KH, int KW, int history, void* stream) {
  COUNTED(conv3d_cl(in, w_packed, bias, out, residual, T, H, W, Cin, Cout, KT, KH, KW, history, static_cast<cudaStream_t>(stream)));
}

int mmpl_vae_norm_act(const void* x, void* out, int64_t rows, int C, const void* gamma, int silu, void* stream) {
  COUNTED(vae_norm_act(x, out, rows, C, gamma, silu, static_cast<cudaStream_t>(stream)));
}
int mmpl_vae_upsample2x(const void* in, void* out, int frames, int H, int W, int C, void* stream) {
  COUNTED(vae_upsample2x(in, out, frames, H, W, C, static_cast<cudaStream_t>(stream)));
}
int mmpl_vae_pick_odd(const void* in, void* out, int frames, int Hin, int Win, int C, void* stream) {
  COUNTED(vae_pick_odd(in, out, frames, Hin, Win, C, static_cast<cudaStream_t>(stream)));
}
int mmpl_softmax_rows(const void* s, int64_t lds, void* p, int64_t ldp, int rows, int L, float scale, void* stream) {
  COUNTED(softmax_rows(s, lds, p, ldp, rows, L, scale, static_cast<cudaStream_t>(stream)));
}

// Anchor hand-off for hosts that own a raw NCCL communicator (the Python host side uses torch.distributed's, see
// mmpl_b200/segment_parallel.py). NCCL is resolved at run time from the process image (torch or the host application has
// loaded it), falling back to dlopen("libnccl.so.2"): the library keeps loading on a machine without NCCL.
int mmpl_anchor_broadcast(void* nccl_comm, void* buf, int64_t bytes, int root, void* stream) {
  MMPL_CHECK(device_is_sm100(), MMPL_ERR_ARCH, "anchor_broadcast: requires an sm_100 device");
  MMPL_CHECK(nccl_comm != nullptr && buf != nullptr && bytes > 0 && root >= 0, MMPL_ERR_ARG, "anchor_broadcast: bad argument");
  using BcastFn = int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t);
  using ErrFn = const char* (*)(int);
  static BcastFn bcast = nullptr;
  static ErrFn errstr = nullptr;
  if (!bcast) {
    void* sym = dlsym(RTLD_DEFAULT, "ncclBroadcast");
    void* handle = nullptr;
    if (!sym && (handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL)) != nullptr) sym = dlsym(handle, "ncclBroadcast");
    MMPL_CHECK(sym != nullptr, MMPL_ERR_STATE, "anchor_broadcast: NCCL is not loaded in this process (ncclBroadcast not found)");
    bcast = reinterpret_cast<BcastFn>(sym);
    errstr = reinterpret_cast<ErrFn>(dlsym(handle ? handle : RTLD_DEFAULT, "ncclGetErrorString"));
  }
  constexpr int kNcclUint8 = 1;  // ncclDataType_t: ncclInt8 = 0, ncclUint8 = 1
  const int rc = bcast(buf, buf, static_cast<size_t>(bytes), kNcclUint8, root, nccl_comm, static_cast<cudaStream_t>(stream));
  MMPL_CHECK(rc == 0, MMPL_ERR_CUDA, "anchor_broadcast: ncclBroadcast failed: %s", errstr ? errstr(rc) : "unknown");
  return MMPL_OK;
}

int mmpl_flash_attn(const void* q, int64_t ldq, int Lq, int H, const void* k0, const void* v0, int64_t ldkv0,
                    int rows0, const void* k1, const void* v1, int64_t ldkv1, int rows1, int nseg,
                    const int* seg_start, const int* seg_rows, const int* seg_src, void* out, int64_t ldo,
                    float softmax_scale, void* stream) {
  MMPL_CHECK(q && k0 && v0 && out && seg_start && seg_rows, MMPL_ERR_ARG, "flash_attn: null argument");
  COUNTED(flash_attn_bf16(q, ldq, Lq, H, k0, v0, ldkv0, rows0, k1, v1, ldkv1, rows1, nseg, seg_start, seg_rows,
                         seg_src, out, ldo, softmax_scale, static_cast<cudaStream_t>(stream)));
}

int mmpl_gemm_set_streamk(int mode) {
  gemm_set_streamk(mode);
  return MMPL_OK;
}

int mmpl_attn_set_split(int split) {
  flash_attn_force_split(split);
  return MMPL_OK;
}

int mmpl_attn_set_ctas(int ctas) {
  flash_attn_force_ctas(ctas);
  return MMPL_OK;
}

int mmpl_attn_plan(int Lq, int H, int kv_tiles, int ctas, int force_split, int* sched, int* pieces, int max_pieces) {
  if (Lq <= 0 || H <= 0 || kv_tiles <= 0 || ctas <= 0 || !pieces || max_pieces <= 0) return MMPL_ERR_ARG;
  return flash_attn_plan(Lq, H, kv_tiles, ctas, force_split, sched, pieces, max_pieces);
}

int mmpl_ln_modulate(const void* x, int64_t ldx, void* out, int64_t ldo, int S, int D, float eps, const void* shift,
                     const void* scale, int64_t mod_stride, int rows_per_frame, void* stream) {
  COUNTED(ln_modulate(x, ldx, out, ldo, S, D, eps, shift, scale, mod_stride, rows_per_frame,
                     static_cast<cudaStream_t>(stream)));
}
int mmpl_ln_affine(const void* x, int64_t ldx, void* out, int64_t ldo, int S, int D, float eps, const void* weight,
                   const void* bias, void* stream) {
  COUNTED(ln_affine(x, ldx, out, ldo, S, D, eps, weight, bias, static_cast<cudaStream_t>(stream)));
}
int mmpl_rmsnorm(const void* x, int64_t ldx, void* out, int64_t ldo, int S, int D, const void* weight, float eps,
                 void* stream) {
  COUNTED(rmsnorm(x, ldx, out, ldo, S, D, weight, eps, static_cast<cudaStream_t>(stream)));
}
int mmpl_qk_norm_rope_kv(const void* q_in, const void* k_in, const void* v_in, int64_t ld_in,
                         const void* norm_q_weight, const void* norm_k_weight, const void* rope_table, void* q_out,
                         int64_t ldq, void* k_dst, void* v_dst, int64_t ldkv, int S, int D, int gh, int gw,
                         int n_frames, const int* frame_pos, const int* kv_row, float eps, void* stream) {
  MMPL_CHECK(q_in && k_in && v_in && norm_q_weight && norm_k_weight && rope_table && q_out && k_dst && v_dst &&
                 frame_pos && kv_row,
             MMPL_ERR_ARG, "qk_norm_rope_kv: null argument");
  COUNTED(qk_norm_rope_kv(q_in, k_in, v_in, ld_in, norm_q_weight, norm_k_weight, rope_table, q_out, ldq, k_dst, v_dst,
                         ldkv, S, D, gh, gw, n_frames, frame_pos, kv_row, eps, static_cast<cudaStream_t>(stream)));
}
int mmpl_modulation_add(const void* mod, const void* src, int64_t src_fstride, int64_t src_jstride, void* out, int F,
                        int J, int D, void* stream) {
  COUNTED(modulation_add(mod, src, src_fstride, src_jstride, out, F, J, D, static_cast<cudaStream_t>(stream)));
}
int mmpl_sinusoid_embedding(const double* t, void* out, int F, int dim, void* stream) {
  COUNTED(sinusoid_embedding(t, out, F, dim, static_cast<cudaStream_t>(stream)));
}
int mmpl_skinny_linear(const void* x, int64_t ldx, const void* w, const void* b, void* out, int64_t ldo, int M, int N,
                       int K, int silu_in, int silu_out, void* stream) {
  COUNTED(skinny_linear(x, ldx, w, b, out, ldo, M, N, K, silu_in, silu_out, static_cast<cudaStream_t>(stream)));
}
int mmpl_patchify(const void* x, int64_t stride_f, int64_t stride_c, void* a, int F, int C, int H, int W,
                  void* stream) {
  COUNTED(patchify(x, stride_f, stride_c, a, F, C, H, W, static_cast<cudaStream_t>(stream)));
}
int mmpl_unpatchify_x0(const void* head, int64_t ldh, const void* xt, int64_t xt_stride_f, int64_t xt_stride_c,
                       const double* sigma, void* flow, void* x0, int F, int C, int H, int W, void* stream) {
  COUNTED(unpatchify_x0(head, ldh, xt, xt_stride_f, xt_stride_c, sigma, flow, x0, F, C, H, W,
                       static_cast<cudaStream_t>(stream)));
}
int mmpl_add_noise(const void* x0, const void* noise, const float* sigma, void* out, int n_frames, int64_t per_frame,
                   void* stream) {
  COUNTED(add_noise(x0, noise, sigma, out, n_frames, per_frame, static_cast<cudaStream_t>(stream)));
}

// ------------------------------------------------------------------------------------------------ context
int mmpl_ctx_create(const mmpl_model_config* cfg, mmpl_ctx** out) {
  MMPL_CHECK(cfg && out, MMPL_ERR_ARG, "ctx_create: null argument");
  MMPL_CHECK(device_is_sm100(), MMPL_ERR_ARCH, "ctx_create: requires an sm_100 device");
  MMPL_CHECK(cfg->dim > 0 && cfg->dim % 256 == 0 && cfg->num_heads * 128 == cfg->dim, MMPL_ERR_SHAPE,
             "ctx_create: dim=%d must be num_heads(%d)*128 and a multiple of 256", cfg->dim, cfg->num_heads);
  MMPL_CHECK(cfg->ffn_dim % 8 == 0 && cfg->num_layers > 0 && cfg->max_tokens > 0 && cfg->text_len > 0 &&
                 cfg->text_dim % 8 == 0 && cfg->freq_dim % 8 == 0,
             MMPL_ERR_SHAPE, "ctx_create: bad config");
  mmpl_ctx* c = new mmpl_ctx;
  c->cfg = *cfg;
  c->layers.resize(cfg->num_layers);
  const size_t D = cfg->dim, Fd = cfg->ffn_dim, S = cfg->max_tokens, T = cfg->text_len;
  const size_t pk = static_cast<size_t>(cfg->in_dim) * 4, po = static_cast<size_t>(cfg->out_dim) * 4;
  auto al = [](size_t b) { return (b + 1023) & ~size_t(1023); };
  const size_t sizes[] = {
      al(S * D * 2),           // x
      al(S * D * 2),           // xm
      al(S * 3 * D * 2),       // qkv
      al(S * D * 2),           // attn
      al(S * Fd * 2),          // h
      al(S * pk * 2),          // patch
      al(S * po * 2),          // hout
      al(32 * cfg->freq_dim * 2), al(32 * D * 2), al(32 * D * 2), al(32 * 6 * D * 2),  // sinus t1 e e0
      al(static_cast<size_t>(cfg->num_layers) * 32 * 6 * D * 2), al(32 * 2 * D * 2),   // emod (all blocks) ehead
      al(T * D * 2), al(T * D * 2),                                                    // ctx_h ctx_e
      al(S * D * 2), al(S * D * 2),                                                    // tail_k tail_v
      al(static_cast<size_t>(cfg->num_layers) * sizeof(void*)),                        // mod_table
  };
  size_t total = 0;
  for (size_t s : sizes) total += s;
  cudaError_t e = cudaMalloc(&c->ws, total);
  if (e != cudaSuccess) {
    set_error("ctx_create: cudaMalloc(%zu) failed: %s", total, cudaGetErrorString(e));
    delete c;
    return MMPL_ERR_CUDA;
  }
  c->ws_bytes = total;
  void** slots[] = {&c->x, &c->xm, &c->qkv, &c->attn, &c->h, &c->patch, &c->hout, &c->sinus, &c->t1,
                    &c->e, &c->e0, &c->emod, &c->ehead, &c->ctx_h, &c->ctx_e, &c->tail_k, &c->tail_v, &c->mod_table};
  size_t off = 0;
  for (size_t i = 0; i < sizeof(sizes) / sizeof(sizes[0]); ++i) {
    *slots[i] = c->ws + off;
    off += sizes[i];
  }
  *out = c;
  return MMPL_OK;
}

void mmpl_ctx_destroy(mmpl_ctx* ctx) {
  if (!ctx) return;
  if (ctx->ws) cudaFree(ctx->ws);
  delete ctx;
}

int mmpl_bind_weight(mmpl_ctx* ctx, const char* name, const void* ptr, int64_t numel) {
  MMPL_CHECK(ctx && name && ptr, MMPL_ERR_ARG, "bind_weight: null argument");
  MMPL_CHECK((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, MMPL_ERR_ARG, "bind_weight: %s is not 16-byte aligned", name);
  const int64_t D = ctx->cfg.dim, Fd = ctx->cfg.ffn_dim;
  const int64_t expect[] = {D * D, D, 3 * D * D, 3 * D, Fd * D, Fd, 6 * D, -1};
  if (std::strncmp(name, "blocks.", 7) == 0) {
    char* end = nullptr;
    const long li = std::strtol(name + 7, &end, 10);
    MMPL_CHECK(end && *end == '.' && li >= 0 && li < ctx->cfg.num_layers, MMPL_ERR_ARG, "bind_weight: bad layer in %s", name);
    const char* suffix = end + 1;
    for (const FieldDesc& f : kLayerFields) {
      if (std::strcmp(suffix, f.name) == 0) {
        MMPL_CHECK(expect[f.kind] < 0 || expect[f.kind] == numel, MMPL_ERR_SHAPE, "bind_weight: %s has %lld elements, expected %lld",
                   name, (long long)numel, (long long)expect[f.kind]);
        *reinterpret_cast<const void**>(reinterpret_cast<char*>(&ctx->layers[li]) + f.offset) = ptr;
        ctx->mod_table_dirty = true;
        return MMPL_OK;
      }
    }
  } else {
    for (const FieldDesc& f : kGlobalFields) {
      if (std::strcmp(name, f.name) == 0) {
        MMPL_CHECK(expect[f.kind] < 0 || expect[f.kind] == numel, MMPL_ERR_SHAPE, "bind_weight: %s has %lld elements, expected %lld",
                   name, (long long)numel, (long long)expect[f.kind]);
        *reinterpret_cast<const void**>(reinterpret_cast<char*>(&ctx->g) + f.offset) = ptr;
        return MMPL_OK;
      }
    }
  }
  set_error("bind_weight: unknown parameter name %s", name);
  return MMPL_ERR_ARG;
}

int mmpl_bind_rope_table(mmpl_ctx* ctx, const void* table) {
  MMPL_CHECK(ctx && table, MMPL_ERR_ARG, "bind_rope_table: null argument");
  ctx->rope_table = table;
  return MMPL_OK;
}

int64_t mmpl_launch_count(mmpl_ctx* ctx, int reset) {
  if (!ctx) return -1;
  const int64_t n = ctx->launches;
  if (reset) ctx->launches = 0;
  return n;
}

static cudaEvent_t prof_event(mmpl_ctx* ctx) {
  if (ctx->ev_used == ctx->ev_pool.size()) {
    cudaEvent_t e;
    cudaEventCreate(&e);
    ctx->ev_pool.push_back(e);
  }
  return ctx->ev_pool[ctx->ev_used++];
}

// RUN(category, algorithmic work, launcher call)
#define RUN(cat, work, call)                                                  \
  do {                                                                        \
    const bool _prof = (ctx->profile_mask >> (cat)) & 1;                      \
    size_t _i0 = 0;                                                           \
    if (_prof) { _i0 = ctx->ev_used; cudaEventRecord(prof_event(ctx), st); }  \
    const int _s = (call);                                                    \
    if (_s != MMPL_OK) return _s;                                             \
    if (_prof) {                                                              \
      const size_t _i1 = ctx->ev_used;                                        \
      cudaEventRecord(prof_event(ctx), st);                                   \
      ctx->ev_spans.push_back({(cat) | (ctx->cur_site << 8), {_i0, _i1}});    \
      ctx->prof_work[(cat)] += (work);                                        \
      ++ctx->prof_launches[(cat)];                                            \
      ctx->site_work[ctx->cur_site] += (work);                                \
      ++ctx->site_launches[ctx->cur_site];                                    \
    }                                                                         \
    ctx->cur_site = MMPL_SITE_OTHER;                                          \
    ++ctx->launches;                                                          \
    ++g_total_launches;                                                       \
  } while (0)

int64_t mmpl_workspace_generation(void) { return workspace_generation(); }

int mmpl_profile_mask(mmpl_ctx* ctx) { return ctx ? ctx->profile_mask : 0; }

int64_t mmpl_launch_credit(mmpl_ctx* ctx, int64_t n) {
  if (!ctx) return 0;
  ctx->launches += n;
  g_total_launches += n;
  return ctx->launches;
}

int mmpl_profile_enable(mmpl_ctx* ctx, int category_mask) {
  MMPL_CHECK(ctx, MMPL_ERR_ARG, "profile_enable: null context");
  ctx->profile_mask = category_mask;
  return MMPL_OK;
}

int mmpl_profile_read(mmpl_ctx* ctx, double* ms, double* work, int64_t* launches, int reset) {
  MMPL_CHECK(ctx && ms && work && launches, MMPL_ERR_ARG, "profile_read: null argument");
  for (auto& sp : ctx->ev_spans) {
    cudaEvent_t a = ctx->ev_pool[sp.second.first], b = ctx->ev_pool[sp.second.second];
    MMPL_CUDA(cudaEventSynchronize(b));
    float t = 0.f;
    MMPL_CUDA(cudaEventElapsedTime(&t, a, b));
    ctx->prof_ms[sp.first & 0xFF] += t;
    ctx->site_ms[sp.first >> 8] += t;
  }
  ctx->ev_spans.clear();
  ctx->ev_used = 0;
  for (int i = 0; i < MMPL_PROF_NCAT; ++i) {
    ms[i] = ctx->prof_ms[i];
    work[i] = ctx->prof_work[i];
    launches[i] = ctx->prof_launches[i];
    if (reset) { ctx->prof_ms[i] = 0; ctx->prof_work[i] = 0; ctx->prof_launches[i] = 0; }
  }
  return MMPL_OK;
}

int mmpl_profile_read_sites(mmpl_ctx* ctx, double* ms, double* work, int64_t* launches, int reset) {
  MMPL_CHECK(ctx && ms && work && launches, MMPL_ERR_ARG, "profile_read_sites: null argument");
  for (int i = 0; i < MMPL_SITE_COUNT; ++i) {
    ms[i] = ctx->site_ms[i];
    work[i] = ctx->site_work[i];
    launches[i] = ctx->site_launches[i];
    if (reset) { ctx->site_ms[i] = 0; ctx->site_work[i] = 0; ctx->site_launches[i] = 0; }
  }
  return MMPL_OK;
}

int64_t mmpl_total_launches(int reset) {
  const int64_t n = g_total_launches;
  if (reset) g_total_launches = 0;
  return n;
}

int mmpl_forward(mmpl_ctx* ctx, const mmpl_forward_args* a, void* stream_v) {
  MMPL_CHECK(ctx && a, MMPL_ERR_ARG, "forward: null argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream_v);
  const mmpl_model_config& c = ctx->cfg;
  const int D = c.dim, Fd = c.ffn_dim, H = c.num_heads, L = c.num_layers, T = c.text_len;
  const int F = a->n_frames, gh = a->lat_h / 2, gw = a->lat_w / 2, fs = gh * gw, S = F * fs;
  const int PK = c.in_dim * 4, PO = c.out_dim * 4;
  MMPL_CHECK(F > 0 && F <= 32 && a->lat_h % 2 == 0 && a->lat_w % 2 == 0 && S > 0 && S <= c.max_tokens, MMPL_ERR_SHAPE,
             "forward: %d frames of %dx%d latents = %d tokens exceeds workspace (%d) or frame limit (32)", F, a->lat_h,
             a->lat_w, S, c.max_tokens);
  MMPL_CHECK(a->latents && a->timesteps && a->kv_k && a->kv_v && a->frame_pos && a->kv_row && a->cross_k && a->cross_v &&
                 a->flow && (a->n_seg == 0 || (a->seg_start && a->seg_rows)),
             MMPL_ERR_ARG, "forward: null argument");
  MMPL_CHECK(a->cross_init || a->context, MMPL_ERR_ARG, "forward: context required to initialise the cross-attention cache");
  MMPL_CHECK(a->x0 == nullptr || a->sigma, MMPL_ERR_ARG, "forward: sigma required for the x0 output");
  MMPL_CHECK(a->n_seg >= 0 && a->n_seg + (a->kv_to_tail ? 1 : 0) <= 8 && a->n_seg + (a->kv_to_tail ? 1 : 0) >= 1, MMPL_ERR_SHAPE,
             "forward: bad segment count %d", a->n_seg);
  MMPL_CHECK(ctx->rope_table, MMPL_ERR_STATE, "forward: RoPE table not bound");
  {
    const void* const* gp = reinterpret_cast<const void* const*>(&ctx->g);
    for (size_t i = 0; i < sizeof(GlobalWeights) / sizeof(void*); ++i)
      MMPL_CHECK(gp[i], MMPL_ERR_STATE, "forward: parameter %s is not bound", kGlobalFields[i].name);
    for (int l = 0; l < L; ++l) {
      const void* const* lp = reinterpret_cast<const void* const*>(&ctx->layers[l]);
      for (size_t i = 0; i < sizeof(LayerWeights) / sizeof(void*); ++i)
        MMPL_CHECK(lp[i], MMPL_ERR_STATE, "forward: parameter blocks.%d.%s is not bound", l, kLayerFields[i].name);
    }
  }
  typedef __nv_bfloat16 bf;
  bf* x = static_cast<bf*>(ctx->x);
  bf* xm = static_cast<bf*>(ctx->xm);
  bf* qkv = static_cast<bf*>(ctx->qkv);
  bf* attn = static_cast<bf*>(ctx->attn);
  bf* hbuf = static_cast<bf*>(ctx->h);
  bf* e0 = static_cast<bf*>(ctx->e0);
  bf* ehead = static_cast<bf*>(ctx->ehead);
  const GlobalWeights& g = ctx->g;
  const float scale = 0.08838834764831845f;  // 1/sqrt(128)

  // patch embedding (causal_model.py:812-818)
  RUN(3, 4.0 * S * PK, patchify(a->latents, a->lat_stride_f, a->lat_stride_c, ctx->patch, F, c.in_dim, a->lat_h, a->lat_w, st));
  RUN(2, 2.0 * S * D * PK, gemm_bf16(ctx->patch, PK, g.patch_w, PK, g.patch_b, x, D, S, D, PK, MMPL_EPI_BIAS, nullptr, 0, nullptr, 0, 0, 0, st));
  // time embedding and projection (causal_model.py:828-831)
  RUN(3, 0.0, sinusoid_embedding(a->timesteps, ctx->sinus, F, c.freq_dim, st));
  RUN(3, 2.0 * D * c.freq_dim, skinny_linear(ctx->sinus, c.freq_dim, g.time0_w, g.time0_b, ctx->t1, D, F, D, c.freq_dim, 0, 1, st));
  RUN(3, 2.0 * D * D, skinny_linear(ctx->t1, D, g.time2_w, g.time2_b, ctx->e, D, F, D, D, 0, 0, st));
  RUN(3, 12.0 * D * D, skinny_linear(ctx->e, D, g.tproj_w, g.tproj_b, e0, 6 * D, F, 6 * D, D, 1, 0, st));
  // text embedding, only when the cross-attention K/V must be (re)computed (causal_model.py:836-841)
  if (!a->cross_init) {
    RUN(2, 2.0 * T * D * c.text_dim, gemm_bf16(a->context, c.text_dim, g.text0_w, c.text_dim, g.text0_b, ctx->ctx_h, D, T, D, c.text_dim,
                  MMPL_EPI_BIAS_GELU, nullptr, 0, nullptr, 0, 0, 0, st));
    RUN(2, 2.0 * T * D * D, gemm_bf16(ctx->ctx_h, D, g.text2_w, D, g.text2_b, ctx->ctx_e, D, T, D, D, MMPL_EPI_BIAS, nullptr, 0, nullptr, 0, 0, 0, st));
  }

  int seg_start[8], seg_rows[8], seg_src[8];
  int nseg = 0;
  for (int i = 0; i < a->n_seg; ++i, ++nseg) {
    seg_start[nseg] = a->seg_start[i];
    seg_rows[nseg] = a->seg_rows[i];
    seg_src[nseg] = 0;
  }
  if (a->kv_to_tail) {
    seg_start[nseg] = 0;
    seg_rows[nseg] = S;
    seg_src[nseg] = 1;
    ++nseg;
  }
  const int cross_start = 0, cross_rows = T;
  double lk_self = 0;
  for (int i = 0; i < nseg; ++i) lk_self += seg_rows[i];

  // e = modulation + e0 of every block (causal_model.py:300), one launch: e0 is fixed for the whole forward
  if (ctx->mod_table_dirty) {
    ctx->mod_table_host.resize(L);
    for (int l = 0; l < L; ++l) ctx->mod_table_host[l] = ctx->layers[l].modulation;
    MMPL_CUDA(cudaMemcpyAsync(ctx->mod_table, ctx->mod_table_host.data(), L * sizeof(void*), cudaMemcpyHostToDevice, st));
    ctx->mod_table_dirty = false;
  }
  ctx->cur_site = MMPL_SITE_MODADD;
  RUN(3, 0.0, modulation_add_layers(static_cast<const void* const*>(ctx->mod_table), e0, ctx->emod, L, F, D, st));

  for (int l = 0; l < L; ++l) {
    const LayerWeights& w = ctx->layers[l];
    bf* kc = static_cast<bf*>(a->kv_k[l]);
    bf* vc = static_cast<bf*>(a->kv_v[l]);
    bf* emod = static_cast<bf*>(ctx->emod) + static_cast<size_t>(l) * F * 6 * D;
    // self-attention (causal_model.py:304-310, 86-231)
    ctx->cur_site = MMPL_SITE_LN_MOD;
    RUN(3, 4.0 * S * D, ln_modulate(x, D, xm, D, S, D, c.eps, emod + 0 * D, emod + 1 * D, 6 * D, fs, st));
    ctx->cur_site = MMPL_SITE_GEMM_QKV;
    RUN(2, 6.0 * S * D * D, gemm_bf16(xm, D, w.qkv_w, D, w.qkv_b, qkv, 3 * D, S, 3 * D, D, MMPL_EPI_BIAS, nullptr, 0, nullptr, 0, 0, 0, st));
    ctx->cur_site = MMPL_SITE_ROPE_KV;
    RUN(3, 12.0 * S * D, qk_norm_rope_kv(qkv, qkv + D, qkv + 2 * D, 3 * D, w.norm_q, w.norm_k, ctx->rope_table, qkv, 3 * D,
                        a->kv_to_tail ? ctx->tail_k : kc, a->kv_to_tail ? ctx->tail_v : vc, D, S, D, gh, gw, F,
                        a->frame_pos, a->kv_row, c.eps, st));
    ctx->cur_site = MMPL_SITE_SELF_ATTN;
    RUN(0, 4.0 * S * lk_self * D, flash_attn_bf16(qkv, 3 * D, S, H, kc, vc, D, static_cast<int>(a->cache_rows), ctx->tail_k, ctx->tail_v, D, S,
                        nseg, seg_start, seg_rows, seg_src, attn, D, scale, st));
    ctx->cur_site = MMPL_SITE_GEMM_O;
    RUN(2, 2.0 * S * D * D, gemm_bf16(attn, D, w.o_w, D, w.o_b, x, D, S, D, D, MMPL_EPI_BIAS_GATE_RES, x, D, emod + 2 * D, 6 * D, fs, 0, st));
    // cross-attention (causal_model.py:314; model.py:159-194)
    ctx->cur_site = MMPL_SITE_LN_AFFINE;
    RUN(3, 4.0 * S * D, ln_affine(x, D, xm, D, S, D, c.eps, w.norm3_w, w.norm3_b, st));
    ctx->cur_site = MMPL_SITE_GEMM_CQ;
    RUN(2, 2.0 * S * D * D, gemm_bf16(xm, D, w.cq_w, D, w.cq_b, qkv, D, S, D, D, MMPL_EPI_BIAS, nullptr, 0, nullptr, 0, 0, 0, st));
    ctx->cur_site = MMPL_SITE_RMSNORM;
    RUN(3, 4.0 * S * D, rmsnorm(qkv, D, qkv, D, S, D, w.cnorm_q, c.eps, st));
    if (!a->cross_init) {
      RUN(2, 2.0 * T * D * D, gemm_bf16(ctx->ctx_e, D, w.ck_w, D, w.ck_b, a->cross_k[l], D, T, D, D, MMPL_EPI_BIAS, nullptr, 0, nullptr, 0, 0, 0, st));
      RUN(3, 4.0 * T * D, rmsnorm(a->cross_k[l], D, a->cross_k[l], D, T, D, w.cnorm_k, c.eps, st));
      RUN(2, 2.0 * T * D * D, gemm_bf16(ctx->ctx_e, D, w.cv_w, D, w.cv_b, a->cross_v[l], D, T, D, D, MMPL_EPI_BIAS, nullptr, 0, nullptr, 0, 0, 0, st));
    }
    ctx->cur_site = MMPL_SITE_CROSS_ATTN;
    RUN(1, 4.0 * S * T * D, flash_attn_bf16(qkv, D, S, H, a->cross_k[l], a->cross_v[l], D, T, nullptr, nullptr, 0, 0, 1, &cross_start,
                        &cross_rows, nullptr, attn, D, scale, st));
    ctx->cur_site = MMPL_SITE_GEMM_CO;
    RUN(2, 2.0 * S * D * D, gemm_bf16(attn, D, w.co_w, D, w.co_b, x, D, S, D, D, MMPL_EPI_BIAS_RES, x, D, nullptr, 0, 0, 0, st));
    // feed-forward (causal_model.py:316-322)
    ctx->cur_site = MMPL_SITE_LN_MOD;
    RUN(3, 4.0 * S * D, ln_modulate(x, D, xm, D, S, D, c.eps, emod + 3 * D, emod + 4 * D, 6 * D, fs, st));
    ctx->cur_site = MMPL_SITE_GEMM_FFN0;
    RUN(2, 2.0 * S * D * Fd, gemm_bf16(xm, D, w.ffn0_w, D, w.ffn0_b, hbuf, Fd, S, Fd, D, MMPL_EPI_BIAS_GELU, nullptr, 0, nullptr, 0, 0, 0, st));
    ctx->cur_site = MMPL_SITE_GEMM_FFN2;
    RUN(2, 2.0 * S * D * Fd, gemm_bf16(hbuf, Fd, w.ffn2_w, Fd, w.ffn2_b, x, D, S, D, Fd, MMPL_EPI_BIAS_GATE_RES, x, D, emod + 5 * D, 6 * D, fs, 0, st));
  }

  // head + unpatchify (+ flow -> x0) (causal_model.py:346-357, 889-892, 1094-1117; wan_wrapper.py:172-196)
  RUN(3, 0.0, modulation_add(g.head_mod, ctx->e, D, 0, ehead, F, 2, D, st));
  RUN(3, 4.0 * S * D, ln_modulate(x, D, xm, D, S, D, c.eps, ehead, ehead + D, 2 * D, fs, st));
  RUN(2, 2.0 * S * D * PO, gemm_bf16(xm, D, g.head_w, D, g.head_b, ctx->hout, PO, S, PO, D, MMPL_EPI_BIAS, nullptr, 0, nullptr, 0, 0, 0, st));
  RUN(3, 6.0 * S * PO, unpatchify_x0(ctx->hout, PO, a->latents, a->lat_stride_f, a->lat_stride_c, a->sigma, a->flow, a->x0, F, c.out_dim,
                    a->lat_h, a->lat_w, st));
  return MMPL_OK;
}

}  // extern "C"
