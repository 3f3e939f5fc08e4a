// HBM-bound kernels of the CausalWanModel hot path: LayerNorm + adaLN modulation, RMSNorm(q,k) +
// 3-D RoPE + KV-cache append, modulation vectors, time embedding, patchify/unpatchify and the
// flow->x0 / add_noise arithmetic. One warp per token row, 16-byte vector accesses, rows kept in
// registers as packed bf16. Every kernel rounds to bf16 exactly where the reference does (SURVEY.md
// appendix A) so that its output is bit-identical to the reference op given identical inputs, up to
// fp32 reduction order.
#include "host_util.h"
#include "mmpl_b200.h"
#include "ptx.cuh"

namespace mmpl {

constexpr int kRowsPerBlock = 8;  // warps per CTA
constexpr int kMaxFrames = 32;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
  const uint32_t* w = reinterpret_cast<const uint32_t*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[2 * i] = bf16_lo(w[i]);
    f[2 * i + 1] = bf16_hi(w[i]);
  }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  return make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]),
                    pack_bf16x2(f[6], f[7]));
}

// Rounds two fp32 values to bf16 and back with ONE conversion instruction (cvt.rn.bf16x2.f32) plus two integer ops,
// instead of one conversion each: the HBM-bound kernels of this file were bound by the conversion pipe, not by memory
// (ncu: pipe_xu 55-68 % of peak at 2-3 TB/s), because every rounding point of the reference costs a conversion.
__device__ __forceinline__ void round2(float& a, float& b) {
  const uint32_t p = pack_bf16x2(a, b);
  a = bf16_lo(p);
  b = bf16_hi(p);
}

// fp32 -> fp64 without the conversion pipe: re-bias the exponent with integer ops (exact for every normal float;
// zeros, subnormals, infinities and NaNs take the conversion instruction).
__device__ __forceinline__ double widen(float f) {
  const uint32_t u = __float_as_uint(f);
  if (((u >> 23) & 0xFFu) - 1u >= 254u) return static_cast<double>(f);
  return __hiloint2double(static_cast<int>((u & 0x80000000u) | (((u & 0x7FFFFFFFu) >> 3) + 0x38000000u)),
                          static_cast<int>(u << 29));
}

// ------------------------------------------------------------------------------------------------
// LayerNorm (eps, no affine) followed by adaLN modulation, or LayerNorm with affine weight/bias.
//   modulated (causal_model.py:305,318): out = bf16( bf16( bf16(LN(x)) * bf16(1 + scale_f) ) + shift_f )
//   affine    (norm3, causal_model.py:314; model.py:89-99): out = bf16( LN(x) * w + b )
// NCH = D / 256 (16-byte chunks per lane).
template <int NCH, bool AFFINE>
__global__ void __launch_bounds__(kRowsPerBlock * 32, NCH <= 8 ? 4 : 1)  // D <= 2048: 64 registers, S = 4680 rows in one wave
ln_kernel(const __nv_bfloat16* __restrict__ x, int64_t ldx, __nv_bfloat16* __restrict__ out, int64_t ldo,
          int S, float eps, const __nv_bfloat16* __restrict__ shift, const __nv_bfloat16* __restrict__ scale,
          int64_t mod_stride, int rows_per_frame) {
  pdl_wait();  // PDL: the previous kernel in the stream has completed, its writes are visible
  pdl_launch_dependents();
  constexpr int D = NCH * 256;
  // The modulation (shift, scale of the CTA's frame) or affine (weight, bias) vectors are the same for every row of the
  // CTA - a frame is hundreds of rows, a CTA eight - so they are staged once per CTA in shared memory instead of being
  // held per warp in 2 x NCH x 4 registers: that was what kept the kernel at 2-3 CTAs per SM and two waves at S = 4680.
  // A row of another frame than the CTA's first (only when rows_per_frame is not a multiple of 8) reads them from global.
  __shared__ uint4 s_a[NCH * 32], s_b[NCH * 32];
  const int row0 = blockIdx.x * kRowsPerBlock;
  const int frame0 = row0 / rows_per_frame;
  {
    const uint4* a0 = reinterpret_cast<const uint4*>(AFFINE ? scale : scale + static_cast<int64_t>(frame0) * mod_stride);
    const uint4* b0 = reinterpret_cast<const uint4*>(AFFINE ? shift : shift + static_cast<int64_t>(frame0) * mod_stride);
    for (int i = threadIdx.x; i < NCH * 32; i += kRowsPerBlock * 32) {
      s_a[i] = __ldg(a0 + i);
      s_b[i] = __ldg(b0 + i);
    }
  }
  const int row = row0 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  uint4 v[NCH];
  if (row < S) {
    const uint4* xr = reinterpret_cast<const uint4*>(x + static_cast<int64_t>(row) * ldx);
#pragma unroll
    for (int i = 0; i < NCH; ++i) v[i] = xr[lane + 32 * i];
  }
  __syncthreads();
  if (row >= S) return;
  const int frame = row / rows_per_frame;
  const bool staged = AFFINE || frame == frame0;
  const uint4* a_ptr = reinterpret_cast<const uint4*>(AFFINE ? scale : scale + static_cast<int64_t>(frame) * mod_stride);
  const uint4* b_ptr = reinterpret_cast<const uint4*>(AFFINE ? shift : shift + static_cast<int64_t>(frame) * mod_stride);
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    float f[8];
    unpack8(v[i], f);
#pragma unroll
    for (int j = 0; j < 8; ++j) sum += f[j];
  }
  const float mean = warp_sum(sum) * (1.0f / D);
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    float f[8];
    unpack8(v[i], f);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float d = f[j] - mean;
      sq += d * d;
    }
  }
  const float rstd = rsqrtf(warp_sum(sq) * (1.0f / D) + eps);
  uint4* orow = reinterpret_cast<uint4*>(out + static_cast<int64_t>(row) * ldo);
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    float f[8], a[8], b[8];
    unpack8(v[i], f);
    unpack8(staged ? s_a[lane + 32 * i] : __ldg(a_ptr + lane + 32 * i), a);
    unpack8(staged ? s_b[lane + 32 * i] : __ldg(b_ptr + lane + 32 * i), b);
    if (AFFINE) {
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = __fadd_rn(__fmul_rn((f[j] - mean) * rstd, a[j]), b[j]);
    } else {
#pragma unroll
      for (int j = 0; j < 8; j += 2) {  // two elements per conversion instruction at every rounding point
        float n0 = (f[j] - mean) * rstd, n1 = (f[j + 1] - mean) * rstd;
        float s0 = 1.0f + a[j], s1 = 1.0f + a[j + 1];
        round2(n0, n1);
        round2(s0, s1);
        float m0 = n0 * s0, m1 = n1 * s1;
        round2(m0, m1);
        f[j] = m0 + b[j];
        f[j + 1] = m1 + b[j + 1];
      }
    }
    orow[lane + 32 * i] = pack8(f);
  }
}

// ------------------------------------------------------------------------------------------------
// WanRMSNorm (model.py:70-86): bf16( bf16( x * rsqrt(mean(x^2 over the full D) + eps) ) * w ).
template <int NCH>
__device__ __forceinline__ void rmsnorm_row(uint4 (&v)[NCH], const __nv_bfloat16* __restrict__ w, int lane, float eps) {
  constexpr int D = NCH * 256;
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    float f[8];
    unpack8(v[i], f);
#pragma unroll
    for (int j = 0; j < 8; ++j) sq += f[j] * f[j];
  }
  const float r = rsqrtf(warp_sum(sq) * (1.0f / D) + eps);
  const uint4* wp = reinterpret_cast<const uint4*>(w);
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    float f[8], g[8];
    unpack8(v[i], f);
    unpack8(__ldg(wp + lane + 32 * i), g);
#pragma unroll
    for (int j = 0; j < 8; j += 2) {
      float t0 = f[j] * r, t1 = f[j + 1] * r;
      round2(t0, t1);
      f[j] = t0 * g[j];          // rounded by the pack below
      f[j + 1] = t1 * g[j + 1];
    }
    v[i] = pack8(f);
  }
}

template <int NCH>
__global__ void __launch_bounds__(kRowsPerBlock * 32)
rmsnorm_kernel(const __nv_bfloat16* __restrict__ x, int64_t ldx, __nv_bfloat16* __restrict__ out, int64_t ldo,
               int S, const __nv_bfloat16* __restrict__ w, float eps) {
  pdl_wait();  // PDL: the previous kernel in the stream has completed, its writes are visible
  pdl_launch_dependents();
  const int row = blockIdx.x * kRowsPerBlock + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= S) return;
  const uint4* xr = reinterpret_cast<const uint4*>(x + static_cast<int64_t>(row) * ldx);
  uint4 v[NCH];
#pragma unroll
  for (int i = 0; i < NCH; ++i) v[i] = xr[lane + 32 * i];
  rmsnorm_row<NCH>(v, w, lane, eps);
  uint4* orow = reinterpret_cast<uint4*>(out + static_cast<int64_t>(row) * ldo);
#pragma unroll
  for (int i = 0; i < NCH; ++i) orow[lane + 32 * i] = v[i];
}

// ------------------------------------------------------------------------------------------------
// Self-attention pre-processing (causal_model.py:111-113,193-217): RMSNorm(q), RMSNorm(k), 3-D RoPE
// on both (causal_rope_apply :27-55, float64 complex multiply, result rounded double->float->bf16 as
// torch does), roped q -> q_out, roped k and v -> their KV-cache rows.
// RoPE table: double [1024][64][2] (cos, sin) = the reference's freqs (22 temporal, 21 height,
// 21 width complex pairs per 128-wide head; model.py:29-36, causal_model.py:473-478).
struct RopeKVParams {
  int S, gh, gw;                 // tokens, latent grid height/width (tokens per frame = gh*gw)
  int frame_pos[kMaxFrames];     // temporal RoPE position of each frame in this call
  int kv_row[kMaxFrames];        // destination row (in k_dst/v_dst) of each frame's first token
  float eps;
};

template <int NCH>
__device__ __forceinline__ void rope_row(uint4 (&v)[NCH], const double2* __restrict__ tab, int lane, int pt, int ph, int pw) {
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    const int chunk = lane + 32 * i;           // 8 elements = 4 complex pairs, inside one head
    const int pair0 = (chunk & 15) * 4;        // pair index within the head, 0..63
    float f[8];
    unpack8(v[i], f);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int pi = pair0 + j;
      const int pos = pi < 22 ? pt : (pi < 43 ? ph : pw);
      const double2 cs = __ldg(tab + pos * 64 + pi);
      const double re = widen(f[2 * j]);
      const double im = widen(f[2 * j + 1]);
      const double ore = __dsub_rn(__dmul_rn(re, cs.x), __dmul_rn(im, cs.y));
      const double oim = __dadd_rn(__dmul_rn(re, cs.y), __dmul_rn(im, cs.x));
      f[2 * j] = __double2float_rn(ore);
      f[2 * j + 1] = __double2float_rn(oim);
    }
    v[i] = pack8(f);
  }
}

// blockIdx.y selects the operand (0 = q, 1 = k, 2 = v), so a warp holds one row of one operand: three times the warps
// of a row-per-warp layout at a third of the registers, which is what hides the load -> fp64 rotate -> store chain
// of this single-wave kernel.
template <int NCH>
__global__ void __launch_bounds__(kRowsPerBlock * 32, NCH <= 8 ? 4 : 1)
qk_norm_rope_kv_kernel(const __nv_bfloat16* __restrict__ q_in, const __nv_bfloat16* __restrict__ k_in,
                       const __nv_bfloat16* __restrict__ v_in, int64_t ld_in,
                       const __nv_bfloat16* __restrict__ wq, const __nv_bfloat16* __restrict__ wk,
                       const double2* __restrict__ rope_tab, __nv_bfloat16* __restrict__ q_out, int64_t ldq,
                       __nv_bfloat16* __restrict__ k_dst, __nv_bfloat16* __restrict__ v_dst, int64_t ldkv,
                       const RopeKVParams p) {
  pdl_wait();  // PDL: the previous kernel in the stream has completed, its writes are visible
  pdl_launch_dependents();
  const int row = blockIdx.x * kRowsPerBlock + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  const int role = blockIdx.y;
  if (row >= p.S) return;
  const int fs = p.gh * p.gw;
  const int fr = row / fs;
  const int rem = row - fr * fs;
  const int64_t dst_row = static_cast<int64_t>(p.kv_row[fr]) + rem;
  const __nv_bfloat16* src = (role == 0 ? q_in : (role == 1 ? k_in : v_in)) + static_cast<int64_t>(row) * ld_in;
  __nv_bfloat16* dst = role == 0 ? q_out + static_cast<int64_t>(row) * ldq : (role == 1 ? k_dst : v_dst) + dst_row * ldkv;
  uint4 v[NCH];
  const uint4* r = reinterpret_cast<const uint4*>(src);
#pragma unroll
  for (int i = 0; i < NCH; ++i) v[i] = r[lane + 32 * i];
  if (role < 2) {
    const int ph = rem / p.gw;
    const int pw = rem - ph * p.gw;
    rmsnorm_row<NCH>(v, role == 0 ? wq : wk, lane, p.eps);
    rope_row<NCH>(v, rope_tab, lane, p.frame_pos[fr], ph, pw);
  }
  uint4* o = reinterpret_cast<uint4*>(dst);
#pragma unroll
  for (int i = 0; i < NCH; ++i) o[lane + 32 * i] = v[i];
}

// ------------------------------------------------------------------------------------------------
// out[f][j][d] = bf16(mod[j][d] + src[f*src_fstride + j*src_jstride + d])   (causal_model.py:300,355)
__global__ void modulation_add_kernel(const __nv_bfloat16* __restrict__ mod, const __nv_bfloat16* __restrict__ src,
                                      int64_t src_fstride, int64_t src_jstride, __nv_bfloat16* __restrict__ out,
                                      int F, int J, int D) {
  pdl_wait();  // PDL: the previous kernel in the stream has completed, its writes are visible
  pdl_launch_dependents();
  const int64_t n = static_cast<int64_t>(F) * J * D;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int d = static_cast<int>(i % D);
    const int j = static_cast<int>((i / D) % J);
    const int f = static_cast<int>(i / (static_cast<int64_t>(D) * J));
    const float a = __bfloat162float(mod[static_cast<int64_t>(j) * D + d]);
    const float b = __bfloat162float(src[f * src_fstride + j * src_jstride + d]);
    out[i] = __float2bfloat16_rn(a + b);
  }
}

// The same for every block of the model in one launch: out[l][f][j][d] = bf16(mods[l][j][d] + src[f][j][d]) with
// J = 6 (causal_model.py:300 evaluated for all CausalWanAttentionBlocks up front; e0 does not change inside a
// forward). `mods` is a device table of the blocks' modulation parameters. blockIdx.y = block index.
__global__ void modulation_add_layers_kernel(const __nv_bfloat16* const* __restrict__ mods,
                                             const __nv_bfloat16* __restrict__ src, __nv_bfloat16* __restrict__ out,
                                             int F, int JD) {
  pdl_wait();
  pdl_launch_dependents();
  const __nv_bfloat16* mod = mods[blockIdx.y];
  __nv_bfloat16* o = out + static_cast<int64_t>(blockIdx.y) * F * JD;
  const int n8 = F * (JD / 8);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += gridDim.x * blockDim.x) {
    const int f = i / (JD / 8);
    const int c = (i - f * (JD / 8)) * 8;
    const uint4 a = *reinterpret_cast<const uint4*>(mod + c);
    const uint4 b = *reinterpret_cast<const uint4*>(src + static_cast<int64_t>(f) * JD + c);
    const uint32_t* aw = reinterpret_cast<const uint32_t*>(&a);
    const uint32_t* bw = reinterpret_cast<const uint32_t*>(&b);
    uint4 r;
    uint32_t* rw = reinterpret_cast<uint32_t*>(&r);
#pragma unroll
    for (int k = 0; k < 4; ++k) rw[k] = pack_bf16x2(bf16_lo(aw[k]) + bf16_lo(bw[k]), bf16_hi(aw[k]) + bf16_hi(bw[k]));
    *reinterpret_cast<uint4*>(o + static_cast<int64_t>(f) * JD + c) = r;
  }
}

// ------------------------------------------------------------------------------------------------
// sinusoidal_embedding_1d (model.py:15-25) in float64, cast to bf16 via float: out[f][0:half]=cos, [half:]=sin
__global__ void sinusoid_kernel(const double* __restrict__ t, __nv_bfloat16* __restrict__ out, int F, int dim) {
  pdl_wait();  // PDL: the previous kernel in the stream has completed, its writes are visible
  pdl_launch_dependents();
  const int half = dim / 2;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= F * half) return;
  const int f = i / half, k = i - f * half;
  const double freq = pow(10000.0, -static_cast<double>(k) / static_cast<double>(half));
  const double a = t[f] * freq;
  out[static_cast<int64_t>(f) * dim + k] = __float2bfloat16_rn(__double2float_rn(cos(a)));
  out[static_cast<int64_t>(f) * dim + half + k] = __float2bfloat16_rn(__double2float_rn(sin(a)));
}

// ------------------------------------------------------------------------------------------------
// Skinny Linear for the time embedding / projection (M <= 32 rows): one warp per output column,
//   out[m][n] = act_out( bf16( sum_k act_in(x[m][k]) * w[n][k] + b[n] ) ),  act = identity | SiLU (-> bf16)
// (causal_model.py:828-831: time_embedding = Linear-SiLU-Linear, time_projection = SiLU-Linear)
template <int MT>
__global__ void __launch_bounds__(256)
skinny_linear_kernel(const __nv_bfloat16* __restrict__ x, int64_t ldx, const __nv_bfloat16* __restrict__ w,
                     const __nv_bfloat16* __restrict__ b, __nv_bfloat16* __restrict__ out, int64_t ldo,
                     int M, int N, int K, int silu_in, int silu_out) {
  pdl_wait();  // PDL: the previous kernel in the stream has completed, its writes are visible
  pdl_launch_dependents();
  const int n = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (n >= N) return;
  const uint4* wr = reinterpret_cast<const uint4*>(w + static_cast<int64_t>(n) * K);
  for (int m0 = 0; m0 < M; m0 += MT) {
    float acc[MT];
#pragma unroll
    for (int m = 0; m < MT; ++m) acc[m] = 0.f;
    for (int c = lane; c < K / 8; c += 32) {
      float wf[8];
      unpack8(__ldg(wr + c), wf);
#pragma unroll
      for (int m = 0; m < MT; ++m) {
        if (m0 + m < M) {
          float xf[8];
          unpack8(*reinterpret_cast<const uint4*>(x + static_cast<int64_t>(m0 + m) * ldx + c * 8), xf);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float xv = xf[j];
            if (silu_in) xv = bf16_round(xv / (1.0f + expf(-xv)));
            acc[m] = fmaf(xv, wf[j], acc[m]);
          }
        }
      }
    }
#pragma unroll
    for (int m = 0; m < MT; ++m) {
      const float s = warp_sum(acc[m]);
      if (lane == 0 && m0 + m < M) {
        float y = bf16_round(s + (b ? __bfloat162float(b[n]) : 0.f));
        if (silu_out) y = y / (1.0f + expf(-y));
        out[static_cast<int64_t>(m0 + m) * ldo + n] = __float2bfloat16_rn(y);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Patchify for patch (1,2,2): A[(f,hh,ww)][c*4 + dy*2 + dx] = x[f][c][2hh+dy][2ww+dx]
// (the im2col of patch_embedding Conv3d, causal_model.py:812; x is the [F,C,H,W] latent chunk)
__global__ void patchify_kernel(const __nv_bfloat16* __restrict__ x, int64_t stride_f, int64_t stride_c,
                                __nv_bfloat16* __restrict__ a, int F, int C, int H, int W) {
  pdl_wait();  // PDL: the previous kernel in the stream has completed, its writes are visible
  pdl_launch_dependents();
  const int gh = H / 2, gw = W / 2;
  const int K = C * 4;
  const int64_t n = static_cast<int64_t>(F) * gh * gw * K;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int col = static_cast<int>(i % K);
    const int64_t row = i / K;
    const int ww = static_cast<int>(row % gw);
    const int hh = static_cast<int>((row / gw) % gh);
    const int f = static_cast<int>(row / (static_cast<int64_t>(gw) * gh));
    const int c = col >> 2, dy = (col >> 1) & 1, dx = col & 1;
    a[i] = x[f * stride_f + c * stride_c + static_cast<int64_t>(2 * hh + dy) * W + (2 * ww + dx)];
  }
}

// Unpatchify (causal_model.py:1094-1117) fused with the flow->x0 conversion of
// WanDiffusionWrapper._convert_flow_pred_to_x0 (utils/wan_wrapper.py:172-196, float64):
//   flow[f][c][2hh+q][2ww+r] = head[(f,hh,ww)][(q*2+r)*C + c];  x0 = bf16(double(xt) - sigma_f*double(flow))
__global__ void unpatchify_x0_kernel(const __nv_bfloat16* __restrict__ head, int64_t ldh,
                                     const __nv_bfloat16* __restrict__ xt, int64_t xt_stride_f, int64_t xt_stride_c,
                                     const double* __restrict__ sigma, __nv_bfloat16* __restrict__ flow,
                                     __nv_bfloat16* __restrict__ x0, int F, int C, int H, int W) {
  pdl_wait();  // PDL: the previous kernel in the stream has completed, its writes are visible
  pdl_launch_dependents();
  const int gh = H / 2, gw = W / 2;
  const int64_t n = static_cast<int64_t>(F) * C * H * W;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int xw = static_cast<int>(i % W);
    const int yh = static_cast<int>((i / W) % H);
    const int c = static_cast<int>((i / (static_cast<int64_t>(W) * H)) % C);
    const int f = static_cast<int>(i / (static_cast<int64_t>(W) * H * C));
    const int64_t row = (static_cast<int64_t>(f) * gh + (yh >> 1)) * gw + (xw >> 1);
    const int col = ((yh & 1) * 2 + (xw & 1)) * C + c;
    const __nv_bfloat16 fl = head[row * ldh + col];
    flow[i] = fl;
    if (x0) {
      const double xv = static_cast<double>(__bfloat162float(
          xt[f * xt_stride_f + c * xt_stride_c + static_cast<int64_t>(yh) * W + xw]));
      const double r = __dsub_rn(xv, __dmul_rn(sigma[f], static_cast<double>(__bfloat162float(fl))));
      x0[i] = __float2bfloat16_rn(__double2float_rn(r));
    }
  }
}

// FlowMatchScheduler.add_noise (utils/scheduler.py:159-176): bf16( (1-sigma)*x0 + sigma*noise ) in fp32
__global__ void add_noise_kernel(const __nv_bfloat16* __restrict__ x0, const __nv_bfloat16* __restrict__ noise,
                                 const float* __restrict__ sigma, __nv_bfloat16* __restrict__ out,
                                 int64_t per_frame, int64_t n) {
  pdl_wait();  // PDL: the previous kernel in the stream has completed, its writes are visible
  pdl_launch_dependents();
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float s = sigma[i / per_frame];
    const float a = __fmul_rn(__fsub_rn(1.0f, s), __bfloat162float(x0[i]));
    const float b = __fmul_rn(s, __bfloat162float(noise[i]));
    out[i] = __float2bfloat16_rn(__fadd_rn(a, b));
  }
}

// ================================================================================================
// host launchers
// ================================================================================================
static inline int grid_for(int64_t n, int block) {
  int64_t g = (n + block - 1) / block;
  const int64_t cap = static_cast<int64_t>(sm_count() > 0 ? sm_count() : 148) * 16;
  return static_cast<int>(g < cap ? (g > 0 ? g : 1) : cap);
}

#define MMPL_DISPATCH_NCH(D, ...)                                                      \
  switch ((D) / 256) {                                                                 \
    case 6: { constexpr int NCH = 6; __VA_ARGS__; break; }   /* 1536 (Wan 1.3B) */     \
    case 20: { constexpr int NCH = 20; __VA_ARGS__; break; } /* 5120 (Wan 14B) */      \
    case 1: { constexpr int NCH = 1; __VA_ARGS__; break; }   /* 256 (tests) */         \
    case 2: { constexpr int NCH = 2; __VA_ARGS__; break; }   /* 512 (tests) */         \
    default: set_error("model dim %d not supported (256, 512, 1536, 5120)", (int)(D)); \
      return MMPL_ERR_SHAPE;                                                           \
  }

int ln_modulate(const void* x, int64_t ldx, void* out, int64_t ldo, int S, int D, float eps, const void* shift,
                const void* scale, int64_t mod_stride, int rows_per_frame, cudaStream_t st) {
  MMPL_CHECK(S > 0 && D % 256 == 0 && ldx % 8 == 0 && ldo % 8 == 0, MMPL_ERR_SHAPE, "ln_modulate: bad shape S=%d D=%d", S, D);
  MMPL_CHECK(shift && scale && rows_per_frame > 0 && mod_stride % 8 == 0, MMPL_ERR_ARG, "ln_modulate: bad modulation args");
  MMPL_CHECK((S + rows_per_frame - 1) / rows_per_frame <= kMaxFrames, MMPL_ERR_SHAPE, "ln_modulate: too many frames");
  const int grid = (S + kRowsPerBlock - 1) / kRowsPerBlock;
  MMPL_DISPATCH_NCH(D, (MMPL_CUDA_LAUNCH(launch_kernel(ln_kernel<NCH, false>, grid, kRowsPerBlock * 32, 0, st, 
                           static_cast<const __nv_bfloat16*>(x), ldx, static_cast<__nv_bfloat16*>(out), ldo, S, eps,
                           static_cast<const __nv_bfloat16*>(shift), static_cast<const __nv_bfloat16*>(scale),
                           mod_stride, rows_per_frame))));
  MMPL_CUDA(cudaGetLastError());
  return MMPL_OK;
}

int ln_affine(const void* x, int64_t ldx, void* out, int64_t ldo, int S, int D, float eps, const void* weight,
              const void* bias, cudaStream_t st) {
  MMPL_CHECK(S > 0 && D % 256 == 0 && ldx % 8 == 0 && ldo % 8 == 0, MMPL_ERR_SHAPE, "ln_affine: bad shape S=%d D=%d", S, D);
  MMPL_CHECK(weight && bias, MMPL_ERR_ARG, "ln_affine: weight and bias required");
  const int grid = (S + kRowsPerBlock - 1) / kRowsPerBlock;
  MMPL_DISPATCH_NCH(D, (MMPL_CUDA_LAUNCH(launch_kernel(ln_kernel<NCH, true>, grid, kRowsPerBlock * 32, 0, st, 
                           static_cast<const __nv_bfloat16*>(x), ldx, static_cast<__nv_bfloat16*>(out), ldo, S, eps,
                           static_cast<const __nv_bfloat16*>(bias), static_cast<const __nv_bfloat16*>(weight), 0, 1 << 30))));
  MMPL_CUDA(cudaGetLastError());
  return MMPL_OK;
}

int rmsnorm(const void* x, int64_t ldx, void* out, int64_t ldo, int S, int D, const void* weight, float eps,
            cudaStream_t st) {
  MMPL_CHECK(S > 0 && D % 256 == 0 && ldx % 8 == 0 && ldo % 8 == 0, MMPL_ERR_SHAPE, "rmsnorm: bad shape S=%d D=%d", S, D);
  const int grid = (S + kRowsPerBlock - 1) / kRowsPerBlock;
  MMPL_DISPATCH_NCH(D, (MMPL_CUDA_LAUNCH(launch_kernel(rmsnorm_kernel<NCH>, grid, kRowsPerBlock * 32, 0, st, 
                           static_cast<const __nv_bfloat16*>(x), ldx, static_cast<__nv_bfloat16*>(out), ldo, S,
                           static_cast<const __nv_bfloat16*>(weight), eps))));
  MMPL_CUDA(cudaGetLastError());
  return MMPL_OK;
}

int qk_norm_rope_kv(const void* q_in, const void* k_in, const void* v_in, int64_t ld_in, const void* wq,
                    const void* wk, const void* rope_table, void* q_out, int64_t ldq, void* k_dst, void* v_dst,
                    int64_t ldkv, int S, int D, int gh, int gw, int n_frames, const int* frame_pos,
                    const int* kv_row, float eps, cudaStream_t st) {
  MMPL_CHECK(S > 0 && D % 256 == 0 && D % 128 == 0, MMPL_ERR_SHAPE, "qk_norm_rope_kv: bad shape S=%d D=%d", S, D);
  MMPL_CHECK(gh > 0 && gw > 0 && n_frames > 0 && n_frames <= kMaxFrames && S == n_frames * gh * gw, MMPL_ERR_SHAPE,
             "qk_norm_rope_kv: S=%d must equal frames(%d)*gh(%d)*gw(%d), frames <= %d", S, n_frames, gh, gw, kMaxFrames);
  MMPL_CHECK(ld_in % 8 == 0 && ldq % 8 == 0 && ldkv % 8 == 0, MMPL_ERR_SHAPE, "qk_norm_rope_kv: strides must be multiples of 8");
  RopeKVParams p{};
  p.S = S; p.gh = gh; p.gw = gw; p.eps = eps;
  for (int f = 0; f < n_frames; ++f) {
    MMPL_CHECK(frame_pos[f] >= 0 && frame_pos[f] < 1024 && gh <= 1024 && gw <= 1024, MMPL_ERR_SHAPE,
               "qk_norm_rope_kv: RoPE position out of the 1024-row table");
    MMPL_CHECK(kv_row[f] >= 0, MMPL_ERR_ARG, "qk_norm_rope_kv: negative cache row");
    p.frame_pos[f] = frame_pos[f];
    p.kv_row[f] = kv_row[f];
  }
  const int grid = (S + kRowsPerBlock - 1) / kRowsPerBlock;
  MMPL_DISPATCH_NCH(D, (MMPL_CUDA_LAUNCH(launch_kernel(qk_norm_rope_kv_kernel<NCH>, dim3(grid, 3), kRowsPerBlock * 32, 0, st, 
                           static_cast<const __nv_bfloat16*>(q_in), static_cast<const __nv_bfloat16*>(k_in),
                           static_cast<const __nv_bfloat16*>(v_in), ld_in, static_cast<const __nv_bfloat16*>(wq),
                           static_cast<const __nv_bfloat16*>(wk), static_cast<const double2*>(rope_table),
                           static_cast<__nv_bfloat16*>(q_out), ldq, static_cast<__nv_bfloat16*>(k_dst),
                           static_cast<__nv_bfloat16*>(v_dst), ldkv, p))));
  MMPL_CUDA(cudaGetLastError());
  return MMPL_OK;
}

int modulation_add(const void* mod, const void* src, int64_t src_fstride, int64_t src_jstride, void* out, int F,
                   int J, int D, cudaStream_t st) {
  MMPL_CHECK(F > 0 && J > 0 && D > 0, MMPL_ERR_SHAPE, "modulation_add: bad shape");
  const int64_t n = static_cast<int64_t>(F) * J * D;
  MMPL_CUDA_LAUNCH(launch_kernel(modulation_add_kernel, grid_for(n, 256), 256, 0, st, static_cast<const __nv_bfloat16*>(mod),
                                                          static_cast<const __nv_bfloat16*>(src), src_fstride,
                                                          src_jstride, static_cast<__nv_bfloat16*>(out), F, J, D));
  MMPL_CUDA(cudaGetLastError());
  return MMPL_OK;
}

int modulation_add_layers(const void* const* mods_dev, const void* src, void* out, int L, int F, int D, cudaStream_t st) {
  MMPL_CHECK(L > 0 && F > 0 && D > 0 && D % 8 == 0, MMPL_ERR_SHAPE, "modulation_add_layers: bad shape");
  const int n8 = F * 6 * D / 8;
  MMPL_CUDA_LAUNCH(launch_kernel(modulation_add_layers_kernel, dim3((n8 + 255) / 256, L), 256, 0, st,
                                 reinterpret_cast<const __nv_bfloat16* const*>(mods_dev), static_cast<const __nv_bfloat16*>(src),
                                 static_cast<__nv_bfloat16*>(out), F, 6 * D));
  MMPL_CUDA(cudaGetLastError());
  return MMPL_OK;
}

int sinusoid_embedding(const double* t, void* out, int F, int dim, cudaStream_t st) {
  MMPL_CHECK(F > 0 && dim > 0 && dim % 2 == 0, MMPL_ERR_SHAPE, "sinusoid_embedding: bad shape");
  const int n = F * dim / 2;
  MMPL_CUDA_LAUNCH(launch_kernel(sinusoid_kernel, (n + 127) / 128, 128, 0, st, t, static_cast<__nv_bfloat16*>(out), F, dim));
  MMPL_CUDA(cudaGetLastError());
  return MMPL_OK;
}

int skinny_linear(const void* x, int64_t ldx, const void* w, const void* b, void* out, int64_t ldo, int M, int N,
                  int K, int silu_in, int silu_out, cudaStream_t st) {
  MMPL_CHECK(M > 0 && M <= kMaxFrames && N > 0 && K % 8 == 0 && ldx % 8 == 0, MMPL_ERR_SHAPE,
             "skinny_linear: bad shape M=%d N=%d K=%d", M, N, K);
  const int grid = (N + 7) / 8;
  if (M <= 4)
    MMPL_CUDA_LAUNCH(launch_kernel(skinny_linear_kernel<4>, grid, 256, 0, st, static_cast<const __nv_bfloat16*>(x), ldx,
                                                  static_cast<const __nv_bfloat16*>(w), static_cast<const __nv_bfloat16*>(b),
                                                  static_cast<__nv_bfloat16*>(out), ldo, M, N, K, silu_in, silu_out));
  else
    MMPL_CUDA_LAUNCH(launch_kernel(skinny_linear_kernel<8>, grid, 256, 0, st, static_cast<const __nv_bfloat16*>(x), ldx,
                                                  static_cast<const __nv_bfloat16*>(w), static_cast<const __nv_bfloat16*>(b),
                                                  static_cast<__nv_bfloat16*>(out), ldo, M, N, K, silu_in, silu_out));
  MMPL_CUDA(cudaGetLastError());
  return MMPL_OK;
}

int patchify(const void* x, int64_t stride_f, int64_t stride_c, void* a, int F, int C, int H, int W, cudaStream_t st) {
  MMPL_CHECK(F > 0 && C > 0 && H % 2 == 0 && W % 2 == 0, MMPL_ERR_SHAPE, "patchify: bad shape");
  const int64_t n = static_cast<int64_t>(F) * C * H * W;
  MMPL_CUDA_LAUNCH(launch_kernel(patchify_kernel, grid_for(n, 256), 256, 0, st, static_cast<const __nv_bfloat16*>(x), stride_f, stride_c,
                                                    static_cast<__nv_bfloat16*>(a), F, C, H, W));
  MMPL_CUDA(cudaGetLastError());
  return MMPL_OK;
}

int unpatchify_x0(const void* head, int64_t ldh, const void* xt, int64_t xt_stride_f, int64_t xt_stride_c,
                  const double* sigma, void* flow, void* x0, int F, int C, int H, int W, cudaStream_t st) {
  MMPL_CHECK(F > 0 && C > 0 && H % 2 == 0 && W % 2 == 0, MMPL_ERR_SHAPE, "unpatchify_x0: bad shape");
  MMPL_CHECK(x0 == nullptr || (xt && sigma), MMPL_ERR_ARG, "unpatchify_x0: x0 output needs xt and sigma");
  const int64_t n = static_cast<int64_t>(F) * C * H * W;
  MMPL_CUDA_LAUNCH(launch_kernel(unpatchify_x0_kernel, grid_for(n, 256), 256, 0, st, 
      static_cast<const __nv_bfloat16*>(head), ldh, static_cast<const __nv_bfloat16*>(xt), xt_stride_f, xt_stride_c,
      sigma, static_cast<__nv_bfloat16*>(flow), static_cast<__nv_bfloat16*>(x0), F, C, H, W));
  MMPL_CUDA(cudaGetLastError());
  return MMPL_OK;
}

int add_noise(const void* x0, const void* noise, const float* sigma, void* out, int n_frames, int64_t per_frame,
              cudaStream_t st) {
  MMPL_CHECK(n_frames > 0 && per_frame > 0, MMPL_ERR_SHAPE, "add_noise: bad shape");
  const int64_t n = per_frame * n_frames;
  MMPL_CUDA_LAUNCH(launch_kernel(add_noise_kernel, grid_for(n, 256), 256, 0, st, static_cast<const __nv_bfloat16*>(x0),
                                                     static_cast<const __nv_bfloat16*>(noise), sigma,
                                                     static_cast<__nv_bfloat16*>(out), per_frame, n));
  MMPL_CUDA(cudaGetLastError());
  return MMPL_OK;
}

}  // namespace mmpl
