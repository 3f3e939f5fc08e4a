// HBM-bound kernels of the VAE "segment connect" (wan/modules/vae.py), all on the zero-haloed
// channels-last grid the tap-GEMM convolution (gemm_tcgen05.cu: conv3d_cl) reads and writes:
//   grid [frames][H + 2][W + 2][C] bf16, halo = 0.
// RMS_norm (+ SiLU), nearest-neighbour 2x up-sampling, the stride-2 pick behind the down-sampling convolutions and the
// row softmax of the single-head middle attention. 16-byte vector accesses; every kernel rounds to bf16 where the
// reference's bf16 tensors round (one rounding per torch operator).
#include <cfloat>

#include "host_util.h"
#include "kernels.h"
#include "mmpl_b200.h"
#include "ptx.cuh"

namespace mmpl {

namespace {

__device__ __forceinline__ void unpack8v(const uint4& v, float (&f)[8]) {
  const uint32_t* w = reinterpret_cast<const uint32_t*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[2 * i] = bf16_lo(w[i]);
    f[2 * i + 1] = bf16_hi(w[i]);
  }
}
__device__ __forceinline__ void round2v(float& a, float& b) {  // both to bf16 and back with one cvt.rn.bf16x2.f32
  const uint32_t p = pack_bf16x2(a, b);
  a = bf16_lo(p);
  b = bf16_hi(p);
}
__device__ __forceinline__ uint4 pack8v(const float (&f)[8]) {
  return make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
}

// RMS_norm.forward (vae.py:39-55) = F.normalize(x, dim=channels) * sqrt(C) * gamma, optionally followed by nn.SiLU, with the
// bf16 rounding points of the reference's operator sequence:
//   n  = bf16(sqrt(sum x^2))            torch.norm: fp32 accumulation, bf16 result
//   y1 = bf16(x / max(n, 1e-12))        F.normalize
//   y2 = bf16(y1 * sqrt(C))             python-float scale (fp32 op-math)
//   y3 = bf16(y2 * gamma)
//   y4 = bf16(y3 / (1 + exp(-y3)))      SiLU
// G lanes share one grid position (row of C channels), NCH 16-byte chunks per lane; a warp handles 32/G rows.
// A halo row is all zero and stays zero (0 / eps = 0, silu(0) = 0), so every row of the grid is processed alike.
template <int G, int NCH>
__global__ void __launch_bounds__(256)
vae_norm_act_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ out, int64_t rows, int C,
                    const __nv_bfloat16* __restrict__ gamma, float scale, int silu) {
  pdl_wait();
  pdl_launch_dependents();
  constexpr int kRowsPerWarp = 32 / G;
  const int lane = threadIdx.x & 31;
  const int sub = lane / G, gl = lane % G;
  const int64_t warp = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t row = warp * kRowsPerWarp + sub;
  const int chunks = C >> 3;
  const bool row_ok = row < rows;
  const uint4* xr = reinterpret_cast<const uint4*>(x + (row_ok ? row : 0) * C);
  uint4 v[NCH];
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    const int c = gl + G * i;
    v[i] = (row_ok && c < chunks) ? xr[c] : make_uint4(0, 0, 0, 0);
    float f[8];
    unpack8v(v[i], f);
#pragma unroll
    for (int j = 0; j < 8; ++j) sq = fmaf(f[j], f[j], sq);
  }
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
  const float n = fmaxf(bf16_round(sqrtf(sq)), 1e-12f);
  if (!row_ok) return;
  // x / n with n shared by the row: q0 = x * RN(1/n), corrected with the exact remainder (two FMAs). For bf16 operands this
  // is the correctly rounded fp32 quotient - checked exhaustively over all 2^7 x 2^7 mantissa pairs - so the bf16 result
  // equals the reference's division; results outside the normal range take the division instruction.
  const float rn = __frcp_rn(n);
  auto quotient = [&](float a) {
    const float q0 = __fmul_rn(a, rn);
    const float q = __fmaf_rn(__fmaf_rn(-q0, n, a), rn, q0);
    const float aq = fabsf(q0);
    return (aq == 0.f || (aq >= 1e-30f && aq <= 1e30f)) ? q : __fdiv_rn(a, n);
  };
  uint4* orow = reinterpret_cast<uint4*>(out + row * C);
  const uint4* gp = reinterpret_cast<const uint4*>(gamma);
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    const int c = gl + G * i;
    if (c >= chunks) continue;
    float f[8], g[8];
    unpack8v(v[i], f);
    unpack8v(__ldg(gp + c), g);
#pragma unroll
    for (int j = 0; j < 8; j += 2) {  // two elements per conversion instruction at every rounding point
      float y0 = quotient(f[j]), y1 = quotient(f[j + 1]);
      round2v(y0, y1);
      y0 *= scale;
      y1 *= scale;
      round2v(y0, y1);
      y0 *= g[j];
      y1 *= g[j + 1];
      if (silu) {
        round2v(y0, y1);
        y0 = y0 / (1.0f + expf(-y0));
        y1 = y1 / (1.0f + expf(-y1));
      }
      f[j] = y0;          // rounded by the pack below
      f[j + 1] = y1;
    }
    orow[c] = pack8v(f);
  }
}

// Upsample(scale_factor=(2,2), mode='nearest') per frame (vae.py:58-64,75-78): out interior (2i+a, 2j+b) = in interior (i, j),
// out halo = 0 (every position of `out` is written: no fill pass). One thread per (output position, 16-byte chunk).
__global__ void __launch_bounds__(256)
vae_upsample2x_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, int frames, int H, int W, int chunks) {
  pdl_wait();
  pdl_launch_dependents();
  const int Ho = 2 * H + 2, Wo = 2 * W + 2;
  const int64_t total = static_cast<int64_t>(frames) * Ho * Wo * chunks;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % chunks);
    int64_t r = i / chunks;
    const int ow = static_cast<int>(r % Wo);
    r /= Wo;
    const int oh = static_cast<int>(r % Ho);
    const int f = static_cast<int>(r / Ho);
    uint4 v = make_uint4(0, 0, 0, 0);
    if (ow >= 1 && ow <= 2 * W && oh >= 1 && oh <= 2 * H)
      v = __ldg(in + ((static_cast<int64_t>(f) * (H + 2) + ((oh - 1) >> 1) + 1) * (W + 2) + ((ow - 1) >> 1) + 1) * chunks + c);
    out[i] = v;
  }
}

// The stride-2 pick that turns a stride-1 "same" 3x3 convolution into nn.ZeroPad2d((0,1,0,1)) + Conv2d(3, stride 2)
// (vae.py:85-88): the strided convolution's output (i, j) has its window centred on input (2i+1, 2j+1), whose bottom /
// right neighbours beyond the image are the halo's zeros. out interior (i, j) = in interior (2i+1, 2j+1), out halo = 0;
// H, W = output size.
__global__ void __launch_bounds__(256)
vae_pick_odd_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, int frames, int H, int W, int Hin, int Win,
                    int chunks) {
  pdl_wait();
  pdl_launch_dependents();
  const int Ho = H + 2, Wo = W + 2;
  const int64_t total = static_cast<int64_t>(frames) * Ho * Wo * chunks;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % chunks);
    int64_t r = i / chunks;
    const int ow = static_cast<int>(r % Wo);
    r /= Wo;
    const int oh = static_cast<int>(r % Ho);
    const int f = static_cast<int>(r / Ho);
    uint4 v = make_uint4(0, 0, 0, 0);
    if (ow >= 1 && ow <= W && oh >= 1 && oh <= H)
      v = __ldg(in + ((static_cast<int64_t>(f) * (Hin + 2) + 2 * oh) * (Win + 2) + 2 * ow) * chunks + c);
    out[i] = v;
  }
}

// Row softmax of the middle attention block (vae.py:253-258, F.scaled_dot_product_attention on one head of C channels):
// p = softmax(scale * s) in fp32 over a row of L bf16 scores, bf16 result. One CTA per row.
__global__ void __launch_bounds__(256)
softmax_rows_kernel(const __nv_bfloat16* __restrict__ s, int64_t lds, __nv_bfloat16* __restrict__ p, int64_t ldp, int L,
                    float scale_log2e) {
  pdl_wait();
  pdl_launch_dependents();
  __shared__ float red[8];
  __shared__ float bcast;
  const __nv_bfloat16* srow = s + static_cast<int64_t>(blockIdx.x) * lds;
  __nv_bfloat16* prow = p + static_cast<int64_t>(blockIdx.x) * ldp;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float m = -FLT_MAX;
  for (int i = tid; i < L; i += 256) m = fmaxf(m, __bfloat162float(srow[i]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if (lane == 0) red[warp] = m;
  __syncthreads();
  if (tid == 0) {
    float t = red[0];
    for (int i = 1; i < 8; ++i) t = fmaxf(t, red[i]);
    bcast = t;
  }
  __syncthreads();
  m = bcast;
  __syncthreads();
  // exp(scale * (s - max)) = exp2(scale_log2e * s - scale_log2e * max)   (scale > 0, so the max commutes)
  const float moff = m * scale_log2e;
  float sum = 0.f;
  for (int i = tid; i < L; i += 256) sum += exp2f(fmaf(__bfloat162float(srow[i]), scale_log2e, -moff));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if (lane == 0) red[warp] = sum;
  __syncthreads();
  if (tid == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += red[i];
    bcast = t;
  }
  __syncthreads();
  const float inv = 1.0f / bcast;
  for (int i = tid; i < L; i += 256)
    prow[i] = __float2bfloat16_rn(exp2f(fmaf(__bfloat162float(srow[i]), scale_log2e, -moff)) * inv);
}

int grid_1d(int64_t n, int block) {
  const int64_t g = (n + block - 1) / block;
  const int64_t cap = static_cast<int64_t>(sm_count() > 0 ? sm_count() : 148) * 16;
  return static_cast<int>(g < cap ? (g > 0 ? g : 1) : cap);
}

}  // namespace

int vae_norm_act(const void* x, void* out, int64_t rows, int C, const void* gamma, int silu, cudaStream_t st) {
  MMPL_CHECK(device_is_sm100(), MMPL_ERR_ARCH, "vae_norm_act: requires an sm_100 device");
  MMPL_CHECK(x && out && gamma, MMPL_ERR_ARG, "vae_norm_act: null argument");
  MMPL_CHECK(rows > 0 && C > 0 && C % 8 == 0 && C <= 1024, MMPL_ERR_SHAPE, "vae_norm_act: bad shape rows=%lld C=%d",
             (long long)rows, C);
  const float scale = sqrtf(static_cast<float>(C));  // float(dim ** 0.5): the python double rounded to fp32 by the scalar op
  const int chunks = C / 8;
  const auto* xp = static_cast<const __nv_bfloat16*>(x);
  auto* op = static_cast<__nv_bfloat16*>(out);
  const auto* gp = static_cast<const __nv_bfloat16*>(gamma);
#define MMPL_VAE_NORM(G, NCH)                                                                                        \
  do {                                                                                                               \
    const int64_t warps = (rows + (32 / G) - 1) / (32 / G);                                                          \
    const int64_t blocks = (warps + 7) / 8;                                                                          \
    MMPL_CHECK(blocks < (int64_t(1) << 31), MMPL_ERR_SHAPE, "vae_norm_act: too many rows");                          \
    MMPL_CUDA_LAUNCH(launch_kernel(vae_norm_act_kernel<G, NCH>, static_cast<int>(blocks), 256, 0, st, xp, op, rows,  \
                                   C, gp, scale, silu));                                                             \
  } while (0)
  if (chunks <= 4) MMPL_VAE_NORM(4, 1);
  else if (chunks <= 8) MMPL_VAE_NORM(8, 1);
  else if (chunks <= 16) MMPL_VAE_NORM(16, 1);
  else if (chunks <= 32) MMPL_VAE_NORM(32, 1);
  else if (chunks <= 64) MMPL_VAE_NORM(32, 2);
  else MMPL_VAE_NORM(32, 4);
#undef MMPL_VAE_NORM
  MMPL_CUDA(cudaGetLastError());
  return MMPL_OK;
}

int vae_upsample2x(const void* in, void* out, int frames, int H, int W, int C, cudaStream_t st) {
  MMPL_CHECK(device_is_sm100(), MMPL_ERR_ARCH, "vae_upsample2x: requires an sm_100 device");
  MMPL_CHECK(in && out && frames > 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0, MMPL_ERR_SHAPE, "vae_upsample2x: bad shape");
  const int64_t total = static_cast<int64_t>(frames) * (2 * H + 2) * (2 * W + 2) * (C / 8);
  MMPL_CUDA_LAUNCH(launch_kernel(vae_upsample2x_kernel, grid_1d(total, 256), 256, 0, st, static_cast<const uint4*>(in),
                                 static_cast<uint4*>(out), frames, H, W, C / 8));
  MMPL_CUDA(cudaGetLastError());
  return MMPL_OK;
}

int vae_pick_odd(const void* in, void* out, int frames, int Hin, int Win, int C, cudaStream_t st) {
  MMPL_CHECK(device_is_sm100(), MMPL_ERR_ARCH, "vae_pick_odd: requires an sm_100 device");
  MMPL_CHECK(in && out && frames > 0 && Hin > 1 && Win > 1 && C > 0 && C % 8 == 0, MMPL_ERR_SHAPE, "vae_pick_odd: bad shape");
  const int H = Hin / 2, W = Win / 2;  // floor((Hin + 1 - 3) / 2) + 1 outputs of the padded stride-2 convolution
  const int64_t total = static_cast<int64_t>(frames) * (H + 2) * (W + 2) * (C / 8);
  MMPL_CUDA_LAUNCH(launch_kernel(vae_pick_odd_kernel, grid_1d(total, 256), 256, 0, st, static_cast<const uint4*>(in),
                                 static_cast<uint4*>(out), frames, H, W, Hin, Win, C / 8));
  MMPL_CUDA(cudaGetLastError());
  return MMPL_OK;
}

int softmax_rows(const void* s, int64_t lds, void* p, int64_t ldp, int rows, int L, float scale, cudaStream_t st) {
  MMPL_CHECK(device_is_sm100(), MMPL_ERR_ARCH, "softmax_rows: requires an sm_100 device");
  MMPL_CHECK(s && p && rows > 0 && L > 0 && scale > 0.f, MMPL_ERR_SHAPE, "softmax_rows: bad shape rows=%d L=%d", rows, L);
  MMPL_CUDA_LAUNCH(launch_kernel(softmax_rows_kernel, rows, 256, 0, st, static_cast<const __nv_bfloat16*>(s), lds,
                                 static_cast<__nv_bfloat16*>(p), ldp, L, scale * 1.4426950408889634f));
  MMPL_CUDA(cudaGetLastError());
  return MMPL_OK;
}

}  // namespace mmpl
