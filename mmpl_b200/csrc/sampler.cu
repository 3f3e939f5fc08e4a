// One launch per denoising step of the MMPL hot loop: classifier-free-guidance combine, flow -> x0,
// UniPC corrector and UniPC predictor (order <= 2, bh2, predict_x0, flow_prediction: the only configuration the reference
// pipeline instantiates, pipeline/casual_fps_inference.py:503-512).
//
//   reference                                                      this kernel, per element
//   casual_fps_inference.py:366-374   u + g * (c - u)               flow
//   fm_solvers_unipc.py:318-321       sample - sigma * flow         x0         (= model_outputs[-1] after the step)
//   fm_solvers_unipc.py:486-626       multistep_uni_c_bh_update     corrected  (= last_sample after the step)
//   fm_solvers_unipc.py:350-484       multistep_uni_p_bh_update     next       (= the sample handed to the next forward)
//
// The reference evaluates this as ~25 element-wise torch operators on bf16 tensors, each computing in fp32 and rounding
// its result to bf16. The kernel performs the same operators in the same order with the same rounding points (every
// line below that ends in r(...) is one torch operator), so its output is bit-identical to the operator sequence; all
// sigma-dependent scalars are computed by the host (mmpl_b200/unipc.py) and arrive in `mmpl_unipc_coeffs`.
// HBM-bound: 6 reads + 3 writes of n bf16 per launch (12.6 MB for the 7-frame anchor stage at 60x104).
#include "host_util.h"
#include "kernels.h"
#include "mmpl_b200.h"
#include "ptx.cuh"

namespace mmpl {

namespace {

__device__ __forceinline__ float r(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }

struct Lane {
  float c, u, x, m1, m2, last;
};
struct Out {
  float next, x0, corrected;
};

// `mul` / `sub` / `add` are spelled with the non-contracting intrinsics: an FMA would skip a rounding point.
__device__ __forceinline__ float scal(float coef, float v) { return r(__fmul_rn(coef, v)); }
__device__ __forceinline__ float diff(float a, float b) { return r(__fsub_rn(a, b)); }

template <bool kCombine>
__device__ __forceinline__ Out unipc_element(const Lane& in, const mmpl_unipc_coeffs& k) {
  Out o;
  float flow = in.c;
  if (kCombine) {
    const float d = diff(in.c, in.u);
    flow = r(__fadd_rn(in.u, scal(k.guidance, d)));
  }
  const float x0 = diff(in.x, scal(k.sigma, flow));
  // corrector: last_sample, model_outputs[-1] = m1, model_outputs[-2] = m2, this step's x0
  float sample = in.x;
  if (k.corr_order > 0) {
    const float xt = diff(scal(k.corr_a, in.last), scal(k.corr_b, in.m1));
    const float d1t = diff(x0, in.m1);
    float res = r(__fmul_rn(k.corr_rho1, d1t));
    if (k.corr_order == 2) {
      const float d = diff(in.m2, in.m1);
      const float d1 = k.true_division ? r(__fdiv_rn(d, k.corr_rk)) : scal(k.corr_rk, d);
      res = r(__fadd_rn(r(__fmul_rn(k.corr_rho0, d1)), res));
    }
    sample = diff(xt, scal(k.corr_c, res));
  }
  // predictor from the corrected sample: model_outputs[-1] = x0, model_outputs[-2] = m1
  const float xt = diff(scal(k.pred_a, sample), scal(k.pred_b, x0));
  float next;
  if (k.pred_order == 2) {
    const float d = diff(in.m1, x0);
    const float d1 = k.true_division ? r(__fdiv_rn(d, k.pred_rk)) : scal(k.pred_rk, d);
    next = diff(xt, scal(k.pred_c, __fmul_rn(0.5f, d1)));  // rhos_p = 0.5: exact in bf16
  } else {
    next = r(__fsub_rn(xt, __fmul_rn(k.pred_c, 0.0f)));  // pred_res = 0
  }
  o.next = next;
  o.x0 = x0;
  o.corrected = sample;
  return o;
}

template <bool kCombine>
__global__ void __launch_bounds__(256)
unipc_cfg_step_kernel(const __nv_bfloat16* __restrict__ fc, const __nv_bfloat16* __restrict__ fu,
                      const __nv_bfloat16* x, const __nv_bfloat16* m1, const __nv_bfloat16* m2, const __nv_bfloat16* last,
                      __nv_bfloat16* next, __nv_bfloat16* x0_out, __nv_bfloat16* corrected, int64_t n,
                      const mmpl_unipc_coeffs k) {
  pdl_wait();
  pdl_launch_dependents();
  const bool need_m1 = k.corr_order > 0 || k.pred_order == 2;
  const bool need_m2 = k.corr_order == 2;
  const int64_t vec = n >> 3;
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i < vec) {
    const uint4 zero = make_uint4(0, 0, 0, 0);
    // all loads first: one round trip to HBM per thread
    const uint4 vc = reinterpret_cast<const uint4*>(fc)[i];
    const uint4 vu = kCombine ? reinterpret_cast<const uint4*>(fu)[i] : zero;
    const uint4 vx = reinterpret_cast<const uint4*>(x)[i];
    const uint4 v1 = need_m1 ? reinterpret_cast<const uint4*>(m1)[i] : zero;
    const uint4 v2 = need_m2 ? reinterpret_cast<const uint4*>(m2)[i] : zero;
    const uint4 vl = k.corr_order > 0 ? reinterpret_cast<const uint4*>(last)[i] : zero;
    const uint32_t *pc = &vc.x, *pu = &vu.x, *px = &vx.x, *p1 = &v1.x, *p2 = &v2.x, *pl = &vl.x;
    uint4 on, o0, os;
    uint32_t *qn = &on.x, *q0 = &o0.x, *qs = &os.x;
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      const Lane lo = {bf16_lo(pc[w]), bf16_lo(pu[w]), bf16_lo(px[w]), bf16_lo(p1[w]), bf16_lo(p2[w]), bf16_lo(pl[w])};
      const Lane hi = {bf16_hi(pc[w]), bf16_hi(pu[w]), bf16_hi(px[w]), bf16_hi(p1[w]), bf16_hi(p2[w]), bf16_hi(pl[w])};
      const Out a = unipc_element<kCombine>(lo, k), b = unipc_element<kCombine>(hi, k);
      qn[w] = pack_bf16x2(a.next, b.next);
      q0[w] = pack_bf16x2(a.x0, b.x0);
      qs[w] = pack_bf16x2(a.corrected, b.corrected);
    }
    reinterpret_cast<uint4*>(next)[i] = on;
    reinterpret_cast<uint4*>(x0_out)[i] = o0;
    reinterpret_cast<uint4*>(corrected)[i] = os;
  } else if (i == vec) {  // tail of n % 8 elements
    for (int64_t j = vec << 3; j < n; ++j) {
      const Lane in = {__bfloat162float(fc[j]), kCombine ? __bfloat162float(fu[j]) : 0.f, __bfloat162float(x[j]),
                       need_m1 ? __bfloat162float(m1[j]) : 0.f, need_m2 ? __bfloat162float(m2[j]) : 0.f,
                       k.corr_order > 0 ? __bfloat162float(last[j]) : 0.f};
      const Out o = unipc_element<kCombine>(in, k);
      next[j] = __float2bfloat16_rn(o.next);
      x0_out[j] = __float2bfloat16_rn(o.x0);
      corrected[j] = __float2bfloat16_rn(o.corrected);
    }
  }
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace

int unipc_cfg_step(const void* flow_cond, const void* flow_uncond, const void* sample, const void* m_prev1, const void* m_prev2,
                   const void* last_sample, void* sample_next, void* x0_out, void* corrected_out, int64_t n,
                   const mmpl_unipc_coeffs* k, cudaStream_t st) {
  MMPL_CHECK(device_is_sm100(), MMPL_ERR_ARCH, "unipc_cfg_step: requires an sm_100 device");
  MMPL_CHECK(flow_cond && sample && sample_next && x0_out && corrected_out && k && n > 0, MMPL_ERR_ARG, "unipc_cfg_step: null argument");
  MMPL_CHECK(k->corr_order >= 0 && k->corr_order <= 2 && k->pred_order >= 1 && k->pred_order <= 2, MMPL_ERR_ARG,
             "unipc_cfg_step: orders must be corrector 0..2, predictor 1..2");
  MMPL_CHECK(k->corr_order == 0 || (m_prev1 && last_sample), MMPL_ERR_ARG, "unipc_cfg_step: the corrector needs m_prev1 and last_sample");
  MMPL_CHECK(k->corr_order < 2 || m_prev2, MMPL_ERR_ARG, "unipc_cfg_step: an order-2 corrector needs m_prev2");
  MMPL_CHECK(k->pred_order < 2 || m_prev1, MMPL_ERR_ARG, "unipc_cfg_step: an order-2 predictor needs m_prev1");
  MMPL_CHECK(aligned16(flow_cond) && aligned16(flow_uncond) && aligned16(sample) && aligned16(m_prev1) && aligned16(m_prev2) &&
                 aligned16(last_sample) && aligned16(sample_next) && aligned16(x0_out) && aligned16(corrected_out),
             MMPL_ERR_ARG, "unipc_cfg_step: tensors must be 16-byte aligned");
  // outputs may alias the input read at the same index (sample_next = sample, corrected_out = last_sample,
  // x0_out = m_prev2); x0_out must not be m_prev1 of a later element's read -- it is not: accesses are per index
  const int64_t threads = (n >> 3) + 1;
  const int grid = static_cast<int>((threads + 255) / 256);
  auto* fc = static_cast<const __nv_bfloat16*>(flow_cond);
  auto* fu = static_cast<const __nv_bfloat16*>(flow_uncond);
  auto* x = static_cast<const __nv_bfloat16*>(sample);
  auto* m1 = static_cast<const __nv_bfloat16*>(m_prev1);
  auto* m2 = static_cast<const __nv_bfloat16*>(m_prev2);
  auto* ls = static_cast<const __nv_bfloat16*>(last_sample);
  auto* nx = static_cast<__nv_bfloat16*>(sample_next);
  auto* xo = static_cast<__nv_bfloat16*>(x0_out);
  auto* co = static_cast<__nv_bfloat16*>(corrected_out);
  if (flow_uncond != nullptr)
    MMPL_CUDA_LAUNCH(launch_kernel(unipc_cfg_step_kernel<true>, grid, 256, 0, st, fc, fu, x, m1, m2, ls, nx, xo, co, n, *k));
  else
    MMPL_CUDA_LAUNCH(launch_kernel(unipc_cfg_step_kernel<false>, grid, 256, 0, st, fc, fu, x, m1, m2, ls, nx, xo, co, n, *k));
  MMPL_CUDA(cudaGetLastError());
  return MMPL_OK;
}

}  // namespace mmpl
