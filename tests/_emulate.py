"""Test helpers: plain-torch emulations of kernel contracts, used to check host-side tables on the CPU (no GPU here)
and as a second reference for the kernels on the GPU."""
import torch


def _r(v: torch.Tensor) -> torch.Tensor:
    return v.to(torch.bfloat16).float()


def emulate_unipc_step(k, c, u, x, m1, m2, last):
    """The element program of mmpl_unipc_cfg_step (include/mmpl_b200.h), one fp32 torch operator per kernel operator.
    `k` is a mmpl_b200.unipc.StepCoeffs; tensors are bf16 (u may be None). Returns (next, x0, corrected) in bf16."""
    f32 = lambda t: None if t is None else t.float()
    c, u, x, m1, m2, last = map(f32, (c, u, x, m1, m2, last))
    mul = lambda coef, v: _r(torch.tensor(coef, dtype=torch.float32) * v)
    flow = c if u is None else _r(u + mul(k.guidance, _r(c - u)))
    x0 = _r(x - mul(k.sigma, flow))

    def D(d, rk):
        return _r(d / torch.tensor(rk, dtype=torch.float32)) if k.true_division else mul(rk, d)

    sample = x
    if k.corr_order > 0:
        xt = _r(mul(k.corr_a, last) - mul(k.corr_b, m1))
        res = mul(k.corr_rho1, _r(x0 - m1))
        if k.corr_order == 2:
            res = _r(mul(k.corr_rho0, D(_r(m2 - m1), k.corr_rk)) + res)
        sample = _r(xt - mul(k.corr_c, res))
    xt = _r(mul(k.pred_a, sample) - mul(k.pred_b, x0))
    if k.pred_order == 2:
        nxt = _r(xt - mul(k.pred_c, 0.5 * D(_r(m1 - x0), k.pred_rk)))
    else:
        nxt = _r(xt - torch.tensor(k.pred_c, dtype=torch.float32) * 0.0)
    bf = lambda t: t.to(torch.bfloat16)
    return bf(nxt), bf(x0), bf(sample)


def emulate_unipc_run(table, flows_c, flows_u, x):
    """All steps of a run; returns the list of samples after each step."""
    m1 = m2 = last = None
    outs = []
    for i, k in enumerate(table.coeffs):
        nxt, x0, corrected = emulate_unipc_step(k, flows_c[i], None if flows_u is None else flows_u[i], x, m1, m2, last)
        m2, m1, last, x = m1, x0, corrected, nxt
        outs.append(nxt)
    return outs
