"""Locate the critical mass of a gauge configuration with the multigrid solver itself.

D(m) = D(0) + m, so "near-critical" means m = m_crit + delta with m_crit = -min Re lambda(D(0)).
For lattices too large for a dense/ARPACK eigen-solve the lowest eigenvalue is found by inverse iteration
x <- D(m0)^-1 x, each solve done by the MG V-cycle iteration; lambda ~ <x, D(0) x>/<x, x>.
Input preparation, not part of the timed hot path.
"""
from __future__ import annotations

import torch


def estimate_critical_mass(U, params_factory, m0: float = 0.0, iters: int = 6, refine: int = 4, tol: float = 1e-6,
                           margin: float = 0.02, verbose: bool = False):
    """params_factory(mass) -> MGParams.  Returns (m_crit_estimate, history of Rayleigh quotients)."""
    from . import setup, solve
    hist = []
    x = None
    lam = None
    for stage, (m, nit) in enumerate(((m0, iters), (None, refine))):
        if m is None:
            m = -lam.real + margin
        p = params_factory(m)
        mg = setup(U, p, init="device")
        lv = mg.LVL[0]
        if x is None:
            g = torch.Generator(device=mg.device); g.manual_seed(99)
            x = torch.randn((lv.S, lv.n), generator=g, dtype=torch.float64, device=mg.device).to(mg.tdtype)
        Dx = torch.empty_like(x)
        for _ in range(nit):
            x = x / torch.linalg.vector_norm(x)
            y, info = solve(mg, rhs=x, tol=tol, max_iters=200, check_every=4)
            x = y.clone()
            lv.apply_D(Dx, x)
            rq = (torch.vdot(x.reshape(-1), Dx.reshape(-1)) / torch.vdot(x.reshape(-1), x.reshape(-1))).item()
            lam = complex(rq) - m            # eigenvalue of D(0)
            # how good is the pair?  |D x - rq x| / |x| (the eigen-residual: with it the estimate is off by at most that much
            # times the condition number of the eigenvector basis) and the change of the estimate in this step
            res = float(torch.linalg.vector_norm(Dx - rq * x) / torch.linalg.vector_norm(x))
            estimate_critical_mass.info = {"eig_residual": res, "last_change": abs(lam - hist[-1]) if hist else None,
                                           "steps": len(hist) + 1}
            hist.append(lam)
            if verbose:
                print(f"  stage {stage} m={m:+.5f} iters={info['iters']} lambda(D0)~{lam:.6f}")
        mg.close()          # tens of GB at 4096^2: give them back before the next hierarchy is built
        del mg, lv, y
        import gc
        gc.collect()
        torch.cuda.empty_cache()
    return -lam.real, hist


estimate_critical_mass.info = {}
