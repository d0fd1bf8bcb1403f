"""U(1) gauge links: the input boundary of the hot path (class Gauge, S6/gauge.h).

Layout: U[L*L, 2] complex, U[s, 0] = U_x(s), U[s, 1] = U_y(s), s = x + y*L (S6/gauge.h:29-37).
The reference only READS configurations (`../gauge_config_files/phase_{L}_b{beta}.dat`, S6/gauge.h:44,88-110)
that it does not ship; the generators below produce inputs of the same kind.
"""
from __future__ import annotations

import math

import numpy as np
import torch


def cold(L: int) -> np.ndarray:
    """U = 1 (the constructor default, S6/gauge.h:29-37)."""
    return np.ones((L * L, 2), dtype=np.complex128)


def from_phases(theta) -> np.ndarray:
    """U = polar(1, phase) (S6/gauge.h:106)."""
    return np.exp(1j * np.asarray(theta, dtype=np.float64)).astype(np.complex128)


def gaussian(L: int, width: float = 0.2, seed: int = 1234) -> np.ndarray:
    """Local phases ~ N(0, width): the option left commented out in S6/gauge.h:25-26,36."""
    rng = np.random.default_rng(seed + L)
    return from_phases(rng.normal(0.0, width, size=(L * L, 2)))


def quenched_phases(L: int, beta: float, sweeps: int = 200, seed: int = 1234, device: str = "cpu") -> torch.Tensor:
    """Compact-U(1) Wilson-action checkerboard Metropolis (beta = 6, 32 are the reference's values:
    S5L/mgrid_laplace.cpp:135, S6/params.h:66).  Returns phases theta[L*L, 2] (float64 torch tensor on `device`).
    Pure tensor ops, so large lattices can be generated on the GPU; this is input generation, not the hot path."""
    g = torch.Generator(device=device)
    g.manual_seed(seed + L)
    th = torch.zeros(L, L, 2, dtype=torch.float64, device=device)      # th[y, x, dir]
    yy, xx = torch.meshgrid(torch.arange(L, device=device), torch.arange(L, device=device), indexing="ij")
    delta = min(math.pi, 2.0 / math.sqrt(beta))

    def sh(a, d, k):  # value at x + k*d_hat
        return torch.roll(a, shifts=-k, dims=1 if d == 0 else 0)

    def action(th, mu, t_mu):
        nu = 1 - mu
        t_nu, o_mu = th[..., nu], th[..., mu]
        p_up = t_mu + sh(t_nu, mu, 1) - sh(o_mu, nu, 1) - t_nu
        p_dn = sh(t_nu, nu, -1) + t_mu - sh(sh(t_nu, nu, -1), mu, 1) - sh(o_mu, nu, -1)
        return -beta * (torch.cos(p_up) + torch.cos(p_dn))

    for _ in range(sweeps):
        for mu in (0, 1):
            for par in (0, 1):
                mask = ((xx + yy) % 2) == par
                old = th[..., mu]
                new = old + (torch.rand(old.shape, generator=g, dtype=torch.float64, device=device) * 2 - 1) * delta
                dS = action(th, mu, new) - action(th, mu, old)
                acc = mask & (torch.rand(old.shape, generator=g, dtype=torch.float64, device=device) < torch.exp(-dS))
                th[..., mu] = torch.where(acc, new, old)
    th = (th + math.pi) % (2 * math.pi) - math.pi
    return th.reshape(L * L, 2)


def quenched_links_device(L: int, beta: float, sweeps: int = 200, seed: int = 1234, device: int | None = None,
                          dtype: str = "complex128", return_phases: bool = False):
    """The same checkerboard Metropolis as CUDA kernels of libmg2d_sm100.so (mg2d_gauge_metropolis: one launch per
    half-update, counter-based random numbers keyed on (seed, half-update, site); mirrored draw for draw by
    oracle.gauge_quenched_phases_counter) followed by mg2d_phases_to_links.  Returns U[L*L, 2] on the device
    (and the phases theta[L*L, 2] when return_phases)."""
    from . import _lib
    from .mg import _DT, _stream
    dev = torch.cuda.current_device() if device is None else device
    ctx = _lib.Context(dev)
    tdtype, dcode = _DT[dtype]
    with torch.cuda.device(dev):
        th = torch.zeros((L * L, 2), dtype=torch.float64, device=f"cuda:{dev}")
        delta = min(math.pi, 2.0 / math.sqrt(beta))
        for sw in range(sweeps):
            for mu in (0, 1):
                for par in (0, 1):
                    ctx.call("mg2d_gauge_metropolis", th.data_ptr(), L, float(beta), float(delta), mu, par, seed,
                             (sw * 2 + mu) * 2 + par, _stream())
        th = torch.remainder(th + math.pi, 2 * math.pi) - math.pi
        U = torch.empty((L * L, 2), dtype=tdtype, device=th.device)
        ctx.call("mg2d_phases_to_links", U.data_ptr(), th.data_ptr(), 2 * L * L, dcode, _stream())
        torch.cuda.synchronize()
    ctx.close()
    return (U, th) if return_phases else U


def plaquette_device(U: torch.Tensor, L: int) -> complex:
    """Gauge::f_plaquette (S6/gauge.h:50-63) as a fused reduction kernel (mg2d_plaquette)."""
    from . import _lib
    from .mg import _stream
    dev = U.device.index
    ctx = _lib.Context(dev)
    out = torch.zeros(2, dtype=torch.float64, device=U.device)
    with torch.cuda.device(dev):
        ctx.call("mg2d_plaquette", U.data_ptr(), L, _lib.C128 if U.dtype == torch.complex128 else _lib.C64, out.data_ptr(), _stream())
        o = out.cpu()
    ctx.close()
    return complex(o[0].item(), o[1].item()) / (L * L)


def plaquette(U, L: int) -> complex:
    """Gauge::f_plaquette (S6/gauge.h:50-63): mean of U_x(s) U_y(s+x) U_x(s+y)^* U_y(s)^*."""
    U = torch.as_tensor(U).reshape(L, L, 2)
    ux, uy = U[..., 0], U[..., 1]
    p = ux * torch.roll(uy, -1, 1) * torch.roll(ux, -1, 0).conj() * uy.conj()
    return complex(p.mean().item())


def write_phase_file(path: str, theta, L: int) -> None:
    """One phase per line, x outer / y inner / dir inner (f_read_gauge_heatbath, S6/gauge.h:88-110)."""
    th = np.asarray(theta, dtype=np.float64).reshape(L, L, 2)           # [y, x, dir]
    np.savetxt(path, th.transpose(1, 0, 2).reshape(-1), fmt="%.17g")


def read_phase_file(path: str, L: int) -> np.ndarray:
    vals = np.loadtxt(path).reshape(L, L, 2)                            # [x, y, dir]
    return np.ascontiguousarray(vals.transpose(1, 0, 2)).reshape(L * L, 2)
