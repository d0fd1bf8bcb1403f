"""numpy restatement of the reference's adaptive (non-)telescoping multigrid -- TEST INFRASTRUCTURE.

Follows, function by function, `S6/` = /root/reference/code/6_ntl-mg_new_code/3_combining_laplace_and_wilson/.
Nothing under the product package may import this module (see oracle/__init__.py).

Conventions (all from S6):
  * site index s = x + y*L, x fastest (S6/level.h:69-75); periodic neighbours
  * stencil slot k: 0 self, 1 (x+1), 2 (x-1), 3 (y+1), 4 (y-1) (S6/level.h:8,71-75)
  * fields `phi[L*L, n]`, operator `D[L*L, 5, n, n]`, near-null/projector `P[L*L, nc, nf]`, complex128
  * lexicographic Gauss-Seidel visits `for x: for y:` (x OUTER, S6/level.h:113-114).  It is evaluated
    here by anti-diagonal wavefronts x+y=c, which reproduces that order exactly: (x-1,y),(x,y-1) lie on
    front c-1 (already new), (x+1,y),(x,y+1) on front c+1 (still old), and the periodic wraps are
    consistent with the visiting order.

Additions with no reference counterpart (mirrored 1:1 by the CUDA path, used for GPU-vs-oracle parity):
  * `relax_mr`  minimal-residual smoother named by BASELINE.json:north_star
  * `n_dof_scale` / per-level block override (config 4: 8 null vectors, 4x4 aggregates)
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

C128 = np.complex128


# --------------------------------------------------------------------------------------------------
# RNG: libstdc++ std::mt19937 + uniform_real_distribution<double>(-pi, pi), bit-exact
# (S6/mgrid_ntl.cpp:35-36, S6/modules_indiv.h:19).  One double = two 32-bit draws, low word first.
# --------------------------------------------------------------------------------------------------
class StdMT19937:
    def __init__(self, seed: int = 4302529):
        self._rs = np.random.RandomState(seed)

    def uniform_pm_pi(self, n: int) -> np.ndarray:
        raw = self._rs.randint(0, 2 ** 32, size=2 * n, dtype=np.uint64)
        lo = raw[0::2].astype(np.float64)
        hi = raw[1::2].astype(np.float64)
        canon = (lo + hi * 4294967296.0) / 18446744073709551616.0
        canon = np.minimum(canon, np.nextafter(1.0, 0.0))
        return canon * (np.pi - (-np.pi)) + (-np.pi)


def counter_uniform(seed: int, stream: int, start: int, n: int, lo: float = 0.0, hi: float = 1.0) -> np.ndarray:
    """Counter-based uniform numbers (ours; mirrored bit for bit by mg2d_fill_uniform / mg2d_gauge_metropolis): element e is
    a splitmix64 hash of (seed, stream, start + e), so any sub-range of a field can be drawn independently -- strips of a
    domain-decomposed lattice get exactly the numbers the whole lattice gets at the same global index."""
    M = (1 << 64) - 1
    with np.errstate(over="ignore"):
        idx = np.arange(start + 1, start + n + 1, dtype=np.uint64)
        base = np.uint64((seed * 0x9E3779B97F4A7C15 + stream * 0xD1B54A32D192ED03) & M)
        z = base + idx * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    u = (z >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)
    return lo + (hi - lo) * u


# --------------------------------------------------------------------------------------------------
# params  (S6/params.h:38-128)
# --------------------------------------------------------------------------------------------------
@dataclass
class Params:
    L: int
    num_iters: int            # smoother sweeps per visit (argv[2])
    block: object             # block_x = block_y (argv[3]); int, or a list with one block size per coarsening step (ours)
    m: float                  # mass (argv[5]); used linearly
    nlevels: int              # number of coarse levels (argv[6]); levels are 0..nlevels
    t_flag: int = 0           # non-telescoping flag (argv[7])
    n_copies: int = 1         # NTL copies 1..4 (argv[8])
    stencil: str = "wilson"   # S6/params.h:68-69
    gs_flag: int = 1          # 1 Gauss-Seidel, 0 Jacobi (S6/params.h:61)
    quad: int = 1             # S6/params.h:63
    max_iters: int = 50000    # S6/params.h:64
    res_threshold: float = 1.0e-13  # S6/params.h:67
    n_dof_scale: int | None = None  # dof on coarse levels: 4 (wilson) / 2 (laplace) (S6/params.h:75,81)
    smoother: str | None = None     # None -> from gs_flag; 'gs' | 'jacobi' | 'mr' | 'rbgs' (mr, rbgs: ours)
    mr_omega: float = 1.0
    null_iters: int = 500     # S6/modules_main.h:193
    null_chunk: int = 4       # S6/level.h:190
    min_res_flag: int = 1     # S6/modules_main.h:391
    n_pre: object = None      # ours: per-level pre / post sweep counts (int or list); default num_iters as the reference
    n_post: object = None
    cycle: str = "V"          # ours: 'K' = Krylov-accelerated coarse solves (k_inner FGCR steps per coarse level)
    k_inner: int = 2
    size: list = field(default_factory=list)
    n_dof: list = field(default_factory=list)

    def __post_init__(self):
        if self.t_flag == 1 and self.nlevels < 2:
            raise ValueError("Need at least 2 levels for non-telescoping")  # S6/params.h:52-55
        if self.stencil == "wilson":
            n0, scale = 2, 4
        elif self.stencil == "laplace":
            n0, scale = 1, 2
        else:
            raise ValueError("Incorrect stencil: need 'laplace' or 'wilson'")
        if self.n_dof_scale is None:
            self.n_dof_scale = scale
        if isinstance(self.block, (list, tuple)):       # ours: per-level block sizes as S5L/setup.h:2-10 `block_x[level]`
            self.blocks = [int(b) for b in self.block]
            self.block = self.blocks[0] if self.blocks else 1
        else:
            self.blocks = [self.block] * self.nlevels
            max_levels = math.ceil(math.log2(self.L) / math.log2(self.block)) if self.block > 1 else 0
            if self.nlevels > max_levels:
                raise ValueError("Too many levels")  # S6/params.h:100-106
        self.size = [self.L]
        self.n_dof = [n0]
        for lvl in range(self.nlevels):
            self.size.append(self.size[-1] // self.blocks[lvl])
            self.n_dof.append(self.n_dof_scale)
        if self.smoother is None:
            self.smoother = "gs" if self.gs_flag == 1 else "jacobi"

        def per_level(v):
            if v is None:
                return [self.num_iters] * (self.nlevels + 1)
            return [v] * (self.nlevels + 1) if isinstance(v, int) else list(v)
        self.pre, self.post = per_level(self.n_pre), per_level(self.n_post)

    @property
    def diag(self) -> float:  # 1/scale[0], S6/params.h:76,82
        return (2.0 + self.m) if self.stencil == "wilson" else (4.0 + self.m)


# --------------------------------------------------------------------------------------------------
# lattice index helpers
# --------------------------------------------------------------------------------------------------
_NBR_CACHE: dict = {}
_FRONT_CACHE: dict = {}
_AGG_CACHE: dict = {}


def neighbours(L: int):
    """xp, xm, yp, ym site-index arrays (S6/level.h:71-74)."""
    if L not in _NBR_CACHE:
        s = np.arange(L * L)
        x, y = s % L, s // L
        _NBR_CACHE[L] = (
            (x + 1) % L + y * L,
            (x - 1 + L) % L + y * L,
            x + ((y + 1) % L) * L,
            x + ((y - 1 + L) % L) * L,
        )
    return _NBR_CACHE[L]


def fronts(L: int):
    """Anti-diagonal wavefronts x+y=c, c=0..2L-2 (see module docstring)."""
    if L not in _FRONT_CACHE:
        out = []
        for c in range(2 * L - 1):
            x = np.arange(max(0, c - L + 1), min(c, L - 1) + 1)
            out.append(x + (c - x) * L)
        _FRONT_CACHE[L] = out
    return _FRONT_CACHE[L]


def base_site(quad: int, xc, yc, Lf: int, block: int):
    """f_get_base_site, S6/modules_indiv.h:6-14."""
    if quad == 1:
        return block * xc, block * yc
    if quad == 2:
        return (block * xc - 1 + Lf) % Lf, block * yc
    if quad == 3:
        return (block * xc - 1 + Lf) % Lf, (block * yc - 1 + Lf) % Lf
    if quad == 4:
        return block * xc, (block * yc - 1 + Lf) % Lf
    raise ValueError("Invalid quad: must be 1-4")


def aggregates(Lf: int, Lc: int, block: int, quad: int) -> np.ndarray:
    """agg[X, x1*block+y1] = fine site of coarse site X=xc+yc*Lc (loop order x1 outer, y1 inner as in
    S6/modules_indiv.h:111-115, S6/near_null.h:231-236)."""
    key = (Lf, Lc, block, quad)
    if key not in _AGG_CACHE:
        X = np.arange(Lc * Lc)
        xc, yc = X % Lc, X // Lc
        bx, by = base_site(quad, xc, yc, Lf, block)
        cols = []
        for x1 in range(block):
            for y1 in range(block):
                cols.append((bx + x1) % Lf + ((by + y1) % Lf) * Lf)
        _AGG_CACHE[key] = np.stack(cols, axis=1)
    return _AGG_CACHE[key]


def mv(A: np.ndarray, v: np.ndarray) -> np.ndarray:
    """batched A[s] @ v[s]."""
    return np.einsum("sij,sj->si", A, v)


# --------------------------------------------------------------------------------------------------
# Gauge links (S6/gauge.h)
# --------------------------------------------------------------------------------------------------
def gauge_cold(L: int) -> np.ndarray:
    return np.ones((L * L, 2), dtype=C128)  # S6/gauge.h:29-37


def gauge_from_phases(theta: np.ndarray) -> np.ndarray:
    """theta[L*L, 2] -> U = exp(i theta) (std::polar(1, phase), S6/gauge.h:106)."""
    return np.exp(1j * theta).astype(C128)


def gauge_gaussian(L: int, width: float = 0.2, seed: int = 1234) -> np.ndarray:
    """The reference's commented-out option: local phase ~ N(0, width) (S6/gauge.h:25-26,36)."""
    rng = np.random.default_rng(seed + L)
    return gauge_from_phases(rng.normal(0.0, width, size=(L * L, 2)))


def gauge_quenched_phases(L: int, beta: float, sweeps: int = 200, seed: int = 1234) -> np.ndarray:
    """Compact-U(1) Wilson-action checkerboard Metropolis.  The reference reads such configurations from
    files it does not ship (S6/gauge.h:44, beta in {6, 32}: S5L/mgrid_laplace.cpp:135, S6/params.h:66)."""
    rng = np.random.default_rng(seed + L)
    th = np.zeros((L, L, 2))  # th[y, x, dir]
    yy, xx = np.meshgrid(np.arange(L), np.arange(L), indexing="ij")
    delta = min(np.pi, 2.0 / math.sqrt(beta))

    def staple_action(th, mu, trial):
        # sum over the two plaquettes containing link (x, mu)
        nu = 1 - mu
        sh = lambda a, d, k: np.roll(a, -k, axis=1 if d == 0 else 0)  # value at x + k*d_hat
        t_mu = trial
        t_nu = th[..., nu]
        p_up = t_mu + sh(t_nu, mu, 1) - sh(th[..., mu], nu, 1) - t_nu
        p_dn = sh(t_nu, nu, -1) + t_mu - sh(sh(t_nu, nu, -1), mu, 1) - sh(th[..., mu], nu, -1)
        return -beta * (np.cos(p_up) + np.cos(p_dn))

    for _ in range(sweeps):
        for mu in (0, 1):
            for par in (0, 1):
                mask = ((xx + yy) % 2) == par
                old = th[..., mu]
                new = old + rng.uniform(-delta, delta, size=old.shape)
                dS = staple_action(th, mu, new) - staple_action(th, mu, old)
                acc = mask & (rng.random(old.shape) < np.exp(-dS))
                th[..., mu] = np.where(acc, new, old)
    th = (th + np.pi) % (2 * np.pi) - np.pi
    return th.reshape(L * L, 2)


def gauge_quenched_phases_counter(L: int, beta: float, sweeps: int = 200, seed: int = 1234) -> np.ndarray:
    """The same checkerboard Metropolis with the counter-based generator (stream 2*tag / 2*tag+1 at index s for the
    proposal / acceptance draw of half-update tag = (sweep*2 + mu)*2 + parity): the algorithm mg2d_gauge_metropolis
    runs on the device, draw for draw."""
    th = np.zeros((L, L, 2))  # th[y, x, dir]
    yy, xx = np.meshgrid(np.arange(L), np.arange(L), indexing="ij")
    delta = min(np.pi, 2.0 / math.sqrt(beta))
    sh = lambda a, d, k: np.roll(a, -k, axis=1 if d == 0 else 0)  # value at x + k*d_hat

    def action(th, mu, t_mu):
        nu = 1 - mu
        t_nu, o_mu = th[..., nu], th[..., mu]
        p_up = t_mu + sh(t_nu, mu, 1) - sh(o_mu, nu, 1) - t_nu
        p_dn = sh(t_nu, nu, -1) + t_mu - sh(sh(t_nu, nu, -1), mu, 1) - sh(o_mu, nu, -1)
        return -beta * (np.cos(p_up) + np.cos(p_dn))

    for sw in range(sweeps):
        for mu in (0, 1):
            for par in (0, 1):
                tag = (sw * 2 + mu) * 2 + par
                mask = ((xx + yy) % 2) == par
                old = th[..., mu]
                u1 = counter_uniform(seed, 2 * tag, 0, L * L).reshape(L, L)
                u2 = counter_uniform(seed, 2 * tag + 1, 0, L * L).reshape(L, L)
                new = old + (2.0 * u1 - 1.0) * delta
                dS = action(th, mu, new) - action(th, mu, old)
                acc = mask & (u2 < np.exp(-dS))
                th[..., mu] = np.where(acc, new, old)
    th = (th + np.pi) % (2 * np.pi) - np.pi
    return th.reshape(L * L, 2)


def plaquette(U: np.ndarray, L: int) -> complex:
    """Gauge::f_plaquette, S6/gauge.h:50-63."""
    xp, _, yp, _ = neighbours(L)
    return complex(np.mean(U[:, 0] * U[xp, 1] * np.conj(U[yp, 0]) * np.conj(U[:, 1])))


def write_phase_file(path: str, theta: np.ndarray, L: int) -> None:
    """`phase_{L}_b{beta}.dat`: one phase per line, x outer / y inner / dir inner (S6/gauge.h:103-107)."""
    with open(path, "w") as f:
        for x in range(L):
            for y in range(L):
                for d in range(2):
                    f.write("%.17g\n" % theta[x + L * y, d])


def read_phase_file(path: str, L: int) -> np.ndarray:
    vals = np.loadtxt(path).reshape(L, L, 2)  # [x, y, dir]
    return np.ascontiguousarray(vals.transpose(1, 0, 2)).reshape(L * L, 2)


# --------------------------------------------------------------------------------------------------
# Level (S6/level.h) + Near_null (S6/near_null.h)
# --------------------------------------------------------------------------------------------------
class Level:
    """class Level : Near_null -- phi, r, D, phi_null (S6/level.h:3-14, S6/near_null.h:10-12)."""

    def __init__(self):
        self.phi = None
        self.r = None
        self.D = None
        self.phi_null = None

    # f_init_level, S6/level.h:42-53 (+ f_init_vectors / f_init_matrix / f_init_near_null_vector,
    # S6/modules_indiv.h:16-68).  `gen` None -> rand=0 (all ones)
    def init_level(self, lvl: int, gen: StdMT19937 | None, p: Params):
        S, n = p.size[lvl] ** 2, p.n_dof[lvl]
        draw = (lambda k: gen.uniform_pm_pi(k)) if gen is not None else (lambda k: np.ones(k))
        self.phi = draw(S * n).reshape(S, n).astype(C128)
        self.r = draw(S * n).reshape(S, n).astype(C128)
        self.D = np.ones((S, 5, n, n), dtype=C128)
        if lvl != p.nlevels:
            nc = p.n_dof[lvl + 1]
            self.phi_null = draw(S * nc * n).reshape(S, nc, n).astype(C128)

    def define_source(self, p: Params):
        self.r[2 + 2 * p.L, 0] = 5.0  # S6/level.h:57

    # f_compute_lvl0_matrix, S6/level.h:131-175
    def compute_lvl0_matrix(self, U: np.ndarray, p: Params):
        L = p.size[0]
        _, xm, _, ym = neighbours(L)
        S = L * L
        if p.stencil == "laplace":
            D = np.zeros((S, 5, 1, 1), dtype=C128)
            D[:, 0, 0, 0] = -p.diag
            D[:, 1, 0, 0] = U[:, 0]
            D[:, 2, 0, 0] = np.conj(U[xm, 0])
            D[:, 3, 0, 0] = U[:, 1]
            D[:, 4, 0, 0] = np.conj(U[ym, 1])
        else:
            g1 = np.array([[0, 1], [1, 0]], dtype=C128)
            g2 = np.array([[0, -1j], [1j, 0]], dtype=C128)
            I2 = np.eye(2, dtype=C128)
            D = np.zeros((S, 5, 2, 2), dtype=C128)
            D[:, 0] = p.diag * I2
            D[:, 1] = U[:, 0, None, None] * (0.5 * (I2 - g1))
            D[:, 2] = np.conj(U[xm, 0])[:, None, None] * (0.5 * (I2 + g1))
            D[:, 3] = U[:, 1, None, None] * (0.5 * (I2 - g2))
            D[:, 4] = np.conj(U[ym, 1])[:, None, None] * (0.5 * (I2 + g2))
        self.D = D

    # f_apply_D, S6/level.h:251-265 (summation order D1,D2,D3,D4,D0)
    def apply_D(self, v: np.ndarray, L: int) -> np.ndarray:
        xp, xm, yp, ym = neighbours(L)
        D = self.D
        return mv(D[:, 1], v[xp]) + mv(D[:, 2], v[xm]) + mv(D[:, 3], v[yp]) + mv(D[:, 4], v[ym]) + mv(D[:, 0], v)

    # f_residue, S6/level.h:61-77
    def residue(self, L: int) -> np.ndarray:
        return self.r - self.apply_D(self.phi, L)

    # f_get_residue_mag, S6/level.h:79-98
    def get_residue_mag(self, L: int) -> float:
        res = np.sum(np.abs(self.residue(L)) ** 2)
        bnorm = np.sum(np.abs(self.r) ** 2)
        return math.sqrt(res) / math.sqrt(bnorm)

    # f_relax, S6/level.h:100-128
    def relax(self, L: int, num_iter: int, gs_flag: int):
        xp, xm, yp, ym = neighbours(L)
        D, phi, r = self.D, self.phi, self.r
        D0inv = -np.linalg.inv(D[:, 0])
        if gs_flag == 1:
            fr = fronts(L)
            for _ in range(num_iter):
                for idx in fr:
                    acc = (mv(D[idx, 1], phi[xp[idx]]) + mv(D[idx, 2], phi[xm[idx]])
                           + mv(D[idx, 3], phi[yp[idx]]) + mv(D[idx, 4], phi[ym[idx]]) - r[idx])
                    phi[idx] = mv(D0inv[idx], acc)
        else:
            for _ in range(num_iter):
                acc = (mv(D[:, 1], phi[xp]) + mv(D[:, 2], phi[xm]) + mv(D[:, 3], phi[yp])
                       + mv(D[:, 4], phi[ym]) - r)
                phi[:] = mv(D0inv, acc)

    # minimal-residual smoother (ours; north_star).  res = r - D phi; repeat: t = D res;
    # alpha = <t,res>/<t,t>; phi += w*alpha*res; res -= w*alpha*t
    def relax_mr(self, L: int, num_iter: int, omega: float = 1.0):
        res = self.residue(L)
        for _ in range(num_iter):
            t = self.apply_D(res, L)
            tt = np.sum(np.abs(t) ** 2)
            if tt == 0.0:
                break
            alpha = omega * (np.vdot(t, res) / tt)
            self.phi += alpha * res
            res -= alpha * t

    # red-black Gauss-Seidel (ours): the parallel ordering of f_relax's update rule; colour (x+y)%2 == 0
    # first, then colour 1, each colour updated from the latest values of the other
    def relax_rb(self, L: int, num_iter: int):
        xp, xm, yp, ym = neighbours(L)
        D, phi, r = self.D, self.phi, self.r
        D0inv = -np.linalg.inv(D[:, 0])
        s = np.arange(L * L)
        par = ((s % L) + (s // L)) % 2
        colours = [np.nonzero(par == c)[0] for c in (0, 1)]
        for _ in range(num_iter):
            for idx in colours:
                acc = (mv(D[idx, 1], phi[xp[idx]]) + mv(D[idx, 2], phi[xm[idx]])
                       + mv(D[idx, 3], phi[yp[idx]]) + mv(D[idx, 4], phi[ym[idx]]) - r[idx])
                phi[idx] = mv(D0inv[idx], acc)

    def smooth(self, L: int, num_iter: int, p: Params):
        if p.smoother == "mr":
            self.relax_mr(L, num_iter, p.mr_omega)
        elif p.smoother == "rbgs":
            self.relax_rb(L, num_iter)
        else:
            self.relax(L, num_iter, 1 if p.smoother == "gs" else 0)

    # f_near_null, S6/level.h:177-249 (+ f_g_norm, S6/modules_indiv.h:70-92)
    def near_null(self, level: int, p: Params):
        L = p.size[level]
        nf, nc = p.n_dof[level], p.n_dof[level + 1]
        num = max(p.null_iters // p.null_chunk, 1)
        tmp = Level()
        tmp.D = self.D
        tmp.r = np.zeros((L * L, nf), dtype=C128)
        nvec = nc if p.stencil == "laplace" else nc // 2
        for d1 in range(nvec):
            tmp.phi = self.phi_null[:, d1, :].copy()
            for _ in range(num):
                tmp.smooth(L, p.null_chunk, p)
                g_norm = math.sqrt(np.sum(np.abs(tmp.phi) ** 2))
                if math.isnan(g_norm):
                    raise FloatingPointError("gnorm is nan")
                tmp.phi /= g_norm
            v = np.conj(tmp.phi)
            if p.stencil == "laplace":
                self.phi_null[:, d1, :] = v
            else:
                h = nf // 2
                self.phi_null[:, d1, :h] = v[:, :h]
                self.phi_null[:, d1, h:] = 0.0
                self.phi_null[:, nc // 2 + d1, h:] = v[:, h:]
                self.phi_null[:, nc // 2 + d1, :h] = 0.0

    # f_block_norm (S6/modules_indiv.h:94-135) applied to every row: f_norm_nn, S6/near_null.h:24-48
    def norm_nn(self, level: int, quad: int, p: Params):
        agg = aggregates(p.size[level], p.size[level + 1], p.blocks[level], quad)
        for d1 in range(p.n_dof[level + 1]):
            self.phi_null[:, d1, :] = _block_norm(self.phi_null[:, d1, :], agg)

    # f_ortho, S6/near_null.h:97-173 (divides by norm, not norm^2: :165)
    def ortho(self, level: int, quad: int, p: Params):
        agg = aggregates(p.size[level], p.size[level + 1], p.blocks[level], quad)
        P = self.phi_null
        for d1 in range(p.n_dof[level + 1]):
            t = P[:, d1, :].copy()
            for d2 in range(d1):
                u = P[:, d2, :]
                ua, ta = u[agg], t[agg]                      # [Lc^2, b^2, nf]
                nrm = np.sqrt(np.sum(np.abs(ua) ** 2, axis=(1, 2)))
                dot = np.sum(np.conj(ua) * ta, axis=(1, 2))
                if np.any(np.isnan(nrm)) or np.any(nrm < 1e-8):
                    raise FloatingPointError("Inside ortho: bad norm")
                t[agg] = ta - (dot / nrm)[:, None, None] * ua
            P[:, d1, :] = _block_norm(t, agg)

    # f_check_ortho, S6/near_null.h:175-214 -> returns the largest |<null_d1, null_d2>| over aggregates
    def check_ortho(self, level: int, quad: int, p: Params) -> float:
        agg = aggregates(p.size[level], p.size[level + 1], p.blocks[level], quad)
        Pa = self.phi_null[agg]                               # [Lc^2, b^2, nc, nf]
        G = np.einsum("Xbif,Xbjf->Xij", np.conj(Pa), Pa)
        worst = 0.0
        for d1 in range(G.shape[1]):
            for d2 in range(d1):
                worst = max(worst, float(np.max(np.abs(G[:, d1, d2]))))
        return worst

    # f_restriction, S6/near_null.h:217-240
    def restriction(self, vec_f: np.ndarray, level: int, p: Params, quad: int) -> np.ndarray:
        agg = aggregates(p.size[level], p.size[level + 1], p.blocks[level], quad)
        Pv = mv(self.phi_null, vec_f)                         # [Lf^2, nc]
        out = np.zeros((agg.shape[0], self.phi_null.shape[1]), dtype=C128)
        for b in range(agg.shape[1]):
            out += Pv[agg[:, b]]
        return out

    # f_prolongation, S6/near_null.h:242-264: vec_f += P^dagger vec_c   (level = COARSE level index)
    def prolongation(self, vec_f: np.ndarray, vec_c: np.ndarray, level: int, p: Params, quad: int):
        agg = aggregates(p.size[level - 1], p.size[level], p.blocks[level - 1], quad)
        for b in range(agg.shape[1]):
            s = agg[:, b]
            vec_f[s] += np.einsum("sij,si->sj", np.conj(self.phi_null[s]), vec_c)


def _block_norm(vec: np.ndarray, agg: np.ndarray) -> np.ndarray:
    """f_block_norm, S6/modules_indiv.h:94-135."""
    va = vec[agg]
    norm = np.sqrt(np.sum(np.abs(va) ** 2, axis=(1, 2)))
    if np.any(np.isnan(norm)):
        raise FloatingPointError("Inside block_norm: nan")
    if np.any(norm < 1e-40):
        raise FloatingPointError("Inside block_norm: very small norm")
    out = vec.copy()
    out[agg] = va / norm[:, None, None]
    return out


# --------------------------------------------------------------------------------------------------
# modules_main.h
# --------------------------------------------------------------------------------------------------
def compute_coarse_matrix(Df: np.ndarray, phi_null: np.ndarray, level: int, quad: int, p: Params) -> np.ndarray:
    """f_compute_coarse_matrix, S6/modules_main.h:81-185.  D_c = P D_f P^dagger."""
    Lf, Lc, blk = p.size[level], p.size[level + 1], p.blocks[level]
    nc = p.n_dof[level + 1]
    agg = aggregates(Lf, Lc, blk, quad)
    xp, xm, yp, ym = neighbours(Lf)
    Dc = np.zeros((Lc * Lc, 5, nc, nc), dtype=C128)
    Pd = np.conj(phi_null).transpose(0, 2, 1)                 # P(s)^dagger  [Lf^2, nf, nc]

    def term(s, k, sp):
        return phi_null[s] @ Df[s, k] @ Pd[sp]

    for x1 in range(blk):
        for y1 in range(blk):
            s = agg[:, x1 * blk + y1]
            Dc[:, 0] += term(s, 0, s)
            # intra-block hops go to the diagonal block (:134-144), faces to the hop blocks (:148-155)
            Dc[:, 0 if x1 != blk - 1 else 1] += term(s, 1, xp[s])
            Dc[:, 0 if x1 != 0 else 2] += term(s, 2, xm[s])
            Dc[:, 0 if y1 != blk - 1 else 3] += term(s, 3, yp[s])
            Dc[:, 0 if y1 != 0 else 4] += term(s, 4, ym[s])
    return Dc


def init_NTL(NTL, p: Params, gen: StdMT19937 | None):
    """f_init_NTL, S6/modules_main.h:7-37.  NTL[lvl][q] are Level objects."""
    if p.t_flag != 0 and p.nlevels > 0:
        lvl = p.nlevels - 1
        S, n, nc = p.size[lvl] ** 2, p.n_dof[lvl], p.n_dof[lvl + 1]
        draw = (lambda k: gen.uniform_pm_pi(k)) if gen is not None else (lambda k: np.ones(k))
        for q in range(p.n_copies):
            NTL[lvl][q].phi = draw(S * n).reshape(S, n).astype(C128)
            NTL[lvl][q].r = draw(S * n).reshape(S, n).astype(C128)
            NTL[lvl][q].phi_null = draw(S * nc * n).reshape(S, nc, n).astype(C128)
        lvl = p.nlevels
        S, n = p.size[lvl] ** 2, p.n_dof[lvl]
        for q in range(p.n_copies):
            NTL[lvl][q].phi = draw(S * n).reshape(S, n).astype(C128)
            NTL[lvl][q].r = draw(S * n).reshape(S, n).astype(C128)
            NTL[lvl][q].D = np.ones((S, 5, n, n), dtype=C128)


def compute_near_null(LVL, NTL, p: Params, quad: int, gen_null: int = 1):
    """f_compute_near_null, S6/modules_main.h:187-222 (gen_null=0: phi_null already supplied)."""
    for lvl in range(p.nlevels):
        if gen_null == 1:
            LVL[lvl].near_null(lvl, p)
        LVL[lvl].norm_nn(lvl, quad, p)
        LVL[lvl].ortho(lvl, quad, p)
        LVL[lvl].ortho(lvl, quad, p)
        LVL[lvl].check_ortho(lvl, quad, p)
        LVL[lvl + 1].D = compute_coarse_matrix(LVL[lvl].D, LVL[lvl].phi_null, lvl, quad, p)
    if p.t_flag == 1:
        lo = p.nlevels - 1
        for q in range(p.n_copies):
            NTL[lo][q].phi_null = LVL[lo].phi_null.copy()
            NTL[lo][q].norm_nn(lo, quad, p)
            NTL[lo][q].ortho(lo, q + 1, p)
            NTL[lo][q].ortho(lo, q + 1, p)
            NTL[lo][q].check_ortho(lo, q + 1, p)
            NTL[p.nlevels][q].D = compute_coarse_matrix(LVL[lo].D, NTL[lo][q].phi_null, lo, q + 1, p)


def restriction_res(L_residue: Level, L_restrict: Level, level: int, p: Params, quad: int) -> np.ndarray:
    """f_restriction_res, S6/modules_main.h:224-241."""
    return L_restrict.restriction(L_residue.residue(p.size[level]), level, p, quad)


def prolongate_phi(phi_f: np.ndarray, phi_c: np.ndarray, LVLP: Level, level: int, p: Params, quad: int):
    """f_prolongate_phi, S6/modules_main.h:243-252."""
    LVLP.prolongation(phi_f, phi_c, level, p, quad)
    phi_c[:] = 0.0


def MG_simple(LVL, p: Params):
    """f_MG_simple, S6/modules_main.h:255-280."""
    if p.nlevels > 0:
        for lvl in range(p.nlevels):
            LVL[lvl].smooth(p.size[lvl], p.pre[lvl], p)
            LVL[lvl + 1].r = restriction_res(LVL[lvl], LVL[lvl], lvl, p, p.quad)
        for lvl in range(p.nlevels, -1, -1):
            LVL[lvl].smooth(p.size[lvl], p.post[lvl], p)
            if lvl > 0:
                prolongate_phi(LVL[lvl - 1].phi, LVL[lvl].phi, LVL[lvl - 1], lvl, p, p.quad)
    else:
        LVL[0].smooth(p.size[0], p.post[0], p)


def _coarse_gcr(LVL, p: Params, lvl: int):
    """K-cycle coarse solve (ours; mirrored 1:1 by the CUDA driver): p.k_inner flexible-GCR steps on D_lvl phi = r_lvl from
    phi = 0, each preconditioned by MG_kcycle(lvl); classical Gram-Schmidt; fixed step count (no residual test)."""
    lv, L = LVL[lvl], p.size[lvl]
    x = np.zeros_like(lv.r)
    Z, W = [], []
    keep = lv.phi
    for _ in range(p.k_inner):
        lv.phi = np.zeros_like(lv.r)
        MG_kcycle(LVL, p, lvl)
        z = lv.phi
        w = lv.apply_D(z, L)
        betas = [np.vdot(wj, w) / np.sum(np.abs(wj) ** 2) for wj in W]
        for zj, wj, beta in zip(Z, W, betas):
            w = w - beta * wj
            z = z - beta * zj
        alpha = np.vdot(w, lv.r) / np.sum(np.abs(w) ** 2)
        x = x + alpha * z
        lv.r = lv.r - alpha * w
        Z.append(z)
        W.append(w)
    lv.phi = keep
    lv.phi[:] = x


def MG_kcycle(LVL, p: Params, lvl: int = 0):
    """One K-cycle on level lvl (ours): f_MG_simple's structure (S6/modules_main.h:255-280) with the single recursive visit
    of the next level replaced by p.k_inner Krylov steps preconditioned by that level's K-cycle; the coarsest level is only
    relaxed (as in the reference)."""
    lv = LVL[lvl]
    if lvl == p.nlevels:
        lv.smooth(p.size[lvl], p.post[lvl], p)
        return
    lv.smooth(p.size[lvl], p.pre[lvl], p)
    LVL[lvl + 1].r = restriction_res(lv, lv, lvl, p, p.quad)
    if lvl + 1 == p.nlevels:
        MG_kcycle(LVL, p, lvl + 1)
    else:
        _coarse_gcr(LVL, p, lvl + 1)
    prolongate_phi(lv.phi, LVL[lvl + 1].phi, lv, lvl + 1, p, p.quad)
    lv.smooth(p.size[lvl], p.post[lvl], p)


def cycle_once(LVL, NTL, p: Params):
    if p.t_flag == 1 and p.nlevels > 0:
        return MG_ntl(LVL, NTL, p)
    if p.cycle == "K" and p.nlevels > 0:
        MG_kcycle(LVL, p, 0)
    else:
        MG_simple(LVL, p)
    return None


def colpiv_householder_qr_solve(A: np.ndarray, b: np.ndarray) -> np.ndarray:
    """x = A.colPivHouseholderQr().solve(b) (S6/modules_main.h:371).  Eigen is absent from the reference tree
    (un-vendored, version-unpinned header dependency); this restates its published algorithm: Householder QR
    with column pivoting on the largest remaining column norm, rank decided by |R_kk| > eps*n*max|R_kk|,
    back-substitution on the leading rank x rank block, remaining unknowns zero."""
    A = np.array(A, dtype=C128)
    n = A.shape[0]
    c = np.array(b, dtype=C128)
    perm = list(range(n))
    diag = np.zeros(n)
    for k in range(n):
        norms = np.sum(np.abs(A[k:, k:]) ** 2, axis=0)
        j = k + int(np.argmax(norms))
        if j != k:
            A[:, [k, j]] = A[:, [j, k]]
            perm[k], perm[j] = perm[j], perm[k]
        x = A[k:, k].copy()
        alpha = np.linalg.norm(x)
        if alpha == 0.0:
            diag[k] = 0.0
            continue
        phase = x[0] / abs(x[0]) if abs(x[0]) != 0 else 1.0
        v = x.copy()
        v[0] += phase * alpha
        v /= np.linalg.norm(v)
        A[k:, k:] -= 2.0 * np.outer(v, np.conj(v) @ A[k:, k:])
        c[k:] -= 2.0 * v * (np.conj(v) @ c[k:])
        diag[k] = abs(A[k, k])
    thresh = np.finfo(float).eps * n * (diag.max() if n else 0.0)
    rank = int(np.sum(diag > thresh))
    y = np.zeros(n, dtype=C128)
    for i in range(rank - 1, -1, -1):
        y[i] = (c[i] - A[i, i + 1:rank] @ y[i + 1:rank]) / A[i, i]
    x = np.zeros(n, dtype=C128)
    for i in range(n):
        x[perm[i]] = y[i]
    return x


def min_res(LVL, NTL, num_copies: int, level: int, p: Params) -> np.ndarray:
    """f_min_res, S6/modules_main.h:283-373.  NB: uses the level RHS r, not the residual (:339,:361)."""
    L = p.size[level]
    t = [LVL[level].apply_D(NTL[level][q].phi, L) for q in range(num_copies)]
    A = np.zeros((num_copies, num_copies), dtype=C128)
    src = np.zeros(num_copies, dtype=C128)
    for q1 in range(num_copies):
        for q2 in range(num_copies):
            A[q1, q2] = np.vdot(NTL[level][q1].phi, t[q2])
    for q1 in range(num_copies):
        if p.stencil == "laplace":
            src[q1] = np.vdot(NTL[level][q1].phi, LVL[level].r)
        else:
            src[q1] = np.vdot(LVL[level].r, t[q1])
    return colpiv_householder_qr_solve(A, src)


def scale_phi(L1: Level, NTL, a_copy, num_copies: int, lvl: int):
    """f_scale_phi, S6/modules_main.h:375-384."""
    for q in range(num_copies):
        L1.phi += a_copy[q] * NTL[lvl][q].phi
        NTL[lvl][q].phi[:] = 0.0


def MG_ntl(LVL, NTL, p: Params) -> np.ndarray:
    """f_MG_ntl, S6/modules_main.h:386-439.  Returns the 4 copy weights."""
    a_copy = np.zeros(4, dtype=C128)
    for lvl in range(p.nlevels):
        LVL[lvl].smooth(p.size[lvl], p.num_iters, p)
        if lvl != p.nlevels - 1:
            LVL[lvl + 1].r = restriction_res(LVL[lvl], LVL[lvl], lvl, p, p.quad)
        else:
            for q in range(p.n_copies):
                NTL[lvl + 1][q].r = restriction_res(LVL[lvl], NTL[lvl][q], lvl, p, q + 1)
    for lvl in range(p.nlevels, -1, -1):
        if lvl == p.nlevels:
            for q in range(p.n_copies):
                NTL[lvl][q].smooth(p.size[lvl], p.num_iters, p)
                prolongate_phi(NTL[lvl - 1][q].phi, NTL[lvl][q].phi, NTL[lvl - 1][q], lvl, p, q + 1)
            if p.min_res_flag == 1:
                a_copy[:p.n_copies] = min_res(LVL, NTL, p.n_copies, lvl - 1, p)
            else:
                a_copy[:p.n_copies] = 1.0 / p.n_copies
            scale_phi(LVL[lvl - 1], NTL, a_copy, p.n_copies, lvl - 1)
        else:
            LVL[lvl].smooth(p.size[lvl], p.num_iters, p)
            if lvl > 0:
                prolongate_phi(LVL[lvl - 1].phi, LVL[lvl].phi, LVL[lvl - 1], lvl, p, p.quad)
    return a_copy


def perform_MG(LVL, NTL, p: Params, record_phi: bool = False):
    """f_perform_MG, S6/modules_main.h:442-481.  Returns dict(iters, resnorms, ntl_weights, converged)."""
    info = {"iters": 0, "resnorms": [], "ntl_weights": [], "converged": False, "diverged": False, "phi_hist": []}
    for it in range(p.max_iters):
        if record_phi:
            info["phi_hist"].append(LVL[0].phi.copy())
        w = cycle_once(LVL, NTL, p)
        if w is not None:
            info["ntl_weights"].append(w)
        resmag = LVL[0].get_residue_mag(p.size[0])
        info["resnorms"].append(resmag)
        info["iters"] = it + 1
        if resmag < p.res_threshold:
            info["converged"] = True
            break
        if resmag > 1e6 or math.isnan(resmag):
            info["diverged"] = True
            break
    return info


def gcr_MG(LVL, NTL, p: Params, b: np.ndarray, x0: np.ndarray | None = None, tol: float = 1e-10,
           max_iters: int = 1000, restart: int = 8):
    """Flexible GCR(restart) with one multigrid cycle (from a zero start) as the preconditioner (ours: the
    reference only iterates the cycle stationarily, f_perform_MG).  Mirrored 1:1 by the CUDA driver.
        z = M(r); w = D z; orthogonalise (w, z) against the stored (w_j, z_j) (classical Gram-Schmidt);
        alpha = <w, r>/<w, w>; x += alpha z; r -= alpha w
    Returns (x, info)."""
    L0 = p.size[0]
    lv0 = LVL[0]
    x = np.zeros_like(b) if x0 is None else x0.copy()
    lv0.phi, lv0.r = x, b
    r = lv0.residue(L0)
    bn = math.sqrt(np.sum(np.abs(b) ** 2))
    Z, W = [], []
    info = {"iters": 0, "resnorms": [], "converged": False, "diverged": False}
    ntl = p.t_flag == 1 and p.nlevels > 0
    for it in range(max_iters):
        lv0.phi = np.zeros_like(b)
        lv0.r = r.copy()
        for lvl in range(1, p.nlevels + 1):
            LVL[lvl].phi[:] = 0.0
        cycle_once(LVL, NTL, p)
        z = lv0.phi
        w = lv0.apply_D(z, L0)
        betas = [np.vdot(wj, w) / np.sum(np.abs(wj) ** 2) for wj in W]      # classical Gram-Schmidt: all from the same w
        for zj, wj, beta in zip(Z, W, betas):
            w = w - beta * wj
            z = z - beta * zj
        alpha = np.vdot(w, r) / np.sum(np.abs(w) ** 2)
        x = x + alpha * z
        r = r - alpha * w
        Z.append(z)
        W.append(w)
        if len(Z) >= restart:
            Z, W = [], []
        resmag = math.sqrt(np.sum(np.abs(r) ** 2)) / bn
        info["resnorms"].append(resmag)
        info["iters"] = it + 1
        if resmag < tol:
            info["converged"] = True
            break
        if resmag > 1e6 or math.isnan(resmag):
            info["diverged"] = True
            break
    lv0.phi, lv0.r = x, b
    return x, info


# --------------------------------------------------------------------------------------------------
# main() flow of S6/mgrid_ntl.cpp:29-73
# --------------------------------------------------------------------------------------------------
def build_reference_problem(p: Params, U: np.ndarray, seed: int = 4302529):
    """Initial data exactly as main() draws them (RNG order: SURVEY appendix A.2)."""
    gen = StdMT19937(seed)
    LVL = [Level() for _ in range(p.nlevels + 1)]
    for lvl in range(p.nlevels + 1):
        LVL[lvl].init_level(lvl, gen, p)
    NTL = [[Level() for _ in range(4)] for _ in range(p.nlevels + 1)]
    init_NTL(NTL, p, gen)
    LVL[0].define_source(p)
    LVL[0].compute_lvl0_matrix(U, p)
    return LVL, NTL


def build_device_problem(p: Params, U: np.ndarray, seed: int = 4302529):
    """Initial data of the large-lattice mode (mirrors MG.init_fields of the CUDA path): phi = 0, r = 0, near-null seeds
    phi_null[s, ic, jf] = counter_uniform(seed, stream=level, global element index) in (-pi, pi), real."""
    LVL = [Level() for _ in range(p.nlevels + 1)]
    for lvl in range(p.nlevels + 1):
        S, n = p.size[lvl] ** 2, p.n_dof[lvl]
        LVL[lvl].phi = np.zeros((S, n), dtype=C128)
        LVL[lvl].r = np.zeros((S, n), dtype=C128)
        if lvl != p.nlevels:
            nc = p.n_dof[lvl + 1]
            LVL[lvl].phi_null = counter_uniform(seed, lvl, 0, S * nc * n, -np.pi, np.pi).reshape(S, nc, n).astype(C128)
    NTL = [[Level() for _ in range(4)] for _ in range(p.nlevels + 1)]
    LVL[0].compute_lvl0_matrix(U, p)
    return LVL, NTL


def run_reference_flow(p: Params, U: np.ndarray, seed: int = 4302529, record_phi: bool = False):
    LVL, NTL = build_reference_problem(p, U, seed)
    if p.nlevels > 0:
        compute_near_null(LVL, NTL, p, p.quad)
    info = perform_MG(LVL, NTL, p, record_phi=record_phi)
    return LVL, NTL, info


# --------------------------------------------------------------------------------------------------
# The reference's in-run property tests (S6/tests.h) as functions returning the worst violation
# --------------------------------------------------------------------------------------------------
def gamma5(n: int) -> np.ndarray:
    return np.diag(np.where(np.arange(n) < n // 2, 1.0, -1.0)).astype(C128)  # S6/tests.h:143-150


def test1_restriction_prolongation(lvlP: Level, vec: np.ndarray, level: int, p: Params, quad: int) -> float:
    """S6/tests.h:5-43: P P^dagger v_c = v_c."""
    vec_f = np.zeros((p.size[level] ** 2, p.n_dof[level]), dtype=C128)
    lvlP.prolongation(vec_f, vec, level + 1, p, quad)
    vec_c = lvlP.restriction(vec_f, level, p, quad)
    return float(np.max(np.abs(vec_c - vec)))


def test2_D(vec, lvl_c: Level, lvl_f: Level, lvl_P: Level, level: int, p: Params, quad: int) -> float:
    """S6/tests.h:46-92: D_c v = P D_f P^dagger v."""
    vec_f1 = np.zeros((p.size[level] ** 2, p.n_dof[level]), dtype=C128)
    lvl_P.prolongation(vec_f1, vec, level + 1, p, quad)
    vec_f2 = lvl_f.apply_D(vec_f1, p.size[level])
    vec_c1 = lvl_P.restriction(vec_f2, level, p, quad)
    vec_c2 = lvl_c.apply_D(vec, p.size[level + 1])
    return float(np.max(np.abs(vec_c1 - vec_c2)))


def test3_hermiticity(lvl: Level, level: int, p: Params) -> float:
    """S6/tests.h:94-182: D_1(s) = G D_2(s+x)^dagger G, D_3(s) = G D_4(s+y)^dagger G, D_0 = G D_0^dagger G."""
    L, n = p.size[level], p.n_dof[level]
    xp, _, yp, _ = neighbours(L)
    G = gamma5(n) if p.stencil == "wilson" else np.eye(n, dtype=C128)
    D = lvl.D
    dag = lambda M: np.conj(M).transpose(0, 2, 1)
    e = np.max(np.abs(D[:, 1] - G @ dag(D[xp, 2]) @ G))
    e = max(e, np.max(np.abs(D[:, 3] - G @ dag(D[yp, 4]) @ G)))
    e = max(e, np.max(np.abs(D[:, 0] - G @ dag(D[:, 0]) @ G)))
    return float(e)


def test4_hermiticity_full(lvl: Level, vec: np.ndarray, level: int, p: Params) -> float:
    """S6/tests.h:184-248: Im <v, D v> = 0 (laplace) / Im <v, D gamma5 v> = 0 (wilson)."""
    L, n = p.size[level], p.n_dof[level]
    if p.stencil == "wilson":
        tmp = Level()
        tmp.D = lvl.D @ gamma5(n)
        w = tmp.apply_D(vec, L)
    else:
        w = lvl.apply_D(vec, L)
    return float(abs(np.vdot(vec, w).imag))


def dense_matrix(lvl: Level, L: int, n: int) -> np.ndarray:
    """Dense (L*L*n)^2 matrix of the stencil, for spectra on small lattices."""
    S = L * L
    M = np.zeros((S * n, S * n), dtype=C128)
    nb = neighbours(L)
    for s in range(S):
        M[s * n:(s + 1) * n, s * n:(s + 1) * n] += lvl.D[s, 0]
        for k in range(4):
            sp = nb[k][s]
            M[s * n:(s + 1) * n, sp * n:(sp + 1) * n] += lvl.D[s, k + 1]
    return M
