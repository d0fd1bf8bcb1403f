/* mg_port.c -- plain C + OpenMP restatement of the solve loop of the reference's adaptive multigrid -- TEST INFRASTRUCTURE.
 *
 * Only bench.py's cpu_baseline / --impl reference legs and tests/ may use it (oracle/__init__.py).  It restates, for the
 * cycle shape the bench runs, the same functions as oracle/mg_oracle.py (which cites them line by line):
 *   Level::f_apply_D   S6/level.h:251-265      Level::f_relax (update rule, red-black order)  S6/level.h:100-128
 *   f_restriction      S6/near_null.h:217-240  f_prolongation + zeroing  S6/near_null.h:242-264, S6/modules_main.h:243-252
 *   f_MG_simple        S6/modules_main.h:255-280 (per-level pre/post sweep counts), wrapped in the flexible GCR(restart)
 *   of oracle.mg_oracle.gcr_MG (classical Gram-Schmidt).
 * The reference is single-threaded C++; this port uses every host thread (OpenMP over sites) so that the CPU number beside
 * the GPU number is not handicapped by numpy.  Layouts are the reference's: D[s][k][i][j], P[s][ic][jf], fields [s][n].
 * The hierarchy (P, D_c, -D0^-1) is built by the numpy oracle and handed over; setup is not timed.
 */
#include <complex.h>
#include <math.h>
#include <omp.h>
#include <stdlib.h>
#include <string.h>

typedef double complex cplx;

typedef struct {
    int L, n, nc;          /* lattice size, dof, dof of the next coarser level (0 on the coarsest) */
    const cplx* D;         /* [S][5][n][n] */
    const cplx* mD0inv;    /* [S][n][n] = -inverse(D0) */
    const cplx* P;         /* [S][nc][n], NULL on the coarsest level */
    cplx* phi;             /* [S][n] */
    cplx* r;               /* [S][n] */
    cplx* tmp;             /* [S][n] */
} level_t;

static inline int nbr(int L, int x, int y, int k) {
    switch (k) {
        case 1: return ((x + 1) % L) + y * L;
        case 2: return ((x - 1 + L) % L) + y * L;
        case 3: return x + ((y + 1) % L) * L;
        default: return x + ((y - 1 + L) % L) * L;
    }
}

/* out = D v (MODE 0) or out = b - D v (MODE 1); summation order of S6/level.h:258-262 */
static void apply_D(const level_t* lv, const cplx* v, const cplx* b, cplx* out) {
    const int L = lv->L, n = lv->n;
#pragma omp parallel for schedule(static)
    for (int s = 0; s < L * L; ++s) {
        const int x = s % L, y = s / L;
        const cplx* Ds = lv->D + (size_t)s * 5 * n * n;
        for (int i = 0; i < n; ++i) {
            cplx acc = 0;
            for (int k = 1; k <= 4; ++k) {
                const cplx* vn = v + (size_t)nbr(L, x, y, k) * n;
                const cplx* row = Ds + (size_t)k * n * n + (size_t)i * n;
                for (int j = 0; j < n; ++j) acc += row[j] * vn[j];
            }
            const cplx* row0 = Ds + (size_t)i * n;
            for (int j = 0; j < n; ++j) acc += row0[j] * v[(size_t)s * n + j];
            out[(size_t)s * n + i] = b ? b[(size_t)s * n + i] - acc : acc;
        }
    }
}

/* red-black ordering of phi(s) = -D0^-1 (sum_k D_k phi(s+d_k) - r(s)): colour (x+y)%2 == 0 first */
static void relax_rb(level_t* lv, int nsweeps) {
    const int L = lv->L, n = lv->n;
    for (int it = 0; it < nsweeps; ++it)
        for (int colour = 0; colour < 2; ++colour) {
#pragma omp parallel for schedule(static)
            for (int y = 0; y < L; ++y) {
                cplx acc[64];
                for (int x = (y + colour) & 1; x < L; x += 2) {
                    const int s = x + y * L;
                    const cplx* Ds = lv->D + (size_t)s * 5 * n * n;
                    for (int i = 0; i < n; ++i) {
                        cplx a = 0;
                        for (int k = 1; k <= 4; ++k) {
                            const cplx* vn = lv->phi + (size_t)nbr(L, x, y, k) * n;
                            const cplx* row = Ds + (size_t)k * n * n + (size_t)i * n;
                            for (int j = 0; j < n; ++j) a += row[j] * vn[j];
                        }
                        acc[i] = a - lv->r[(size_t)s * n + i];
                    }
                    const cplx* Is = lv->mD0inv + (size_t)s * n * n;
                    for (int i = 0; i < n; ++i) {
                        cplx o = 0;
                        for (int j = 0; j < n; ++j) o += Is[(size_t)i * n + j] * acc[j];
                        lv->phi[(size_t)s * n + i] = o;
                    }
                }
            }
        }
}

/* vc(X) = sum_{s in agg(X)} P(s) vf(s), quadrant 1, x1 outer / y1 inner */
static void restrict_(const level_t* fine, const cplx* vf, cplx* vc, int block) {
    const int Lf = fine->L, Lc = Lf / block, nf = fine->n, nc = fine->nc;
#pragma omp parallel for schedule(static)
    for (int X = 0; X < Lc * Lc; ++X) {
        const int xc = X % Lc, yc = X / Lc;
        cplx* o = vc + (size_t)X * nc;
        for (int i = 0; i < nc; ++i) o[i] = 0;
        for (int x1 = 0; x1 < block; ++x1)
            for (int y1 = 0; y1 < block; ++y1) {
                const int s = (block * xc + x1) + (block * yc + y1) * Lf;
                const cplx* Ps = fine->P + (size_t)s * nc * nf;
                for (int i = 0; i < nc; ++i) {
                    cplx a = 0;
                    for (int j = 0; j < nf; ++j) a += Ps[(size_t)i * nf + j] * vf[(size_t)s * nf + j];
                    o[i] += a;
                }
            }
    }
}

/* vf(s) += P(s)^dagger vc(X(s)); vc = 0 */
static void prolong_add(const level_t* fine, cplx* vf, cplx* vc, int block) {
    const int Lf = fine->L, Lc = Lf / block, nf = fine->n, nc = fine->nc;
#pragma omp parallel for schedule(static)
    for (int X = 0; X < Lc * Lc; ++X) {
        const int xc = X % Lc, yc = X / Lc;
        const cplx* c = vc + (size_t)X * nc;
        for (int x1 = 0; x1 < block; ++x1)
            for (int y1 = 0; y1 < block; ++y1) {
                const int s = (block * xc + x1) + (block * yc + y1) * Lf;
                const cplx* Ps = fine->P + (size_t)s * nc * nf;
                for (int j = 0; j < nf; ++j) {
                    cplx a = 0;
                    for (int i = 0; i < nc; ++i) a += conj(Ps[(size_t)i * nf + j]) * c[i];
                    vf[(size_t)s * nf + j] += a;
                }
            }
    }
#pragma omp parallel for schedule(static)
    for (long k = 0; k < (long)Lc * Lc * nc; ++k) vc[k] = 0;
}

static void mg_cycle(int nlevels, level_t* lv, const int* pre, const int* post, int block) {
    for (int l = 0; l < nlevels; ++l) {
        relax_rb(&lv[l], pre[l]);
        apply_D(&lv[l], lv[l].phi, lv[l].r, lv[l].tmp);             /* residual */
        restrict_(&lv[l], lv[l].tmp, lv[l + 1].r, block);
    }
    for (int l = nlevels; l >= 0; --l) {
        relax_rb(&lv[l], post[l]);
        if (l > 0) prolong_add(&lv[l - 1], lv[l - 1].phi, lv[l].phi, block);
    }
}

static cplx cdot(const cplx* a, const cplx* b, long n) {
    double re = 0, im = 0;
#pragma omp parallel for schedule(static) reduction(+ : re, im)
    for (long k = 0; k < n; ++k) { const cplx z = conj(a[k]) * b[k]; re += creal(z); im += cimag(z); }
    return re + im * I;
}
static void axpy(cplx* y, cplx a, const cplx* x, long n) {
#pragma omp parallel for schedule(static)
    for (long k = 0; k < n; ++k) y[k] += a * x[k];
}

/* flexible GCR(restart) around the cycle; x starts at 0.  Returns the iterations done; resnorms[it] = |r|/|b|. */
int mgport_gcr_solve(int nlevels, level_t* lv, const int* pre, const int* post, int block, const cplx* b, cplx* x, double tol,
                     int max_iters, int restart, double* resnorms) {
    const long n0 = (long)lv[0].L * lv[0].L * lv[0].n;
    cplx* r = malloc(sizeof(cplx) * n0);
    cplx* Z = malloc(sizeof(cplx) * n0 * restart);
    cplx* W = malloc(sizeof(cplx) * n0 * restart);
    double* wn = malloc(sizeof(double) * restart);
    memcpy(r, b, sizeof(cplx) * n0);
    memset(x, 0, sizeof(cplx) * n0);
    const double bn = sqrt(creal(cdot(b, b, n0)));
    for (int l = 1; l <= nlevels; ++l) memset(lv[l].phi, 0, sizeof(cplx) * (size_t)lv[l].L * lv[l].L * lv[l].n);
    int slot = 0, it = 0;
    for (; it < max_iters; ++it) {
        memset(lv[0].phi, 0, sizeof(cplx) * n0);
        memcpy(lv[0].r, r, sizeof(cplx) * n0);
        mg_cycle(nlevels, lv, pre, post, block);
        cplx* z = Z + (size_t)slot * n0;
        cplx* w = W + (size_t)slot * n0;
        memcpy(z, lv[0].phi, sizeof(cplx) * n0);
        apply_D(&lv[0], z, NULL, w);
        cplx beta[64];
        for (int j = 0; j < slot; ++j) beta[j] = cdot(W + (size_t)j * n0, w, n0) / wn[j];   /* classical Gram-Schmidt */
        for (int j = 0; j < slot; ++j) { axpy(w, -beta[j], W + (size_t)j * n0, n0); axpy(z, -beta[j], Z + (size_t)j * n0, n0); }
        wn[slot] = creal(cdot(w, w, n0));
        const cplx alpha = cdot(w, r, n0) / wn[slot];
        axpy(x, alpha, z, n0);
        axpy(r, -alpha, w, n0);
        if (++slot >= restart) slot = 0;
        const double res = sqrt(creal(cdot(r, r, n0))) / bn;
        resnorms[it] = res;
        if (res < tol || res > 1e6 || res != res) { ++it; break; }
    }
    free(r); free(Z); free(W); free(wn);
    return it;
}

int mgport_level_size(void) { return (int)sizeof(level_t); }


/* Thread count of the solve loops: n > 0 sets it explicitly (launchers such as torch.distributed.run export
 * OMP_NUM_THREADS=1, which would silently serialise the CPU arm), n <= 0 uses every processor.  Returns the count in force. */
int mgport_set_threads(int n) {
    omp_set_dynamic(0);
    omp_set_num_threads(n > 0 ? n : omp_get_num_procs());
    return omp_get_max_threads();
}
