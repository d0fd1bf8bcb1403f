// scalar.cu -- real scalar Laplace geometric multigrid (BASELINE config 1), restating the operators of
// S2 = code/2_scalar_2d_nontelescoping/telescoping_2d_laplace_Mgrid.cpp (and S1/2D_laplace_Mgrid.cpp:25-106):
//   relax            S2:46-72    phi(s) = scale (phi(x+1)+phi(x-1)+phi(y+1)+phi(y-1) - b(s) a^2), lexicographic GS
//   f_projection     S2:74-110   res_c = 1/4 sum over the quadrant's 2x2 block of  b - (1/a^2)(sum nbrs - phi/scale)
//   f_interpolate    S2:112-143  phi_f(4 sites) += phi_c ; phi_c = 0
//   f_get_residue_mag S2:23-44   sum |res|
// Fields are real fp64, 8 B/site; all kernels are HBM/L2-streaming except the GS wavefront, which is
// latency-bound by construction (2L-1 dependent fronts per sweep reproduce the sequential order exactly).
#include "common.cuh"
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

namespace {

constexpr int S2_THREADS = 256;

__device__ __forceinline__ int wrap(int v, int L) { return v < 0 ? v + L : (v >= L ? v - L : v); }

__global__ void __launch_bounds__(S2_THREADS)
s2_gs_kernel(double* phi, const double* __restrict__ b, int L, double scale, double a2, int num_iter) {
    cg::grid_group grid = cg::this_grid();
    const bool single = gridDim.x == 1;
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    const int nthreads = gridDim.x * blockDim.x;
    for (int it = 0; it < num_iter; ++it) {
        for (int c = 0; c <= 2 * L - 2; ++c) {
            const int x0 = max(0, c - L + 1), x1 = min(c, L - 1);
            for (int x = x0 + tid; x <= x1; x += nthreads) {
                const int y = c - x;
                const int xp = wrap(x + 1, L), xm = wrap(x - 1, L), yp = wrap(y + 1, L), ym = wrap(y - 1, L);
                // summation order of S2:58-59
                const double v = scale * (__ldcg(phi + xp + (size_t)y * L) + __ldcg(phi + xm + (size_t)y * L)
                                          + __ldcg(phi + x + (size_t)yp * L) + __ldcg(phi + x + (size_t)ym * L)
                                          - __ldg(b + x + (size_t)y * L) * a2);
                __stcg(phi + x + (size_t)y * L, v);
            }
            if (single) __syncthreads(); else grid.sync();
        }
    }
}

__device__ __forceinline__ double s2_res(const double* __restrict__ phi, const double* __restrict__ b, int x, int y, int L,
                                         double inv_a2, double inv_scale) {
    const int xp = wrap(x + 1, L), xm = wrap(x - 1, L), yp = wrap(y + 1, L), ym = wrap(y - 1, L);
    return b[x + (size_t)y * L] - inv_a2 * (phi[xp + (size_t)y * L] + phi[xm + (size_t)y * L] + phi[x + (size_t)yp * L]
                                           + phi[x + (size_t)ym * L] - phi[x + (size_t)y * L] * inv_scale);
}

__global__ void __launch_bounds__(S2_THREADS)
s2_project_kernel(double* __restrict__ res_c, const double* __restrict__ b, const double* __restrict__ phi, int L,
                  double inv_a2, double scale, int sx, int sy) {
    const int Lc = L / 2;
    const long long n = (long long)Lc * Lc;
    for (long long X = blockIdx.x * (long long)blockDim.x + threadIdx.x; X < n; X += (long long)gridDim.x * blockDim.x) {
        const int y = (int)(X / Lc), x = (int)(X - (long long)y * Lc);
        const int xa = 2 * x, ya = 2 * y, xb = wrap(2 * x + sx, L), yb = wrap(2 * y + sy, L);
        // phi/scale as a true division, like S2:89; order of the four terms as S2:107
        const double r00 = b[xa + (size_t)ya * L] - inv_a2 * (phi[wrap(xa + 1, L) + (size_t)ya * L] + phi[wrap(xa - 1, L) + (size_t)ya * L] + phi[xa + (size_t)wrap(ya + 1, L) * L] + phi[xa + (size_t)wrap(ya - 1, L) * L] - phi[xa + (size_t)ya * L] / scale);
        const double r01 = b[xa + (size_t)yb * L] - inv_a2 * (phi[wrap(xa + 1, L) + (size_t)yb * L] + phi[wrap(xa - 1, L) + (size_t)yb * L] + phi[xa + (size_t)wrap(yb + 1, L) * L] + phi[xa + (size_t)wrap(yb - 1, L) * L] - phi[xa + (size_t)yb * L] / scale);
        const double r10 = b[xb + (size_t)ya * L] - inv_a2 * (phi[wrap(xb + 1, L) + (size_t)ya * L] + phi[wrap(xb - 1, L) + (size_t)ya * L] + phi[xb + (size_t)wrap(ya + 1, L) * L] + phi[xb + (size_t)wrap(ya - 1, L) * L] - phi[xb + (size_t)ya * L] / scale);
        const double r11 = b[xb + (size_t)yb * L] - inv_a2 * (phi[wrap(xb + 1, L) + (size_t)yb * L] + phi[wrap(xb - 1, L) + (size_t)yb * L] + phi[xb + (size_t)wrap(yb + 1, L) * L] + phi[xb + (size_t)wrap(yb - 1, L) * L] - phi[xb + (size_t)yb * L] / scale);
        res_c[X] = 0.25 * (r00 + r01 + r10 + r11);
    }
}

__global__ void __launch_bounds__(S2_THREADS)
s2_interpolate_kernel(double* __restrict__ phi_f, double* __restrict__ phi_c, int Lc, int sx, int sy) {
    const int L = 2 * Lc;
    const long long n = (long long)Lc * Lc;
    for (long long X = blockIdx.x * (long long)blockDim.x + threadIdx.x; X < n; X += (long long)gridDim.x * blockDim.x) {
        const int y = (int)(X / Lc), x = (int)(X - (long long)y * Lc);
        const int xa = 2 * x, ya = 2 * y, xb = wrap(2 * x + sx, L), yb = wrap(2 * y + sy, L);
        const double c = phi_c[X];
        phi_f[xa + (size_t)ya * L] += c;
        phi_f[xa + (size_t)yb * L] += c;
        phi_f[xb + (size_t)ya * L] += c;
        phi_f[xb + (size_t)yb * L] += c;
        phi_c[X] = 0.0;
    }
}

__global__ void __launch_bounds__(S2_THREADS)
s2_residue_mag_kernel(const double* __restrict__ phi, const double* __restrict__ b, int L, double inv_a2, double scale,
                      double* __restrict__ partials, unsigned int* __restrict__ counter, double* __restrict__ out) {
    const long long n = (long long)L * L;
    double red[1] = {0.0};
    for (long long s = blockIdx.x * (long long)blockDim.x + threadIdx.x; s < n; s += (long long)gridDim.x * blockDim.x) {
        const int y = (int)(s / L), x = (int)(s - (long long)y * L);
        const int xp = wrap(x + 1, L), xm = wrap(x - 1, L), yp = wrap(y + 1, L), ym = wrap(y - 1, L);
        const double r = b[s] - inv_a2 * (phi[xp + (size_t)y * L] + phi[xm + (size_t)y * L] + phi[x + (size_t)yp * L]
                                          + phi[x + (size_t)ym * L] - phi[s] / scale);
        red[0] += fabs(r);
    }
    grid_reduce<1, S2_THREADS>(red, partials, counter, out, blockIdx.x, gridDim.x);
}

__global__ void s2_scale_kernel(double* __restrict__ x, double s, long long n) {
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) x[e] *= s;
}

inline void quad_shift(int quad, int& sx, int& sy) {   // S2:100-103
    sx = (quad == 1 || quad == 4) ? 1 : -1;
    sy = (quad == 1 || quad == 2) ? 1 : -1;
}
inline int sgrid(mg2d_ctx* ctx, long long n) {
    long long nb = (n + S2_THREADS - 1) / S2_THREADS;
    long long cap = (long long)ctx->num_sms * 8;
    if (nb > cap) nb = cap;
    if (nb > MG2D_MAX_PARTIALS) nb = MG2D_MAX_PARTIALS;
    return (int)(nb < 1 ? 1 : nb);
}

}  // namespace

extern "C" int mg2d_s2_relax(mg2d_ctx* ctx, double* phi, const double* b, int L, double scale, double a, int num_iter,
                             int gs_flag, void* stream) {
    if (!ctx) return MG2D_EINVAL;
    if (!phi || !b || L < 2 || num_iter < 0) return mg2d_fail(ctx, MG2D_EINVAL, "mg2d_s2_relax: bad argument");
    if (gs_flag != 1) return mg2d_fail(ctx, MG2D_EUNSUPPORTED, "mg2d_s2_relax: only gs_flag=1 (the reference's hard-coded choice, S2:194) is built");
    if (num_iter == 0) return MG2D_OK;
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, s2_gs_kernel, S2_THREADS, 0) != cudaSuccess || per_sm < 1)
        return mg2d_fail(ctx, MG2D_ECUDA, "mg2d_s2_relax: occupancy query failed");
    int grid = (L + S2_THREADS - 1) / S2_THREADS;
    if (grid > per_sm * ctx->num_sms) grid = per_sm * ctx->num_sms;
    double a2 = a * a;
    void* args[] = {&phi, &b, &L, &scale, &a2, &num_iter};
    cudaError_t e = cudaLaunchCooperativeKernel((const void*)s2_gs_kernel, dim3(grid), dim3(S2_THREADS), args, 0, (cudaStream_t)stream);
    if (e != cudaSuccess) { snprintf(ctx->err, sizeof(ctx->err), "mg2d_s2_relax: %s", cudaGetErrorString(e)); return MG2D_ECUDA; }
    return mg2d_check_launch(ctx, "mg2d_s2_relax");
}

extern "C" int mg2d_s2_project(mg2d_ctx* ctx, double* res_c, const double* b, const double* phi, int L, double scale, double a,
                               int quad, void* stream) {
    if (!ctx) return MG2D_EINVAL;
    if (!res_c || !b || !phi || L < 2 || (L & 1) || quad < 1 || quad > 4) return mg2d_fail(ctx, MG2D_EINVAL, "mg2d_s2_project: bad argument");
    int sx, sy; quad_shift(quad, sx, sy);
    s2_project_kernel<<<sgrid(ctx, (long long)(L / 2) * (L / 2)), S2_THREADS, 0, (cudaStream_t)stream>>>(res_c, b, phi, L, 1.0 / (a * a), scale, sx, sy);
    return mg2d_check_launch(ctx, "mg2d_s2_project");
}

extern "C" int mg2d_s2_interpolate(mg2d_ctx* ctx, double* phi_f, double* phi_c, int Lc, int quad, void* stream) {
    if (!ctx) return MG2D_EINVAL;
    if (!phi_f || !phi_c || Lc < 1 || quad < 1 || quad > 4) return mg2d_fail(ctx, MG2D_EINVAL, "mg2d_s2_interpolate: bad argument");
    int sx, sy; quad_shift(quad, sx, sy);
    s2_interpolate_kernel<<<sgrid(ctx, (long long)Lc * Lc), S2_THREADS, 0, (cudaStream_t)stream>>>(phi_f, phi_c, Lc, sx, sy);
    return mg2d_check_launch(ctx, "mg2d_s2_interpolate");
}

extern "C" int mg2d_s2_residue_mag(mg2d_ctx* ctx, const double* phi, const double* b, int L, double scale, double a,
                                   double* out, void* stream) {
    if (!ctx) return MG2D_EINVAL;
    if (!phi || !b || !out || L < 2) return mg2d_fail(ctx, MG2D_EINVAL, "mg2d_s2_residue_mag: bad argument");
    s2_residue_mag_kernel<<<sgrid(ctx, (long long)L * L), S2_THREADS, 0, (cudaStream_t)stream>>>(phi, b, L, 1.0 / (a * a), scale, ctx->partials, ctx->counter, out);
    return mg2d_check_launch(ctx, "mg2d_s2_residue_mag");
}

extern "C" int mg2d_s2_scale(mg2d_ctx* ctx, double* x, double s, long long n, void* stream) {
    if (!ctx) return MG2D_EINVAL;
    if (!x || n < 1) return mg2d_fail(ctx, MG2D_EINVAL, "mg2d_s2_scale: bad argument");
    s2_scale_kernel<<<sgrid(ctx, n), S2_THREADS, 0, (cudaStream_t)stream>>>(x, s, n);
    return mg2d_check_launch(ctx, "mg2d_s2_scale");
}
