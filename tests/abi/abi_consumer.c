/* abi_consumer.c -- a plain C program that uses libmg2d_sm100.so through include/mg2d.h only, the way a maintainer of the
 * reference would bind it from C/C++: no Python, no torch, cudaMalloc'ed buffers.
 *
 *   abi_consumer in.bin out.bin
 * in.bin  (little endian): int32 L, int32 n, int32 nc, int32 block, float64 mass, then complex128 arrays
 *         U[L*L][2], v[L*L][2], D[L*L][5][n][n] (device layout: column-major blocks), w[L*L][n], P[L*L][nc][n]
 * out.bin: complex128 arrays  Dv[L*L][2] (mg2d_wilson_apply), b - Dv with the four fused reductions (4 float64),
 *          Dw[L*L][n] (mg2d_stencil_apply), Pw[(L/block)^2][nc] (mg2d_restrict)
 * Replaces Level::f_apply_D (S6/level.h:251-265), f_residue (:61-77) and Near_null::f_restriction (S6/near_null.h:217-240).
 * Exit code 0 on success; any MG2D error prints mg2d_last_error and exits 1.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <cuda_runtime_api.h>
#include "mg2d.h"

#define CK(call) do { int rc_ = (call); if (rc_ != MG2D_OK) { fprintf(stderr, "%s failed (%d): %s\n", #call, rc_, mg2d_last_error(ctx)); return 1; } } while (0)
#define CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #call, cudaGetErrorString(e_)); return 1; } } while (0)

static void* upload(const void* host, size_t bytes) {
    void* d = NULL;
    if (cudaMalloc(&d, bytes) != cudaSuccess) return NULL;
    if (cudaMemcpy(d, host, bytes, cudaMemcpyHostToDevice) != cudaSuccess) return NULL;
    return d;
}

int main(int argc, char** argv) {
    if (argc != 3) { fprintf(stderr, "usage: %s in.bin out.bin\n", argv[0]); return 2; }
    FILE* f = fopen(argv[1], "rb");
    if (!f) { perror("in.bin"); return 2; }
    int hdr[4]; double mass;
    if (fread(hdr, sizeof(int), 4, f) != 4 || fread(&mass, sizeof(double), 1, f) != 1) return 2;
    const int L = hdr[0], n = hdr[1], nc = hdr[2], block = hdr[3];
    const size_t S = (size_t)L * L, C16 = 16;
    const size_t bU = S * 2 * C16, bv = S * 2 * C16, bD = S * 5 * n * n * C16, bw = S * n * C16, bP = S * nc * n * C16;
    char* host = (char*)malloc(bU + bv + bD + bw + bP);
    if (fread(host, 1, bU + bv + bD + bw + bP, f) != bU + bv + bD + bw + bP) { fprintf(stderr, "short read\n"); return 2; }
    fclose(f);

    mg2d_ctx* ctx = NULL;
    if (mg2d_create(&ctx, 0) != MG2D_OK) { fprintf(stderr, "mg2d_create failed (needs an sm_100 device)\n"); return 1; }
    if (mg2d_version() < 100) return 1;
    char* dU = (char*)upload(host, bU);
    char* dv = (char*)upload(host + bU, bv);
    char* dD = (char*)upload(host + bU + bv, bD);
    char* dw = (char*)upload(host + bU + bv + bD, bw);
    char* dP = (char*)upload(host + bU + bv + bD + bw, bP);
    if (!dU || !dv || !dD || !dw || !dP) { fprintf(stderr, "upload failed\n"); return 1; }
    const size_t Sc = (size_t)(L / block) * (L / block);
    char *dDv, *dres, *dDw, *dPw; double* ddots;
    CU(cudaMalloc((void**)&dDv, bv)); CU(cudaMalloc((void**)&dres, bv)); CU(cudaMalloc((void**)&dDw, bw));
    CU(cudaMalloc((void**)&dPw, Sc * nc * C16)); CU(cudaMalloc((void**)&ddots, 4 * sizeof(double)));

    /* one GPU: the halo rows are the periodic wrap rows of the arrays themselves (include/mg2d.h, "Strip decomposition") */
    const size_t row2 = (size_t)L * 2 * C16, rown = (size_t)L * n * C16;
    CK(mg2d_wilson_apply(ctx, dDv, dv, dv + (L - 1) * row2, dv, dU, dU + (L - 1) * row2, NULL, mass, L, L, MG2D_MODE_APPLY, MG2D_C128, NULL, NULL));
    CK(mg2d_wilson_apply(ctx, dres, dDv, dDv + (L - 1) * row2, dDv, dU, dU + (L - 1) * row2, dv, mass, L, L, MG2D_MODE_RESID, MG2D_C128, ddots, NULL));
    CK(mg2d_stencil_apply(ctx, dDw, dw, dw + (L - 1) * rown, dw, dD, NULL, n, L, L, MG2D_MODE_APPLY, MG2D_C128, 1, (long long)S * n, (long long)S * n, NULL, NULL));
    CK(mg2d_restrict(ctx, dPw, dw, dP, n, nc, L, L, block, 1, MG2D_C128, NULL));
    /* error behaviour: bad arguments return MG2D_EINVAL and leave a message, they never abort */
    if (mg2d_restrict(ctx, dPw, dw, dP, n, nc, L, L, 5, 1, MG2D_C128, NULL) != MG2D_EINVAL || strlen(mg2d_last_error(ctx)) == 0) return 1;
    CU(cudaDeviceSynchronize());
    if (mg2d_launch_count(ctx) != 4) { fprintf(stderr, "launch count %d\n", mg2d_launch_count(ctx)); return 1; }

    char* out = (char*)malloc(2 * bv + 4 * sizeof(double) + bw + Sc * nc * C16);
    char* q = out;
    CU(cudaMemcpy(q, dDv, bv, cudaMemcpyDeviceToHost)); q += bv;
    CU(cudaMemcpy(q, dres, bv, cudaMemcpyDeviceToHost)); q += bv;
    CU(cudaMemcpy(q, ddots, 4 * sizeof(double), cudaMemcpyDeviceToHost)); q += 4 * sizeof(double);
    CU(cudaMemcpy(q, dDw, bw, cudaMemcpyDeviceToHost)); q += bw;
    CU(cudaMemcpy(q, dPw, Sc * nc * C16, cudaMemcpyDeviceToHost)); q += Sc * nc * C16;
    f = fopen(argv[2], "wb");
    if (!f || fwrite(out, 1, (size_t)(q - out), f) != (size_t)(q - out)) return 2;
    fclose(f);
    CK(mg2d_destroy(ctx));
    printf("abi_consumer ok\n");
    return 0;
}
