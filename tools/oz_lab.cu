// Development bench for the INT8-sliced FP64 GEMM (oz_gemm.cuh / oz_split.cuh); not part of the product library.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o tools/oz_lab tools/oz_lab.cu
//   ./tools/oz_lab [check|perf|all]
// check: digit planes against a host re-computation, the integer GEMM against an exact host sum over the planes,
//        the scaled result against a long-double product, transposed split against the row split, the
//        lower-triangular K-range map (K^-1 = M^T M shape) and the batched form.
// perf:  timings of square / SYRK / LAUUM shapes against the DMMA kernel.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../gp-plus_b200/csrc/oz_split.cuh"

using namespace gpp;

#define CHECK(x)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (x);                                                                      \
        if (e_ != cudaSuccess) {                                                                   \
            fprintf(stderr, "%s:%d %s -> %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); \
            exit(1);                                                                               \
        }                                                                                          \
    } while (0)

static double urand(unsigned long long& s) {
    s = s * 6364136223846793005ull + 1442695040888963407ull;
    return ((double)(s >> 11) / 9007199254740992.0) * 2.0 - 1.0;
}

static OzPlanes alloc_planes(long long rows, long long pitch) {
    OzPlanes P;
    P.rows = rows;
    P.pitch = pitch;
    CHECK(cudaMalloc(&P.planes, (size_t)OZ_S * rows * pitch));
    CHECK(cudaMemset(P.planes, 0, (size_t)OZ_S * rows * pitch));
    CHECK(cudaMalloc(&P.scale, sizeof(double) * rows));
    CHECK(cudaMemset(P.scale, 0, sizeof(double) * rows));
    CHECK(cudaMalloc(&P.colmax, sizeof(unsigned long long) * rows));
    return P;
}
static void free_planes(OzPlanes& P) {
    cudaFree(P.planes);
    cudaFree(P.scale);
    cudaFree(P.colmax);
}

// host digits of one value
static void host_digits(double x, double scale, int8_t* d) {
    long long v = (long long)floor(ldexp(x / scale, 56));
    for (int p = OZ_S - 1; p >= 0; p--) {
        long long lo = (long long)(int8_t)(v & 0xff);
        d[p] = (int8_t)lo;
        v = (v - lo) >> 8;
    }
}

static int check_small() {
    int fails = 0;
    const int M = 256, N = 256, K = 256;
    std::vector<double> A((size_t)M * K), B((size_t)N * K), At((size_t)K * M);
    unsigned long long seed = 12345;
    for (int r = 0; r < M; r++) {
        const double rs = pow(10.0, 3.0 * urand(seed));
        for (int k = 0; k < K; k++) A[(size_t)r * K + k] = rs * urand(seed) * ((k % 7 == 0) ? 1e-6 : 1.0);
    }
    for (int r = 0; r < N; r++) {
        const double rs = pow(10.0, 3.0 * urand(seed));
        for (int k = 0; k < K; k++) B[(size_t)r * K + k] = rs * urand(seed);
    }
    for (int r = 0; r < M; r++)
        for (int k = 0; k < K; k++) At[(size_t)k * M + r] = A[(size_t)r * K + k];
    double *dA, *dB, *dAt, *dC;
    CHECK(cudaMalloc(&dA, sizeof(double) * M * K));
    CHECK(cudaMalloc(&dB, sizeof(double) * N * K));
    CHECK(cudaMalloc(&dAt, sizeof(double) * M * K));
    CHECK(cudaMalloc(&dC, sizeof(double) * M * N));
    CHECK(cudaMemcpy(dA, A.data(), sizeof(double) * M * K, cudaMemcpyHostToDevice));
    CHECK(cudaMemcpy(dB, B.data(), sizeof(double) * N * K, cudaMemcpyHostToDevice));
    CHECK(cudaMemcpy(dAt, At.data(), sizeof(double) * M * K, cudaMemcpyHostToDevice));
    OzPlanes PA = alloc_planes(M, K), PB = alloc_planes(N, K), PAt = alloc_planes(M, K);
    CHECK(oz_split_rows(dA, K, 0, M, M, K, PA, 0, 0, 0, 0, 1, 0));
    CHECK(oz_split_rows(dB, K, 0, N, N, K, PB, 0, 0, 0, 0, 1, 0));
    CHECK(oz_split_cols(dAt, M, 0, K, K, M, 0, PAt, 0, 0, 0, 0, 1, 0));
    CHECK(cudaDeviceSynchronize());
    std::vector<int8_t> hA((size_t)OZ_S * M * K), hB((size_t)OZ_S * N * K), hAt((size_t)OZ_S * M * K);
    std::vector<double> sA(M), sB(N), sAt(M);
    CHECK(cudaMemcpy(hA.data(), PA.planes, hA.size(), cudaMemcpyDeviceToHost));
    CHECK(cudaMemcpy(hB.data(), PB.planes, hB.size(), cudaMemcpyDeviceToHost));
    CHECK(cudaMemcpy(hAt.data(), PAt.planes, hAt.size(), cudaMemcpyDeviceToHost));
    CHECK(cudaMemcpy(sA.data(), PA.scale, sizeof(double) * M, cudaMemcpyDeviceToHost));
    CHECK(cudaMemcpy(sB.data(), PB.scale, sizeof(double) * N, cudaMemcpyDeviceToHost));
    CHECK(cudaMemcpy(sAt.data(), PAt.scale, sizeof(double) * M, cudaMemcpyDeviceToHost));
    // (1) planes against the host digits
    long long bad_digits = 0, bad_t = 0;
    double worst_rep = 0.0;
    for (int r = 0; r < M; r++) {
        double amax = 0.0;
        for (int k = 0; k < K; k++) amax = fmax(amax, fabs(A[(size_t)r * K + k]));
        int ex;
        frexp(amax, &ex);
        const double sc = ldexp(1.0, ex + 2);
        if (sc != sA[r]) bad_digits++;
        if (sAt[r] != sA[r]) bad_t++;
        for (int k = 0; k < K; k++) {
            int8_t d[OZ_S];
            host_digits(A[(size_t)r * K + k], sc, d);
            long double rep = 0.0L;
            for (int p = 0; p < OZ_S; p++) {
                const int8_t g = hA[((size_t)p * M + r) * K + k];
                if (g != d[p]) bad_digits++;
                if (hAt[((size_t)p * M + r) * K + k] != g) bad_t++;
                rep += (long double)g * powl(256.0L, -(p + 1));
            }
            worst_rep = fmax(worst_rep, (double)fabsl(rep * sc - A[(size_t)r * K + k]) / sc);
        }
    }
    printf("split_rows: digit mismatches %lld, representation error / scale %.3e (bound 2^-56 = %.3e)\n", bad_digits,
           worst_rep, ldexp(1.0, -56));
    printf("split_cols: mismatches against split_rows %lld\n", bad_t);
    if (bad_digits || bad_t || worst_rep > ldexp(1.0, -56)) fails++;
    fflush(stdout);

    // (2) GEMM
    CUtensorMap tmA, tmB;
    CHECK(oz_make_map(&tmA, PA.planes, K, (long long)OZ_S * M, K, OZ_BM));
    CHECK(oz_make_map(&tmB, PB.planes, K, (long long)OZ_S * N, K, OZ_BN));
    CHECK(oz_set_attributes());
    OzGemmOp op = oz_default();
    op.a_plane_rows = M;
    op.b_plane_rows = N;
    op.a_scale = PA.scale;
    op.b_scale = PB.scale;
    op.C = dC;
    op.ldc = N;
    op.tiles_m = op.tiles_m_last = M / TILE;
    op.tiles_n = N / TILE;
    op.klo_c = 0;
    op.khi_c = K / TILE;
    CHECK(cudaMemset(dC, 0xff, sizeof(double) * M * N));
    CHECK(launch_oz_gemm(tmA, tmB, op, 1, 0));
    CHECK(cudaDeviceSynchronize());
    printf("oz_gemm small launch done\n");
    fflush(stdout);
    std::vector<double> C((size_t)M * N);
    CHECK(cudaMemcpy(C.data(), dC, sizeof(double) * M * N, cudaMemcpyDeviceToHost));
    double worst_int = 0.0, worst_true = 0.0;
    for (int r = 0; r < M; r++)
        for (int c = 0; c < N; c++) {
            long double lv[OZ_S];
            for (int l = 0; l < OZ_S; l++) lv[l] = 0.0L;
            long double tru = 0.0L, mag = 0.0L;
            for (int k = 0; k < K; k++) {
                tru += (long double)A[(size_t)r * K + k] * (long double)B[(size_t)c * K + k];
                mag += fabsl((long double)A[(size_t)r * K + k] * (long double)B[(size_t)c * K + k]);
            }
            for (int i = 0; i < OZ_S; i++)
                for (int j = 0; i + j < OZ_S; j++) {
                    long long acc = 0;
                    const int8_t* pa = &hA[((size_t)i * M + r) * K];
                    const int8_t* pb = &hB[((size_t)j * N + c) * K];
                    for (int k = 0; k < K; k++) acc += (long long)pa[k] * (long long)pb[k];
                    lv[i + j] += (long double)acc;
                }
            long double s = 0.0L;
            for (int l = OZ_S - 1; l >= 0; l--) s = s / 256.0L + lv[l];
            s = s / 65536.0L * (long double)sA[r] * (long double)sB[c];
            const double got = C[(size_t)r * N + c];
            worst_int = fmax(worst_int, (double)(fabsl(got - s) / (mag + 1e-300L)));
            worst_true = fmax(worst_true, (double)(fabsl(got - tru) / (mag + 1e-300L)));
        }
    printf("oz_gemm 256x256x256 (both / sum|a||b|): vs exact plane sum %.3e; vs long-double product %.3e\n",
           worst_int, worst_true);
    if (!(worst_int < 1e-15) || !(worst_true < 1e-15)) fails++;
    fflush(stdout);

    // (3) alpha / beta and a K sub-range
    {
        std::vector<double> C0((size_t)M * N);
        for (auto& v : C0) v = urand(seed);
        CHECK(cudaMemcpy(dC, C0.data(), sizeof(double) * M * N, cudaMemcpyHostToDevice));
        OzGemmOp o2 = op;
        o2.alpha = -1.0;
        o2.beta = 1.0;
        o2.klo_c = 1;
        o2.khi_c = 2;
        CHECK(launch_oz_gemm(tmA, tmB, o2, 1, 0));
        CHECK(cudaDeviceSynchronize());
        CHECK(cudaMemcpy(C.data(), dC, sizeof(double) * M * N, cudaMemcpyDeviceToHost));
        double worst = 0.0;
        for (int r = 0; r < M; r++)
            for (int c = 0; c < N; c++) {
                long double tru = C0[(size_t)r * N + c], mag = fabs(C0[(size_t)r * N + c]);
                for (int k = 128; k < 256; k++) {
                    tru -= (long double)A[(size_t)r * K + k] * (long double)B[(size_t)c * K + k];
                    mag += fabsl((long double)A[(size_t)r * K + k] * (long double)B[(size_t)c * K + k]);
                }
                worst = fmax(worst, (double)(fabsl(C[(size_t)r * N + c] - tru) / mag));
            }
        printf("alpha=-1 beta=1 K block [1,2): rel %.3e\n", worst);
        if (!(worst < 1e-13)) fails++;
        fflush(stdout);
    }
    cudaFree(dA); cudaFree(dB); cudaFree(dAt); cudaFree(dC);
    free_planes(PA); free_planes(PB); free_planes(PAt);
    return fails;
}

// K^-1 = M^T M shape: M lower triangular (zeros above the diagonal), transposed planes, lower tiles, K range [ti, T)
static int check_lauum(int n) {
    int fails = 0;
    const int T = n / TILE;
    std::vector<double> Mh((size_t)n * n, 0.0);
    unsigned long long seed = 777;
    for (int i = 0; i < n; i++)
        for (int j = 0; j <= i; j++) Mh[(size_t)i * n + j] = urand(seed) * exp(-0.01 * (i - j)) * (i == j ? 3.0 : 1.0);
    double *dM, *dC, *dR;
    CHECK(cudaMalloc(&dM, sizeof(double) * n * n));
    CHECK(cudaMalloc(&dC, sizeof(double) * n * n));
    CHECK(cudaMalloc(&dR, sizeof(double) * n * n));
    CHECK(cudaMemcpy(dM, Mh.data(), sizeof(double) * n * n, cudaMemcpyHostToDevice));
    CHECK(cudaMemset(dC, 0, sizeof(double) * n * n));
    CHECK(cudaMemset(dR, 0, sizeof(double) * n * n));
    OzPlanes P = alloc_planes(n, n);
    CHECK(oz_split_cols(dM, n, 0, n, n, n, 1, P, 0, 0, 0, 0, 1, 0));
    CUtensorMap tmA, tmB;
    CHECK(oz_make_map(&tmA, P.planes, n, (long long)OZ_S * n, n, OZ_BM));
    CHECK(oz_make_map(&tmB, P.planes, n, (long long)OZ_S * n, n, OZ_BN));
    OzGemmOp op = oz_default();
    op.a_plane_rows = op.b_plane_rows = n;
    op.a_scale = op.b_scale = P.scale;
    op.C = dC;
    op.ldc = n;
    op.map = MAP_TRI;
    op.tiles_m = op.tiles_m_last = T;
    op.tiles_n = T;
    op.klo_sel = KSEL_TI;
    op.klo_c = 0;
    op.khi_sel = KSEL_CONST;
    op.khi_c = T;
    CHECK(launch_oz_gemm(tmA, tmB, op, 1, 0));
    // DMMA reference (chol.cuh's lauum shape)
    CHECK(gemm_set_attributes());
    GemmOp g = gemm_default();
    g.A = dM; g.lda = n; g.B = dM; g.ldb = n; g.C = dR; g.ldc = n;
    g.map = MAP_TRI; g.tiles_m = g.tiles_m_last = T; g.tiles_n = T;
    g.klo_sel = KSEL_TI; g.klo_c = 0; g.khi_sel = KSEL_CONST; g.khi_c = T;
    CHECK(launch_gemm(g, false, false, 1, 0));
    CHECK(cudaDeviceSynchronize());
    std::vector<double> C((size_t)n * n), R((size_t)n * n);
    CHECK(cudaMemcpy(C.data(), dC, sizeof(double) * n * n, cudaMemcpyDeviceToHost));
    CHECK(cudaMemcpy(R.data(), dR, sizeof(double) * n * n, cudaMemcpyDeviceToHost));
    double worst = 0.0, cmax = 0.0;
    for (int i = 0; i < n; i++)
        for (int j = 0; j <= (i / TILE) * TILE + TILE - 1 && j < n; j++) {
            if (j / TILE > i / TILE) continue;
            worst = fmax(worst, fabs(C[(size_t)i * n + j] - R[(size_t)i * n + j]));
            cmax = fmax(cmax, fabs(R[(size_t)i * n + j]));
        }
    printf("lauum shape n=%d: |oz - dmma|max / |dmma|max = %.3e\n", n, worst / cmax);
    if (!(worst / cmax < 1e-13)) fails++;
    fflush(stdout);
    for (int lv = 6; lv >= 5; lv--) {
        // fewer significance levels (the K^-1 product of the engine keeps 6: it only feeds the gradient trace)
        OzGemmOp o2 = op;
        o2.levels = lv;
        CHECK(launch_oz_gemm(tmA, tmB, o2, 1, 0));
        CHECK(cudaDeviceSynchronize());
        CHECK(cudaMemcpy(C.data(), dC, sizeof(double) * n * n, cudaMemcpyDeviceToHost));
        double w2 = 0.0;
        for (int i = 0; i < n; i++)
            for (int j = 0; j < n; j++) {
                if (j / TILE > i / TILE) continue;
                w2 = fmax(w2, fabs(C[(size_t)i * n + j] - R[(size_t)i * n + j]));
            }
        printf("   %d levels: |oz - dmma|max / |dmma|max = %.3e\n", lv, w2 / cmax);
        if (!(w2 / cmax < (lv == 6 ? 1e-10 : 1e-8))) fails++;
    }
    cudaFree(dM); cudaFree(dC); cudaFree(dR);
    free_planes(P);
    return fails;
}

// batched rectangular products with per-entry strides (trtri level shape): C_z = A_z B_z^T, z = 0..nb-1, blocks on
// the diagonal of an n x n matrix
static int check_batched() {
    int fails = 0;
    const int n = 1024, hb = 2, nb = 2;   // two groups of 2*hb tiles: A_z = X[(z*4+2)*128 .. +256, z*512 .. +256]
    std::vector<double> Xh((size_t)n * n), Yh((size_t)n * n);
    unsigned long long seed = 99;
    for (auto& v : Xh) v = urand(seed);
    for (auto& v : Yh) v = urand(seed);
    double *dX, *dY, *dC, *dR;
    CHECK(cudaMalloc(&dX, sizeof(double) * n * n));
    CHECK(cudaMalloc(&dY, sizeof(double) * n * n));
    CHECK(cudaMalloc(&dC, sizeof(double) * n * n));
    CHECK(cudaMalloc(&dR, sizeof(double) * n * n));
    CHECK(cudaMemcpy(dX, Xh.data(), sizeof(double) * n * n, cudaMemcpyHostToDevice));
    CHECK(cudaMemcpy(dY, Yh.data(), sizeof(double) * n * n, cudaMemcpyHostToDevice));
    CHECK(cudaMemset(dC, 0, sizeof(double) * n * n));
    CHECK(cudaMemset(dR, 0, sizeof(double) * n * n));
    OzPlanes PX = alloc_planes(n, n), PY = alloc_planes(n, n);
    const long long zs = (long long)2 * hb * TILE * n + (long long)2 * hb * TILE;   // next group, diagonal step
    const long long off21 = (long long)hb * TILE * n;
    // A_z = X21 block (k-contiguous), B_z = Y11 block given as [k][c] (row-contiguous) -> transposed planes
    CHECK(oz_split_rows(dX + off21, n, zs, hb * TILE, hb * TILE, hb * TILE, PX, hb * TILE, 0, 2 * hb * TILE, 2 * hb * TILE, nb, 0));
    CHECK(oz_split_cols(dY, n, zs, hb * TILE, hb * TILE, hb * TILE, 0, PY, 0, 0, 2 * hb * TILE, 2 * hb * TILE, nb, 0));
    CUtensorMap tmA, tmB;
    CHECK(oz_make_map(&tmA, PX.planes, n, (long long)OZ_S * n, n, OZ_BM));
    CHECK(oz_make_map(&tmB, PY.planes, n, (long long)OZ_S * n, n, OZ_BN));
    OzGemmOp op = oz_default();
    op.a_plane_rows = op.b_plane_rows = n;
    op.a_row0 = hb * TILE; op.a_k0 = 0; op.a_zs_row = 2 * hb * TILE; op.a_zs_k = 2 * hb * TILE;
    op.b_row0 = 0; op.b_k0 = 0; op.b_zs_row = 2 * hb * TILE; op.b_zs_k = 2 * hb * TILE;
    op.a_scale = PX.scale; op.b_scale = PY.scale;
    op.C = dC + off21; op.ldc = n; op.c_zs = zs;
    op.tiles_m = op.tiles_m_last = hb; op.tiles_n = hb;
    op.klo_sel = KSEL_TJ; op.klo_c = 0; op.khi_sel = KSEL_CONST; op.khi_c = hb;
    CHECK(launch_oz_gemm(tmA, tmB, op, nb, 0));
    GemmOp g = gemm_default();
    g.A = dX + off21; g.lda = n; g.a_zs = zs; g.B = dY; g.ldb = n; g.b_zs = zs; g.C = dR + off21; g.ldc = n; g.c_zs = zs;
    g.tiles_m = g.tiles_m_last = hb; g.tiles_n = hb;
    g.klo_sel = KSEL_TJ; g.klo_c = 0; g.khi_sel = KSEL_CONST; g.khi_c = hb;
    CHECK(launch_gemm(g, true, false, nb, 0));
    CHECK(cudaDeviceSynchronize());
    std::vector<double> C((size_t)n * n), R((size_t)n * n);
    CHECK(cudaMemcpy(C.data(), dC, sizeof(double) * n * n, cudaMemcpyDeviceToHost));
    CHECK(cudaMemcpy(R.data(), dR, sizeof(double) * n * n, cudaMemcpyDeviceToHost));
    double worst = 0.0, cmax = 0.0;
    for (size_t i = 0; i < C.size(); i++) {
        worst = fmax(worst, fabs(C[i] - R[i]));
        cmax = fmax(cmax, fabs(R[i]));
    }
    printf("batched KSEL_TJ shape: |oz - dmma|max / |dmma|max = %.3e (|dmma|max %.3e)\n", worst / cmax, cmax);
    if (!(worst / cmax < 1e-13) || !(cmax > 0.0)) fails++;
    fflush(stdout);
    cudaFree(dX); cudaFree(dY); cudaFree(dC); cudaFree(dR);
    free_planes(PX); free_planes(PY);
    return fails;
}

__global__ void fill_rand(double* A, long long n, unsigned seed) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned long long s = (unsigned long long)i * 2654435761ull + seed;
    s ^= s >> 17; s *= 0xed5ad4bbull; s ^= s >> 11; s *= 0xac4c1b51ull; s ^= s >> 15;
    A[i] = ((double)(s & 0xffffffffull) / 4294967296.0) * 2.0 - 1.0;
}

static void perf(int n, int kblocks, int tri, int reps) {
    // C (n x n, or lower tiles) = A[:, :K] B[:, :K]^T with K = kblocks * 128 (tri: K range [ti, T) like K^-1 = M^T M)
    const int T = n / TILE;
    double *dA, *dC;
    CHECK(cudaMalloc(&dA, sizeof(double) * n * n));
    CHECK(cudaMalloc(&dC, sizeof(double) * n * n));
    fill_rand<<<(unsigned)(((long long)n * n + 255) / 256), 256>>>(dA, (long long)n * n, 7u);
    CHECK(cudaMemset(dC, 0, sizeof(double) * n * n));
    OzPlanes P = alloc_planes(n, n);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float ms_split_r = 0, ms_split_c = 0, ms_oz = 0, ms_dmma = 0;
    CHECK(oz_split_rows(dA, n, 0, n, n, kblocks * TILE, P, 0, 0, 0, 0, 1, 0));
    cudaEventRecord(e0);
    for (int i = 0; i < reps; i++) CHECK(oz_split_rows(dA, n, 0, n, n, kblocks * TILE, P, 0, 0, 0, 0, 1, 0));
    cudaEventRecord(e1);
    CHECK(cudaEventSynchronize(e1));
    cudaEventElapsedTime(&ms_split_r, e0, e1);
    ms_split_r /= reps;
    if (kblocks == T) {
        cudaEventRecord(e0);
        for (int i = 0; i < reps; i++) CHECK(oz_split_cols(dA, n, 0, n, n, n, tri, P, 0, 0, 0, 0, 1, 0));
        cudaEventRecord(e1);
        CHECK(cudaEventSynchronize(e1));
        cudaEventElapsedTime(&ms_split_c, e0, e1);
        ms_split_c /= reps;
    }
    CUtensorMap tmA, tmB;
    CHECK(oz_make_map(&tmA, P.planes, n, (long long)OZ_S * n, n, OZ_BM));
    CHECK(oz_make_map(&tmB, P.planes, n, (long long)OZ_S * n, n, OZ_BN));
    OzGemmOp op = oz_default();
    op.a_plane_rows = op.b_plane_rows = n;
    op.a_scale = op.b_scale = P.scale;
    op.C = dC; op.ldc = n;
    op.tiles_m = op.tiles_m_last = T; op.tiles_n = T;
    op.map = MAP_TRI;
    if (tri) { op.klo_sel = KSEL_TI; op.klo_c = 0; op.khi_c = T; }
    else { op.klo_c = 0; op.khi_c = kblocks; op.alpha = -1.0; op.beta = 1.0; }
    CHECK(launch_oz_gemm(tmA, tmB, op, 1, 0));
    CHECK(cudaDeviceSynchronize());
    {
        long long* dprof;
        CHECK(cudaMalloc(&dprof, 8 * sizeof(long long)));
        CHECK(cudaMemset(dprof, 0, 8 * sizeof(long long)));
        OzGemmOp o2 = op;
        o2.prof = dprof;
        CHECK(launch_oz_gemm(tmA, tmB, o2, 1, 0));
        CHECK(cudaDeviceSynchronize());
        long long hp[8];
        CHECK(cudaMemcpy(hp, dprof, sizeof(hp), cudaMemcpyDeviceToHost));
        printf("   CTA 0, first tile (cycles since pass-0 start): p0 issued %lld  complete %lld  epilogue done %lld | p1 start %lld issued %lld"
               " complete %lld epilogue done %lld\n", hp[1] - hp[0], hp[2] - hp[0], hp[3] - hp[0], hp[4] - hp[0], hp[5] - hp[0],
               hp[6] - hp[0], hp[7] - hp[0]);
        cudaFree(dprof);
    }
    cudaEventRecord(e0);
    for (int i = 0; i < reps; i++) CHECK(launch_oz_gemm(tmA, tmB, op, 1, 0));
    cudaEventRecord(e1);
    CHECK(cudaEventSynchronize(e1));
    cudaEventElapsedTime(&ms_oz, e0, e1);
    ms_oz /= reps;
    float ms_lv[2] = {0.f, 0.f};
    if (tri) {
        for (int q = 0; q < 2; q++) {
            OzGemmOp o2 = op;
            o2.levels = 6 - q;
            CHECK(launch_oz_gemm(tmA, tmB, o2, 1, 0));
            cudaEventRecord(e0);
            for (int i = 0; i < reps; i++) CHECK(launch_oz_gemm(tmA, tmB, o2, 1, 0));
            cudaEventRecord(e1);
            CHECK(cudaEventSynchronize(e1));
            cudaEventElapsedTime(&ms_lv[q], e0, e1);
            ms_lv[q] /= reps;
        }
        printf("   K^-1 shape with 6 levels (21 pairs): %.3f ms, 5 levels (15 pairs): %.3f ms\n", ms_lv[0], ms_lv[1]);
    }
    GemmOp g = gemm_default();
    g.A = dA; g.lda = n; g.B = dA; g.ldb = n; g.C = dC; g.ldc = n;
    g.map = MAP_TRI; g.tiles_m = g.tiles_m_last = T; g.tiles_n = T;
    if (tri) { g.klo_sel = KSEL_TI; g.klo_c = 0; g.khi_c = T; }
    else { g.klo_c = 0; g.khi_c = kblocks; g.alpha = -1.0; g.beta = 1.0; }
    CHECK(launch_gemm(g, !tri, !tri, 1, 0));
    CHECK(cudaDeviceSynchronize());
    cudaEventRecord(e0);
    for (int i = 0; i < reps; i++) CHECK(launch_gemm(g, !tri, !tri, 1, 0));
    cudaEventRecord(e1);
    CHECK(cudaEventSynchronize(e1));
    cudaEventElapsedTime(&ms_dmma, e0, e1);
    ms_dmma /= reps;
    // flops: lower tiles x K range
    double flops = 0.0;
    for (int ti = 0; ti < T; ti++)
        for (int tj = 0; tj <= ti; tj++) flops += 2.0 * TILE * TILE * (double)TILE * (tri ? (T - ti) : kblocks);
    printf("n=%d %s K=%d: oz %.3f ms = %.1f TFLOP/s-eq (int8 %.0f TOPS), dmma %.3f ms = %.1f TFLOP/s, split_rows %.3f ms"
           " split_cols %.3f ms\n", n, tri ? "lauum" : "syrk", tri ? n : kblocks * TILE, ms_oz, flops / ms_oz * 1e-9,
           flops * 28.0 / ms_oz * 1e-9, ms_dmma, flops / ms_dmma * 1e-9, ms_split_r, ms_split_c);
    fflush(stdout);
    cudaFree(dA); cudaFree(dC);
    free_planes(P);
}


// ---- raw tcgen05 kind::i8 issue rate: one CTA per SM, MMAs back to back on resident shared-memory tiles ----------
template <int N, int KB>
__global__ void __launch_bounds__(128, 1) mma_rate_kernel(int iters, int a_tiles, int b_tiles, long long* cycles) {
    extern __shared__ uint8_t raw[];
    uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < (a_tiles * 128 * KB + b_tiles * N * KB) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(base)[i] = 0x01010101u;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc(&slot, 512);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = slot;
    constexpr uint64_t layout = (KB == 128) ? 2 : (KB == 64 ? 4 : 6);
    constexpr uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    if (warp == 1 && lane == 0) {
        const uint32_t sa = smem_u32(base), sb = sa + a_tiles * 128 * KB;
        const long long t0 = clock64();
        const int nacc = 512 / N;
        const uint64_t hi = ((uint64_t)((8 * KB) >> 4) << 32) | (1ull << 46) | (layout << 61);
        for (int it = 0; it < iters; it += 8) {
#pragma unroll
            for (int u = 0; u < 8; u++) {
                // operand tiles rotate with compile-time offsets; a_tiles / b_tiles are powers of two or 7 (masked to < 8)
                const int ia = (u * 3) % 8 < a_tiles ? (u * 3) % 8 : 0, ib = (u * 5) % 8 < b_tiles ? (u * 5) % 8 : 0;
#pragma unroll
                for (int ks = 0; ks < KB / 32; ks++) {
                    const uint64_t ad = (uint64_t)(((sa + ia * 128 * KB + ks * 32) & 0x3FFFF) >> 4) | hi;
                    const uint64_t bd = (uint64_t)(((sb + ib * N * KB + ks * 32) & 0x3FFFF) >> 4) | hi;
                    umma_i8(tm + (u % nacc) * N, ad, bd, idesc, 1u);
                }
            }
        }
        umma_commit(&bar);
        mbar_wait(&bar, 0);
        const long long t1 = clock64();
        if (blockIdx.x == 0) cycles[0] = t1 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tm, 512);
}

template <int N, int KB>
static void mma_rate(int a_tiles, int b_tiles) {
    long long* dc;
    CHECK(cudaMalloc(&dc, 8));
    const int smem = a_tiles * 128 * KB + b_tiles * N * KB + 2048;
    CHECK(cudaFuncSetAttribute(mma_rate_kernel<N, KB>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const int iters = 8192;
    mma_rate_kernel<N, KB><<<148, 128, smem>>>(iters, a_tiles, b_tiles, dc);
    CHECK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    mma_rate_kernel<N, KB><<<148, 128, smem>>>(iters, a_tiles, b_tiles, dc);
    cudaEventRecord(e1);
    CHECK(cudaEventSynchronize(e1));
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long cyc; CHECK(cudaMemcpy(&cyc, dc, 8, cudaMemcpyDeviceToHost));
    const double nm = (double)iters * (KB / 32);
    printf("i8 MMA M=128 N=%d (K chunk %d B, %d A tiles, %d B tiles): %.1f cycles per MMA (ideal %.1f), %.0f TOPS on 148 SMs\n", N, KB,
           a_tiles, b_tiles, (double)cyc / nm, 128.0 * N * 32 / 7736.0 * 1.0, 148.0 * nm * 2.0 * 128 * N * 32 / (ms * 1e-3) * 1e-12);
    fflush(stdout);
    cudaFree(dc);
}

int main(int argc, char** argv) {
    const char* what = argc > 1 ? argv[1] : "all";
    int fails = 0;
    CHECK(oz_set_attributes());
    CHECK(gemm_set_attributes());
    if (!strcmp(what, "rate")) {
        mma_rate<64, 64>(7, 7);
        mma_rate<64, 64>(1, 1);
        mma_rate<128, 64>(7, 7);
        mma_rate<128, 64>(1, 1);
        mma_rate<256, 64>(4, 4);
        mma_rate<256, 64>(1, 1);
        mma_rate<64, 128>(4, 4);
        mma_rate<128, 128>(4, 4);
        mma_rate<256, 128>(2, 2);
        mma_rate<64, 32>(7, 7);
        mma_rate<128, 32>(7, 7);
        return 0;
    }
    if (!strcmp(what, "check") || !strcmp(what, "all")) {
        fails += check_small();
        fails += check_lauum(512);
        fails += check_lauum(1152);
        fails += check_batched();
        printf("check: %s (%d failing groups)\n", fails ? "FAILED" : "ok", fails);
        fflush(stdout);
    }
    if ((!strcmp(what, "perf") || !strcmp(what, "all")) && fails == 0) {
        if (argc > 4) {
            perf(atoi(argv[2]), atoi(argv[3]), atoi(argv[4]), argc > 5 ? atoi(argv[5]) : 3);
        } else {
            perf(4096, 8, 0, 5);
            perf(8192, 12, 0, 5);
            perf(16384, 12, 0, 3);
            perf(16384, 24, 0, 3);
            perf(8192, 64, 1, 3);
            perf(16384, 128, 1, 2);
        }
    }
    return fails ? 1 : 0;
}
