// Timing-only comparison of the experimental GEMM forms (pair tiles / persistent CTAs) on the shapes of one train step.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "tc_gemm_x.cuh"
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 1; } } while (0)
static int bench(int M, int N, int K, int mode, int pair, int persist) {
    __nv_bfloat16 *dA, *dB; void* dC; unsigned int* dbg;
    CK(cudaMalloc(&dA, (size_t)M * K * 2)); CK(cudaMalloc(&dB, (size_t)N * K * 2)); CK(cudaMalloc(&dC, (size_t)M * N * 4)); CK(cudaMalloc(&dbg, 64));
    CK(cudaMemset(dA, 0x11, (size_t)M * K * 2)); CK(cudaMemset(dB, 0x11, (size_t)N * K * 2)); CK(cudaMemset(dbg, 0, 64));
    CUtensorMap tA, tB;
    if (tcg::make_operand_map(&tA, dA, M, K, K) || tcg::make_operand_map(&tB, dB, N, K, K)) { printf("map failed\n"); return 1; }
    tcg::Params p{};
    p.M = M; p.N = N; p.K = K; p.batch = 1; p.splitk = 1; p.mode = mode; p.C = dC; p.ldc = N; p.dbg = dbg; p.pair = pair; p.persist = persist; p.m_fast = M < N;
    CK(tcg::launch(tA, tB, p, 0)); CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    for (int i = 0; i < 10; ++i) CK(tcg::launch(tA, tB, p, 0));
    cudaEventRecord(e1); CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 10;
    unsigned int h[4]; CK(cudaMemcpy(h, dbg, 16, cudaMemcpyDeviceToHost));
    printf("M=%d N=%d K=%d mode=%d pair=%d persist=%d: %.1f us  %.0f TFLOP/s dbg=%x\n", M, N, K, mode, pair, persist, ms * 1e3, 2.0 * M * N * K / ms / 1e9, h[0]);
    cudaFree(dA); cudaFree(dB); cudaFree(dC); cudaFree(dbg);
    return 0;
}
int main() {
    for (int pair = 0; pair < 2; ++pair)
        for (int persist = 0; persist < 3; ++persist) {
            bench(1536, 65536, 512, tcg::OUT_BF16, pair, persist);
            bench(512, 65536, 1536, tcg::OUT_F32, pair, persist);
            bench(1536, 65536, 64, tcg::OUT_BF16, pair, persist);
        }
    return 0;
}
