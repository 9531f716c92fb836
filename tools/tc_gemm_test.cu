// Stand-alone bring-up test of the tcgen05/TMA GEMM (run on the GPU box):  tools/_bin/tc_gemm_test
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "../financial_market_data_analysis_b200/csrc/tc_gemm.cuh"

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 1; } } while (0)

static float bf(float x) { return __bfloat162float(__float2bfloat16(x)); }

static int run_case(int M, int N, int K, int mode, int splitk, int kshift, bool use_bias, int a_off, int b_off, int a_mn = 0, int b_mn = 0) {
    const int Arows = M + a_off, Brows = N + b_off;
    std::vector<float> A((size_t)Arows * K), B((size_t)Brows * K), bias(N);
    srand(M * 31 + N * 7 + K);
    for (auto& v : A) v = bf((rand() % 2001 - 1000) / 1000.f);
    for (auto& v : B) v = bf((rand() % 2001 - 1000) / 1000.f);
    for (auto& v : bias) v = (rand() % 2001 - 1000) / 500.f;
    std::vector<__nv_bfloat16> Ah(A.size()), Bh(B.size());
    // device copies: K-major [rows][K], or MN-major (transposed) [K][rows]
    for (int r = 0; r < Arows; ++r) for (int k = 0; k < K; ++k)
        Ah[a_mn ? (size_t)k * Arows + r : (size_t)r * K + k] = __float2bfloat16(A[(size_t)r * K + k]);
    for (int r = 0; r < Brows; ++r) for (int k = 0; k < K; ++k)
        Bh[b_mn ? (size_t)k * Brows + r : (size_t)r * K + k] = __float2bfloat16(B[(size_t)r * K + k]);
    __nv_bfloat16 *dA, *dB; float* dbias; void* dC; unsigned int* dbg;
    CK(cudaMalloc(&dA, Ah.size() * 2)); CK(cudaMalloc(&dB, Bh.size() * 2)); CK(cudaMalloc(&dbias, N * 4));
    CK(cudaMalloc(&dC, (size_t)M * N * 4)); CK(cudaMalloc(&dbg, 64));
    CK(cudaMemcpy(dA, Ah.data(), Ah.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, Bh.data(), Bh.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dbias, bias.data(), N * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(dC, 0, (size_t)M * N * 4)); CK(cudaMemset(dbg, 0, 64));
    CUtensorMap tA, tB;
    int me1 = a_mn ? tcg::make_operand_map_mn(&tA, dA, K, Arows, Arows) : tcg::make_operand_map(&tA, dA, Arows, K, K);
    int me2 = b_mn ? tcg::make_operand_map_mn(&tB, dB, K, Brows, Brows) : tcg::make_operand_map(&tB, dB, Brows, K, K);
    if (me1 || me2) { printf("tensor map failed\n"); return 1; }
    tcg::Params p{};
    p.M = M; p.N = N; p.K = K; p.batch = 1; p.splitk = splitk; p.mode = mode; p.C = dC; p.ldc = N; p.zC = 0;
    p.a_row_off[0] = a_off; p.b_row_off[0] = b_off; p.b_k_off[0] = kshift; p.a_mn = a_mn; p.b_mn = b_mn; p.bias = use_bias ? dbias : nullptr; p.zBias = 0; p.dbg = dbg;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    CK(tcg::launch(tA, tB, p, 0));
    CK(cudaDeviceSynchronize());
    float ms = 0.f;
    if (mode != tcg::OUT_ATOMIC_F32 || M * (long)N * K > (1L << 34)) {
        cudaEventRecord(e0);
        for (int i = 0; i < 5; ++i) CK(tcg::launch(tA, tB, p, 0));
        cudaEventRecord(e1);
        CK(cudaDeviceSynchronize());
        cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
    }
    unsigned int h[8]; CK(cudaMemcpy(h, dbg, 32, cudaMemcpyDeviceToHost));
    std::vector<float> C((size_t)M * N);
    if (mode == tcg::OUT_BF16) {
        std::vector<__nv_bfloat16> Cb((size_t)M * N);
        CK(cudaMemcpy(Cb.data(), dC, Cb.size() * 2, cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < C.size(); ++i) C[i] = __bfloat162float(Cb[i]);
    } else CK(cudaMemcpy(C.data(), dC, C.size() * 4, cudaMemcpyDeviceToHost));
    // reference on a sample of entries (all for small problems)
    double maxerr = 0, maxref = 0; long checked = 0;
    const long total = (long)M * N; const long stride = total > 400000 ? total / 200003 : 1;
    for (long idx = 0; idx < total; idx += stride) {
        const int m = idx / N, n = idx % N;
        double s = 0;
        for (int k = 0; k < K; ++k) {
            const int kb = k + kshift;
            if (kb < 0 || kb >= K) continue;
            s += (double)A[(size_t)(m + a_off) * K + k] * B[(size_t)(n + b_off) * K + kb];
        }
        if (use_bias) s += bias[n];
        const double err = fabs(s - C[idx]);
        if (err > maxerr) maxerr = err;
        if (fabs(s) > maxref) maxref = fabs(s);
        ++checked;
    }
    const double tol = mode == tcg::OUT_BF16 ? 1e-2 * (maxref + 1) : 2e-4 * (maxref + 1);
    const bool pass = h[0] == 0 && maxerr < tol;
    printf("%s mn=%d%d M=%d N=%d K=%d mode=%d splitk=%d kshift=%d bias=%d off=(%d,%d): maxerr=%.3e (ref max %.2f, %ld checked) dbg=%x/%u/%u  %.3f ms %.1f TFLOP/s\n",
           pass ? "PASS" : "FAIL", a_mn, b_mn, M, N, K, mode, splitk, kshift, (int)use_bias, a_off, b_off, maxerr, maxref, checked, h[0], h[1], h[2], ms,
           ms > 0 ? 2.0 * M * N * K / ms / 1e9 : 0.0);
    cudaFree(dA); cudaFree(dB); cudaFree(dbias); cudaFree(dC); cudaFree(dbg);
    return pass ? 0 : 2;
}

// x3 product: fp32 operands split into (hi, lo) bf16 pairs, A_hi B_hi + A_hi B_lo + A_lo B_hi  (nsplit = 3)
static int run_split_case(int M, int N, int K, int mode, int splitk, int a_mn, int b_mn) {
    std::vector<float> A((size_t)M * K), B((size_t)N * K);
    srand(M * 13 + N * 5 + K);
    for (auto& v : A) v = (rand() % 200001 - 100000) / 100000.f;
    for (auto& v : B) v = (rand() % 200001 - 100000) / 100000.f;
    std::vector<__nv_bfloat16> Ah(A.size()), Al(A.size()), Bh(B.size()), Bl(B.size());
    auto put = [](std::vector<__nv_bfloat16>& hi, std::vector<__nv_bfloat16>& lo, size_t i, float x) {
        hi[i] = __float2bfloat16(x); lo[i] = __float2bfloat16(x - __bfloat162float(hi[i]));
    };
    for (int r = 0; r < M; ++r) for (int k = 0; k < K; ++k) put(Ah, Al, a_mn ? (size_t)k * M + r : (size_t)r * K + k, A[(size_t)r * K + k]);
    for (int r = 0; r < N; ++r) for (int k = 0; k < K; ++k) put(Bh, Bl, b_mn ? (size_t)k * N + r : (size_t)r * K + k, B[(size_t)r * K + k]);
    __nv_bfloat16 *dAh, *dAl, *dBh, *dBl; float* dC; unsigned int* dbg;
    CK(cudaMalloc(&dAh, Ah.size() * 2)); CK(cudaMalloc(&dAl, Ah.size() * 2)); CK(cudaMalloc(&dBh, Bh.size() * 2)); CK(cudaMalloc(&dBl, Bh.size() * 2));
    CK(cudaMalloc(&dC, (size_t)M * N * 4)); CK(cudaMalloc(&dbg, 64));
    CK(cudaMemcpy(dAh, Ah.data(), Ah.size() * 2, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dAl, Al.data(), Al.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dBh, Bh.data(), Bh.size() * 2, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dBl, Bl.data(), Bl.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemset(dC, 0, (size_t)M * N * 4)); CK(cudaMemset(dbg, 0, 64));
    CUtensorMap tA, tB, tAl, tBl;
    int e = a_mn ? tcg::make_operand_map_mn(&tA, dAh, K, M, M) | tcg::make_operand_map_mn(&tAl, dAl, K, M, M)
                 : tcg::make_operand_map(&tA, dAh, M, K, K) | tcg::make_operand_map(&tAl, dAl, M, K, K);
    e |= b_mn ? tcg::make_operand_map_mn(&tB, dBh, K, N, N) | tcg::make_operand_map_mn(&tBl, dBl, K, N, N)
              : tcg::make_operand_map(&tB, dBh, N, K, K) | tcg::make_operand_map(&tBl, dBl, N, K, K);
    if (e) { printf("tensor map failed\n"); return 1; }
    tcg::Params p{};
    p.M = M; p.N = N; p.K = K; p.batch = 1; p.splitk = splitk; p.mode = mode; p.C = dC; p.ldc = N; p.a_mn = a_mn; p.b_mn = b_mn; p.dbg = dbg; p.nsplit = 3;
    CK(tcg::launch(tA, tB, p, 0, &tAl, &tBl));
    CK(cudaDeviceSynchronize());
    float ms = 0.f;
    if (mode != tcg::OUT_ATOMIC_F32) {
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        for (int i = 0; i < 5; ++i) CK(tcg::launch(tA, tB, p, 0, &tAl, &tBl));
        cudaEventRecord(e1); CK(cudaDeviceSynchronize());
        cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
    }
    unsigned int h[8]; CK(cudaMemcpy(h, dbg, 32, cudaMemcpyDeviceToHost));
    std::vector<float> C((size_t)M * N);
    CK(cudaMemcpy(C.data(), dC, C.size() * 4, cudaMemcpyDeviceToHost));
    double maxerr = 0, maxref = 0, rms = 0; long checked = 0;
    const long total = (long)M * N; const long stride = total > 200000 ? total / 100003 : 1;
    for (long idx = 0; idx < total; idx += stride) {
        const int m = idx / N, n = idx % N;
        double s = 0, s2 = 0;
        for (int k = 0; k < K; ++k) { const double t = (double)A[(size_t)m * K + k] * B[(size_t)n * K + k]; s += t; s2 += t * t; }
        maxerr = fmax(maxerr, fabs(s - C[idx])); maxref = fmax(maxref, fabs(s)); rms = fmax(rms, sqrt(s2)); ++checked;
    }
    const bool pass = h[0] == 0 && maxerr < 2e-4 * (rms + 1e-3);     // fp32 accumulation over K up to 65536 per split (measured ~1e-4 of the term rms at K = 6.5 k per CTA)
    printf("%s x3 mn=%d%d M=%d N=%d K=%d mode=%d splitk=%d: maxerr=%.3e (ref max %.2f, term rms %.2f, %ld checked) dbg=%x  %.3f ms %.1f fp32-class TFLOP/s\n",
           pass ? "PASS" : "FAIL", a_mn, b_mn, M, N, K, mode, splitk, maxerr, maxref, rms, checked, h[0], ms, ms > 0 ? 2.0 * M * N * K / ms / 1e9 : 0.0);
    cudaFree(dAh); cudaFree(dAl); cudaFree(dBh); cudaFree(dBl); cudaFree(dC); cudaFree(dbg);
    return pass ? 0 : 2;
}

// determinism of the blocked-layout (OUT_SCAN_*) epilogue at the train step's shapes: the same GEMM twice, outputs compared
// byte for byte (no atomics on this path: any difference is a race)
static int run_scan_determinism(int M, int N, int K, int mode, int nsplit, int U, int NBt, int Bb, int Tt, int H, int G, int reps) {
    std::vector<__nv_bfloat16> Ah((size_t)M * K), Bh((size_t)N * K);
    srand(M + N + K);
    for (auto& v : Ah) v = __float2bfloat16((rand() % 2001 - 1000) / 1000.f);
    for (auto& v : Bh) v = __float2bfloat16((rand() % 2001 - 1000) / 1000.f);
    std::vector<float> bias(M, 0.5f);
    __nv_bfloat16 *dA, *dB; float* dbias; uint8_t *dC1, *dC2; unsigned int* dbg;
    const size_t es = mode == tcg::OUT_SCAN_BF16 ? 2 : 4, cbytes = (size_t)M * N * es;
    CK(cudaMalloc(&dA, Ah.size() * 2)); CK(cudaMalloc(&dB, Bh.size() * 2)); CK(cudaMalloc(&dbias, M * 4));
    CK(cudaMalloc(&dC1, cbytes)); CK(cudaMalloc(&dC2, cbytes)); CK(cudaMalloc(&dbg, 64));
    CK(cudaMemcpy(dA, Ah.data(), Ah.size() * 2, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, Bh.data(), Bh.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dbias, bias.data(), M * 4, cudaMemcpyHostToDevice)); CK(cudaMemset(dbg, 0, 64));
    CUtensorMap tA, tB;
    if (tcg::make_operand_map(&tA, dA, M, K, K) | tcg::make_operand_map(&tB, dB, N, K, K)) { printf("tensor map failed\n"); return 1; }
    tcg::Params p{};
    p.M = M; p.N = N; p.K = K; p.batch = 1; p.splitk = 1; p.mode = mode; p.ldc = N; p.bias = dbias; p.bias_per_row = 1; p.m_fast = 1; p.dbg = dbg;
    p.blk = tcg::ScanBlk{Tt, Bb, H, G, U, NBt}; p.nsplit = nsplit;
    std::vector<uint8_t> h1(cbytes), h2(cbytes);
    long worst = 0; int badruns = 0;
    for (int r = 0; r < reps; ++r) {
        CK(cudaMemset(dC1, 0xff, cbytes)); CK(cudaMemset(dC2, 0xff, cbytes));
        p.C = dC1; CK(tcg::launch(tA, tB, p, 0, nsplit == 3 ? &tA : nullptr, nsplit == 3 ? &tB : nullptr));
        p.C = dC2; CK(tcg::launch(tA, tB, p, 0, nsplit == 3 ? &tA : nullptr, nsplit == 3 ? &tB : nullptr));
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(h1.data(), dC1, cbytes, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(h2.data(), dC2, cbytes, cudaMemcpyDeviceToHost));
        long nd = 0, unwritten = 0;
        for (size_t i = 0; i < cbytes; ++i) nd += h1[i] != h2[i];
        for (size_t i = 0; i + 3 < cbytes; i += 4) unwritten += (h1[i] == 0xff && h1[i + 1] == 0xff && h1[i + 2] == 0xff && h1[i + 3] == 0xff);
        if (nd || unwritten) { ++badruns; if (nd > worst) worst = nd; printf("   rep %d: %ld bytes differ, %ld words never written\n", r, nd, unwritten); }
    }
    unsigned int h[8]; CK(cudaMemcpy(h, dbg, 32, cudaMemcpyDeviceToHost));
    printf("%s scan-layout determinism M=%d N=%d K=%d mode=%d nsplit=%d U=%d NB=%d: %d of %d repetitions differ (worst %ld bytes) dbg=%x\n",
           badruns ? "FAIL" : "PASS", M, N, K, mode, nsplit, U, NBt, badruns, reps, worst, h[0]);
    cudaFree(dA); cudaFree(dB); cudaFree(dbias); cudaFree(dC1); cudaFree(dC2); cudaFree(dbg);
    return badruns ? 2 : 0;
}

int main() {
    setvbuf(stdout, NULL, _IONBF, 0);
    int bad = 0;
    if (getenv("GEMM_DET")) {
        bad += run_scan_determinism(1536, 65536, 64, tcg::OUT_SCAN_F32, 3, 64, 32, 512, 128, 256, 3, 6);
        bad += run_scan_determinism(1536, 65536, 512, tcg::OUT_SCAN_F32, 3, 64, 32, 512, 128, 256, 3, 4);
        bad += run_scan_determinism(1536, 65536, 512, tcg::OUT_SCAN_BF16, 1, 128, 16, 512, 128, 256, 3, 4);
        bad += run_scan_determinism(512, 65536, 1536, tcg::OUT_SCAN_F32, 3, 64, 32, 512, 128, 256, 1, 4);
        printf(bad ? "SOME FAILED\n" : "ALL PASSED\n");
        return bad ? 1 : 0;
    }
    bad += run_split_case(128, 128, 64, tcg::OUT_F32, 1, 0, 0);
    bad += run_split_case(256, 384, 512, tcg::OUT_F32, 1, 0, 0);
    bad += run_split_case(768, 256, 8192, tcg::OUT_ATOMIC_F32, 8, 1, 1);
    bad += run_split_case(1536, 65536, 512, tcg::OUT_F32, 1, 0, 0);
    bad += run_split_case(512, 65536, 1536, tcg::OUT_F32, 1, 0, 0);
    bad += run_split_case(768, 512, 65536, tcg::OUT_ATOMIC_F32, 10, 1, 1);
    bad += run_case(128, 128, 64, tcg::OUT_F32, 1, 0, false, 0, 0);
    bad += run_case(128, 128, 256, tcg::OUT_F32, 1, 0, true, 0, 0);
    bad += run_case(256, 384, 512, tcg::OUT_BF16, 1, 0, true, 0, 0);
    bad += run_case(300, 200, 136, tcg::OUT_F32, 1, 0, true, 0, 0);
    bad += run_case(300, 200, 136, tcg::OUT_BF16, 1, 0, false, 128, 256);
    bad += run_case(768, 256, 8192, tcg::OUT_ATOMIC_F32, 8, 0, false, 0, 0);
    bad += run_case(768, 256, 8192, tcg::OUT_ATOMIC_F32, 16, -512, false, 768, 256);
    bad += run_case(768, 256, 8192, tcg::OUT_ATOMIC_F32, 16, 512, false, 0, 0);
    bad += run_case(128, 128, 64, tcg::OUT_F32, 1, 0, false, 0, 0, 1, 0);
    bad += run_case(128, 128, 64, tcg::OUT_F32, 1, 0, false, 0, 0, 0, 1);
    bad += run_case(256, 384, 320, tcg::OUT_F32, 1, 0, true, 0, 0, 1, 1);
    bad += run_case(768, 512, 8192, tcg::OUT_ATOMIC_F32, 16, 0, false, 768, 0, 1, 1);
    bad += run_case(768, 256, 8192, tcg::OUT_ATOMIC_F32, 16, -512, false, 0, 256, 1, 1);
    bad += run_case(768, 256, 65536, tcg::OUT_ATOMIC_F32, 37, 512, false, 0, 0, 1, 1);
    bad += run_case(768, 512, 65536, tcg::OUT_ATOMIC_F32, 10, 0, false, 0, 0, 1, 1);
    bad += run_case(768, 512, 65536, tcg::OUT_ATOMIC_F32, 10, 0, false, 0, 0, 0, 0);
    bad += run_case(768, 512, 65536, tcg::OUT_ATOMIC_F32, 20, 0, false, 0, 0, 1, 1);
    bad += run_case(512, 256, 65536, tcg::OUT_ATOMIC_F32, 37, 0, false, 0, 0, 1, 1);
    bad += run_case(65536, 1536, 64, tcg::OUT_BF16, 1, 0, true, 0, 0);
    bad += run_case(65536, 1536, 512, tcg::OUT_BF16, 1, 0, true, 0, 0);
    bad += run_case(65536, 512, 1536, tcg::OUT_F32, 1, 0, false, 0, 0);
    bad += run_case(8192, 8192, 8192, tcg::OUT_BF16, 1, 0, false, 0, 0);
    printf(bad ? "SOME FAILED\n" : "ALL PASSED\n");
    return bad ? 1 : 0;
}
