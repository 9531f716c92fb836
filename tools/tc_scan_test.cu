// Stand-alone bring-up test of the persistent tcgen05 GRU scan (forward).  tools/_bin/tc_scan_test
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "../financial_market_data_analysis_b200/csrc/tc_scan.cuh"

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 1; } } while (0)
static float bf(float x) { return __bfloat162float(__float2bfloat16(x)); }
static float frand() { return (rand() % 20001 - 10000) / 10000.f; }

static int run_case(int B, int T, int H, int D, int reps, int fuse = 0) {
    const int CS = H / 128; const long R = (long)T * B;
    srand(B + 3 * T + H);
    std::vector<float> whh((size_t)D * 3 * H * H), bhn((size_t)D * H), gi((size_t)R * D * 3 * H);
    const float sc = 1.f / sqrtf((float)H);
    for (auto& v : whh) v = frand() * sc;
    for (auto& v : bhn) v = frand() * sc;
    for (auto& v : gi) v = bf(frand() * 1.5f);
    // fused input projection: gi = W_ih x + b is formed inside the kernel from x [R][64], W_ih [D*3H][64], b [D*3H]
    std::vector<float> xs, wih, bfold;
    __nv_bfloat16 *d_x = nullptr, *d_wih = nullptr; float* d_bfold = nullptr;
    if (fuse) {
        xs.resize((size_t)R * 64); wih.resize((size_t)D * 3 * H * 64); bfold.resize((size_t)D * 3 * H);
        for (auto& v : xs) v = bf(frand());
        for (auto& v : wih) v = bf(frand() * 0.2f);
        for (auto& v : bfold) v = frand() * 0.3f;
        for (long r = 0; r < R; ++r)
            for (int q = 0; q < D * 3 * H; ++q) {
                double a = bfold[q];
                for (int k = 0; k < 64; ++k) a += (double)wih[(size_t)q * 64 + k] * xs[(size_t)r * 64 + k];
                gi[(size_t)r * D * 3 * H + q] = (float)a;
            }
        std::vector<__nv_bfloat16> xb(xs.size()), wb(wih.size());
        for (size_t i = 0; i < xs.size(); ++i) xb[i] = __float2bfloat16(xs[i]);
        for (size_t i = 0; i < wih.size(); ++i) wb[i] = __float2bfloat16(wih[i]);
        CK(cudaMalloc(&d_x, xb.size() * 2)); CK(cudaMalloc(&d_wih, wb.size() * 2)); CK(cudaMalloc(&d_bfold, bfold.size() * 4));
        CK(cudaMemcpy(d_x, xb.data(), xb.size() * 2, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(d_wih, wb.data(), wb.size() * 2, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(d_bfold, bfold.data(), bfold.size() * 4, cudaMemcpyHostToDevice));
    }
    float* d_whh; float* d_bhn; __nv_bfloat16 *d_img, *d_gi, *d_Y, *d_YT, *d_G; float* d_hn; unsigned int* dbg;
    const size_t img_elems = (size_t)D * 3 * H * H;
    CK(cudaMalloc(&d_whh, whh.size() * 4)); CK(cudaMalloc(&d_bhn, bhn.size() * 4)); CK(cudaMalloc(&d_img, img_elems * 2));
    CK(cudaMalloc(&d_gi, gi.size() * 2)); CK(cudaMalloc(&d_Y, (size_t)R * D * H * 2)); CK(cudaMalloc(&d_YT, (size_t)R * D * H * 2));
    CK(cudaMalloc(&d_G, (size_t)R * D * 4 * H * 2)); CK(cudaMalloc(&d_hn, (size_t)D * B * H * 4)); CK(cudaMalloc(&dbg, 64));
    std::vector<__nv_bfloat16> gih(gi.size());        // device copy is BLOCKED: [d][tile][t][cta][thread][gate][8]
    for (long r = 0; r < R; ++r)
        for (int q = 0; q < D * 3 * H; ++q) {
            const int dd = q / (3 * H), g = (q / H) % 3, unit = q % H, t = (int)(r / B), b = (int)(r % B);
            const int tile = b / 16, half = (b % 16) / 8, i = b % 8, cta = unit / 128, ju = unit % 128;
            const int tid = ((ju / 32) + 4 * half) * 32 + (ju % 32);
            const size_t e = ((((size_t)dd * (B / 16) + tile) * T + t) * CS + cta) * 256 + tid;
            gih[(((e / 256) * 3 + g) * 256 + tid) * 8 + i] = __float2bfloat16(gi[(size_t)r * D * 3 * H + q]);
        }
    CK(cudaMemcpy(d_whh, whh.data(), whh.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_bhn, bhn.data(), bhn.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_gi, gih.data(), gih.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemset(dbg, 0, 64)); CK(cudaMemset(d_Y, 0, (size_t)R * D * H * 2));
    for (int d = 0; d < D; ++d) {
        tcs::pack_whh_image_kernel<<<256, 256>>>(d_whh + (size_t)d * 3 * H * H, d_img + (size_t)d * 3 * H * H, H);
        CK(cudaGetLastError());
    }
    tcs::FwdParams p{};
    p.B = B; p.T = T; p.H = H; p.D = D; p.Wimg = d_img; p.giB = d_gi; p.b_hn = d_bhn; p.Yrow = d_Y; p.G = d_G;
    __nv_bfloat16* d_YB; CK(cudaMalloc(&d_YB, (size_t)R * D * H * 2)); p.YB = d_YB;
    p.hn_out = d_hn; p.dbg = dbg;
    p.fuse_x = fuse; p.Xrow = d_x; p.Wih = d_wih; p.bfold = d_bfold;
#ifdef BIGRU_SCAN_TIMING
    unsigned long long* d_ts; CK(cudaMalloc(&d_ts, 8 * 16 * 8)); CK(cudaMemset(d_ts, 0, 8 * 16 * 8)); p.ts = d_ts;
#endif
    CK(tcs::launch_fwd(p, 0));
    CK(cudaDeviceSynchronize());
    float ms = 0;
    if (reps > 0) {
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        for (int i = 0; i < reps; ++i) CK(tcs::launch_fwd(p, 0));
        cudaEventRecord(e1); CK(cudaDeviceSynchronize());
        cudaEventElapsedTime(&ms, e0, e1); ms /= reps;
    }
    unsigned int h[8]; CK(cudaMemcpy(h, dbg, 32, cudaMemcpyDeviceToHost));
#ifdef BIGRU_SCAN_TIMING
    if (T >= 80) {
        unsigned long long ts[8 * 16]; CK(cudaMemcpy(ts, d_ts, sizeof(ts), cudaMemcpyDeviceToHost));
        // slots: control 0 epi_done seen, 1 local MMAs issued, 2 h_full seen, 3 all MMAs issued, 4 committed;
        // epilogue thread 0: 5 loop top, 6 gi ready, 7 mma_done seen, 8 tmem loaded, 9 h written, 10 arrived, 11 stores issued;
        // epilogue thread 224: 12 mma_done seen, 13 arrived
        printf("timing (cycles, CTA 0, H=%d): step | c:wait->loc | loc->hfull | hfull->issued | commit || e: commit->mma_done | ld | math+sts | fence+arrive | arrive->c.wake | stores | step total\n", H);
        for (int k = 1; k < 7; ++k) {
            const unsigned long long* a = ts + k * 16; const unsigned long long* nx = ts + (k + 1) * 16;
            printf("  s=%d | %5lld | %5lld | %5lld | %5lld || %5lld (w7 %5lld) | %5lld | %5lld | %5lld (w7 %5lld) | %5lld | %5lld | %5lld   [gi wait %lld]\n", 64 + k,
                   (long long)(a[1] - a[0]), (long long)(a[2] - a[1]), (long long)(a[3] - a[2]), (long long)(a[4] - a[3]),
                   (long long)(a[7] - a[4]), (long long)(a[12] - a[4]), (long long)(a[8] - a[7]), (long long)(a[9] - a[8]), (long long)(a[10] - a[9]),
                   (long long)(a[13] - a[9]), (long long)(nx[0] - a[10]), (long long)(a[11] - a[10]), (long long)(nx[0] - a[0]), (long long)(a[6] - a[5]));
        }
    }
#endif
    std::vector<__nv_bfloat16> Y((size_t)R * D * H), YT((size_t)R * D * H), G((size_t)R * D * 4 * H);
    std::vector<float> hn((size_t)D * B * H);
    CK(cudaMemcpy(Y.data(), d_Y, Y.size() * 2, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(YT.data(), d_YT, YT.size() * 2, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(G.data(), d_G, G.size() * 2, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hn.data(), d_hn, hn.size() * 4, cudaMemcpyDeviceToHost));
    // CPU reference with the same roundings: bf16 weights, bf16 h into the product, fp32 state
    double eY = 0, eYT = 0, eG = 0, eHn = 0;
    const int bcheck = B > 48 ? 48 : B;                     // first rows of the batch (several tiles) are enough
    std::vector<float> wq(whh.size());
    for (size_t i = 0; i < whh.size(); ++i) wq[i] = bf(whh[i]);
    for (int d = 0; d < D; ++d)
        for (int b = 0; b < bcheck; ++b) {
            std::vector<float> hs(H, 0.f), hq(H, 0.f), hnew(H);
            for (int s = 0; s < T; ++s) {
                const int t = d == 0 ? s : T - 1 - s;
                const long row = (long)t * B + b;
                for (int j = 0; j < H; ++j) {
                    double a[3] = {0, 0, 0};
                    for (int g = 0; g < 3; ++g) {
                        const float* w = &wq[((size_t)d * 3 * H + g * H + j) * H];
                        double acc = 0;
                        for (int k = 0; k < H; ++k) acc += (double)w[k] * hq[k];
                        a[g] = acc;
                    }
                    const float* gp = &gi[row * D * 3 * H + d * 3 * H];
                    const float r = 1.f / (1.f + expf(-(gp[j] + (float)a[0])));
                    const float z = 1.f / (1.f + expf(-(gp[H + j] + (float)a[1])));
                    const float hnv = (float)a[2] + bhn[d * H + j];
                    const float n = tanhf(gp[2 * H + j] + r * hnv);
                    hnew[j] = n + z * (hs[j] - n);
                    // thread-private stash: [d][tile][t][cta][thread = warp*32+lane][gate][8 columns]
                    const int cta = j / 128, jj = j % 128, tile = b / 16, bb = b % 16;
                    const int tid = ((jj / 32) + 4 * (bb / 8)) * 32 + (jj % 32);
                    const size_t blk = (((size_t)d * (B / 16) + tile) * T + t) * CS + cta;       // [block][gate][thread][8]
                    const size_t gidx = (blk * 4 * 256 + tid) * 8 + (bb % 8);
                    eG = fmax(eG, fabs(r - __bfloat162float(G[gidx])));
                    eG = fmax(eG, fabs(z - __bfloat162float(G[gidx + 256 * 8])));
                    eG = fmax(eG, fabs(n - __bfloat162float(G[gidx + 2 * 256 * 8])));
                    eG = fmax(eG, fabs(hnv - __bfloat162float(G[gidx + 3 * 256 * 8])));
                }
                for (int j = 0; j < H; ++j) {
                    hs[j] = hnew[j]; hq[j] = bf(hnew[j]);
                    eY = fmax(eY, fabs(hnew[j] - __bfloat162float(Y[(size_t)row * D * H + d * H + j])));

                }
            }
            for (int j = 0; j < H; ++j) eHn = fmax(eHn, fabs(hs[j] - hn[((size_t)d * B + b) * H + j]));
        }
    const bool pass = h[0] == 0 && eY < 2e-2 && eYT < 2e-2 && eG < 3e-2 && eHn < 2e-2;
    printf("%s scan_fwd%s B=%d T=%d H=%d D=%d (cluster %d, grid %d): errY=%.2e errYT=%.2e errG=%.2e errHn=%.2e dbg=%x blk=%u thr=%u a=%u  %.3f ms (%.2f us/step)\n",
           pass ? "PASS" : "FAIL", fuse ? "+x" : "", B, T, H, D, CS, D * (B / 16) * CS, eY, eYT, eG, eHn, h[0], h[1], h[2], h[3], ms, ms * 1e3 / T);
    cudaFree(d_whh); cudaFree(d_bhn); cudaFree(d_img); cudaFree(d_gi); cudaFree(d_Y); cudaFree(d_YT); cudaFree(d_G); cudaFree(d_hn); cudaFree(dbg);
    return pass ? 0 : 2;
}

// timing-only run of the backward scan (zero-filled inputs: there is no data-dependent control flow)
static int run_bwd_timing(int B, int T, int H, int D, int top) {
    const long R = (long)T * B;
    __nv_bfloat16 *WT, *G, *YB, *dgi, *dghn; float *dYB, *db, *dlog, *linw; int* arg; unsigned int* dbg;
    CK(cudaMalloc(&WT, (size_t)D * 3 * H * H * 2)); CK(cudaMemset(WT, 0, (size_t)D * 3 * H * H * 2));
    CK(cudaMalloc(&G, (size_t)R * D * 4 * H * 2)); CK(cudaMemset(G, 0, (size_t)R * D * 4 * H * 2));
    CK(cudaMalloc(&YB, (size_t)R * D * H * 2)); CK(cudaMemset(YB, 0, (size_t)R * D * H * 2));
    CK(cudaMalloc(&dYB, (size_t)R * D * H * 4)); CK(cudaMemset(dYB, 0, (size_t)R * D * H * 4));
    CK(cudaMalloc(&dgi, (size_t)R * D * 3 * H * 2)); CK(cudaMalloc(&dghn, (size_t)R * D * H * 2));
    CK(cudaMalloc(&db, (size_t)4 * D * 3 * H * 4)); CK(cudaMemset(db, 0, (size_t)4 * D * 3 * H * 4));
    CK(cudaMalloc(&dlog, (size_t)B * 3 * 4)); CK(cudaMemset(dlog, 0, (size_t)B * 3 * 4));
    CK(cudaMalloc(&linw, (size_t)3 * 3 * H * 4)); CK(cudaMemset(linw, 0, (size_t)3 * 3 * H * 4));
    CK(cudaMalloc(&arg, (size_t)B * H * 4)); CK(cudaMemset(arg, 0, (size_t)B * H * 4));
    CK(cudaMalloc(&dbg, 64)); CK(cudaMemset(dbg, 0, 64));
    tcs::BwdParams p{};
    p.B = B; p.T = T; p.H = H; p.D = D; p.WTimg = WT; p.G = G; p.YB = YB; p.dYB = dYB;
    if (top) { p.dlogits = dlog; p.lin_w = linw; p.arg = arg; p.C = 3; }
    p.dgi_row = dgi; p.dghn_row = dghn; p.db_ih = db; p.db_hh = db + (size_t)2 * D * 3 * H; p.dir_stride = 3 * H; p.dbg = dbg;
#ifdef BIGRU_SCAN_TIMING
    unsigned long long* d_ts; CK(cudaMalloc(&d_ts, 8 * 16 * 8)); CK(cudaMemset(d_ts, 0, 8 * 16 * 8)); p.ts = d_ts;
#endif
    CK(tcs::launch_bwd(p, 0)); CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    for (int i = 0; i < 10; ++i) CK(tcs::launch_bwd(p, 0));
    cudaEventRecord(e1); CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 10;
    unsigned int h[8]; CK(cudaMemcpy(h, dbg, 32, cudaMemcpyDeviceToHost));
    printf("scan_bwd timing B=%d T=%d H=%d D=%d top=%d: %.3f ms (%.2f us/step) dbg=%x\n", B, T, H, D, top, ms, ms * 1e3 / T, h[0]);
#ifdef BIGRU_SCAN_TIMING
    if (T >= 80) {
        unsigned long long ts[8 * 16]; CK(cudaMemcpy(ts, d_ts, sizeof(ts), cudaMemcpyDeviceToHost));
        printf("  bwd cycles: step | c:wake->loc | loc->dfull | dfull->commit || e: commit->wake | ld | math+sts | st.async+fence+arrive | arrive->c.wake | tail | total  [in wait %s, mma wait]\n", "");
        for (int k = 1; k < 7; ++k) {
            const unsigned long long* a = ts + k * 16; const unsigned long long* nx = ts + (k + 1) * 16;
            printf("   s=%d | %5lld | %5lld | %5lld || %5lld | %5lld | %5lld | %5lld | %5lld | %5lld | %5lld  [%lld, %lld]\n", 64 + k,
                   (long long)(a[1] - a[0]), (long long)(a[2] - a[1]), (long long)(a[3] - a[2]), (long long)(a[7] - a[3]), (long long)(a[8] - a[7]),
                   (long long)(a[9] - a[8]), (long long)(a[10] - a[9]), (long long)(nx[0] - a[10]), (long long)(a[11] - a[10]), (long long)(nx[0] - a[0]),
                   (long long)(a[6] - a[5]), (long long)(a[7] - a[4]));
        }
    }
#endif
    return h[0] ? 1 : 0;
}

int main() {
    setvbuf(stdout, NULL, _IONBF, 0);
    int bad = 0;
    bad += run_case(16, 1, 128, 1, 0);
    bad += run_case(16, 4, 128, 1, 0);
    bad += run_case(32, 6, 128, 2, 0);
    bad += run_case(16, 2, 256, 1, 0);
    bad += run_case(16, 5, 256, 1, 0);
    bad += run_case(64, 9, 256, 2, 0);
    bad += run_case(512, 128, 256, 2, 10);
    bad += run_case(512, 64, 128, 2, 10);
    bad += run_case(16, 1, 128, 1, 0, 1);
    bad += run_case(16, 3, 128, 1, 0, 1);
    bad += run_case(32, 12, 256, 2, 0, 1);
    bad += run_case(512, 128, 256, 2, 10, 1);
    bad += run_case(512, 64, 128, 2, 10, 1);
    bad += run_bwd_timing(512, 128, 256, 2, 0);
    bad += run_bwd_timing(512, 128, 256, 2, 1);
    printf(bad ? "SOME FAILED\n" : "ALL PASSED\n");
    return bad ? 1 : 0;
}
