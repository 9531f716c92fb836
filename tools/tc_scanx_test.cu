// Stand-alone bring-up test of the fp32-class (x3) persistent scans of tc_scan_x.cuh against a double-precision CPU
// recurrence / BPTT.  tools/_bin/tc_scanx_test
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "../financial_market_data_analysis_b200/csrc/tc_scan_x.cuh"

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 1; } } while (0)
static float frand() { return (rand() % 20001 - 10000) / 10000.f; }
static float b2f(__nv_bfloat16 v) { return __bfloat162float(v); }

// blocked index of (d, row b, t, unit, gate g of G)
static size_t bidx(int d, int b, int t, int unit, int g, int G, int B, int T, int H) {
    const int CS = H / 64, tile = b / 32, cb = b % 32, c = unit / 64, j = unit % 64;
    const size_t blk = (((size_t)d * (B / 32) + tile) * T + t) * CS + c;
    return ((blk * G + g) * 256 + j + 64 * (cb / 8)) * 8 + cb % 8;
}

template <class T_> static T_* dev(const std::vector<T_>& v) {
    T_* p = nullptr;
    if (cudaMalloc(&p, v.size() * sizeof(T_) + 16) != cudaSuccess) return nullptr;
    cudaMemcpy(p, v.data(), v.size() * sizeof(T_), cudaMemcpyHostToDevice);
    return p;
}

static int run_case(int B, int T, int H, int D, int reps, int use_h0, int top) {
    const long R = (long)T * B;
    const int CS = H / 64, C = 3;
    srand(B + 3 * T + H + 7 * use_h0);
    std::vector<float> whh((size_t)D * 3 * H * H), bhn((size_t)D * H), gi((size_t)R * D * 3 * H), h0((size_t)D * B * H), dY((size_t)R * D * H);
    const float sc = 1.f / sqrtf((float)H);
    for (auto& v : whh) v = frand() * sc;
    for (auto& v : bhn) v = frand() * sc;
    for (auto& v : gi) v = frand() * 1.5f;
    for (auto& v : h0) v = frand() * 0.8f;
    for (auto& v : dY) v = frand() * 0.01f;
    std::vector<float> dlog((size_t)B * C), linw((size_t)C * 3 * H);
    std::vector<int> arg((size_t)B * H);
    for (auto& v : dlog) v = frand() * 0.01f;
    for (auto& v : linw) v = frand() * 0.2f;
    for (auto& v : arg) v = rand() % T;
    // blocked device images
    std::vector<float> giB((size_t)R * D * 3 * H), dYB((size_t)R * D * H);
    for (int d = 0; d < D; ++d)
        for (int t = 0; t < T; ++t)
            for (int b = 0; b < B; ++b)
                for (int u = 0; u < H; ++u) {
                    for (int g = 0; g < 3; ++g) giB[bidx(d, b, t, u, g, 3, B, T, H)] = gi[((size_t)t * B + b) * D * 3 * H + d * 3 * H + g * H + u];
                    dYB[bidx(d, b, t, u, 0, 1, B, T, H)] = dY[((size_t)t * B + b) * D * H + d * H + u];
                }
    float *d_whh = dev(whh), *d_bhn = dev(bhn), *d_gi = dev(giB), *d_h0 = dev(h0), *d_dYB = dev(dYB), *d_dlog = dev(dlog), *d_linw = dev(linw);
    int* d_arg = dev(arg);
    const size_t img = (size_t)D * CS * 128 * 3 * H;
    __nv_bfloat16 *d_f, *d_b, *d_Yh, *d_Yl, *d_gih, *d_gil, *d_gnh, *d_gnl;
    float *d_G, *d_YB, *d_hn, *d_db, *d_dh0; unsigned int* dbg;
    CK(cudaMalloc(&d_f, img * 2)); CK(cudaMalloc(&d_b, img * 2));
    CK(cudaMalloc(&d_Yh, (size_t)R * D * H * 2)); CK(cudaMalloc(&d_Yl, (size_t)R * D * H * 2));
    CK(cudaMalloc(&d_gih, (size_t)R * D * 3 * H * 2)); CK(cudaMalloc(&d_gil, (size_t)R * D * 3 * H * 2));
    CK(cudaMalloc(&d_gnh, (size_t)R * D * H * 2)); CK(cudaMalloc(&d_gnl, (size_t)R * D * H * 2));
    CK(cudaMalloc(&d_G, (size_t)R * D * 4 * H * 4)); CK(cudaMalloc(&d_YB, (size_t)R * D * H * 4)); CK(cudaMalloc(&d_hn, (size_t)D * B * H * 4));
    CK(cudaMalloc(&d_db, (size_t)2 * D * 3 * H * 4)); CK(cudaMalloc(&d_dh0, (size_t)D * B * H * 4)); CK(cudaMalloc(&dbg, 64));
    CK(cudaMemset(dbg, 0, 64)); CK(cudaMemset(d_db, 0, (size_t)2 * D * 3 * H * 4));
    for (int d = 0; d < D; ++d) {
        tcx::pack_whh_images_kernel<<<256, 256>>>(d_whh + (size_t)d * 3 * H * H, d_f + (size_t)d * CS * 128 * 3 * H, d_b + (size_t)d * CS * 128 * 3 * H, H);
        CK(cudaGetLastError());
    }
    tcx::FwdParams p{};
    p.B = B; p.T = T; p.H = H; p.D = D; p.Wimg = d_f; p.giX = d_gi; p.b_hn = d_bhn; p.h0 = use_h0 ? d_h0 : nullptr;
    std::vector<float> gh0((size_t)D * B * 3 * H, 0.f);
    if (use_h0)
        for (int d = 0; d < D; ++d) for (int b = 0; b < B; ++b) for (int q = 0; q < 3 * H; ++q) {
            double a = 0; for (int k = 0; k < H; ++k) a += (double)whh[((size_t)d * 3 * H + q) * H + k] * h0[((size_t)d * B + b) * H + k];
            gh0[((size_t)d * B + b) * 3 * H + q] = (float)a;
        }
    float* d_gh0 = dev(gh0);
    p.gh0 = use_h0 ? d_gh0 : nullptr;
    p.GX = d_G; p.YBX = d_YB; p.hn_out = d_hn; p.Yhi = d_Yh; p.Ylo = d_Yl; p.dbg = dbg;
#ifdef BIGRU_SCAN_TIMING
    unsigned long long* d_ts; CK(cudaMalloc(&d_ts, 8 * 16 * 8)); CK(cudaMemset(d_ts, 0, 8 * 16 * 8)); p.ts = d_ts;
#endif
    CK(tcx::launch_fwd(p, 0));
    CK(cudaDeviceSynchronize());
    float ms = 0;
    if (reps > 0) {
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        for (int i = 0; i < reps; ++i) CK(tcx::launch_fwd(p, 0));
        cudaEventRecord(e1); CK(cudaDeviceSynchronize());
        cudaEventElapsedTime(&ms, e0, e1); ms /= reps;
    }
    unsigned int hdbg[8]; CK(cudaMemcpy(hdbg, dbg, 32, cudaMemcpyDeviceToHost));
    std::vector<__nv_bfloat16> Yh((size_t)R * D * H), Yl((size_t)R * D * H);
    std::vector<float> G((size_t)R * D * 4 * H), hn((size_t)D * B * H);
    CK(cudaMemcpy(Yh.data(), d_Yh, Yh.size() * 2, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(Yl.data(), d_Yl, Yl.size() * 2, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(G.data(), d_G, G.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hn.data(), d_hn, hn.size() * 4, cudaMemcpyDeviceToHost));
    // ---- CPU forward in double, full stash for the BPTT below (first rows of the batch only: they cover several tiles)
    // rows checked: the first 40 (several tiles) and, with SCANX_ALLTILES, one row of every 5 across the whole batch
    std::vector<int> rows_chk;
    for (int b = 0; b < (B > 40 ? 40 : B); ++b) rows_chk.push_back(b);
    if (getenv("SCANX_ALLTILES")) for (int b = 40; b < B; b += 5) rows_chk.push_back(b);
    const int bcheck = (int)rows_chk.size();
    double eY = 0, eG = 0, eHn = 0;
    int nbad = 0; std::vector<int> bad_b(B, 0), bad_s(T, 0);
    // stash[d][b][s] -> r, z, n, hn, hprev per unit
    std::vector<double> sr((size_t)D * bcheck * T * H), sz(sr.size()), sn(sr.size()), shn(sr.size()), shp(sr.size());
    for (int d = 0; d < D; ++d)
        for (int bi = 0; bi < bcheck; ++bi) {
            const int b = rows_chk[bi];
            std::vector<double> hs(H, 0.0), hnew(H);
            if (use_h0) for (int j = 0; j < H; ++j) hs[j] = h0[((size_t)d * B + b) * H + j];
            for (int s = 0; s < T; ++s) {
                const int t = d == 0 ? s : T - 1 - s;
                const long row = (long)t * B + b;
                for (int j = 0; j < H; ++j) {
                    double a[3] = {0, 0, 0};
                    for (int g = 0; g < 3; ++g) {
                        const float* w = &whh[((size_t)d * 3 * H + g * H + j) * H];
                        double acc = 0;
                        for (int k = 0; k < H; ++k) acc += (double)w[k] * hs[k];
                        a[g] = acc;
                    }
                    const float* gp = &gi[row * D * 3 * H + d * 3 * H];
                    const double r = 1.0 / (1.0 + exp(-(gp[j] + a[0])));
                    const double z = 1.0 / (1.0 + exp(-(gp[H + j] + a[1])));
                    const double hnv = a[2] + bhn[d * H + j];
                    const double n = tanh(gp[2 * H + j] + r * hnv);
                    hnew[j] = n + z * (hs[j] - n);
                    const size_t si = (((size_t)d * bcheck + bi) * T + t) * H + j;
                    sr[si] = r; sz[si] = z; sn[si] = n; shn[si] = hnv; shp[si] = hs[j];
                    eG = fmax(eG, fabs(r - G[bidx(d, b, t, j, 0, 4, B, T, H)]));
                    eG = fmax(eG, fabs(z - G[bidx(d, b, t, j, 1, 4, B, T, H)]));
                    eG = fmax(eG, fabs(n - G[bidx(d, b, t, j, 2, 4, B, T, H)]));
                    eG = fmax(eG, fabs(hnv - G[bidx(d, b, t, j, 3, 4, B, T, H)]));
                }
                for (int j = 0; j < H; ++j) {
                    hs[j] = hnew[j];
                    const size_t yi = (size_t)row * D * H + d * H + j;
                    const double ey = fabs(hnew[j] - ((double)b2f(Yh[yi]) + (double)b2f(Yl[yi])));
                    if (ey > 1e-3 && nbad < 12) { ++nbad; printf("   bad Y d=%d b=%d t=%d (s=%d) unit=%d: got %.6f want %.6f\n", d, b, t, s, j, (double)b2f(Yh[yi]) + (double)b2f(Yl[yi]), hnew[j]); }
                    if (ey > 1e-3) { bad_b[b]++; bad_s[s]++; }
                    eY = fmax(eY, ey);
                }
            }
            for (int j = 0; j < H; ++j) eHn = fmax(eHn, fabs(hs[j] - hn[((size_t)d * B + b) * H + j]));
        }
    if (nbad) { printf("   bad per batch row:"); for (int b = 0; b < B; ++b) if (bad_b[b]) printf(" %d:%d", b, bad_b[b]); printf("\n   bad per step:"); for (int t = 0; t < T; ++t) printf(" %d", bad_s[t]); printf("\n"); }
    const bool fpass = hdbg[0] == 0 && eY < 2e-5 && eG < 2e-5 && eHn < 2e-5;
    printf("%s scanx_fwd B=%d T=%d H=%d D=%d h0=%d (cluster %d, grid %d): errY=%.2e errG=%.2e errHn=%.2e dbg=%x blk=%u thr=%u  %.3f ms (%.2f us/step)\n",
           fpass ? "PASS" : "FAIL", B, T, H, D, use_h0, CS, D * (B / 32) * CS, eY, eG, eHn, hdbg[0], hdbg[1], hdbg[2], ms, ms * 1e3 / T);
#ifdef BIGRU_SCAN_TIMING
    if (T >= 80) {
        unsigned long long ts[8 * 16]; CK(cudaMemcpy(ts, d_ts, sizeof(ts), cudaMemcpyDeviceToHost));
        printf("  fwd cycles: c: wake->own issued | ->all issued || e: issue-end->mma_done | tmem ld | exchange | math+sts | st.async+arrive | arrive->c.wake | total\n");
        for (int k = 1; k < 5; ++k) {
            const unsigned long long* a = ts + k * 16; const unsigned long long* nx = ts + (k + 1) * 16;
            printf("   s=%d | %5lld | %5lld || %5lld | %5lld | %5lld | %5lld | %5lld | %5lld | %5lld\n", 64 + k, (long long)(a[1] - a[0]), (long long)(a[3] - a[1]),
                   (long long)(a[7] - a[3]), (long long)(a[8] - a[7]), (long long)(a[9] - a[8]), (long long)(a[10] - a[9]), (long long)(a[11] - a[10]),
                   (long long)(nx[0] - a[11]), (long long)(nx[0] - a[0]));
        }
        CK(cudaMemset(d_ts, 0, 8 * 16 * 8));
    }
#endif
    // ---- backward on the GPU stash
    tcx::BwdParams q{};
    q.B = B; q.T = T; q.H = H; q.D = D; q.WTimg = d_b; q.GX = d_G; q.YBX = d_YB; q.dYBX = d_dYB; q.h0 = use_h0 ? d_h0 : nullptr;
    if (top) { q.dlogits = d_dlog; q.lin_w = d_linw; q.arg = d_arg; q.C = C; }
    q.dgi_hi = d_gih; q.dgi_lo = d_gil; q.dghn_hi = d_gnh; q.dghn_lo = d_gnl;
    q.db_ih = d_db; q.db_hh = d_db + (size_t)D * 3 * H; q.dir_stride = 3 * H; q.dh0 = use_h0 ? d_dh0 : nullptr; q.dbg = dbg;
#ifdef BIGRU_SCAN_TIMING
    q.ts = d_ts;
#endif
    CK(tcx::launch_bwd(q, 0));
    CK(cudaDeviceSynchronize());
    float msb = 0;
    if (reps > 0) {
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        for (int i = 0; i < reps; ++i) CK(tcx::launch_bwd(q, 0));
        cudaEventRecord(e1); CK(cudaDeviceSynchronize());
        cudaEventElapsedTime(&msb, e0, e1); msb /= reps;
    }
    CK(cudaMemcpy(hdbg, dbg, 32, cudaMemcpyDeviceToHost));
    std::vector<__nv_bfloat16> gih((size_t)R * D * 3 * H), gil(gih.size()), gnh((size_t)R * D * H), gnl(gnh.size());
    std::vector<float> dh0((size_t)D * B * H);
    CK(cudaMemcpy(gih.data(), d_gih, gih.size() * 2, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(gil.data(), d_gil, gil.size() * 2, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(gnh.data(), d_gnh, gnh.size() * 2, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(gnl.data(), d_gnl, gnl.size() * 2, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(dh0.data(), d_dh0, dh0.size() * 4, cudaMemcpyDeviceToHost));
    double eD = 0, eH0 = 0, mD = 0;
    for (int d = 0; d < D; ++d)
        for (int bi = 0; bi < bcheck; ++bi) {
            const int b = rows_chk[bi];
            std::vector<double> carry(H, 0.0), rec(H, 0.0), dgh(3 * H);
            if (top)
                for (int j = 0; j < H; ++j) { double dl = 0; for (int c = 0; c < C; ++c) dl += (double)dlog[(size_t)b * C + c] * linw[(size_t)c * 3 * H + j]; carry[j] = dl; }
            for (int s = 0; s < T; ++s) {
                const int t = d == 0 ? T - 1 - s : s;
                const long row = (long)t * B + b;
                for (int j = 0; j < H; ++j) {
                    const size_t si = (((size_t)d * bcheck + bi) * T + t) * H + j;
                    double dy;
                    if (top) {
                        double dm = 0, da = 0;
                        for (int c = 0; c < C; ++c) { dm += (double)dlog[(size_t)b * C + c] * linw[(size_t)c * 3 * H + H + j]; da += (double)dlog[(size_t)b * C + c] * linw[(size_t)c * 3 * H + 2 * H + j]; }
                        dy = da / T + (arg[(size_t)b * H + j] == t ? dm : 0.0);
                    } else dy = dY[(size_t)row * D * H + d * H + j];
                    const double dh = carry[j] + rec[j] + dy;
                    const double r = sr[si], z = sz[si], n = sn[si];
                    const double dan = dh * (1 - z) * (1 - n * n), dar = dan * shn[si] * r * (1 - r), daz = dh * (shp[si] - n) * z * (1 - z);
                    dgh[j] = dar; dgh[H + j] = daz; dgh[2 * H + j] = dan * r;
                    carry[j] = dh * z;
                    const size_t gi_i = (size_t)row * D * 3 * H + d * 3 * H + j;
                    const size_t gn_i = (size_t)row * D * H + d * H + j;
                    const double g0 = (double)b2f(gih[gi_i]) + b2f(gil[gi_i]), g1 = (double)b2f(gih[gi_i + H]) + b2f(gil[gi_i + H]);
                    const double g2 = (double)b2f(gih[gi_i + 2 * H]) + b2f(gil[gi_i + 2 * H]), g3 = (double)b2f(gnh[gn_i]) + b2f(gnl[gn_i]);
                    eD = fmax(eD, fabs(g0 - dar)); eD = fmax(eD, fabs(g1 - daz)); eD = fmax(eD, fabs(g2 - dan)); eD = fmax(eD, fabs(g3 - dan * r));
                    mD = fmax(mD, fabs(dar)); mD = fmax(mD, fabs(daz)); mD = fmax(mD, fabs(dan));
                }
                for (int k = 0; k < H; ++k) {
                    double acc = 0;
                    for (int qq = 0; qq < 3 * H; ++qq) acc += (double)whh[((size_t)d * 3 * H + qq) * H + k] * dgh[qq];
                    rec[k] = acc;
                }
            }
            if (use_h0) for (int j = 0; j < H; ++j) eH0 = fmax(eH0, fabs(carry[j] + rec[j] - dh0[((size_t)d * B + b) * H + j]));
        }
    const bool bpass = hdbg[0] == 0 && eD < 2e-5 * fmax(mD, 1e-6) + 1e-9 && eH0 < 2e-5 * fmax(mD, 1e-6) * 4 + 1e-9;
    printf("%s scanx_bwd B=%d T=%d H=%d D=%d h0=%d top=%d: err(dgi,dghn)=%.2e (max |dg| %.2e) err(dh0)=%.2e dbg=%x blk=%u thr=%u  %.3f ms (%.2f us/step)\n",
           bpass ? "PASS" : "FAIL", B, T, H, D, use_h0, top, eD, mD, eH0, hdbg[0], hdbg[1], hdbg[2], msb, msb * 1e3 / T);
#ifdef BIGRU_SCAN_TIMING
    if (T >= 80) {
        unsigned long long ts[8 * 16]; CK(cudaMemcpy(ts, d_ts, sizeof(ts), cudaMemcpyDeviceToHost));
        printf("  bwd cycles: c: wake->A issued | ->B issued || e: issue-end->mma_b seen | route+recv | math+sts | arrive->c.wake | total\n");
        for (int k = 1; k < 5; ++k) {
            const unsigned long long* a = ts + k * 16; const unsigned long long* nx = ts + (k + 1) * 16;
            printf("   s=%d | %5lld | %5lld || %5lld | %5lld | %5lld | %5lld | %5lld\n", 64 + k, (long long)(a[1] - a[0]), (long long)(a[3] - a[1]),
                   (long long)(a[7] - a[3]), (long long)(a[8] - a[7]), (long long)(a[9] - a[8]), (long long)(nx[0] - a[10]), (long long)(nx[0] - a[0]));
        }
    }
#endif
    cudaFree(d_whh); cudaFree(d_bhn); cudaFree(d_gi); cudaFree(d_h0); cudaFree(d_dYB); cudaFree(d_dlog); cudaFree(d_linw); cudaFree(d_arg);
    cudaFree(d_f); cudaFree(d_b); cudaFree(d_Yh); cudaFree(d_Yl); cudaFree(d_gih); cudaFree(d_gil); cudaFree(d_gnh); cudaFree(d_gnl);
    cudaFree(d_G); cudaFree(d_YB); cudaFree(d_hn); cudaFree(d_db); cudaFree(d_dh0); cudaFree(dbg);
    return (fpass && bpass) ? 0 : 2;
}

int main() {
    setvbuf(stdout, NULL, _IONBF, 0);
    int bad = 0;
    if (getenv("SCANX_BIG")) {
        const int n = atoi(getenv("SCANX_BIG"));
        for (int i = 0; i < n; ++i) bad += run_case(512, 128, 256, 2, 0, 0, 1);
        printf("big: %d failures of %d\n", bad / 2, n);
        return bad ? 1 : 0;
    }
    if (getenv("SCANX_STRESS")) {
        const int h0 = atoi(getenv("SCANX_STRESS"));
        for (int i = 0; i < 25; ++i) bad += run_case(64, 11, 256, 2, 0, h0, 1);
        for (int i = 0; i < 10; ++i) bad += run_case(128, 6, 256, 1, 0, h0, 0);
        for (int i = 0; i < 10; ++i) bad += run_case(64, 7, 128, 2, 0, h0, 0);
        printf("stress h0=%d: %d failures\n", h0, bad / 2);
        return bad ? 1 : 0;
    }
    if (getenv("SCANX_H0")) {
        for (int i = 0; i < 4; ++i) bad += run_case(64, 4, 256, 1, 0, 1, 0);
        for (int i = 0; i < 2; ++i) bad += run_case(64, 11, 256, 2, 0, 1, 1);
        for (int i = 0; i < 2; ++i) bad += run_case(128, 6, 256, 1, 0, 1, 0);
        for (int i = 0; i < 2; ++i) bad += run_case(32, 4, 256, 1, 0, 1, 0);
        printf(bad ? "SOME FAILED\n" : "ALL PASSED\n");
        return bad ? 1 : 0;
    }
    bad += run_case(32, 1, 128, 1, 0, 0, 0);
    bad += run_case(32, 3, 128, 1, 0, 0, 0);
    bad += run_case(64, 5, 128, 2, 0, 1, 1);
    bad += run_case(32, 2, 256, 1, 0, 0, 0);
    bad += run_case(32, 4, 256, 1, 0, 1, 0);
    bad += run_case(64, 9, 256, 2, 0, 0, 1);
    bad += run_case(512, 128, 256, 2, 10, 0, 0);
    bad += run_case(512, 128, 256, 2, 10, 0, 1);
    bad += run_case(512, 64, 128, 2, 10, 0, 1);
    printf(bad ? "SOME FAILED\n" : "ALL PASSED\n");
    return bad ? 1 : 0;
}
