// features.cuh - SURVEY.md 8(f) N4: the SQL window-function features of the reference (create_database.py:76-190) as one
// row-parallel kernel over the joined table's columns.  Every output row i depends on rows [i - w + 1, i] (moving
// averages, Bollinger bands, stochastic oscillator, ATR), on row i - 1 (price change) or on rows i + 8 / i + 15 (targets):
// block = 256 consecutive rows staged in shared memory with their halo, thread = row.  The SQL server evaluates AVG / STD
// of FLOAT columns in double; here the frame sums run in fp32 over DIFFERENCES to the current row (exact to ~1e-7 of the
// spread) and the targets compare exact fp32 price differences: no FP64 anywhere.  SQL NULL is NaN.  HBM-bound: 4 * (5 + n_out + 4) bytes per row.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <math.h>

struct FeatureCfg {
    int n_vol, n_price, n_delta;
    int vol_p[8], price_p[8], delta_p[8];      // AVG(col) OVER (ROWS BETWEEN p-1 PRECEDING AND CURRENT ROW)
    int bb_period; float bb_std;               // 0: no Bollinger columns
    int stochastic;                            // ROWS BETWEEN 14 PRECEDING AND CURRENT ROW (15 rows)
    float n1, n2;                              // ATR factors of the targets
    int n_out;
};

constexpr int FEAT_TR = 256;            // table rows per block (= threads)

// mean of the staged column c over the frame [jlo, j], as c[j] + mean(c[k] - c[j]): the differences are small, so fp32
// sums keep ~1e-7 of the spread (not of the level).  (Non-tensor FP64 is a 1/64-rate pipe on this part: a double
// version of this kernel - and one with double prefix sums - measured 1.10 / 1.21 ms against 0.8 for fp32.)
__device__ __forceinline__ float frame_mean(const float* __restrict__ c, int jlo, int j) {
    const float ref = c[j];
    float s = 0.f;
#pragma unroll 4
    for (int k = jlo; k < j; ++k) s += c[k] - ref;
    return ref + s / (float)(j - jlo + 1);
}

// One block = FEAT_TR consecutive table rows: the columns of rows [r0 - halo, r0 + TR + 15) are staged in shared memory
// with coalesced loads, every thread forms the frames of its row from there, and the outputs leave through shared
// memory as coalesced rows.
// FAST: the periods are those of the reference's config.py:40-49 (vol 6 / 20, price 20, delta 12, Bollinger 20, stochastic
// on) as compile-time constants - the frames of every row past the head of the table are then straight-line code (the
// generic loops are instruction-issue bound: ~1200 instructions per row against ~350).
template <bool FAST>
__global__ void __launch_bounds__(FEAT_TR) window_features_kernel(
        const float* __restrict__ close, const float* __restrict__ high, const float* __restrict__ low, const float* __restrict__ volume,
        const float* __restrict__ delta, int64_t n, FeatureCfg cfg, int halo, float* __restrict__ out, float* __restrict__ targets) {
    extern __shared__ float fsm[];
    const int span = FEAT_TR + halo;
    float* sc = fsm;                          // [span + 16] close, incl. the 15 rows the targets look ahead
    float* sh = sc + span + 16;               // [span] high - low
    float* sv = sh + span;                    // [span] volume
    float* sd = sv + span;                    // [span] delta
    float* so = sd + span;                    // [TR][n_out]
    float* st = so + FEAT_TR * cfg.n_out;     // [TR][4]
    const int tid = threadIdx.x;
    const int64_t ntiles = (n + FEAT_TR - 1) / FEAT_TR;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t r0 = tile * FEAT_TR, base = r0 - halo;      // staged index j <-> table row base + j
        for (int j = tid; j < span + 16; j += FEAT_TR) {
            const int64_t row = base + j;
            const bool ok = row >= 0 && row < n;
            sc[j] = ok ? close[row] : 0.f;
            if (j < span) {
                sh[j] = ok ? high[row] - low[row] : 0.f;
                sv[j] = (ok && volume) ? volume[row] : 0.f;
                sd[j] = (ok && delta) ? delta[row] : 0.f;
            }
        }
        __syncthreads();
        const int64_t i = r0 + tid;
        if (i < n) {
            const int j = halo + tid;
            float* o = so + tid * cfg.n_out;
            int c = 0;
            const float pcf = sc[j];
            // frame [max(0, i - w + 1), i] -> first staged index
            auto first = [&](int w) { const int64_t lo = i - w + 1 < 0 ? 0 : i - w + 1; return j - (int)(i - lo); };
            float atrf;
            if (FAST && i >= 19) {
                // ---- straight-line frames (all 20 / 15 / 12 / 6 rows exist)
                float s1 = 0.f;
#pragma unroll
                for (int k = 1; k < 20; ++k) s1 += sc[j - k] - pcf;
                const float md = s1 * (1.f / 20.f);
                float v = md * md;
#pragma unroll
                for (int k = 1; k < 20; ++k) { const float d = (sc[j - k] - pcf) - md; v = fmaf(d, d, v); }
                const float sdv = sqrtf(v * (1.f / 20.f));
                o[c++] = md + cfg.bb_std * sdv;
                o[c++] = cfg.bb_std * sdv - md;
                const float vref = sv[j];
                float v6 = 0.f;
#pragma unroll
                for (int k = 1; k < 6; ++k) v6 += sv[j - k] - vref;
                float v20 = v6;
#pragma unroll
                for (int k = 6; k < 20; ++k) v20 += sv[j - k] - vref;
                o[c++] = vref + v6 * (1.f / 6.f);
                o[c++] = vref + v20 * (1.f / 20.f);
                o[c++] = pcf + md;                                                                   // price_MA20 = BB_avg
                const float dref = sd[j];
                float d12 = 0.f;
#pragma unroll
                for (int k = 1; k < 12; ++k) d12 += sd[j - k] - dref;
                o[c++] = dref + d12 * (1.f / 12.f);
                float mn = pcf, mx = pcf;
#pragma unroll
                for (int k = 1; k < 15; ++k) { mn = fminf(mn, sc[j - k]); mx = fmaxf(mx, sc[j - k]); }
                o[c++] = mx > mn ? (pcf - mn) / (mx - mn) : nanf("");
                float sa = 0.f;
#pragma unroll
                for (int k = 0; k < 15; ++k) sa += sh[j - k];
                atrf = sa * (1.f / 15.f);
                o[c++] = atrf;
            } else {
            float bb_mean = 0.f;
            if (cfg.bb_period > 0) {
                // (BB_avg + k * BB_std) - close, close - (BB_avg - k * BB_std); STD = population standard deviation.
                // With d_k = close[k] - close[i]: avg - close = mean(d), std = sqrt(mean((d - mean d)^2)) (two passes)
                const int jlo = first(cfg.bb_period);
                const float cnt = (float)(j - jlo + 1);
                float s1 = 0.f;
#pragma unroll 4
                for (int k = jlo; k < j; ++k) s1 += sc[k] - pcf;
                const float md = s1 / cnt;
                float v = md * md;                                 // the k == i term: (0 - md)^2
#pragma unroll 4
                for (int k = jlo; k < j; ++k) { const float d = (sc[k] - pcf) - md; v = fmaf(d, d, v); }
                const float sdv = sqrtf(v / cnt);
                o[c++] = md + cfg.bb_std * sdv;
                o[c++] = cfg.bb_std * sdv - md;
                bb_mean = pcf + md;
            }
            for (int q = 0; q < cfg.n_vol; ++q) o[c++] = frame_mean(sv, first(cfg.vol_p[q]), j);
            for (int q = 0; q < cfg.n_price; ++q)
                o[c++] = (cfg.price_p[q] == cfg.bb_period) ? bb_mean : frame_mean(sc, first(cfg.price_p[q]), j);
            for (int q = 0; q < cfg.n_delta; ++q) o[c++] = frame_mean(sd, first(cfg.delta_p[q]), j);
            const int j15 = first(15);
            if (cfg.stochastic) {
                float mn = sc[j15], mx = sc[j15];
#pragma unroll 4
                for (int k = j15 + 1; k <= j; ++k) { mn = fminf(mn, sc[k]); mx = fmaxf(mx, sc[k]); }
                o[c++] = mx > mn ? (pcf - mn) / (mx - mn) : nanf("");     // x / 0 is NULL in SQL; both differences are exact in fp32
            }
            float sa = 0.f;                                                                            // ATR = AVG(high - low), 15 rows
#pragma unroll 4
            for (int k = j15; k <= j; ++k) sa += sh[k];
            atrf = sa / (float)(j - j15 + 1);
            o[c++] = atrf;
            }
            o[c++] = i > 0 ? pcf - sc[j - 1] : nanf("");                                               // LAG(close, 1): NULL on the first row
            if (targets) {
                // LEAD(close, 8 / 15): NULL past the end, and a comparison with NULL is not true -> 0
                float* t = st + tid * 4;
                // (price differences are exact in fp32, so the comparison is made on them; no FP64 in this kernel)
                const bool h8 = i + 8 < n, h15 = i + 15 < n;
                const float d8 = sc[j + 8] - pcf, d15 = sc[j + 15] - pcf, a1 = cfg.n1 * atrf, a2 = cfg.n2 * atrf;
                t[0] = (h8 && d8 >= a1) ? 1.f : 0.f;
                t[1] = (h15 && d15 >= a2) ? 1.f : 0.f;
                t[2] = (h8 && d8 <= -a1) ? 1.f : 0.f;
                t[3] = (h15 && d15 <= -a2) ? 1.f : 0.f;
            }
        }
        __syncthreads();
        const int rows = (int)((n - r0) < FEAT_TR ? (n - r0) : FEAT_TR);
        for (int k = tid; k < rows * cfg.n_out; k += FEAT_TR) out[r0 * cfg.n_out + k] = so[k];
        if (targets)
            for (int k = tid; k < rows * 4; k += FEAT_TR) targets[r0 * 4 + k] = st[k];
        __syncthreads();
    }
}
