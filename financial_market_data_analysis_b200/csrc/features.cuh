// features.cuh - SURVEY.md 8(f) N4: the SQL window-function features of the reference (create_database.py:76-190) as one
// row-parallel kernel over the joined table's columns.  Every output row i depends on rows [i - w + 1, i] (moving
// averages, Bollinger bands, stochastic oscillator, ATR), on row i - 1 (price change) or on rows i + 8 / i + 15 (targets):
// thread = row, the window is re-read from L1/L2 (w <= a few hundred rows), arithmetic in double like the SQL server's
// AVG / STD over FLOAT columns.  SQL NULL is NaN.  HBM-bound: 4 * (5 + n_out + 4) bytes per row.
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

__device__ __forceinline__ double win_mean(const float* __restrict__ c, int64_t i, int w) {
    const int64_t lo = i - w + 1 < 0 ? 0 : i - w + 1;
    double s = 0.0;
    for (int64_t k = lo; k <= i; ++k) s += (double)c[k];
    return s / (double)(i - lo + 1);
}

__global__ void window_features_kernel(const float* __restrict__ close, const float* __restrict__ high, const float* __restrict__ low,
                                       const float* __restrict__ volume, const float* __restrict__ delta, int64_t n, FeatureCfg cfg,
                                       float* __restrict__ out, float* __restrict__ targets) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float* o = out + i * cfg.n_out;
        int c = 0;
        const double pc = (double)close[i];
        if (cfg.bb_period > 0) {
            // (BB_avg + k * BB_std) - close, close - (BB_avg - k * BB_std); STD = population standard deviation
            const int64_t lo = i - cfg.bb_period + 1 < 0 ? 0 : i - cfg.bb_period + 1;
            const double m = win_mean(close, i, cfg.bb_period);
            double v = 0.0;
            for (int64_t k = lo; k <= i; ++k) { const double d = (double)close[k] - m; v += d * d; }
            const double sd = sqrt(v / (double)(i - lo + 1));
            o[c++] = (float)((m + (double)cfg.bb_std * sd) - pc);
            o[c++] = (float)(pc - (m - (double)cfg.bb_std * sd));
        }
        for (int j = 0; j < cfg.n_vol; ++j) o[c++] = (float)win_mean(volume, i, cfg.vol_p[j]);
        for (int j = 0; j < cfg.n_price; ++j) o[c++] = (float)win_mean(close, i, cfg.price_p[j]);
        for (int j = 0; j < cfg.n_delta; ++j) o[c++] = (float)win_mean(delta, i, cfg.delta_p[j]);
        if (cfg.stochastic) {
            const int64_t lo = i - 14 < 0 ? 0 : i - 14;
            float mn = close[lo], mx = close[lo];
            for (int64_t k = lo + 1; k <= i; ++k) { mn = fminf(mn, close[k]); mx = fmaxf(mx, close[k]); }
            o[c++] = mx > mn ? (float)((pc - (double)mn) / ((double)mx - (double)mn)) : nanf("");     // x / 0 is NULL in SQL
        }
        // ATR = AVG(high - low) over 15 rows (the subtraction is done in double, as the server evaluates the expression)
        double atr;
        {
            const int64_t lo = i - 14 < 0 ? 0 : i - 14;
            double s = 0.0;
            for (int64_t k = lo; k <= i; ++k) s += (double)high[k] - (double)low[k];
            atr = s / (double)(i - lo + 1);
        }
        o[c++] = (float)atr;
        o[c++] = i > 0 ? (float)(pc - (double)close[i - 1]) : nanf("");                               // LAG(close, 1): NULL on the first row
        if (targets) {
            // LEAD(close, 8 / 15): NULL past the end, and a comparison with NULL is not true -> 0
            float* t = targets + i * 4;
            const bool h8 = i + 8 < n, h15 = i + 15 < n;
            const double p8 = h8 ? (double)close[i + 8] : 0.0, p15 = h15 ? (double)close[i + 15] : 0.0;
            t[0] = (h8 && p8 >= pc + (double)cfg.n1 * atr) ? 1.f : 0.f;
            t[1] = (h15 && p15 >= pc + (double)cfg.n2 * atr) ? 1.f : 0.f;
            t[2] = (h8 && p8 <= pc - (double)cfg.n1 * atr) ? 1.f : 0.f;
            t[3] = (h15 && p15 <= pc - (double)cfg.n2 * atr) ? 1.f : 0.f;
        }
    }
}
