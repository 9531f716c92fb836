"""TEST INFRASTRUCTURE ONLY - CPU restatement of the reference's SQL window-function features.

Follows /root/reference/create_database.py:76-190 (the CREATE VIEW statements) and :239-256 (column order of the join):
  vol_MA / price_MA / delta_MA  :76-118   AVG(col) OVER (ORDER BY Timestamp ROWS BETWEEN p-1 PRECEDING AND CURRENT ROW)
  bollinger_bands               :120-135  (avg + k*std) - close, close - (avg - k*std); MariaDB STD = population std
  stochastic_oscillator         :137-147  (close - min) / (max - min) over ROWS BETWEEN 14 PRECEDING AND CURRENT ROW; x/0 = NULL
  price_change                  :150-154  close - LAG(close, 1): NULL on the first row
  ATR                           :156-161  AVG(high - low) over ROWS BETWEEN 14 PRECEDING AND CURRENT ROW
  target                        :163-185  LEAD(close, 8 / 15) vs close +- n1 / n2 * ATR; a comparison with NULL is not true -> 0
Frames at the head of the table are shorter (SQL window frames clip at the partition start).  NULL is NaN.

PINNED to the reference's own SQL: tests/golden/make_features_golden.py imports the UNMODIFIED create_database.py with a
`mysql.connector` stub that forwards its statements to sqlite3 (window functions; MariaDB's STD registered as a window
aggregate), fills the table with a seed-fixed synthetic market and stores what the reference's views return in
tests/golden/features.npz; tests/test_oracle_cpu.py checks this restatement against it (exact) and, independently,
against pandas.rolling and hand-computed rows."""
from __future__ import annotations

import numpy as np


def _frames(n, w):
    i = np.arange(n)
    return np.maximum(0, i - w + 1), i


def rolling_mean(x, w):
    x = np.asarray(x, dtype=np.float64)
    cs = np.concatenate([[0.0], np.cumsum(x)])
    lo, hi = _frames(len(x), w)
    return (cs[hi + 1] - cs[lo]) / (hi - lo + 1)


def rolling_std_pop(x, w):
    x = np.asarray(x, dtype=np.float64)
    out = np.empty(len(x))
    for i in range(len(x)):                      # two-pass per frame: no cancellation
        f = x[max(0, i - w + 1): i + 1]
        out[i] = np.sqrt(np.mean((f - f.mean()) ** 2))
    return out


def rolling_minmax(x, w):
    x = np.asarray(x, dtype=np.float64)
    mn, mx = np.empty(len(x)), np.empty(len(x))
    for i in range(len(x)):
        f = x[max(0, i - w + 1): i + 1]
        mn[i], mx[i] = f.min(), f.max()
    return mn, mx


def window_features(close, high, low, volume=None, delta=None, volume_MA_periods=(6, 20), price_MA_periods=(20,),
                    delta_MA_periods=(12,), bollinger_bands_period=20, bollinger_bands_std=2, stochastic_oscillator=True,
                    n1=1.5, n2=3.0):
    """Returns (features [n, n_out] float64, targets [n, 4] float64), columns as create_database.py:239-240."""
    close = np.asarray(close, dtype=np.float64)
    n = len(close)
    cols = []
    if bollinger_bands_period and bollinger_bands_std:
        avg, sd = rolling_mean(close, bollinger_bands_period), rolling_std_pop(close, bollinger_bands_period)
        cols += [(avg + bollinger_bands_std * sd) - close, close - (avg - bollinger_bands_std * sd)]
    cols += [rolling_mean(volume, p) for p in (volume_MA_periods or [])]
    cols += [rolling_mean(close, p) for p in (price_MA_periods or [])]
    cols += [rolling_mean(delta, p) for p in (delta_MA_periods or [])]
    if stochastic_oscillator:
        mn, mx = rolling_minmax(close, 15)
        with np.errstate(divide="ignore", invalid="ignore"):
            cols.append(np.where(mx > mn, (close - mn) / (mx - mn), np.nan))
    atr = rolling_mean(np.asarray(high, dtype=np.float64) - np.asarray(low, dtype=np.float64), 15)
    cols.append(atr)
    pc = np.full(n, np.nan)
    if n > 1:
        pc[1:] = close[1:] - close[:-1]
    cols.append(pc)
    feats = np.stack(cols, axis=1) if n else np.zeros((0, len(cols)))
    tgt = np.zeros((n, 4))
    for i in range(n):
        if i + 8 < n:
            tgt[i, 0] = close[i + 8] >= close[i] + n1 * atr[i]
            tgt[i, 2] = close[i + 8] <= close[i] - n1 * atr[i]
        if i + 15 < n:
            tgt[i, 1] = close[i + 15] >= close[i] + n2 * atr[i]
            tgt[i, 3] = close[i + 15] <= close[i] - n2 * atr[i]
    return feats, tgt
