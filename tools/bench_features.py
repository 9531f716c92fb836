"""Window-function features: GPU kernel time against the HBM roofline and the CPU oracle on a sample.
Run on the GPU box:  python tools/bench_features.py [rows]"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from financial_market_data_analysis_b200.features import window_features, feature_names   # noqa: E402
from oracle import features_oracle as fo                                                  # noqa: E402  (checker / CPU baseline only)

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16_000_000
g = torch.Generator(device="cuda").manual_seed(1)
close = 3000 + torch.cumsum(torch.randn(n, device="cuda", generator=g), 0)
sp = torch.rand(n, device="cuda", generator=g) + 0.5
cols = [close, close + sp, close - sp, torch.rand(n, device="cuda", generator=g) * 1e4, torch.randn(n, device="cuda", generator=g)]
for _ in range(3):
    f, t = window_features(*cols)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    f, t = window_features(*cols)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
n_out = len(feature_names())
bytes_alg = 4.0 * n * (5 + n_out + 4)
peak = 6569.6
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
m = 200_000
host = [c[:m].double().cpu().numpy() for c in cols]
t0 = time.perf_counter()
rf, rt = fo.window_features(*host)
cpu_s = time.perf_counter() - t0
err = float(np.nanmax(np.abs(f[:m - 20].cpu().numpy() - rf[:m - 20])))
print(json.dumps({"kernel": "window_features", "rows": n, "ms": ms, "rows_per_s": n / ms * 1e3,
                  "roofline": {"bound": "hbm", "achieved": bytes_alg / ms / 1e6, "peak": peak, "unit": "GB/s",
                               "frac": bytes_alg / ms / 1e6 / peak, "algorithmic_bytes_per_row": 4 * (5 + n_out + 4)},
                  "cpu_baseline": {"kind": "port", "cores": 1, "rows_per_s": m / cpu_s, "sample": f"{m} rows through oracle/features_oracle.py (numpy)"},
                  "max_abs_err_vs_oracle_on_sample": err}))
