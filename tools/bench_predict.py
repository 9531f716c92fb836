"""Live-window latency: predict.py's per-message model work on the shipped checkpoint's shape
(window 5 x 108 features, hidden 8, 1 layer, bidirectional, 4 labels).  python tools/bench_predict.py"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from financial_market_data_analysis_b200.predict import LivePredictor   # noqa: E402

z = np.load(os.path.join(ROOT, "tests", "golden", "kat.npz"))
state = {k[2:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("p:")}
mn, mx = z["norm_min"].astype(np.float32), z["norm_max"].astype(np.float32)
lp = LivePredictor(state, (mn, mx))
raw = (mn + np.random.default_rng(0).uniform(0, 1, (5, mn.size)) * (mx - mn)).astype(np.float32)
xd = torch.from_numpy(raw).cuda()[None]


def timed(fn, n=200):
    for _ in range(20):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


xn = ((xd - lp.x_min) / (lp.x_max - lp.x_min)).contiguous()
fused_us = timed(lambda: lp.forward_windows(xd))
steps_us = timed(lambda: torch.sigmoid(lp.model((xd - lp.x_min) / (lp.x_max - lp.x_min))))
t0 = time.perf_counter()
for _ in range(200):
    lp.predict(raw)
e2e_us = (time.perf_counter() - t0) / 200 * 1e6
# the reference's own arithmetic on the host cores (torch.nn.GRU CPU): oracle = checker / CPU baseline only
from oracle import bigru_oracle as bo                                    # noqa: E402
ref = bo.OracleBiGRU(8, mn.size, 4, 1, 50, 0.2, False, True)
ref.load_state_dict(state)
ref.eval()
xc = torch.from_numpy((raw - mn) / (mx - mn))[None]
with torch.no_grad():
    for _ in range(20):
        ref(xc)
    t0 = time.perf_counter()
    for _ in range(200):
        torch.sigmoid(ref(xc))
    cpu_us = (time.perf_counter() - t0) / 200 * 1e6
print(json.dumps({"workload": "one live window 5 x 108, hidden 8, 1 layer bidirectional, 4 labels (shipped checkpoint)",
                  "gpu_one_launch_us": fused_us, "gpu_training_path_launches_us": steps_us, "predict_call_wall_us_incl_d2h": e2e_us,
                  "cpu_reference_math_us": cpu_us}))
