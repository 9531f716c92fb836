"""One training step of the two cluster-scan paths for an `ncu --set full` capture (GPU box):
   bf16x3 at BASELINE configs[1] (tcx::gru_scanx_*), bf16 at hidden 512 (tcw::gru_scanw_*, configs[4] with T = 256).
   BIGRU_B200_CUDA_GRAPH=0 ncu --set full --clock-control none --import-source on -k regex:'scanx|scanw' -o out python tools/ncu_step.py"""
import os, sys
import torch, torch.nn as nn
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import financial_market_data_analysis_b200 as pkg

for prec, (B, T, F, H) in (("bf16x3", (512, 128, 64, 256)), ("bf16", (256, 256, 128, 512))):
    torch.manual_seed(0)
    m = pkg.BiGRU(H, F, 3, 2, 50, 0.0, False, True, precision=prec).cuda().train()
    m.add_loss_fn(nn.CrossEntropyLoss()); m.add_optimizer(torch.optim.Adam(m.parameters(), lr=1e-3))
    m.use_cuda_graph = False
    x = torch.randn(B, T, F, device="cuda"); y = torch.randint(0, 3, (B,), device="cuda")
    for _ in range(2):
        loss, _ = m.train_step(x, y)
    torch.cuda.synchronize()
    print(prec, float(loss))
