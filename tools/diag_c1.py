"""GPU diagnostic: C1-shape parity of a precision against a float64 torch reference, separating max-pool argmax flips."""
import os, sys
import numpy as np, torch, torch.nn as nn
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import financial_market_data_analysis_b200 as pkg
from oracle import bigru_oracle as bo

def rel_l2(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)

def ref_forward(ref, x, idx=None, hidden=None):
    B, T = x.shape[:2]; H = ref.hidden_size
    out, h_n = ref.gru(x, hidden)
    last = h_n.view(ref.n_layers, ref.n_directions, B, H)[-1].sum(0)
    s = out[..., :H] + out[..., H:] if ref.bidirectional else out
    mx = s.max(dim=1).values if idx is None else s.gather(1, idx.unsqueeze(1)).squeeze(1)
    return ref.linear(torch.cat([last, mx, s.sum(1) / float(T)], dim=1)), s

def main():
    cfgs = [(512, 128, 64, 256, 2, 3, False), (64, 11, 24, 256, 2, 4, True)]
    if os.environ.get('DIAG_H0'):
        cfgs = [(64, 11, 24, 256, 1, 4, True), (64, 11, 24, 256, 2, 4, False), (32, 11, 24, 256, 2, 4, True), (64, 4, 24, 256, 2, 4, True), (64, 11, 24, 128, 2, 4, True)]
    for (B, T, F, H, L, C, use_h0) in cfgs:
        torch.manual_seed(0)
        ref = bo.OracleBiGRU(H, F, C, L, 50, 0.0, False, True).double()
        g = torch.Generator().manual_seed(1234)
        x = torch.randn(B, T, F, generator=g)
        target = torch.randint(0, C, (B,), generator=g)
        h0 = torch.randn(L * 2, B, H, generator=g) * 0.5 if use_h0 else None
        for precision in sys.argv[1:] or ["bf16x3"]:
            m = pkg.BiGRU(H, F, C, L, 50, 0.0, False, True, precision=precision)
            m.load_state_dict({k: v.float() for k, v in ref.state_dict().items()})
            m = m.cuda(); m.train()
            xg = x.cuda().requires_grad_(True)
            hg = h0.cuda().requires_grad_(True) if use_h0 else None
            y = m(xg, hg)
            arg = m.pooled_argmax().cpu().long()
            nn.CrossEntropyLoss()(y, target.cuda()).backward()
            ref.zero_grad()
            xr = x.double().requires_grad_(True)
            hr = h0.double().requires_grad_(True) if use_h0 else None
            pred, s = ref_forward(ref, xr, None, hr)
            top2 = s.detach().topk(2, dim=1).values
            gap = (top2[:, 0] - top2[:, 1])
            ref_arg = s.detach().argmax(dim=1)
            flips = (ref_arg != arg)
            print(f"[{precision}] B{B} T{T} H{H} h0={use_h0}: logits rel {np.abs(y.detach().cpu().numpy()-pred.detach().numpy()).max()/np.abs(pred.detach().numpy()).max():.2e}; "
                  f"argmax flips {int(flips.sum())} of {flips.numel()}, max gap at a flip {float(gap[flips].max()) if flips.any() else 0:.2e}; pairs with gap<1e-5: {int((gap<1e-5).sum())}")
            for name, idx in (("ref routing", None), ("our routing", arg)):
                ref.zero_grad()
                xr = x.double().requires_grad_(True)
                hr = h0.double().requires_grad_(True) if use_h0 else None
                pred, _ = ref_forward(ref, xr, idx, hr)
                nn.CrossEntropyLoss()(pred, target).backward()
                errs = {k: rel_l2(p.grad.cpu().numpy(), q.grad.numpy()) for (k, p), (_, q) in zip(m.named_parameters(), ref.named_parameters())}
                worst = max(errs, key=errs.get)
                print(f"    grads vs fp64 reference ({name}): worst tensor {worst} {errs[worst]:.2e}; dx {rel_l2(xg.grad.cpu().numpy(), xr.grad.numpy()):.2e}"
                      + (f"; dh0 {rel_l2(hg.grad.cpu().numpy(), hr.grad.numpy()):.2e}" if use_h0 else ""))
                print("      " + " ".join(f"{k.split('.')[-1]}={v:.1e}" for k, v in errs.items()))

if __name__ == "__main__":
    main()
