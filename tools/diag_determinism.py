"""GPU diagnostic: run the x3 forward of the C1 shape several times and compare the stash byte for byte between runs
(the forward has no atomics: any difference is a race); reports which tensor of the layout first differs, and for the first
layer where (direction, time step, tile, warp, lane, column) the first differing values sit.

History: this found the ring-slot release race of the scan kernels (an mbarrier arrive issued right behind the shared-memory
loads of a prefetch slot overtook them; the slot was refilled before it had been read) - see DESIGN.md.
Usage (GPU box): python tools/diag_determinism.py [bf16x3|bf16|fp32] [runs]"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import financial_market_data_analysis_b200 as pkg

def al(x): return (x + 1023) & ~1023

def layout(B, T, F, H, L, D=2):
    R, DH, o, out = B * T, D * H, 0, []
    for l in range(L):
        I = F if l == 0 else DH
        for name, n in (("Yhi", R * DH * 2), ("Ylo", R * DH * 2), ("YB", R * DH * 4), ("G", R * D * 4 * H * 4), ("Xhi", R * I * 2), ("Xlo", R * I * 2),
                        ("Wih_hi", D * 3 * H * I * 2), ("Wih_lo", D * 3 * H * I * 2), ("WihT_hi", D * 3 * H * I * 2), ("WihT_lo", D * 3 * H * I * 2),
                        ("Wimg", D * 2 * 3 * H * H * 2), ("WTimg", D * 2 * 3 * H * H * 2), ("bfold", D * 3 * H * 4), ("bhn", D * H * 4)):
            out.append((f"{name}[{l}]", o, n)); o = al(o + n)
    for name, n in (("cat", B * 3 * H * 4), ("arg", B * H * 4)):
        out.append((name, o, n)); o = al(o + n)
    return out

def main():
    B, T, F, H, L, C = 512, 128, 64, 256, 2, 3
    prec = sys.argv[1] if len(sys.argv) > 1 else "bf16x3"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 6
    if len(sys.argv) > 3:                                   # B,T,F,H   e.g. 256,256,128,512 (the hidden-512 kernels, precision bf16)
        B, T, F, H = (int(v) for v in sys.argv[3].split(","))
    torch.manual_seed(0)
    m = pkg.BiGRU(H, F, C, L, 50, 0.0, False, True, precision=prec).cuda().eval()
    x = torch.randn(B, T, F, generator=torch.Generator().manual_seed(1234)).cuda()
    lay = layout(B, T, F, H, L) if prec == "bf16x3" else [("stash", 0, None)]
    ref = None
    for it in range(n):
        with torch.no_grad():
            y = m(x)
        torch.cuda.synchronize()
        plan, stash = m._last_plan_stash
        snap = stash.clone()
        if ref is None:
            ref, y0 = snap, y.clone()
            continue
        msgs = []
        for name, off, nb in lay:
            nb = ref.numel() if nb is None else nb
            a, b = ref[off:off + nb], snap[off:off + nb]
            if not torch.equal(a, b):
                nd = int((a != b).sum())
                first = int((a != b).nonzero()[0])
                msgs.append(f"{name}: {nd} bytes differ (first at +{first} of {nb})")
        for name, off, nb in lay:
            if name in ("Yhi[0]", "Ylo[0]") and not torch.equal(ref[off:off + nb], snap[off:off + nb]):
                a = ref[off:off + nb].view(torch.int16).view(B * T, 2 * H); b2 = snap[off:off + nb].view(torch.int16).view(B * T, 2 * H)
                rows, cols = (a != b2).nonzero(as_tuple=True)
                t, bb = rows // B, rows % B
                d = cols // H
                for dd in (0, 1):
                    sel = d == dd
                    if sel.any():
                        tt = t[sel]; tiles = torch.unique(bb[sel] // 32); units = cols[sel] % H
                        first_t = int(tt.min()) if dd == 0 else int(tt.max())
                        at_first = sel & (t == first_t)
                        print(f"   {name} dir {dd}: first bad t={first_t}, tiles {tiles.tolist()[:8]}, at first t: rows-in-tile {sorted(set((bb[at_first] % 32).tolist()))[:16]} "
                              f"units {sorted(set((cols[at_first] % H).tolist()))[:12]}.. ({int(at_first.sum())} values)")
        for name, off, nb in lay:
            if name == "G[0]" and not torch.equal(ref[off:off + nb], snap[off:off + nb]):
                CS = H // 64
                a = ref[off:off + nb].view(torch.float32).view(2, B // 32, T, CS, 4, 256, 8); b2 = snap[off:off + nb].view(torch.float32).view(2, B // 32, T, CS, 4, 256, 8)
                dd = (a != b2)
                idx = dd.nonzero()
                # earliest step per direction (dir 0: min t, dir 1: max t)
                for d_ in (0, 1):
                    sel = idx[idx[:, 0] == d_]
                    if len(sel) == 0: continue
                    t0 = int(sel[:, 2].min()) if d_ == 0 else int(sel[:, 2].max())
                    s0 = sel[sel[:, 2] == t0]
                    for g_ in range(4):
                        sg = s0[s0[:, 4] == g_]
                        if len(sg):
                            tids = sorted(set(sg[:, 5].tolist())); cols_ = sorted(set(sg[:, 6].tolist()))
                            k = sg[0]
                            print(f"   G dir {d_} t={t0} gate {'rznh'[g_]}: {len(sg)} values, tiles {sorted(set(sg[:,1].tolist()))} ctas {sorted(set(sg[:,3].tolist()))} warps {sorted(set(t_//32 for t_ in tids))} "
                                  f"lanes {sorted(set(t_%32 for t_ in tids))[:8]}.. cols {cols_}; e.g. {float(a[tuple(k.tolist())]):.6f} vs {float(b2[tuple(k.tolist())]):.6f}")
        print(f"run {it}: logits max diff {float((y - y0).abs().max()):.3e}; " + ("; ".join(msgs[:6]) if msgs else "stash identical"))

if __name__ == "__main__":
    main()
