"""CPU oracle for the biGRU hot path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this file.  The product (financial_market_data_analysis_b200) never
imports anything under oracle/ and has no CPU fallback.

Two restatements live here:

* ``OracleBiGRU`` - the wrapper of /root/reference/biGRU_model.py:32-138 restated on top of
  ``torch.nn.GRU`` on CPU.  The recurrent arithmetic of the reference is not in the
  reference's own source: it is the third-party dependency ``torch`` (pinned
  ``torch==1.2.0`` in /root/reference/requirements.txt:11; this image has 2.11.0), reached at
  biGRU_model.py:54-56 (construct) and :102 (call).  This class makes the same library
  call, so it is also what bench.py times as the reference CPU arm (kind "port").
* ``gru_forward_np`` / ``gru_backward_np`` - the published GRU equations and their BPTT
  written out in float64 numpy (no torch), used to pin gate order, the position of b_hn,
  reverse-direction indexing, head concat order and the max-pool tie rule independently
  of torch.  oracle/bigru_ref.c is the same algorithm in plain C.

Pinning: tests/golden/*.npz were produced by tests/golden/make_golden.py, which imports the
UNMODIFIED reference class from /root/reference in the build container (known-answer
vectors from the shipped model_params.pt plus seed-fixed synthetic cases).
tests/test_oracle_cpu.py checks both restatements against those fixtures.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn


class OracleBiGRU(nn.Module):
    """Restatement of the reference model wrapper (biGRU_model.py:8-138).

    Same constructor argument order, same submodule names (``gru``, ``linear``) so the
    reference state_dict loads, same forward semantics:
    dropout -> nn.GRU(batch_first, bidirectional) -> direction-sum ->
    cat[last_hidden, max_t, mean_t] -> Linear(3H -> C).
    """

    def __init__(self, hidden_size, n_features, output_size, n_layers=1, clip=50,
                 dropout=0.2, spatial_dropout=True, bidirectional=True):
        super().__init__()
        self.hidden_size, self.n_features, self.output_size = hidden_size, n_features, output_size
        self.n_layers, self.clip, self.dropout_p = n_layers, clip, dropout
        self.spatial_dropout, self.bidirectional = spatial_dropout, bidirectional
        self.n_directions = 2 if bidirectional else 1
        self.dropout = nn.Dropout(dropout)                       # biGRU_model.py:50
        if spatial_dropout:
            self.spatial_dropout1d = nn.Dropout2d(dropout)       # :51-52
        self.gru = nn.GRU(n_features, hidden_size, num_layers=n_layers,
                          dropout=(0 if n_layers == 1 else dropout),
                          batch_first=True, bidirectional=bidirectional)   # :54-56
        self.linear = nn.Linear(3 * hidden_size, output_size)    # :60

    def forward(self, x, hidden=None):
        B, T = x.size(0), x.size(1)
        if self.spatial_dropout:                                 # :87-92 channel-wise over T
            x = self.spatial_dropout1d(x.transpose(1, 2)).transpose(1, 2)
        else:
            x = self.dropout(x)                                  # :94
        out, h_n = self.gru(x, hidden)                           # :102
        last = h_n.view(self.n_layers, self.n_directions, B, self.hidden_size)[-1].sum(0)  # :111-115
        if self.bidirectional:
            out = out[..., :self.hidden_size] + out[..., self.hidden_size:]               # :119-120
        mx = out.max(dim=1).values                               # :125 adaptive_max_pool1d over T
        av = out.sum(dim=1) / float(T)                           # :130
        return self.linear(torch.cat([last, mx, av], dim=1))     # :133-137


def forward_routed(model: OracleBiGRU, x, hidden=None, idx=None):
    """``OracleBiGRU.forward`` (eval / dropout-free) with the max-pool's gradient routing made explicit: with ``idx``
    ([B, H] time indices) the pooled maximum is read at those steps (``gather``), otherwise at the arg-max.  The max-pool
    (biGRU_model.py:125) is discontinuous where two steps tie to within rounding, so parity tests compare gradients under
    the routing the kernel actually took and separately require every disagreement to be such a tie.
    Returns (logits, s) with s[B, T, H] the direction-summed GRU output."""
    B, T = x.size(0), x.size(1)
    H = model.hidden_size
    out, h_n = model.gru(x, hidden)
    last = h_n.view(model.n_layers, model.n_directions, B, H)[-1].sum(0)
    s = out[..., :H] + out[..., H:] if model.bidirectional else out
    mx = s.max(dim=1).values if idx is None else s.gather(1, idx.unsqueeze(1)).squeeze(1)
    av = s.sum(dim=1) / float(T)
    return model.linear(torch.cat([last, mx, av], dim=1)), s


def train_step(model: OracleBiGRU, optimizer, loss_fn, x, target):
    """Body of biGRU_model.py:198-210 (without the sklearn metrics)."""
    optimizer.zero_grad()
    pred = model(x)
    loss = loss_fn(pred, target)
    loss.backward()
    nn.utils.clip_grad_norm_(model.parameters(), model.clip)
    optimizer.step()
    return loss.detach()


# --------------------------------------------------------------------------------------
# Explicit equations (float64 numpy).  Parameter dict uses the nn.GRU / nn.Linear names.
# --------------------------------------------------------------------------------------

def _sig(a):
    return 1.0 / (1.0 + np.exp(-a))


def _suffix(layer, d):
    return f"l{layer}" + ("_reverse" if d == 1 else "")


def gru_forward_np(params, x, H, L, bidirectional=True, h0=None, keep=False):
    """Forward of biGRU_model.py:63-138 in eval mode, float64.

    GRU cell (torch.nn.GRU docs; gate row order of weight_ih/hh is r | z | n):
        r = s(W_ir x + b_ir + W_hr h + b_hr);  z = s(W_iz x + b_iz + W_hz h + b_hz)
        n = tanh(W_in x + b_in + r * (W_hn h + b_hn));  h' = (1 - z) * n + z * h
    Returns logits [B,C] (and a cache for gru_backward_np if keep).
    """
    x = np.asarray(x, np.float64)
    B, T, _ = x.shape
    D = 2 if bidirectional else 1
    P = {k: np.asarray(v, np.float64) for k, v in params.items()}
    inp = x
    cache = {"layers": []}
    h_last = None
    for l in range(L):
        out = np.zeros((B, T, D * H))
        lc = []
        h_last = []
        for d in range(D):
            sfx = _suffix(l, d)
            Wi, Wh = P[f"gru.weight_ih_{sfx}"], P[f"gru.weight_hh_{sfx}"]
            bi, bh = P[f"gru.bias_ih_{sfx}"], P[f"gru.bias_hh_{sfx}"]
            h = np.zeros((B, H)) if h0 is None else np.asarray(h0, np.float64)[l * D + d]
            order = range(T) if d == 0 else range(T - 1, -1, -1)
            rs, zs, ns, hns, hps = {}, {}, {}, {}, {}
            for t in order:
                gi = inp[:, t] @ Wi.T + bi
                gh = h @ Wh.T + bh
                r = _sig(gi[:, :H] + gh[:, :H])
                z = _sig(gi[:, H:2 * H] + gh[:, H:2 * H])
                n = np.tanh(gi[:, 2 * H:] + r * gh[:, 2 * H:])
                hps[t] = h
                h = (1.0 - z) * n + z * h
                out[:, t, d * H:(d + 1) * H] = h
                rs[t], zs[t], ns[t], hns[t] = r, z, n, gh[:, 2 * H:]
            h_last.append(h)
            lc.append((rs, zs, ns, hns, hps))
        cache["layers"].append((inp, lc))
        inp = out
    s = out[..., :H] + out[..., H:] if D == 2 else out
    last = sum(h_last)                       # top layer, sum over directions (:111-115)
    arg = s.argmax(axis=1)                   # first maximum wins (adaptive_max_pool1d rule)
    mx = np.take_along_axis(s, arg[:, None, :], 1)[:, 0]
    av = s.sum(1) / float(T)
    cat = np.concatenate([last, mx, av], 1)
    logits = cat @ P["linear.weight"].T + P["linear.bias"]
    if keep:
        cache.update(cat=cat, arg=arg, T=T, B=B, D=D, H=H, L=L, P=P, h0=h0)
        return logits, cache
    return logits


def gru_backward_np(cache, dlogits):
    """BPTT of gru_forward_np.  Returns {param name: grad}, dx [B,T,F], dh0 [L*D,B,H]."""
    P, H, L, D, T, B = cache["P"], cache["H"], cache["L"], cache["D"], cache["T"], cache["B"]
    dlogits = np.asarray(dlogits, np.float64)
    g = {"linear.weight": dlogits.T @ cache["cat"], "linear.bias": dlogits.sum(0)}
    dcat = dlogits @ P["linear.weight"]
    dlast, dmax, davg = dcat[:, :H], dcat[:, H:2 * H], dcat[:, 2 * H:]
    ds = np.repeat((davg / T)[:, None, :], T, 1)
    bi_, ji_ = np.meshgrid(np.arange(B), np.arange(H), indexing="ij")
    ds[bi_, cache["arg"], ji_] += dmax
    dout = np.concatenate([ds] * D, 2)
    dh0 = np.zeros((L * D, B, H))
    for l in range(L - 1, -1, -1):
        inp, lc = cache["layers"][l]
        dinp = np.zeros_like(inp)
        for d in range(D):
            sfx = _suffix(l, d)
            Wi, Wh = P[f"gru.weight_ih_{sfx}"], P[f"gru.weight_hh_{sfx}"]
            rs, zs, ns, hns, hps = lc[d]
            dWi, dWh = np.zeros_like(Wi), np.zeros_like(Wh)
            dbi, dbh = np.zeros(3 * H), np.zeros(3 * H)
            dh = dlast.copy() if l == L - 1 else np.zeros((B, H))
            order = range(T - 1, -1, -1) if d == 0 else range(T)
            for t in order:
                dh = dh + dout[:, t, d * H:(d + 1) * H]
                r, z, n, hn, hp = rs[t], zs[t], ns[t], hns[t], hps[t]
                dn = dh * (1.0 - z)
                dz = dh * (hp - n)
                dan = dn * (1.0 - n * n)
                dar = dan * hn * r * (1.0 - r)
                daz = dz * z * (1.0 - z)
                dgi = np.concatenate([dar, daz, dan], 1)
                dgh = np.concatenate([dar, daz, dan * r], 1)
                dWi += dgi.T @ inp[:, t]
                dbi += dgi.sum(0)
                dWh += dgh.T @ hp
                dbh += dgh.sum(0)
                dinp[:, t] += dgi @ Wi
                dh = dh * z + dgh @ Wh
            dh0[l * D + d] = dh
            g[f"gru.weight_ih_{sfx}"], g[f"gru.weight_hh_{sfx}"] = dWi, dWh
            g[f"gru.bias_ih_{sfx}"], g[f"gru.bias_hh_{sfx}"] = dbi, dbh
        dout = dinp
    return g, dout, dh0


def ce_loss_np(logits, target):
    """Mean softmax cross-entropy and its gradient wrt logits (torch.nn.CrossEntropyLoss)."""
    lg = np.asarray(logits, np.float64)
    m = lg.max(1, keepdims=True)
    lse = m[:, 0] + np.log(np.exp(lg - m).sum(1))
    B = lg.shape[0]
    loss = (lse - lg[np.arange(B), target]).mean()
    p = np.exp(lg - lse[:, None])
    p[np.arange(B), target] -= 1.0
    return loss, p / B


def state_dict_to_np(sd):
    return {k: v.detach().cpu().numpy() for k, v in sd.items()}
