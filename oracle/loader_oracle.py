"""CPU oracle for the windowed collation path.  TEST INFRASTRUCTURE ONLY (see bigru_oracle.py).

Restates, on plain numpy arrays, the rules of /root/reference/sql_pytorch_dataloader.py:
  window_indices :8-18, chunk ranges :72-78, min==max guard :108-113, order-book level
  sharing :119-144, normalisation :239, window/target pairing :243-245, chunk split :287-320,
and the iteration behaviour of ``DataLoader(MySQLBatchLoader(...), batch_size)`` (the
exhausted-generator quirk: __len__ is the row count, :247-248, so the batch in flight when
the window generator runs dry is silently lost).
Pinned by tests/golden/loader_*.npz, produced by running the unmodified reference classes
against tests/fake_db.FakeCursor (tests/golden/make_golden.py).
"""
from __future__ import annotations

import numpy as np


def window_indices(n_rows: int, window: int):
    """Stride-1 windows over range(n_rows): tuples (i, ..., i+window-1)  (:8-18)."""
    return [tuple(range(i, i + window)) for i in range(0, n_rows - window + 1)]


def chunk_ranges(db_length: int, chunk_size: int, window: int):
    """Database ID ranges of each chunk, consecutive chunks overlapping by window-1 rows (:68-78)."""
    n = db_length // chunk_size
    out = []
    for c in range(n + 1):
        if c == 0:
            out.append(range(window, chunk_size))
        elif c < n:
            out.append(range(chunk_size * c - window + 1, chunk_size * (c + 1)))
        else:
            out.append(range(chunk_size * c - window + 1, db_length + 1))
    return out


def guard_min_max(x_min: np.ndarray, x_max: np.ndarray):
    """min == max would divide by zero: widen max by 0.1 % (or by 1e-3 when it is 0)  (:108-113).
    float32 arithmetic, as the reference does it on torch.Tensor."""
    x_min = np.array(x_min, np.float32)
    x_max = np.array(x_max, np.float32)
    eq = x_min == x_max
    nz = eq & (x_max != 0)
    x_max[nz] = x_max[nz] + x_max[nz] * np.float32(0.001)
    x_max[eq & ~nz] = x_max[eq & ~nz] + np.float32(0.001)
    return x_min, x_max


def share_order_book(x_fields, x_min, x_max, bid_levels, ask_levels):
    """All ask (resp. bid) size levels share the min/max taken over the levels (:119-144)."""
    if "sd.bid_0_size" not in x_fields:
        return x_min, x_max
    for side, levels in (("ask", ask_levels), ("bid", bid_levels)):
        idx = [x_fields.index(f"sd.{side}_{i}_size") for i in range(levels) if f"sd.{side}_{i}_size" in x_fields]
        if idx:
            x_min[idx] = x_min[idx].min()
            x_max[idx] = x_max[idx].max()
    return x_min, x_max


def normalise(x, x_min, x_max):
    return (np.asarray(x, np.float32) - x_min) / (x_max - x_min)          # (:239)


def delivered_batches(n_rows: int, window: int, batch_size: int):
    """Start rows of the windows each DataLoader batch contains, as the reference delivers them."""
    n_win = max(n_rows - window + 1, 0)
    starts = list(range(n_win))
    batches = [starts[i:i + batch_size] for i in range(0, n_win, batch_size)]
    if batches and len(batches[-1]) < batch_size and n_win < n_rows:
        batches.pop()            # generator ran dry mid-batch -> StopIteration ends the epoch
    elif batches and len(batches[-1]) == batch_size and n_win < n_rows:
        pass                     # next fetch raises on its first sample; nothing in flight is lost
    return batches


def collate(x_norm, y, starts, window):
    """x[B,T,F], y[B,1,C] for the given window start rows (:243-245 + default_collate)."""
    xb = np.stack([x_norm[s:s + window] for s in starts])
    yb = np.stack([y[s + window - 1:s + window] for s in starts])
    return xb, yb


def split_sizes(n_chunks: int, val_size=0.1, test_size=0.1):
    """Chunk-granular contiguous split (:287-320): returns (train, val, test) index slices."""
    train_end = int((1 - val_size - test_size) * n_chunks)
    val_end = train_end + int(val_size * n_chunks) + 1
    test_end = val_end + int(test_size * n_chunks) + 1
    return slice(0, train_end), slice(train_end, val_end), slice(val_end, test_end)
