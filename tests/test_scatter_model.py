"""Executable model of the windowed embedding-gradient scatter (ader_b200/csrc/encoder.cu: k_scatter_apply) on CPU.

The CUDA kernel is checked against the oracle on the GPU (tests/test_gpu_parity.py); this file pins the ALGORITHM it
implements -- window / run / piece bookkeeping, the segment bounds the plan works out per window (k_seg_bounds: head and
end probing), which slot a piece is parked in, who combines the pieces and in which order -- on the edge cases a GPU batch rarely hits: segments that end exactly on a window edge,
windows that lie completely inside one segment, a last window shorter than 16, one token, one item for every token.
The model walks the sorted (item id, token) pairs exactly like a warp does and must (a) reproduce a plain scatter-add
up to fp32 re-association, (b) touch every table row exactly once (the kernel's single-contributor reduction), (c) give
the same bits for any arrival order of the pieces (determinism), (d) see every spanning segment combined exactly once.
"""
import numpy as np
import pytest

SW = 16   # sorted positions per warp window (encoder.cu: SW)


def _stable_sort(ids):
    order = np.argsort(ids, kind="stable")          # k_sort_hist/scan/scatter: stable LSD radix sort of (id, token)
    return ids[order].astype(np.int64), order.astype(np.int64)


def seg_bounds_model(keys):
    """k_seg_bounds (scatter plan, a warp per window): bounds[w] = (head of the segment of the window's FIRST run if that run
    continues from the previous window, else p0; end (exclusive) of the segment of its LAST run if it continues into the next
    window, else p0 + n).  Probes in steps of 32 positions like the warp's ballots."""
    T = len(keys)
    n_win = (T + SW - 1) // SW
    bounds = np.zeros((n_win, 2), np.int64)
    for w in range(n_win):
        p0 = w * SW
        n = min(SW, T - p0)
        kf, kl = keys[p0], keys[p0 + n - 1]
        head, end = p0, p0 + n
        if p0 > 0 and keys[p0 - 1] == kf:
            base = p0 - 32
            while True:
                q = np.arange(base, base + 32)
                diff = (q < 0) | (keys[np.clip(q, 0, T - 1)] != kf)
                if diff.any():
                    head = base + int(np.flatnonzero(diff).max()) + 1
                    break
                base -= 32
        if p0 + n < T and keys[p0 + n] == kl:
            base = p0 + n
            while True:
                q = np.arange(base, base + 32)
                diff = (q >= T) | (keys[np.clip(q, 0, T - 1)] != kl)
                if diff.any():
                    end = base + int(np.flatnonzero(diff).min())
                    break
                base += 32
        bounds[w] = (head, end)
    return bounds


def scatter_model(ids, gx, n_rows, scale, window_order=None):
    """Returns (table [n_rows, d] fp32, writes per row, combines per spanning segment head-window)."""
    keys, vals = _stable_sort(np.asarray(ids))
    T, d = len(keys), gx.shape[1]
    table = np.zeros((n_rows, d), np.float32)
    writes = np.zeros(n_rows, np.int64)
    n_win = (T + SW - 1) // SW
    part = np.full((n_win, 2, d), np.nan, np.float32)            # partial slots; NaN = never written
    counter = np.zeros(n_win, np.int64)
    combines = {}
    bounds = seg_bounds_model(keys)
    windows = list(range(n_win)) if window_order is None else list(window_order)
    for w in windows:                                              # any order: warps run concurrently
        p0 = w * SW
        n = min(SW, T - p0)
        key_prev = keys[p0 - 1] if p0 > 0 else -1
        key_next = keys[p0 + n] if p0 + n < T else -1
        acc = np.zeros(d, np.float32)
        run_start = 0
        for u in range(n):
            ku = keys[p0 + u]
            acc = acc + gx[vals[p0 + u]].astype(np.float32)       # rows of a run are added in sorted (= token) order
            last = (u + 1 == n)
            if last or keys[p0 + u + 1] != ku:
                cont_after = last and key_next == ku
                cont_before = run_start == 0 and key_prev == ku
                if not cont_after and not cont_before:              # whole segment inside the window: one reduction
                    table[ku] += np.float32(scale) * acc
                    writes[ku] += 1
                else:
                    part[w, 0 if cont_before else 1] = acc
                    head, end = p0 + run_start, p0 + u + 1
                    if cont_before:                                 # the plan's bounds (k_seg_bounds), not a probe here
                        head = int(bounds[w, 0])
                    if cont_after:
                        end = int(bounds[w, 1])
                    assert keys[head] == ku and (head == 0 or keys[head - 1] != ku)
                    assert keys[end - 1] == ku and (end == T or keys[end] != ku)
                    w1, w2 = head // SW, (end - 1) // SW
                    old = counter[w1]
                    counter[w1] += 1
                    if old == w2 - w1:                              # last piece to arrive combines, in window order
                        tot = np.zeros(d, np.float32)
                        for pw in range(w1, w2 + 1):
                            piece = part[pw, 1 if pw == w1 else 0]
                            assert not np.isnan(piece).any(), "piece read before it was written"
                            tot = tot + piece
                        table[ku] += np.float32(scale) * tot
                        writes[ku] += 1
                        counter[w1] = 0
                        combines[w1] = combines.get(w1, 0) + 1
                acc = np.zeros(d, np.float32)
                run_start = u + 1
    assert (counter == 0).all(), "an arrival counter was left armed"
    return table, writes, combines


def _reference(ids, gx, n_rows, scale):
    out = np.zeros((n_rows, gx.shape[1]), np.float64)
    np.add.at(out, np.asarray(ids), gx.astype(np.float64))
    return out * scale


CASES = {
    "one_token": [5],
    "all_distinct_31": list(range(1, 32)),
    "one_item_16": [3] * 16,                       # exactly one window, complete segment
    "one_item_17": [3] * 17,                       # spills one position into the next window
    "one_item_32": [3] * 32,                       # two full windows, ends on the edge
    "one_item_100": [3] * 100,                     # windows completely inside the segment + short last window
    "edge_aligned": [1] * 16 + [2] * 16 + [3] * 48 + [4] * 1,
    "edge_straddle": [1] * 15 + [2] * 2 + [3] * 31 + [4] * 40 + [5] * 3,
    "two_hot_neighbours": [7] * 70 + [8] * 70,
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_windowed_scatter_edge_cases(name):
    import zlib
    rng = np.random.RandomState(zlib.crc32(name.encode()))
    ids = np.array(CASES[name])
    n_rows = int(ids.max()) + 1
    rng.shuffle(ids)                                               # token order is arbitrary; the sort groups the items
    gx = rng.randn(len(ids), 7).astype(np.float32)
    table, writes, combines = scatter_model(ids, gx, n_rows, 12.25)
    np.testing.assert_allclose(table, _reference(ids, gx, n_rows, 12.25), rtol=2e-5, atol=2e-5)
    present = np.unique(ids)
    assert (writes[present] == 1).all() and writes.sum() == len(present)      # one contributor per row
    assert all(c == 1 for c in combines.values())


def test_windowed_scatter_random_batches_and_arrival_orders():
    rng = np.random.RandomState(3)
    for trial in range(25):
        T = int(rng.randint(1, 400))
        vmax = int(rng.choice([1, 2, 5, 40, 400]))
        u = rng.random_sample(T)
        ids = np.maximum(1, np.minimum(vmax, np.floor(vmax ** u))).astype(np.int64)   # Zipf-like: a few hot items
        gx = rng.randn(T, 5).astype(np.float32)
        base, writes, _ = scatter_model(ids, gx, vmax + 1, 3.0)
        np.testing.assert_allclose(base, _reference(ids, gx, vmax + 1, 3.0), rtol=5e-5, atol=5e-5)
        assert (writes[np.unique(ids)] == 1).all()
        n_win = (T + SW - 1) // SW
        for _ in range(3):                                         # pieces may arrive in any order: same bits
            order = rng.permutation(n_win)
            again, _, _ = scatter_model(ids, gx, vmax + 1, 3.0, window_order=order)
            assert np.array_equal(base, again)


def test_segment_bounds_of_the_plan():
    """bounds[w] against a direct computation from the run-length structure of the sorted keys."""
    rng = np.random.RandomState(11)
    for trial in range(40):
        T = int(rng.randint(1, 500))
        vmax = int(rng.choice([1, 2, 3, 30]))
        keys = np.sort(rng.randint(1, vmax + 1, T)).astype(np.int64)
        change = np.flatnonzero(np.diff(keys)) + 1
        starts = np.concatenate([[0], change]); ends = np.concatenate([change, [T]])
        seg_of = np.repeat(np.arange(len(starts)), ends - starts)
        b = seg_bounds_model(keys)
        for w in range(len(b)):
            p0 = w * SW; n = min(SW, T - p0)
            first, last = seg_of[p0], seg_of[p0 + n - 1]
            want_head = starts[first] if (p0 > 0 and keys[p0 - 1] == keys[p0]) else p0
            want_end = ends[last] if (p0 + n < T and keys[p0 + n] == keys[p0 + n - 1]) else p0 + n
            assert (b[w, 0], b[w, 1]) == (want_head, want_end), (trial, w)
