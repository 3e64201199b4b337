"""Host matcher replay (csrc/match_replay.cpp) vs the oracle's RLE matcher/tracker on random
label stacks. Tables are built with numpy here (the GPU kernels that normally build them are
covered by the -m gpu tests); no GPU is needed."""
import numpy as np
import pytest

from oracle import tracking as otr


def tables_from_pan_stack(pan_stack, class_id, div):
    n = len(pan_stack)
    ccs = []
    for pan in pan_stack:
        ins = pan.copy()
        ins[(pan < class_id * div) | (pan >= (class_id + 1) * div)] = 0
        ccs.append(otr.connected_components(ins))
    cap = max(1, max(int(c.max()) for c in ccs))
    n_cc = np.array([int(c.max()) for c in ccs], dtype=np.int32)
    table = np.zeros((n, cap, 5), dtype=np.int32)
    for s, c in enumerate(ccs):
        for i in range(1, n_cc[s] + 1):
            ys, xs = np.nonzero(c == i)
            table[s, i - 1] = (len(ys), ys.min(), xs.min(), ys.max() + 1, xs.max() + 1)
    keys, vals = [], []
    for s in range(1, n):
        m = (ccs[s - 1] > 0) & (ccs[s] > 0)
        if m.any():
            k = (ccs[s - 1][m].astype(np.uint64) << np.uint64(20)) | ccs[s][m].astype(np.uint64)
            u, cnt = np.unique(k, return_counts=True)
            keys.append((np.uint64(s) << np.uint64(40)) | u)
            vals.append(cnt.astype(np.int32))
    keys = np.concatenate(keys) if keys else np.zeros(0, np.uint64)
    vals = np.concatenate(vals) if vals else np.zeros(0, np.int32)
    return ccs, n_cc, table, keys, vals


def random_stack(rng, n, h, w, mode):
    """Sequences of label images with persistent, splitting, merging and vanishing blobs."""
    stack = []
    yy, xx = np.mgrid[0:h, 0:w]
    # "crowded": many small blobs that keep touching, splitting and merging, so that most IoU
    # blocks hold several rows and columns (the assignment path) and labels merge in every slice
    k = rng.integers(18, 32) if mode == "crowded" else rng.integers(3, 9)
    cy, cx = rng.uniform(0, h, k), rng.uniform(0, w, k)
    r = rng.uniform(1.5, 4.5, k) if mode == "crowded" else rng.uniform(2, 7, k)
    for t in range(n):
        cy += rng.normal(0, 1.2, k); cx += rng.normal(0, 1.2, k); r = np.clip(r + rng.normal(0, 0.6, k), 1.0, 9)
        pan = np.zeros((h, w), dtype=np.int64)
        order = rng.permutation(k)
        for j, i in enumerate(order):
            if mode == "flicker" and rng.random() < 0.25:
                continue
            m = (yy - cy[i]) ** 2 + (xx - cx[i]) ** 2 <= r[i] ** 2
            pan[m] = 1000 + j + 1
        if mode == "symmetric":  # exact ties: mirrored blobs
            pan = np.zeros((h, w), dtype=np.int64)
            off = t % 3
            pan[4:10, 4 + off:10 + off] = 1001
            pan[4:10, w - 10 - off:w - 4 - off] = 1002
            if t % 2:
                pan[4:10, 10 + off:w - 10 - off] = 1003
        if mode == "lattice":    # a row of identical squares, shifted by half a period every slice:
            # each square overlaps two squares of the neighbouring slice EQUALLY, so the IoU matrix
            # has several optimal assignments (linear_sum_assignment tie order, matcher.py:213)
            pan = np.zeros((h, w), dtype=np.int64)
            period, side = 8, 6
            shift = (period // 2) * (t % 2)
            for j, x0 in enumerate(range(shift, w - side, period)):
                pan[3:3 + side, x0:x0 + side] = 1001 + j
                pan[h - 3 - side:h - 3, x0:x0 + side] = 1100 + j
        if mode == "noise":
            pan[rng.random((h, w)) < 0.08] = 0
        stack.append(pan)
    return stack


@pytest.mark.parametrize("mode", ["plain", "flicker", "noise", "symmetric", "lattice", "crowded"])
@pytest.mark.parametrize("axis_name", ["xy", "xz", "yz"])
def test_replay_matches_oracle(mode, axis_name):
    from empanada_napari_b200 import tracking
    rng = np.random.default_rng(hash((mode, axis_name)) % 2**32)
    for trial in range(6):
        n, h, w = int(rng.integers(3, 12)), int(rng.integers(16, 40)), int(rng.integers(16, 40))
        stack = random_stack(rng, n, h, w, mode)
        if trial == 0:
            stack[0][:] = 0  # empty first slice
        if trial == 1:
            stack[n // 2][:] = 0  # empty middle slice
        # oracle
        shape3d = {"xy": (n, h, w), "xz": (h, n, w), "yz": (h, w, n)}[axis_name]
        matchers = [otr.RLEMatcher(1, 1000, 0.25, 0.25)]
        rle_stack = otr.forward_matching([p.copy() for p in stack], matchers, [1], 1000, [1])
        tr = otr.InstanceTracker(1, 1000, shape3d, axis_name)
        per_slice_labels = {}
        for idx, rs in otr.backward_matching(rle_stack, matchers, n):
            tr.update(rs[1], idx)
            per_slice_labels[idx] = {lab: otr.rle_decode(a["starts"], a["runs"]) for lab, a in rs[1].items()}
        tr.finish()
        # product replay
        ccs, n_cc, table, keys, vals = tables_from_pan_stack(stack, 1, 1000)
        lut, labels, sizes, boxes = tracking.match_replay(n_cc, table, keys, vals, 1, 1000, axis_name)
        assert labels.tolist() == list(tr.instances.keys())
        for i, lab in enumerate(labels):
            inst = tr.instances[int(lab)]
            assert tuple(boxes[i].tolist()) == tuple(int(v) for v in inst["box"])
            assert int(sizes[i]) == int(np.sum(inst["runs"]))
        for s in range(n):
            final = lut[s][ccs[s]]
            expect = np.zeros(h * w, dtype=np.int64)
            for lab, flat in per_slice_labels[s].items():
                expect[flat] = lab
            assert np.array_equal(final.ravel(), expect), (mode, axis_name, trial, s)
