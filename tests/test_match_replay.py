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


# ---------------------------------------------------------------------------------------------
# Table-level restatement of the reference's matcher / tracker with SciPy's assignment on the
# FULL zero-padded IoU matrix (matcher.py:136-232,234-326; patterns.py:55-121; tracker.py:61-100).
# The native replay solves the assignment per connected block of the sparse matrix; this checks,
# on random tables full of exact ties, merges and splits, that both give the same decisions.
def table_level_replay(n_cc, table, keys, vals, class_id, div, axis_name, iou_thr, ioa_thr):
    from scipy.optimize import linear_sum_assignment
    n = len(n_cc)
    base = class_id * div
    inter = [dict() for _ in range(n + 1)]        # inter[s][(prev cc, cur cc)] between s-1 and s
    for k, v in zip(keys.tolist(), vals.tolist()):
        inter[k >> 40][((k >> 20) & 0xFFFFF, k & 0xFFFFF)] = v

    def slice_objects(s):
        return {base + c + 1: {"area": int(table[s, c, 0]), "box": tuple(int(v) for v in table[s, c, 1:5]), "cc": [c + 1]}
                for c in range(int(n_cc[s]))}

    def merge(a, b):
        return {"area": a["area"] + b["area"], "cc": a["cc"] + b["cc"],
                "box": (min(a["box"][0], b["box"][0]), min(a["box"][1], b["box"][1]),
                        max(a["box"][2], b["box"][2]), max(a["box"][3], b["box"][3]))}

    state = {"next": base + 1}

    def step(target, match, pair_inter, assign_new):
        """pair_inter(t_attrs, m_attrs) -> intersecting pixels of two (merged) objects."""
        t_labels, m_labels = list(target.keys()), list(match.keys())
        matches, ioa = {}, np.zeros((0, 0), np.float32)
        if t_labels and m_labels:
            iou = np.zeros((len(t_labels), len(m_labels)), dtype="float")
            ioa = np.zeros((len(t_labels), len(m_labels)), dtype=np.float32)
            for r, ta in enumerate(target.values()):
                for c, ma in enumerate(match.values()):
                    it = pair_inter(ta, ma)
                    if it > 0:
                        iou[r, c] = it / (ta["area"] + ma["area"] - it)
                        ioa[r, c] = it / ma["area"]
            rows, cols = linear_sum_assignment(iou, maximize=True)
            keep = iou[rows, cols] >= iou_thr
            matches = {m_labels[c]: t_labels[r] for r, c in zip(rows[keep], cols[keep])}
        out = {}
        for i, (ml, attrs) in enumerate(match.items()):
            if ml in matches:
                new = matches[ml]
            else:
                ioa_max = ioa[:, i].max() if len(ioa) > 0 else 0
                if ioa_max >= np.float32(ioa_thr):
                    new = t_labels[int(ioa[:, i].argmax())]
                elif assign_new:
                    new = state["next"]
                    state["next"] += 1
                else:
                    new = ml
            out[new] = attrs if new not in out else merge(out[new], attrs)
        return out

    fwd = []
    for s in range(n):
        cur = slice_objects(s)
        if s == 0:
            if cur:
                state["next"] = max(cur.keys()) + 1
            fwd.append(cur)
        else:
            tab = inter[s]
            fwd.append(step(fwd[-1], cur, lambda ta, ma: sum(tab.get((q, c), 0) for q in ta["cc"] for c in ma["cc"]), True))
    lut = np.zeros((n, table.shape[1] + 1), dtype=np.int32)
    instances = {}
    nxt = None
    for s in range(n - 1, -1, -1):
        if nxt is None:
            cur = fwd[s]
        else:
            tab = inter[s + 1]
            cur = step(nxt, fwd[s], lambda ta, ma: sum(tab.get((q, c), 0) for c in ta["cc"] for q in ma["cc"]), False)
        for label, a in cur.items():
            for c in a["cc"]:
                lut[s, c] = label
            y0, x0, y1, x1 = a["box"]
            b3 = {"xy": (s, y0, x0, s + 1, y1, x1), "xz": (y0, s, x0, y1, s + 1, x1), "yz": (y0, x0, s, y1, x1, s + 1)}[axis_name]
            if label not in instances:
                instances[label] = [a["area"], b3]
            else:
                o = instances[label][1]
                instances[label] = [instances[label][0] + a["area"],
                                    tuple(min(o[d], b3[d]) if d < 3 else max(o[d], b3[d]) for d in range(6))]
        nxt = cur
    return lut, instances


def random_tables(rng, n, maxc, density, tie):
    """Random component / overlap tables (no geometry behind them; the replay only sees tables)."""
    n_cc = rng.integers(0, maxc + 1, n).astype(np.int32)
    if rng.random() < 0.3:
        n_cc[rng.integers(0, n)] = 0
    cap = max(1, int(n_cc.max()))
    table = np.zeros((n, cap, 5), np.int32)
    for s in range(n):
        c = n_cc[s]
        table[s, :c, 0] = rng.integers(2, 5, c) if tie else rng.integers(1, 200, c)
        y0, x0 = rng.integers(0, 50, c), rng.integers(0, 50, c)
        table[s, :c, 1], table[s, :c, 2] = y0, x0
        table[s, :c, 3], table[s, :c, 4] = y0 + rng.integers(1, 20, c), x0 + rng.integers(1, 20, c)
    keys, vals = [], []
    for s in range(1, n):
        a, b = int(n_cc[s - 1]), int(n_cc[s])
        if a == 0 or b == 0:
            continue
        npair = min(a * b, rng.binomial(a * b, min(1.0, density / max(a, b))))
        if npair == 0:
            continue
        flat = rng.choice(a * b, size=npair, replace=False)
        q, c = flat // b + 1, flat % b + 1
        lim = np.minimum(table[s - 1, q - 1, 0], table[s, c - 1, 0])
        v = np.ones(npair, np.int64) if tie else np.maximum(1, (lim * rng.uniform(0.05, 1.0, npair) / 3).astype(np.int64))
        keys.append((np.uint64(s) << np.uint64(40)) | (q.astype(np.uint64) << np.uint64(20)) | c.astype(np.uint64))
        vals.append(v.astype(np.int32))
    keys = np.concatenate(keys) if keys else np.zeros(0, np.uint64)
    vals = np.concatenate(vals) if vals else np.zeros(0, np.int32)
    perm = rng.permutation(len(keys))
    return n_cc, table, np.ascontiguousarray(keys[perm]), np.ascontiguousarray(vals[perm])


@pytest.mark.parametrize("seed", range(4))
def test_replay_vs_full_matrix_scipy_on_random_tables(seed):
    from empanada_napari_b200 import tracking
    rng = np.random.default_rng(1000 + seed)
    for trial in range(150):
        n = int(rng.integers(1, 14))
        maxc = int(rng.choice([1, 3, 6, 12]))
        density = float(rng.choice([0.5, 1.0, 2.0, 3.5]))
        tie = bool(rng.random() < 0.5)
        n_cc, table, keys, vals = random_tables(rng, n, maxc, density, tie)
        axis_name = ("xy", "xz", "yz")[int(rng.integers(0, 3))]
        thr = float(rng.choice([0.25, 0.25, 0.5, 0.125]))
        lut, labels, sizes, boxes = tracking.match_replay(n_cc, table, keys, vals, 1, 1000, axis_name, thr, thr)
        want_lut, want = table_level_replay(n_cc, table, keys, vals, 1, 1000, axis_name, thr, thr)
        ctx = (seed, trial, n, maxc, density, tie, thr)
        assert labels.tolist() == list(want.keys()), ctx
        assert np.array_equal(lut, want_lut), ctx
        for i, lab in enumerate(labels.tolist()):
            assert int(sizes[i]) == want[lab][0] and tuple(boxes[i].tolist()) == tuple(want[lab][1]), ctx


def test_replay_diagnostics_count_steps_and_full_matrix_replays():
    """`be_match_replay_stats`: matcher steps = 2 (n - 1); a slice pair in which two objects tie
    exactly for one target must be replayed on the full matrix, a tie-free one must not."""
    import ctypes
    from empanada_napari_b200 import _lib, tracking

    def stats():
        out = (ctypes.c_longlong * 3)()
        _lib.lib().be_match_replay_stats(ctypes.cast(out, ctypes.c_void_p))
        return list(out)

    def tables(inter_b):
        n_cc = np.array([1, 2], dtype=np.int32)
        table = np.zeros((2, 2, 5), dtype=np.int32)
        table[0, 0] = (10, 0, 0, 4, 4)
        table[1, 0] = (6, 0, 0, 2, 3)
        table[1, 1] = (6, 2, 0, 4, 3)
        keys = np.array([(1 << 40) | (1 << 20) | 1, (1 << 40) | (1 << 20) | 2], dtype=np.uint64)
        return n_cc, table, keys, np.array([3, inter_b], dtype=np.int32)

    tracking.match_replay(*tables(3), 1, 1000, "xy")        # both candidates: IoU 3 / 13, an exact tie
    steps, multi, full = stats()
    assert steps == 2 and full >= 1
    tracking.match_replay(*tables(2), 1, 1000, "xy")        # 3 / 13 against 2 / 14: unique optimum
    steps, multi, full = stats()
    assert steps == 2 and full == 0
