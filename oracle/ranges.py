"""CPU restatement of the reference's box / run-length / range arithmetic
(empanada/array_utils.py). TEST INFRASTRUCTURE ONLY (see oracle/post.py header).

Pinned against the reference's own known-answer tests, tests/test_array_utils.py:8-154
(re-stated in tests/test_oracle_golden.py), including the two quirks those tests pin:
`rle_voting([(10,20),(7,26)]) -> [[10,20],[23,26]]` and
`invert_ranges([(2,6),(4,12)], 15) -> [[0,2],[6,4],[12,15]]`.
The loop bodies follow the reference's control flow so the quirks reproduce; numba (present in
this image, and the reference's own accelerator) is used when importable, else plain Python.
"""
import numpy as np

try:  # same accelerator the reference uses; optional
    import numba
    _jit = numba.jit(nopython=True, cache=False)
except Exception:  # pragma: no cover
    def _jit(f):
        return f


def merge_boxes(box1, box2):
    """array_utils.py:105-129."""
    n = len(box1)
    nd = n // 2
    return tuple(min(box1[i], box2[i]) if i < nd else max(box1[i], box2[i]) for i in range(n))


@_jit
def _box_pairs(boxes1, boxes2):
    """array_utils.py:148-176 (_box_iou): pairs with positive box intersection."""
    ndim = boxes1.shape[1] // 2
    rows = []
    cols = []
    ious = []
    inters = []
    for x in range(boxes1.shape[0]):
        for y in range(boxes2.shape[0]):
            inter = 1
            a1 = 1
            a2 = 1
            for i in range(ndim):
                lo = max(boxes1[x, i], boxes2[y, i])
                hi = min(boxes1[x, i + ndim], boxes2[y, i + ndim])
                inter *= max(0, hi - lo)
                a1 *= boxes1[x, i + ndim] - boxes1[x, i]
                a2 *= boxes2[y, i + ndim] - boxes2[y, i]
                if inter == 0:
                    break
            if inter > 0:
                rows.append(x)
                cols.append(y)
                ious.append(inter / (a1 + a2 - inter))
                inters.append(inter)
    return rows, cols, ious, inters


def box_iou_pairs(boxes1, boxes2=None):
    """array_utils.py:178-211 (box_iou(...).nonzero()): (k,2) index pairs, row-major order."""
    if boxes2 is None:
        boxes2 = boxes1
    b1 = np.asarray(boxes1, dtype=np.int64)
    b2 = np.asarray(boxes2, dtype=np.int64)
    if len(b1) == 0 or len(b2) == 0:
        return np.zeros((0, 2), dtype=np.int64), [], []
    rows, cols, ious, inters = _box_pairs(b1, b2)
    pairs = np.array([list(rows), list(cols)], dtype=np.int64).T.reshape(-1, 2)
    return pairs, list(ious), list(inters)


def rle_encode(indices):
    """array_utils.py:213-239."""
    indices = np.asarray(indices)
    changes = np.where(indices[1:] != indices[:-1] + 1)[0] + 1
    changes = np.insert(changes, 0, [0], axis=0)
    changes = np.append(changes, [len(indices)], axis=0)
    runs = changes[1:] - changes[:-1]
    changes = changes[:-1]
    return indices[changes], runs


def rle_decode(starts, runs):
    """array_utils.py:241-256."""
    ends = starts + runs
    return np.concatenate([np.arange(s, e) for s, e in zip(starts, ends)])


def rle_to_string(starts, runs):
    """array_utils.py:258-271."""
    return " ".join(f"{i} {r}" for i, r in zip(starts, runs))


def string_to_rle(encoding):
    """array_utils.py:273-287."""
    enc = np.array([int(i) for i in encoding.split(" ")])
    return enc[::2], enc[1::2]


@_jit
def intersection_from_ranges(merged_runs, changes):
    """array_utils.py:344-373."""
    total = 0
    have = False
    c0 = 0
    c1 = 0
    for i in range(len(changes)):
        r2s = merged_runs[i + 1, 0]
        r2e = merged_runs[i + 1, 1]
        if changes[i]:
            c0 = merged_runs[i, 0]
            c1 = merged_runs[i, 1]
            have = True
        elif not have:
            continue
        if c1 < r2s:
            continue
        total += min(c1, r2e) - max(c0, r2s)
    return total


def rle_intersection(starts_a, runs_a, starts_b, runs_b):
    """array_utils.py:375-407."""
    ra = np.stack([starts_a, starts_a + runs_a], axis=1)
    rb = np.stack([starts_b, starts_b + runs_b], axis=1)
    merged = np.concatenate([ra, rb], axis=0).astype(np.int64)
    ids = np.concatenate([np.repeat([0], len(ra)), np.repeat([1], len(rb))])
    order = np.argsort(merged, axis=0, kind="stable")[:, 0]
    merged = merged[order]
    ids = ids[order]
    changes = ids[:-1] != ids[1:]
    if len(changes) == 0:
        return 0
    return int(intersection_from_ranges(merged, changes))


def rle_iou(starts_a, runs_a, starts_b, runs_b, return_intersection=False):
    """array_utils.py:409-433 (int64 / int64 -> float64)."""
    inter = rle_intersection(starts_a, runs_a, starts_b, runs_b)
    union = runs_a.sum() + runs_b.sum() - inter
    iou = inter / union
    return (iou, inter) if return_intersection else iou


def rle_ioa(starts_a, runs_a, starts_b, runs_b, return_intersection=False):
    """array_utils.py:435-459 (area of b)."""
    inter = rle_intersection(starts_a, runs_a, starts_b, runs_b)
    area = runs_b.sum()
    ioa = inter / area
    return (ioa, inter) if return_intersection else ioa


@_jit
def split_range_by_votes(running_range, num_votes, vote_thr=2):
    """array_utils.py:461-519."""
    out = np.empty((0, 2), dtype=np.int64)
    s_assigned = False
    e_assigned = False
    s = 0
    e = 0
    for ix in range(len(num_votes)):
        n = num_votes[ix]
        if n >= vote_thr:
            if not s_assigned:
                s = running_range[0] + ix
                s_assigned = True
            else:
                e = running_range[0] + ix + 1
                e_assigned = True
        elif s_assigned:
            if not e_assigned:
                e = s + 1
            cur = np.empty((1, 2), dtype=np.int64)
            cur[0, 0] = s
            cur[0, 1] = e
            out = np.vstack((out, cur))
            s_assigned = False
            e_assigned = False
    if s_assigned:
        if not e_assigned:
            e = s + 1
        cur = np.empty((1, 2), dtype=np.int64)
        cur[0, 0] = s
        cur[0, 1] = e
        out = np.vstack((out, cur))
    return out


@_jit
def extend_range(range1, range2, num_votes):
    """array_utils.py:521-561 (negative first_idx wraps, exactly as numpy/numba indexing)."""
    first_idx = range2[0] - range1[0]
    last_idx = len(num_votes)
    end_offset = range2[1] - range1[1]
    if end_offset > 0:
        range1[1] = range2[1]
        num_votes = np.concatenate((num_votes, np.ones(end_offset, dtype=np.int64)))
    elif end_offset < 0:
        last_idx += end_offset
    for i in range(first_idx, last_idx):
        num_votes[i] += 1
    return range1, num_votes


@_jit
def rle_voting(ranges, vote_thr=2):
    """array_utils.py:563-625 (init_index / term_index unused on the hot path)."""
    voted = np.empty((0, 2), dtype=np.int64)
    running = np.empty(0, dtype=np.int64)
    votes = np.empty(0, dtype=np.int64)
    for i in range(len(ranges) - 1):
        range1 = ranges[i]
        range2 = ranges[i + 1]
        if running.shape[0] == 0:
            running = range1
            votes = np.ones(range1[1] - range1[0], dtype=np.int64)
        if running[1] < range2[0]:
            voted = np.vstack((voted, split_range_by_votes(running, votes, vote_thr)))
            running = np.empty(0, dtype=np.int64)
            votes = np.empty(0, dtype=np.int64)
        else:
            running, votes = extend_range(running, range2, votes)
    if running.shape[0] != 0:
        voted = np.vstack((voted, split_range_by_votes(running, votes, vote_thr)))
    return voted


def concat_sort_ranges(list_of_ranges):
    """array_utils.py:649-656."""
    lst = [r for r in list_of_ranges if len(r) > 0]
    ranges = np.concatenate(lst, axis=0)
    order = np.argsort(ranges[:, 0], kind="stable")
    return ranges[order]


@_jit
def _join_ranges(ranges):
    """array_utils.py:658-691 (union of sorted ranges; touching ranges coalesce)."""
    joined = np.empty((0, 2), dtype=np.int64)
    running = np.empty(0, dtype=np.int64)
    last = np.empty(0, dtype=np.int64)
    for i in range(len(ranges) - 1):
        range1 = ranges[i]
        range2 = ranges[i + 1]
        last = range2
        if running.shape[0] == 0:
            running = range1
        if running[1] >= range2[0]:
            running[1] = max(running[1], range2[1])
        else:
            joined = np.vstack((joined, running.reshape(1, 2)))
            running = np.empty(0, dtype=np.int64)
    if running.shape[0] != 0:
        joined = np.vstack((joined, running.reshape(1, 2)))
    else:
        # the reference reads an unbound `range2` here when given a single range and raises
        if last.shape[0] == 0:
            raise ValueError("_join_ranges needs at least two ranges (reference behaviour)")
        joined = np.vstack((joined, last.reshape(1, 2)))
    return joined


def join_ranges(list_of_ranges):
    """array_utils.py:693-699."""
    lst = [r for r in list_of_ranges if len(r) > 0]
    ranges = concat_sort_ranges(lst).astype(np.int64)
    return np.array(_join_ranges(ranges))


def vote_by_ranges(list_of_ranges, vote_thr=2):
    """array_utils.py:627-639."""
    lst = [r for r in list_of_ranges if len(r) > 0]
    if vote_thr == 1:
        return join_ranges(lst)
    if len(lst) >= vote_thr:
        ranges = concat_sort_ranges(lst).astype(np.int64)
        return np.array(rle_voting(ranges, vote_thr))
    return np.array([])


@_jit
def invert_ranges(ranges, size):
    """array_utils.py:701-717."""
    inv = np.empty((0, 2), dtype=np.int64)
    if ranges[0][0] > 0:
        first = np.empty((1, 2), dtype=np.int64)
        first[0, 0] = 0
        first[0, 1] = ranges[0][0]
        inv = np.vstack((inv, first))
    for i in range(len(ranges) - 1):
        s = ranges[i][1]
        e = ranges[i + 1][0]
        if s != e:
            cur = np.empty((1, 2), dtype=np.int64)
            cur[0, 0] = s
            cur[0, 1] = e
            inv = np.vstack((inv, cur))
    if ranges[-1][1] < size:
        cur = np.empty((1, 2), dtype=np.int64)
        cur[0, 0] = ranges[-1][1]
        cur[0, 1] = size
        inv = np.vstack((inv, cur))
    return inv


def merge_rles(starts_a, runs_a, starts_b=None, runs_b=None):
    """array_utils.py:719-752."""
    lst = [np.stack([starts_a, starts_a + runs_a], axis=1)]
    if starts_b is not None and runs_b is not None:
        lst.append(np.stack([starts_b, starts_b + runs_b], axis=1))
    joined = join_ranges(lst)
    joined = joined.copy()
    joined[:, 1] = joined[:, 1] - joined[:, 0]
    return joined[:, 0], joined[:, 1]


def numpy_fill_instances(volume, instances):
    """array_utils.py:754-765: paint runs in dict order (later ids overwrite)."""
    shape = volume.shape
    flat = volume.reshape(-1)
    for iid, attrs in instances.items():
        starts = attrs["starts"]
        ends = starts + attrs["runs"]
        for s, e in zip(starts, ends):
            flat[s:e] = iid
    return flat.reshape(shape)
