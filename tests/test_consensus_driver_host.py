"""Host side of the orthoplane consensus (`consensus.consensus_driver`: graph decisions, candidate
tables, merge_overlapping, filters, instance tables, slab joins) driven on the CPU: `NumpyShard`
re-states what the CUDA kernels of csrc/consensus_runs.cu hand to the host (pair overlaps, vote
sizes, claims, painted ids, joined ranges) with dense numpy arithmetic, so the whole driver can be
compared with the oracle (= the reference's RLE / networkx path) on random disagreeing
segmentations of the three planes. The kernels themselves are covered by the -m gpu tests."""
import numpy as np
import pytest

from oracle import consensus as ocons
from oracle.ranges import numpy_fill_instances, rle_encode
from oracle.tracking import InstanceTracker, remove_pancakes, remove_small_objects


class NumpyShard:
    """CPU stand-in for `consensus.ConsensusShard` (one z-slab [z0, z0 + dz))."""

    def __init__(self, vols, luts, z0=0):
        self.vols = [np.ascontiguousarray(v, dtype=np.int32) for v in vols]
        self.luts = luts
        self.shape = self.vols[0].shape
        self.z0 = int(z0)
        self.painted = None

    def _nodes(self):
        out = []
        for v, lut in zip(self.vols, self.luts):
            lut = np.asarray(lut)
            ok = (v > 0) & (v < len(lut))
            out.append(np.where(ok, lut[np.where(ok, v, 0)], 0).astype(np.int64).ravel())
        return out

    def pairs(self, cap, luts=None):
        if luts is not None:
            self.luts = luts
        a, b, c = self.nodes = self._nodes()
        ka, kb, kv = [], [], []
        for x, y in ((a, b), (a, c), (b, c)):
            m = (x != 0) & (y != 0)
            key, cnt = np.unique((x[m] << 32) | y[m], return_counts=True)
            ka.append(key >> 32); kb.append(key & 0xFFFFFFFF); kv.append(cnt)
        return np.concatenate(ka), np.concatenate(kb), np.concatenate(kv).astype(np.int64)

    def _claims(self, memb_off, memb_list, vote_thr):
        """Unique node triples of the slab, their voxel counts, and the candidate ids claiming
        each (a candidate claims a voxel when at least vote_thr of its nodes are members)."""
        a, b, c = self.nodes
        tri = np.stack([a, b, c], axis=1)
        tri = tri[(tri != 0).any(axis=1)]
        uniq, inv, cnt = np.unique(tri, axis=0, return_inverse=True, return_counts=True)
        claims = []
        for t in uniq.tolist():
            votes = {}
            for nd in t:
                if nd:
                    for cid in memb_list[memb_off[nd]:memb_off[nd + 1]].tolist():
                        votes[cid] = votes.get(cid, 0) + 1
            claims.append([cid for cid, v in votes.items() if v >= vote_thr])
        return uniq, cnt, claims

    def stats(self, memb_off, memb_list, vote_thr, n_cands, cap):
        self.memb, self.vote_thr = (memb_off, memb_list), int(vote_thr)
        _, cnt, claims = self._claims(memb_off, memb_list, vote_thr)
        sizes = np.zeros(n_cands + 1, dtype=np.int64)
        pair = {}
        for n, cl in zip(cnt.tolist(), claims):
            for u, cu in enumerate(cl):
                sizes[cu] += n
                for cv in cl[u + 1:]:
                    k = (min(cu, cv), max(cu, cv))
                    pair[k] = pair.get(k, 0) + n
        ks = sorted(pair)
        return (sizes, np.array([k[0] for k in ks], np.int64), np.array([k[1] for k in ks], np.int64),
                np.array([pair[k] for k in ks], np.int64))

    def final_sizes(self, cid_final, n_final):
        self.cid_final = np.asarray(cid_final)
        _, cnt, claims = self._claims(*self.memb, self.vote_thr)
        fsize = np.zeros(n_final + 1, dtype=np.int64)
        for n, cl in zip(cnt.tolist(), claims):
            for f in {int(self.cid_final[c]) for c in cl} - {0}:
                fsize[f] += n
        return fsize

    def paint(self, keep, on_volume_ready=None):
        a, b, c = self.nodes
        uniq, _, claims = self._claims(*self.memb, self.vote_thr)
        fin = {tuple(t): sorted({int(self.cid_final[c]) for c in cl} - {0}) for t, cl in zip(uniq.tolist(), claims)}
        n = a.size
        painted = np.zeros(n, dtype=np.int32)
        per_id = {}
        tri = np.stack([a, b, c], axis=1)
        nz = np.flatnonzero((tri != 0).any(axis=1))
        for i, t in zip(nz.tolist(), tri[nz].tolist()):
            ids = fin[tuple(t)]
            for f in ids:
                per_id.setdefault(f, []).append(i)
                if keep[f]:
                    painted[i] = max(painted[i], f)       # later ids overwrite earlier ones
        self.painted = painted.reshape(self.shape)
        flat0 = self.z0 * self.shape[1] * self.shape[2]
        ids, starts, lens = [], [], []
        for f in sorted(per_id):
            s, r = rle_encode(np.array(per_id[f], dtype=np.int64))
            ids.append(np.full(len(s), f, np.int32)); starts.append(s + flat0); lens.append(r)
        if not ids:
            return np.zeros(0, np.int32), np.zeros(0, np.int64), np.zeros(0, np.int64)
        return np.concatenate(ids), np.concatenate(starts).astype(np.int64), np.concatenate(lens).astype(np.int64)

    def zero(self):
        self.painted = np.zeros(self.shape, dtype=np.int32)


def disagreeing_planes(rng, shape, n_objects):
    """Three label volumes of the same scene that disagree the way the planes of a real run do:
    shifted, eroded, dropped, merged and split objects."""
    import empanada_napari_b200.synthetic as syn
    _, lab, _ = syn.make_volume(shape, seed=int(rng.integers(1 << 30)), n_objects=n_objects, scale=1.0)
    planes = []
    for p in range(3):
        v = lab.copy()
        ids = [i for i in np.unique(v) if i]
        out = np.zeros_like(v)
        for i in ids:
            m = v == i
            mode = rng.random()
            if mode < 0.12:
                continue                                        # missed in this plane
            if mode < 0.35:                                     # shifted by a voxel or two
                m = np.roll(m, (int(rng.integers(-2, 3)), int(rng.integers(-2, 3)), int(rng.integers(-2, 3))), (0, 1, 2))
            elif mode < 0.5:                                    # split along a random axis
                ax = int(rng.integers(0, 3))
                idx = np.nonzero(m)[ax]
                cut = int(np.median(idx))
                half = np.zeros_like(m)
                sl = [slice(None)] * 3
                sl[ax] = slice(cut + 1, None)
                half[tuple(sl)] = m[tuple(sl)]
                out[half] = 1000 + 100 * (p + 1) + int(i)
                m = m & ~half
            elif mode < 0.6 and len(ids) > 1:                   # merged with another object
                j = int(rng.choice([k for k in ids if k != i]))
                m = m | (v == j)
            out[m & (out == 0)] = 1000 + int(i)
        planes.append(out.astype(np.int32))
    return planes


def tracker_from_dense(vol, axis_name):
    tr = InstanceTracker(1, 1000, vol.shape, axis_name)
    flat = vol.ravel()
    order = np.argsort(flat, kind="stable")
    vals = flat[order]
    bounds = np.flatnonzero(np.r_[True, vals[1:] != vals[:-1], True])
    sizes = {}
    for a, b in zip(bounds[:-1], bounds[1:]):
        label = int(vals[a])
        if label == 0:
            continue
        idx = np.sort(order[a:b])
        zz, yy, xx = np.unravel_index(idx, vol.shape)
        starts, runs = rle_encode(idx)
        tr.instances[label] = {"box": (int(zz.min()), int(yy.min()), int(xx.min()), int(zz.max()) + 1,
                                       int(yy.max()) + 1, int(xx.max()) + 1), "starts": starts, "runs": runs}
        sizes[label] = int(b - a)
    tr.finished = True
    tr._b200_sizes = sizes
    return tr


def run_driver(vols, trackers, n_shards, vote, iou_thr, bypass, min_size, min_extent):
    from empanada_napari_b200 import consensus
    n_nodes, node_sizes, node_boxes, luts = consensus.tracker_node_tables(trackers)
    min_cluster = 1 if bypass else (len(trackers) // 2) + 1
    if vote < min_cluster:
        iou_thr = 0
    D = vols[0].shape[0]
    cuts = [round(i * D / n_shards) for i in range(n_shards + 1)]
    shards = [NumpyShard([v[a:b] for v in vols], luts, z0=a) for a, b in zip(cuts[:-1], cuts[1:])]

    def each(method, *args):
        return [getattr(sh, method)(*args) for sh in shards]

    inst = consensus.consensus_driver(each, n_shards, n_nodes, node_sizes, node_boxes, luts, vote, iou_thr,
                                      min_cluster, min_size, min_extent)
    return inst, np.concatenate([sh.painted for sh in shards])


@pytest.mark.parametrize("seed", range(10))
def test_consensus_driver_on_numpy_shards_equals_oracle(seed):
    from conftest import assert_instances_equal
    rng = np.random.default_rng(300 + seed)
    shape = (int(rng.integers(10, 20)), int(rng.integers(20, 36)), int(rng.integers(20, 36)))
    vols = disagreeing_planes(rng, shape, int(rng.integers(4, 14)))
    trackers = [tracker_from_dense(v, a) for v, a in zip(vols, ("xy", "xz", "yz"))]
    vote = int(rng.choice([1, 2, 2, 3]))
    bypass = bool(rng.random() < 0.3)
    iou_thr = float(rng.choice([0.75, 0.5, 0.9]))
    min_size, min_extent = int(rng.choice([1, 30, 100])), int(rng.choice([1, 3]))
    want = InstanceTracker(1, 1000, shape, "xy")
    want.instances = ocons.merge_objects_from_trackers(trackers, vote, iou_thr, bypass)
    remove_small_objects(want, min_size=min_size)
    remove_pancakes(want, min_span=min_extent)
    want_vol = np.zeros(shape, dtype=np.int32)
    numpy_fill_instances(want_vol, want.instances)
    for n_shards in (1, 3):
        got, painted = run_driver(vols, trackers, n_shards, vote, iou_thr, bypass, min_size, min_extent)
        assert_instances_equal(got, want.instances)
        assert np.array_equal(painted, want_vol), (seed, n_shards)
