"""Host-side consensus graph logic (no GPU): the table-driven component graph must have exactly
the container orders of the reference's `graph.subgraph(comp)` (networkx view), because cluster
ids and merge decisions depend on them (empanada/consensus.py:35-142,427-431)."""
import networkx as nx
import numpy as np
import pytest


def test_component_subgraph_matches_networkx_view():
    from empanada_napari_b200.consensus import component_subgraph, create_graph_of_clusters, merge_clusters
    rng = np.random.default_rng(1)
    checked = 0
    for trial in range(150):
        n = int(rng.integers(5, 400))
        m = int(rng.integers(3, 300))
        ea, eb = rng.integers(0, n, m), rng.integers(0, n, m)
        keep = ea < eb
        key = np.unique(ea[keep] * 100000 + eb[keep])
        ea, eb = key // 100000, key % 100000
        iou, ov = rng.random(len(ea)), rng.integers(1, 300, len(ea))
        G = nx.Graph()
        for i in range(n):
            G.add_node(i)
        for a, b, i, o in zip(ea, eb, iou, ov):
            G.add_edge(int(a), int(b), iou=float(i), overlap=int(o))
        for comp in nx.connected_components(G):
            if len(comp) < 3:
                continue
            view = G.subgraph(comp)
            edges = [(int(ea[k]), int(eb[k]), float(iou[k]), int(ov[k])) for k in range(len(ea)) if int(ea[k]) in comp]
            sub = component_subgraph(sorted(comp), edges, n)
            assert list(sub.nodes) == list(view.nodes)
            assert [list(sub.adj[v]) for v in sub.nodes] == [list(view.adj[v]) for v in view.nodes]
            ref = merge_clusters(create_graph_of_clusters(view, 0.75))
            got = merge_clusters(create_graph_of_clusters(sub, 0.75))
            assert [sorted(ref.nodes[x]["cluster"]) for x in ref.nodes] == [sorted(got.nodes[x]["cluster"]) for x in got.nodes]
            checked += 1
    assert checked > 500


def test_join_shard_ranges_merges_runs_across_slab_boundaries():
    """Per-slab joined ranges -> whole-volume ranges: a run that continues from the last voxel of
    one z-slab into the first voxel of the next is one range, exactly as in a one-piece pass."""
    from empanada_napari_b200.consensus import join_shard_ranges
    rng = np.random.default_rng(5)
    n, n_final, cut = 4000, 6, (1000, 2500)
    vol = np.zeros(n, dtype=np.int32)
    pos = 0
    while pos < n:                                    # random runs, some crossing the cuts
        ln = int(rng.integers(1, 60))
        vol[pos:pos + ln] = rng.integers(0, n_final + 1)
        pos += ln
    vol[990:1010] = 3
    vol[2499:2501] = 5

    def ranges(a, off):
        ids, starts, lens = [], [], []
        for fid in range(1, n_final + 1):
            m = np.r_[0, (a == fid).astype(np.int8), 0]
            d = np.diff(m)
            st, en = np.flatnonzero(d == 1), np.flatnonzero(d == -1)
            ids += [fid] * len(st); starts += list(st + off); lens += list(en - st)
        return np.array(ids, np.int32), np.array(starts, np.int64), np.array(lens, np.int64)

    bounds = (0,) + cut + (n,)
    parts = [ranges(vol[a:b], a) for a, b in zip(bounds[:-1], bounds[1:])]
    got = join_shard_ranges(parts, n_final)
    ids, starts, lens = ranges(vol, 0)
    for fid in range(1, n_final + 1):
        m = ids == fid
        assert np.array_equal(got[fid][0], starts[m]) and np.array_equal(got[fid][1], lens[m]), fid


def test_networkx_free_clustering_equals_networkx():
    """`component_clusters` (plain dict graphs with networkx's container semantics) against the
    networkx path on the reference's subgraph VIEW: same clusters, same order, same member
    iteration order (the float sums of `_avg_edge` run over it), for many thresholds and shapes."""
    from empanada_napari_b200.consensus import component_clusters, create_graph_of_clusters, merge_clusters
    rng = np.random.default_rng(7)
    checked = 0
    for trial in range(400):
        n = int(rng.integers(4, 120))
        m = int(rng.integers(3, 200))
        ea, eb = rng.integers(0, n, m), rng.integers(0, n, m)
        keep = ea < eb
        key = np.unique(ea[keep] * 100000 + eb[keep])
        ea, eb = key // 100000, key % 100000
        # a mix of strong, weak and tiny overlaps so that every branch of the merge runs
        iou = np.where(rng.random(len(ea)) < 0.5, rng.uniform(0.7, 1.0, len(ea)), rng.uniform(0.0, 0.05, len(ea)))
        ov = np.where(rng.random(len(ea)) < 0.5, rng.integers(1, 90, len(ea)), rng.integers(90, 400, len(ea)))
        thr = float(rng.choice([0.0, 0.5, 0.75, 0.9]))
        G = nx.Graph()
        for i in range(n):
            G.add_node(i)
        for a, b, i, o in zip(ea, eb, iou, ov):
            G.add_edge(int(a), int(b), iou=float(i), overlap=int(o))
        for comp in nx.connected_components(G):
            if len(comp) < 2:
                continue
            view = G.subgraph(comp)
            edges = [(int(ea[k]), int(eb[k]), float(iou[k]), int(ov[k])) for k in range(len(ea)) if int(ea[k]) in comp]
            ref = merge_clusters(create_graph_of_clusters(view, thr))
            want = [list(ref.nodes[x]["cluster"]) for x in ref.nodes]
            got = component_clusters(sorted(comp), edges, n, thr)
            assert got == want, (trial, sorted(comp))
            checked += 1
    assert checked > 1500


def _random_tables(rng, n_obj, weak_links):
    """Node sizes / boxes and a pair table shaped like three planes' instances of n_obj objects."""
    sizes = rng.integers(2000, 200000, size=3 * n_obj).astype(np.int64)
    boxes = [tuple(int(v) for v in np.r_[rng.integers(0, 900, 3), rng.integers(900, 1024, 3)]) for _ in range(3 * n_obj)]
    pa, pb, inter = [], [], []
    for i in range(n_obj):
        ids = [p * n_obj + i + 1 for p in range(3)]
        for a in range(3):
            for b in range(a + 1, 3):
                if rng.random() < 0.9:
                    pa.append(ids[a]); pb.append(ids[b])
                    inter.append(int(rng.uniform(0.05, 1.0) * min(sizes[ids[a] - 1], sizes[ids[b] - 1])))
    for _ in range(weak_links):
        i = int(rng.integers(0, n_obj - 1)); j = i + int(rng.integers(1, 3))
        if j >= n_obj:
            continue
        a, b = rng.integers(0, 3, 2)
        if a == b:
            continue
        x, y = sorted((a * n_obj + i + 1, b * n_obj + j + 1))
        pa.append(x); pb.append(y); inter.append(int(rng.integers(1, 3000)))
    pa, pb, inter = np.array(pa), np.array(pb), np.array(inter)
    _, idx = np.unique((pa.astype(np.int64) << 32) | pb, return_index=True)
    return sizes, boxes, pa[idx], pb[idx], inter[idx]


def test_native_cluster_decisions_equal_python_graph():
    """csrc/cluster_graph.cpp (CPython set iteration order, networkx container orders and the
    interpreter's float summation re-stated natively) against the Python `_Graph` path, which the
    test above ties to networkx: identical candidate lists for whole instance graphs."""
    from empanada_napari_b200 import consensus as C
    rng = np.random.default_rng(11)
    total = 0
    for trial in range(12):
        n_obj = int(rng.integers(20, 400))
        sizes, boxes, pa, pb, inter = _random_tables(rng, n_obj, weak_links=int(rng.integers(0, n_obj)))
        for thr, min_cluster in ((0.75, 2), (0.0, 1), (0.5, 2)):
            C.NATIVE_CLUSTERS = False
            want = C.cluster_candidates(3 * n_obj, boxes, sizes, pa, pb, inter, thr, min_cluster)
            C.NATIVE_CLUSTERS = True
            got = C.cluster_candidates(3 * n_obj, boxes, sizes, pa, pb, inter, thr, min_cluster)
            assert len(got) == len(want)
            for g, w in zip(got, want):
                assert g[0] == w[0] and sorted(g[1]) == sorted(w[1]) and tuple(g[2]) == tuple(w[2])
            total += len(want)
    assert total > 3000


def test_native_component_clusters_over_id_ranges():
    """Single components with node ids from 0 to millions (set iteration order depends on the ids'
    low bits, table growth and perturbation): native clusters == Python `component_clusters`,
    including clusters that share nodes and components that ARE the whole graph."""
    import sys
    from empanada_napari_b200 import _lib, consensus as C
    fs = 1 if sys.version_info >= (3, 12) else 0

    def native(members, edges, n_nodes, thr):
        mem = np.ascontiguousarray(members, dtype=np.int32)
        ea = np.array([e[0] for e in edges], dtype=np.int32); eb = np.array([e[1] for e in edges], dtype=np.int32)
        ei = np.array([e[2] for e in edges], dtype=np.float64); eo = np.array([e[3] for e in edges], dtype=np.int64)
        noff, eoff = np.array([0, len(mem)], np.int32), np.array([0, len(ea)], np.int32)
        ncl, tot = np.zeros(1, np.int32), np.zeros(2, np.int64)
        _lib.call("be_components_clusters", 1, _lib.ptr(noff), _lib.ptr(mem), _lib.ptr(eoff), _lib.ptr(ea), _lib.ptr(eb),
                  _lib.ptr(ei), _lib.ptr(eo), int(n_nodes), float(thr), C.MIN_IOU, float(C.MIN_OVERLAP), fs,
                  _lib.ptr(ncl), _lib.ptr(tot))
        sz, out = np.zeros(max(1, int(tot[0])), np.int32), np.zeros(max(1, int(tot[1])), np.int32)
        _lib.call("be_components_clusters_fetch", _lib.ptr(sz), _lib.ptr(out))
        res, pos = [], 0
        for i in range(int(ncl[0])):
            res.append(sorted(out[pos:pos + sz[i]].tolist()))
            pos += sz[i]
        return res

    rng = np.random.default_rng(3)
    checked = shared = 0
    for trial in range(500):
        n, m = int(rng.integers(4, 200)), int(rng.integers(3, 300))
        base = int(rng.choice([0, 1000, 40000, 3000000]))
        ids = np.sort(rng.choice(np.arange(base, base + 5 * n), size=n, replace=False))
        a, b = rng.integers(0, n, m), rng.integers(0, n, m)
        keep = a < b
        key = np.unique(a[keep] * 100000 + b[keep])
        a, b = key // 100000, key % 100000
        iou = np.where(rng.random(len(a)) < 0.5, rng.uniform(0.7, 1.0, len(a)), rng.uniform(0.0, 0.05, len(a)))
        ov = np.where(rng.random(len(a)) < 0.5, rng.integers(1, 90, len(a)), rng.integers(90, 400, len(a)))
        thr = float(rng.choice([0.0, 0.5, 0.75, 0.9]))
        whole = rng.random() < 0.2
        G = nx.Graph()
        G.add_nodes_from(range(n))
        G.add_edges_from(zip(a.tolist(), b.tolist()))
        for comp in nx.connected_components(G):
            if len(comp) < 2 or (whole and len(comp) != n):
                continue
            n_total = n if whole else int(ids.max()) + 1 + int(rng.integers(0, 50))
            members = ids[sorted(comp)].tolist()
            edges = [(int(ids[a[k]]), int(ids[b[k]]), float(iou[k]), int(ov[k])) for k in range(len(a)) if int(a[k]) in comp]
            want = [sorted(c) for c in C.component_clusters(members, edges, n_total, thr)]
            assert native(members, edges, n_total, thr) == want, (trial, members[:6], thr)
            checked += 1
            shared += sum(len(c) for c in want) > len(members)
    assert checked > 2000 and shared > 0


def _numpy_runs(vol):
    """Maximal flat-index runs of equal non-zero value (what `consensus.extract_runs` returns)."""
    import torch
    flat = vol.reshape(-1).numpy()
    edges = np.flatnonzero(np.diff(np.concatenate([[0], flat, [0]])) != 0)
    starts, ends = edges[:-1], edges[1:]
    keep = flat[starts] != 0
    starts, ends = starts[keep], ends[keep]
    return (torch.from_numpy(flat[starts].astype(np.int32)), torch.from_numpy(starts.astype(np.int64)),
            torch.from_numpy((ends - starts).astype(np.int32)))


@pytest.mark.parametrize("tag", ["stuff_class", "stuff_class_vote3"])
def test_semantic_consensus_vote_against_reference_fixture(tag):
    """Stuff-class consensus (consensus.py:289-346): the voting / box logic of the product on the
    reference's own per-plane trackers (rasterised on the CPU, numpy run extraction standing in for
    the CUDA kernel) against the reference's consensus."""
    import os
    import torch
    from conftest import GOLDEN, assert_instances_equal, unpack_instances
    from empanada_napari_b200 import consensus
    from empanada_napari_b200.tracking import InstanceTracker
    z = np.load(os.path.join(GOLDEN, f"volume_{tag}.npz"))
    shape = tuple(int(v) for v in z["shape"])
    trackers = []
    for axis_name in ("xy", "xz", "yz"):
        tr = InstanceTracker(1, 1000, shape, axis_name)
        tr.instances = unpack_instances(z, f"{axis_name}_tr_")
        trackers.append(tr)
    vol, inst = consensus.merge_semantic_from_trackers(trackers, int(z["pixel_vote_thr"]), dev=torch.device("cpu"),
                                                       runs_fn=_numpy_runs)
    assert_instances_equal(inst, unpack_instances(z, "consensus_"))
    assert np.array_equal(vol.numpy(), z["consensus_vol"])
    # threshold 1 = join_ranges (array_utils.py:690-697) against the oracle's restatement
    from oracle import consensus as ocons
    vol1, inst1 = consensus.merge_semantic_from_trackers(trackers, 1, dev=torch.device("cpu"), runs_fn=_numpy_runs)
    assert_instances_equal(inst1, ocons.merge_semantic_from_trackers(trackers, 1))
    # no plane found anything -> no instance, empty volume
    empty = [InstanceTracker(1, 1000, shape, a) for a in ("xy", "xz", "yz")]
    vol0, inst0 = consensus.merge_semantic_from_trackers(empty, 2, dev=torch.device("cpu"), runs_fn=_numpy_runs)
    assert inst0 == {} and not vol0.any()
    # fewer voting planes than the threshold: the reference fails on the empty 1-D vote result
    with pytest.raises(IndexError):
        consensus.merge_semantic_from_trackers(trackers[:1], 2, dev=torch.device("cpu"), runs_fn=_numpy_runs)
    with pytest.raises(IndexError):
        ocons.merge_semantic_from_trackers(trackers[:1], 2)


def test_semantic_consensus_counts_self_overlapping_xz_runs_like_the_reference():
    """A stuff class that spans the full width of xz slices: the reference lifts such 2-D runs to
    3-D unsplit (tracker.py:80-84), the xz table then overlaps itself and `vote_by_ranges` counts
    every covering range as a vote. The product votes from the table for such a plane (and from
    the label volumes for the others, as on the GPU): identical to the oracle."""
    import torch
    from conftest import assert_instances_equal
    from empanada_napari_b200 import consensus
    from oracle import consensus as ocons, pipeline
    from oracle.ranges import numpy_fill_instances
    from test_gpu_z_random_options import MODEL_CONFIG, random_engine_case
    vol, heads, o = random_engine_case(6)            # stuff config, inference_scale 2
    assert o["stuff_config"]
    cfg = dict(MODEL_CONFIG, thing_list=[])
    trackers = []
    for a, axis_name in enumerate(("xy", "xz", "yz")):
        sem, ctr, off = heads[a]
        stack, trs = pipeline.infer_on_axis(
            vol, axis_name, lambda i, x: (sem[i], ctr[i], off[i]), cfg, median_kernel_size=o["ks"],
            nms_kernel=o["nms_kernel"], confidence_thr=o["conf"], min_size=o["min_size"],
            min_extent=o["min_extent"], semantic_only=o["semantic_only"], inference_scale=o["scale"])
        tr = trs[0]
        # what Engine3d leaves on a tracker: the label volume; the xz plane is flagged when a run
        # wraps around a row end and its volume is the rasterised table
        tr._b200_dense = torch.from_numpy(stack.astype(np.int32))
        covered = np.zeros(vol.size + 4096, np.int64)
        for attrs in tr.instances.values():
            for s, r in zip(attrs["starts"], attrs["runs"]):
                covered[s:s + r] += 1
        tr._b200_xz_wrap = bool(covered.max() > 1)
        trackers.append(tr)
    assert trackers[1]._b200_xz_wrap and not trackers[0]._b200_xz_wrap and not trackers[2]._b200_xz_wrap
    for thr in (1, 2, 3):
        want = ocons.merge_semantic_from_trackers(trackers, thr)
        got_vol, got = consensus.merge_semantic_from_trackers(trackers, thr, dev=torch.device("cpu"), runs_fn=_numpy_runs)
        assert_instances_equal(got, want)
        want_vol = np.zeros(vol.shape, dtype=np.int32)
        numpy_fill_instances(want_vol, want)
        assert np.array_equal(got_vol.numpy(), want_vol)
    # a plain dense vote would differ here (one vote per plane): the case is a real one
    dense_votes = sum((t._b200_dense.numpy() != 0).astype(np.int32) for t in trackers)
    assert int((dense_votes >= 2).sum()) != int(ocons.merge_semantic_from_trackers(trackers, 2)[1]["runs"].sum())
