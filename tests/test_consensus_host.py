"""Host-side consensus graph logic (no GPU): the table-driven component graph must have exactly
the container orders of the reference's `graph.subgraph(comp)` (networkx view), because cluster
ids and merge decisions depend on them (empanada/consensus.py:35-142,427-431)."""
import networkx as nx
import numpy as np


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
