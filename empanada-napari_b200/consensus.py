"""Orthoplane instance consensus (piece 6a): voxel arithmetic on the GPU, graph decisions on
the host. Follows empanada/consensus.py:35-469 step by step:

  object_iou_graph (:233-287)        -> be_plane_pairs on three dense label volumes
  connected components / clusters    -> networkx, same calls and insertion orders as the
  create_graph_of_clusters (:35-74)     reference (networkx is the reference's own un-vendored
  merge_clusters (:86-142)              dependency; its container iteration orders define the ids)
  vote_by_ranges (:449-460)          -> be_vote_stats / be_vote_paint (votes per voxel)
  merge_overlapping (:166-195)       -> pair table from be_vote_stats + networkx components
  filters + fill (inference.py:149-167)
"""
from itertools import combinations

import networkx as nx
import numpy as np
import torch
from scipy.sparse import coo_matrix
from scipy.sparse.csgraph import connected_components

from . import _lib
from ._lib import call, ptr, stream_ptr

MIN_OVERLAP = 100
MIN_IOU = 1e-2
LAST_PROFILE = {}
LAST_LAUNCHES = 0  # kernels launched by the last merge_objects_from_trackers call
# initial capacities of the growable device tables (None: sized from the node count); every table
# regrows on overflow, tests start them tiny to exercise that
CAPS = {"pairs": None, "votes": 1 << 16, "side": 1 << 20}
# cluster decisions of the non-trivial components in native code (0: the Python `_Graph` path)
NATIVE_CLUSTERS = True
# the reference averages edge weights with the builtin sum(), which is Neumaier-compensated on
# floats from Python 3.12 on; the native path follows the running interpreter
_COMPENSATED_SUM = __import__("sys").version_info >= (3, 12)


def merge_boxes(box1, box2):
    """array_utils.py:105-129."""
    n = len(box1)
    nd = n // 2
    return tuple(min(box1[i], box2[i]) if i < nd else max(box1[i], box2[i]) for i in range(n))


def _next_pow2(n):
    p = 1
    while p < n:
        p <<= 1
    return p


# ------------------------------------------------------------------------- graph logic (host)
def _avg_edge(G, c1, c2, key):
    w = []
    for a in c1:
        for b in c2:
            w.append(G[a][b][key] if G.has_edge(a, b) else 0)
    return sum(w) / len(w)


def create_graph_of_clusters(G, cluster_iou_thr):
    H = G.copy()
    for (u, v, d) in G.edges(data=True):
        if d["iou"] <= cluster_iou_thr:
            H.remove_edge(u, v)
    CG = nx.Graph()
    owner = {}
    for i, cluster in enumerate(nx.connected_components(H)):
        CG.add_node(i, cluster=cluster)
        for n in cluster:
            owner[n] = i
    # The reference tests every pair of clusters (combinations(CG.nodes, 2)); pairs without a
    # single edge between them average to exactly 0 and never pass the thresholds, so only pairs
    # joined by at least one edge are evaluated - in the same lexicographic order, with the same
    # per-pair arithmetic.
    linked = set()
    for u, v in G.edges():
        a, b = owner[u], owner[v]
        if a != b:
            linked.add((a, b) if a < b else (b, a))
    for n1, n2 in sorted(linked):
        c1, c2 = CG.nodes[n1]["cluster"], CG.nodes[n2]["cluster"]
        iw = _avg_edge(G, c1, c2, "iou")
        ow = _avg_edge(G, c1, c2, "overlap")
        if iw > MIN_IOU or ow > MIN_OVERLAP:
            CG.add_edge(n1, n2, iou=iw, overlap=ow)
    return CG


def merge_clusters(G):
    H = G.copy()
    while len(H.edges()) > 0:
        mc = max(H.nodes, key=H.degree)  # first node of maximal degree == sorted(..., reverse=True)[0]
        nbrs = sorted(H.neighbors(mc), key=lambda x: len(H.nodes[x]["cluster"]), reverse=True)
        if len(H.nodes[nbrs[0]]["cluster"]) > len(H.nodes[mc]["cluster"]):
            for nb in nbrs:
                H.nodes[nb]["cluster"] = H.nodes[nb]["cluster"].union(H.nodes[mc]["cluster"])
                H.remove_edge(mc, nb)
            H.remove_node(mc)
        else:
            for nb in nbrs:
                H.nodes[mc]["cluster"] = H.nodes[mc]["cluster"].union(H.nodes[nb]["cluster"])
                H.remove_edge(nb, mc)
                for sn in list(H.neighbors(nb)):
                    if not H.has_edge(mc, sn):
                        H.add_edge(mc, nb, iou=H[nb][sn]["iou"])
                H.remove_node(nb)
    return H


def component_subgraph(members, edges, n_nodes):
    """The graph `G.subgraph(comp)` of the reference's full instance graph G (consensus.py:427-431)
    for one connected component, rebuilt from tables with IDENTICAL container orders - cluster ids
    depend on them. members: ascending node ids; edges: (a, b, iou, overlap) in G's edge-insertion
    order; n_nodes: len(G). Adjacency lists keep edge-insertion order; nodes come in the iteration
    order of the set networkx builds for the view (the BFS set of `connected_components`,
    re-inserted by `show_nodes`) when that set is less than half of G (FilterAtlas), else ascending."""
    g0 = nx.Graph()
    g0.add_nodes_from(members)
    g0.add_edges_from((a, b) for a, b, _, _ in edges)
    comp_set = next(nx.connected_components(g0))
    node_seq = list(set(v for v in comp_set)) if 2 * len(comp_set) < n_nodes else list(members)
    sub = nx.Graph()
    sub.add_nodes_from(node_seq)
    for a, b, iou, overlap in edges:
        sub.add_edge(a, b, iou=iou, overlap=overlap)
    return sub


# ------------------------------------------------------------------------- the same, without networkx
class _Graph:
    """Undirected graph with EXACTLY networkx.Graph's container semantics (dict of node ->
    attribute dict, dict of node -> dict of neighbour -> shared edge attribute dict, all in
    insertion order; `copy` re-inserts edges node by node as `Graph.copy` does), so that every
    order-dependent step of create_graph_of_clusters / merge_clusters gives what networkx gives,
    at a fraction of its per-call overhead. `tests/test_consensus_host.py` checks it against the
    networkx versions above on thousands of random components."""
    __slots__ = ("node", "adj")

    def __init__(self):
        self.node, self.adj = {}, {}

    def add_node(self, n, attr=None):
        if n not in self.node:
            self.adj[n] = {}
            self.node[n] = {}
        if attr:
            self.node[n].update(attr)

    def add_edge(self, u, v, attr=None):
        if u not in self.node:
            self.adj[u] = {}
            self.node[u] = {}
        if v not in self.node:
            self.adj[v] = {}
            self.node[v] = {}
        dd = self.adj[u].get(v, {})
        if attr:
            dd.update(attr)
        self.adj[u][v] = dd
        self.adj[v][u] = dd

    def remove_edge(self, u, v):
        del self.adj[u][v]
        if u != v:
            del self.adj[v][u]

    def remove_node(self, n):
        nbrs = list(self.adj[n])
        del self.node[n]
        for u in nbrs:
            del self.adj[u][n]
        del self.adj[n]

    def copy(self):
        g = _Graph()
        for n, d in self.node.items():
            g.add_node(n, d)
        for u, nbrs in self.adj.items():
            for v, dd in nbrs.items():
                g.add_edge(u, v, dd)
        return g

    def edges(self):
        seen = set()
        for n, nbrs in self.adj.items():
            for nbr, dd in nbrs.items():
                if nbr not in seen:
                    yield n, nbr, dd
            seen.add(n)

    def n_edges(self):
        return sum(len(nbrs) + (n in nbrs) for n, nbrs in self.adj.items()) // 2

    def degree(self, n):
        nbrs = self.adj[n]
        return len(nbrs) + (n in nbrs)

    def has_edge(self, u, v):
        return u in self.adj and v in self.adj[u]

    def components(self):
        """networkx.connected_components: sets built by `_plain_bfs`, in node order."""
        seen = set()
        n = len(self.adj)
        for v in self.adj:
            if v not in seen:
                c = _plain_bfs(self.adj, n, v)
                seen.update(c)
                yield c


def _plain_bfs(adj, n, source):
    """networkx/algorithms/components/connected.py `_plain_bfs` (same insertions, same order:
    the iteration order of the returned set is part of the contract)."""
    seen = {source}
    nextlevel = [source]
    while nextlevel:
        thislevel = nextlevel
        nextlevel = []
        for v in thislevel:
            for w in adj[v]:
                if w not in seen:
                    seen.add(w)
                    nextlevel.append(w)
            if len(seen) == n:
                return seen
    return seen


def _avg_edge_fast(G, c1, c2, key):
    w = []
    adj = G.adj
    for a in c1:
        na = adj[a]
        for b in c2:
            w.append(na[b][key] if b in na else 0)
    return sum(w) / len(w)


def create_graph_of_clusters_fast(G, cluster_iou_thr):
    H = G.copy()
    for (u, v, d) in list(G.edges()):
        if d["iou"] <= cluster_iou_thr:
            H.remove_edge(u, v)
    CG = _Graph()
    owner = {}
    for i, cluster in enumerate(H.components()):
        CG.add_node(i, {"cluster": cluster})
        for n in cluster:
            owner[n] = i
    linked = set()
    for u, v, _ in G.edges():
        a, b = owner[u], owner[v]
        if a != b:
            linked.add((a, b) if a < b else (b, a))
    for n1, n2 in sorted(linked):
        c1, c2 = CG.node[n1]["cluster"], CG.node[n2]["cluster"]
        iw = _avg_edge_fast(G, c1, c2, "iou")
        ow = _avg_edge_fast(G, c1, c2, "overlap")
        if iw > MIN_IOU or ow > MIN_OVERLAP:
            CG.add_edge(n1, n2, {"iou": iw, "overlap": ow})
    return CG


def merge_clusters_fast(G):
    H = G.copy()
    node = H.node
    while H.n_edges() > 0:
        mc = max(node, key=H.degree)
        nbrs = sorted(H.adj[mc], key=lambda x: len(node[x]["cluster"]), reverse=True)
        if len(node[nbrs[0]]["cluster"]) > len(node[mc]["cluster"]):
            for nb in nbrs:
                node[nb]["cluster"] = node[nb]["cluster"].union(node[mc]["cluster"])
                H.remove_edge(mc, nb)
            H.remove_node(mc)
        else:
            for nb in nbrs:
                node[mc]["cluster"] = node[mc]["cluster"].union(node[nb]["cluster"])
                H.remove_edge(nb, mc)
                for sn in list(H.adj[nb]):
                    if not H.has_edge(mc, sn):
                        H.add_edge(mc, nb, {"iou": H.adj[nb][sn]["iou"]})
                H.remove_node(nb)
    return H


def component_subgraph_fast(members, edges, n_nodes):
    """`component_subgraph` as a `_Graph` (no networkx objects)."""
    adj = {m: {} for m in members}
    for a, b, _, _ in edges:
        adj[a][b] = None
        adj[b][a] = None
    comp_set = _plain_bfs(adj, n_nodes, members[0])
    node_seq = list(set(v for v in comp_set)) if 2 * len(comp_set) < n_nodes else list(members)
    sub = _Graph()
    for n in node_seq:
        sub.add_node(n)
    for a, b, iou, overlap in edges:
        sub.add_edge(a, b, {"iou": iou, "overlap": overlap})
    return sub


def component_clusters(members, edges, n_nodes, cluster_iou_thr):
    """Clusters (lists of nodes, in the cluster graph's node order) of one connected component
    whose edges do not all pass the IoU cut."""
    cg = merge_clusters_fast(create_graph_of_clusters_fast(component_subgraph_fast(members, edges, n_nodes), cluster_iou_thr))
    return [list(cg.node[n]["cluster"]) for n in cg.node]


def component_clusters_nx(members, edges, n_nodes, cluster_iou_thr):
    """The same through networkx (the reference's own dependency): the checker of the fast path."""
    cg = merge_clusters(create_graph_of_clusters(component_subgraph(members, edges, n_nodes), cluster_iou_thr))
    return [list(cg.nodes[n]["cluster"]) for n in cg.nodes]


# ------------------------------------------------------------------------- device helpers
def _hash_table(cap, dev):
    keys = torch.empty(cap, dtype=torch.int64, device=dev)
    vals = torch.empty(cap, dtype=torch.int32, device=dev)
    call("be_hash_clear", ptr(keys), ptr(vals), cap, stream_ptr())
    return keys, vals


def _hash_items(keys, vals, cap, dev):
    out_keys = torch.empty(cap, dtype=torch.int64, device=dev)
    out_vals = torch.empty(cap, dtype=torch.int32, device=dev)
    cursor = torch.zeros(1, dtype=torch.int32, device=dev)
    call("be_hash_compact", ptr(keys), ptr(vals), cap, ptr(out_keys), ptr(out_vals), cap,
         ptr(cursor), stream_ptr())
    n = int(cursor.item())
    k = out_keys[:n].cpu().numpy().view(np.uint64)
    v = out_vals[:n].cpu().numpy().astype(np.int64)
    return (k >> np.uint64(32)).astype(np.int64), (k & np.uint64(0xFFFFFFFF)).astype(np.int64), v


def rasterize_instances(instances, shape3d, dev):
    """`numpy_fill_instances` (array_utils.py:754-765) on the device: every instance's runs
    painted into a zero (D,H,W) int32 volume in dictionary order (later instances overwrite
    earlier ones; runs are clipped to the volume, as numpy slicing clips them)."""
    shape3d = tuple(int(s) for s in shape3d)
    vol = torch.zeros(shape3d, dtype=torch.int32, device=dev)
    flat = vol.view(-1)
    n = flat.numel()
    for label, attrs in instances.items():
        starts = torch.from_numpy(np.asarray(attrs["starts"], dtype=np.int64)).to(dev)
        runs = torch.from_numpy(np.asarray(attrs["runs"], dtype=np.int64)).to(dev)
        if starts.numel() == 0:
            continue
        total = int(runs.sum().item())
        rep = torch.repeat_interleave(starts - torch.cumsum(runs, 0) + runs, runs)
        idx = rep + torch.arange(total, device=dev)
        flat[idx[idx < n]] = int(label)
    return vol


def dense_volume(tracker, dev):
    """Device (D,H,W) int32 label volume of a tracker: the one `Engine3d.infer_on_axis` left on
    the GPU, or a rasterisation of the RLE (trackers loaded from JSON)."""
    vol = getattr(tracker, "_b200_dense", None)
    if vol is not None:
        return vol
    return rasterize_instances(tracker.instances, tracker.shape3d, dev)


def extract_runs(vol):
    """(labels, starts, lens) of maximal flat-index runs of a (D,H,W) int32 device volume."""
    dev = vol.device
    n = vol.numel()
    chunks = (n + _lib.RUN_CHUNK - 1) // _lib.RUN_CHUNK
    counts = torch.zeros(2 * (chunks + 1), dtype=torch.int32, device=dev)
    st = stream_ptr()
    call("be_runs_count", ptr(vol), n, n, ptr(counts), st)
    c2 = counts.view(2, chunks + 1)
    offsets = torch.zeros((2, chunks + 1), dtype=torch.int64, device=dev)
    torch.cumsum(c2[:, :-1], 1, out=offsets[:, 1:])
    total = int(offsets[0, -1].item())
    labels = torch.empty(total, dtype=torch.int32, device=dev)
    starts = torch.empty(total, dtype=torch.int64, device=dev)
    ends = torch.empty(total, dtype=torch.int64, device=dev)
    if total:
        call("be_runs_write", ptr(vol), n, n, ptr(offsets), ptr(labels), ptr(starts), ptr(ends),
             total, st)
    lens = (ends - starts).to(torch.int32)
    return labels, starts, lens


# ------------------------------------------------------------------------- triple runs (device)
def _scan_i32_to_i64(counts, n):
    """Exclusive scan of counts[0:n] (int32, device) into int64 offsets[0:n+1] (offsets[n] = total;
    counts must have n + 1 readable entries) on the library's scan."""
    dev = counts.device
    offsets = torch.empty(n + 1, dtype=torch.int64, device=dev)
    need = _lib.SZ(0)
    _lib.lib().be_scan_i32_to_i64(None, None, n, None, 0, need, None)
    temp = torch.empty(max(int(need.value), 1), dtype=torch.uint8, device=dev)
    call("be_scan_i32_to_i64", ptr(counts), ptr(offsets), n, ptr(temp), temp.numel(), None, stream_ptr())
    return offsets


class TripleRuns:
    """Maximal x-runs of identical (xy, xz, yz) node triples of three dense label volumes
    (csrc/consensus_runs.cu): two dense passes build them, everything afterwards is per run."""

    def __init__(self, vols, luts, shape3d, flat0=0):
        self.dev = next(v for v in vols if v is not None).device
        D, H, W = shape3d
        self.rows, self.W, self.flat0 = D * H, W, flat0
        n = D * H * W
        self.vargs = [ptr(v) for v in vols] + [ptr(l) for l in luts] + [int(l.numel()) if l is not None else 0 for l in luts]
        chunks = (n + 1023) // 1024
        st = stream_ptr()
        counts = torch.zeros(2 * (chunks + 1), dtype=torch.int32, device=self.dev)
        call("be_triple_count", *self.vargs, n, W, ptr(counts), st)
        offs = torch.cat([_scan_i32_to_i64(counts[:chunks + 1], chunks), _scan_i32_to_i64(counts[chunks + 1:], chunks)])
        self.n = int(offs[chunks].item())
        R = max(self.n, 1)
        self.row_ptr = torch.empty(self.rows + 1, dtype=torch.int32, device=self.dev)
        self.row_ptr[self.rows] = self.n
        self.yx = torch.empty((R, 2), dtype=torch.int32, device=self.dev)
        self.x1 = torch.empty(R, dtype=torch.int32, device=self.dev)
        self.abc = torch.empty((R, 4), dtype=torch.int32, device=self.dev)
        call("be_triple_write", *self.vargs, n, W, ptr(offs), ptr(self.row_ptr), ptr(self.yx), ptr(self.x1),
             ptr(self.abc), st)
        self.launches = 4
        self.rargs = (ptr(self.yx), ptr(self.x1), ptr(self.abc), self.n)

    def _hash_pass(self, name, cap, *mid):
        """Runs `name` with a growing (key, count) table until nothing overflows."""
        while True:
            keys, vals = _hash_table(cap, self.dev)
            overflow = torch.zeros(1, dtype=torch.int32, device=self.dev)
            call(name, *self.rargs, *mid, ptr(keys), ptr(vals), cap, ptr(overflow), stream_ptr())
            self.launches += 3
            ov = int(overflow.item())
            if ov == 2:
                raise _lib.B200EmpanadaError("a voxel is claimed by more than 32 consensus clusters")
            if ov == 0:
                return keys, vals, cap
            cap *= 4

    def pairs(self, cap):
        keys, vals, cap = self._hash_pass("be_triple_pairs", cap)
        return _hash_items(keys, vals, cap, self.dev)

    def stats(self, memb_off_d, memb_list_d, vote_thr, n_cands, cap):
        csize_d = torch.zeros(n_cands + 1, dtype=torch.int32, device=self.dev)
        keys, vals, cap = None, None, cap
        while True:       # sizes accumulate inside the pass: start from zero on every attempt
            csize_d.zero_()
            ktab, vtab = _hash_table(cap, self.dev)
            overflow = torch.zeros(1, dtype=torch.int32, device=self.dev)
            call("be_triple_stats", *self.rargs, ptr(memb_off_d), ptr(memb_list_d), int(vote_thr), ptr(csize_d),
                 ptr(ktab), ptr(vtab), cap, ptr(overflow), stream_ptr())
            self.launches += 3
            ov = int(overflow.item())
            if ov == 2:
                raise _lib.B200EmpanadaError("a voxel is claimed by more than 32 consensus clusters")
            if ov == 0:
                break
            cap *= 4
        ca, cb, cinter = _hash_items(ktab, vtab, cap, self.dev)
        return csize_d.cpu().numpy().astype(np.int64), ca, cb, cinter

    def final_sizes(self, memb_off_d, memb_list_d, vote_thr, cid_final_d, n_final):
        """Voxels claimed by every final instance (overlapped voxels count for each claimant) and
        the per-run number of claiming instances."""
        self.rec_count = torch.zeros(self.n + 1, dtype=torch.int32, device=self.dev)
        fsize = torch.zeros(n_final + 1, dtype=torch.int32, device=self.dev)
        call("be_triple_rec_count", *self.rargs, ptr(memb_off_d), ptr(memb_list_d), int(vote_thr),
             ptr(cid_final_d), ptr(self.rec_count), ptr(fsize), stream_ptr())
        self.launches += 1
        return fsize.cpu().numpy().astype(np.int64)

    def records(self, memb_off_d, memb_list_d, vote_thr, cid_final_d, keep_d):
        """(run_val: painted id per run, joined per-instance ranges (id, start, length) sorted by
        (id, start))."""
        dev, st = self.dev, stream_ptr()
        rec_off = _scan_i32_to_i64(self.rec_count, self.n)
        n_rec = int(rec_off[self.n].item())
        run_val = torch.zeros(max(self.n, 1), dtype=torch.int32, device=dev)
        key = torch.empty(max(n_rec, 1), dtype=torch.int64, device=dev)
        ln = torch.empty(max(n_rec, 1), dtype=torch.int32, device=dev)
        call("be_triple_rec_write", *self.rargs, self.W, int(self.flat0), ptr(memb_off_d), ptr(memb_list_d),
             int(vote_thr), ptr(cid_final_d), ptr(rec_off), ptr(keep_d), ptr(key), ptr(ln), ptr(run_val), st)
        self.launches += 2
        self.run_val = run_val
        self.rec_key, self.rec_len, self.n_rec = key, ln, n_rec
        return run_val

    def joined_ranges(self):
        """Host arrays (ids int32, starts int64, lengths int64) of the joined ranges."""
        dev, st, n = self.dev, stream_ptr(), self.n_rec
        if n == 0:
            return np.zeros(0, np.int32), np.zeros(0, np.int64), np.zeros(0, np.int64)
        key2, len2 = torch.empty_like(self.rec_key), torch.empty_like(self.rec_len)
        need = _lib.SZ(0)
        _lib.lib().be_sort_records(None, None, None, None, n, None, 0, need, None)
        temp = torch.empty(max(int(need.value), 1), dtype=torch.uint8, device=dev)
        call("be_sort_records", ptr(self.rec_key), ptr(key2), ptr(self.rec_len), ptr(len2), n, ptr(temp),
             temp.numel(), None, st)
        head = torch.empty(n, dtype=torch.int32, device=dev)
        rank = torch.empty(n, dtype=torch.int32, device=dev)
        _lib.lib().be_join_flags(None, None, n, None, None, None, 0, need, None)
        temp2 = torch.empty(max(int(need.value), 1), dtype=torch.uint8, device=dev)
        call("be_join_flags", ptr(key2), ptr(len2), n, ptr(head), ptr(rank), ptr(temp2), temp2.numel(), None, st)
        m = int(rank[n - 1].item())
        out_start = torch.empty(m, dtype=torch.int64, device=dev)
        out_len = torch.zeros(m, dtype=torch.int64, device=dev)
        out_id = torch.empty(m, dtype=torch.int32, device=dev)
        call("be_join_write", ptr(key2), ptr(len2), ptr(head), ptr(rank), n, ptr(out_start), ptr(out_len),
             ptr(out_id), st)
        self.launches += 8
        return out_id.cpu().numpy(), out_start.cpu().numpy(), out_len.cpu().numpy()

    def paint(self, out):
        """Writes every voxel of `out` (rows x W int32, contiguous) from run_val."""
        slice_off = torch.tensor([0, self.n], dtype=torch.int32, device=self.dev)
        call("be_runs_paint", ptr(self.row_ptr), ptr(self.yx), ptr(self.x1), ptr(self.run_val), ptr(slice_off),
             None, 0, 0, 0, 1, self.rows, self.W, ptr(out), 0, self.W, 1, stream_ptr())
        self.launches += 1
        return out


# ------------------------------------------------------------------------- consensus
class Candidates:
    """Candidate output instances in discovery order, as columns: `comp` (component index of the
    instance graph), the member nodes of candidate i = `members[mstart[i]:mstart[i + 1]]` (0-based
    node ids) and its merged box `boxes[i]`. Reads like the list of (component, members, box)
    tuples it replaces (len, iteration, indexing)."""

    def __init__(self, comp, mstart, members, boxes):
        self.comp = np.asarray(comp, dtype=np.int64)
        self.mstart = np.asarray(mstart, dtype=np.int64)
        self.members = np.asarray(members, dtype=np.int64)
        self.boxes = np.asarray(boxes, dtype=np.int64).reshape(-1, 6)

    def __len__(self):
        return len(self.comp)

    def __getitem__(self, i):
        return (int(self.comp[i]), self.members[self.mstart[i]:self.mstart[i + 1]].tolist(), tuple(self.boxes[i].tolist()))

    def __iter__(self):
        return (self[i] for i in range(len(self)))


def _segment_boxes(boxes, starts):
    """Merged box (min of the lower corner, max of the upper corner) of every segment of `boxes`
    that begins at `starts` (non-empty segments, ascending)."""
    out = np.zeros((len(starts), 6), dtype=np.int64)
    if len(starts):
        out[:, :3] = np.minimum.reduceat(boxes[:, :3], starts, axis=0)
        out[:, 3:] = np.maximum.reduceat(boxes[:, 3:], starts, axis=0)
    return out


def cluster_candidates(n_nodes, node_boxes, sizes, pa, pb, inter, cluster_iou_thr, min_cluster):
    """Host graph logic of consensus.py:400-447: instance graph from the pair table, connected
    components in networkx order, IoU sub-clustering, cluster merging. Returns the `Candidates`
    (component index, member nodes, merged box) in the reference's discovery order."""
    order = np.lexsort((pb, pa))
    ea, eb, eit = pa[order] - 1, pb[order] - 1, inter[order]
    eiou = eit / (sizes[ea] + sizes[eb] - eit)          # float64, as `intersection / union`
    pos = eiou > 0
    ea, eb, eit, eiou = ea[pos], eb[pos], eit[pos], eiou[pos]
    # Connected components of the instance graph. networkx yields them in order of their first
    # node in insertion order (0..n-1), i.e. by smallest node id; scipy's labelling is renumbered
    # to that order. Member order inside a component never reaches the output (boxes are
    # min/max merges, memberships are sets).
    adj = coo_matrix((np.ones(len(ea), dtype=np.int8), (ea, eb)), shape=(n_nodes, n_nodes))
    _, lab = connected_components(adj, directed=False)
    first = np.full(int(lab.max()) + 1, n_nodes, dtype=np.int64)
    np.minimum.at(first, lab, np.arange(n_nodes))
    rank = np.empty_like(first)
    rank[np.argsort(first, kind="stable")] = np.arange(len(first))
    comp_of = rank[lab]                                   # node -> component index (nx order)
    n_comp = len(first)
    comp_size = np.bincount(comp_of, minlength=n_comp)
    comp_min_iou = np.full(n_comp, np.inf)
    np.minimum.at(comp_min_iou, comp_of[ea], eiou)
    node_order = np.argsort(comp_of, kind="stable")       # nodes grouped by component, ascending ids
    node_start = np.concatenate([[0], np.cumsum(comp_size)])
    edge_order = np.argsort(comp_of[ea], kind="stable")   # edges grouped by component, lexsorted inside
    edge_start = np.concatenate([[0], np.cumsum(np.bincount(comp_of[ea], minlength=n_comp))])
    boxes = np.asarray(node_boxes, dtype=np.int64).reshape(-1, 6)
    one_cluster = comp_min_iou > cluster_iou_thr
    eligible = comp_size >= min_cluster
    # (a) every edge survives the IoU cut: the component is one cluster and the cluster graph has
    # no edges, so create_graph_of_clusters / merge_clusters reduce to the identity
    one = np.flatnonzero(eligible & one_cluster)
    one_box = _segment_boxes(boxes[node_order], node_start[:-1][comp_size > 0])
    comp_box = np.zeros((n_comp, 6), dtype=np.int64)
    comp_box[comp_size > 0] = one_box
    # (b) components with an edge at or below the IoU cut: create_graph_of_clusters + merge_clusters,
    # natively and in one call (csrc/cluster_graph.cpp; `component_clusters` is the same in Python)
    slow = np.flatnonzero(eligible & ~one_cluster)
    ncl = np.zeros(len(slow), dtype=np.int64)
    csz, cmem = np.zeros(0, dtype=np.int64), np.zeros(0, dtype=np.int64)
    if len(slow) and NATIVE_CLUSTERS:
        n_nodes_c = comp_size[slow]
        n_edges_c = (edge_start[1:] - edge_start[:-1])[slow]
        noff = np.concatenate([[0], np.cumsum(n_nodes_c)]).astype(np.int32)
        eoff = np.concatenate([[0], np.cumsum(n_edges_c)]).astype(np.int32)
        nsel = np.repeat(node_start[slow] - noff[:-1], n_nodes_c) + np.arange(noff[-1])
        esel = np.repeat(edge_start[slow] - eoff[:-1], n_edges_c) + np.arange(eoff[-1])
        nodes_c = np.ascontiguousarray(node_order[nsel], dtype=np.int32)
        eo = edge_order[esel]
        ea_c = np.ascontiguousarray(ea[eo], dtype=np.int32)
        eb_c = np.ascontiguousarray(eb[eo], dtype=np.int32)
        ei_c = np.ascontiguousarray(eiou[eo], dtype=np.float64)
        et_c = np.ascontiguousarray(eit[eo], dtype=np.int64)
        ncl32 = np.zeros(len(slow), dtype=np.int32)
        totals = np.zeros(2, dtype=np.int64)
        call("be_components_clusters", len(slow), ptr(noff), ptr(nodes_c), ptr(eoff), ptr(ea_c), ptr(eb_c), ptr(ei_c),
             ptr(et_c), int(n_nodes), float(cluster_iou_thr), float(MIN_IOU), float(MIN_OVERLAP),
             1 if _COMPENSATED_SUM else 0, ptr(ncl32), ptr(totals))
        csz32 = np.zeros(max(1, int(totals[0])), dtype=np.int32)
        cmem32 = np.zeros(max(1, int(totals[1])), dtype=np.int32)
        call("be_components_clusters_fetch", ptr(csz32), ptr(cmem32))
        ncl = ncl32.astype(np.int64)
        csz, cmem = csz32[:int(totals[0])].astype(np.int64), cmem32[:int(totals[1])].astype(np.int64)
    elif len(slow):
        ea_l, eb_l = ea[edge_order].tolist(), eb[edge_order].tolist()
        eiou_l, eit_l = eiou[edge_order].tolist(), eit[edge_order].tolist()
        node_order_l = node_order.tolist()
        sz, mem = [], []
        for k, ci in enumerate(slow.tolist()):
            members = node_order_l[node_start[ci]:node_start[ci + 1]]
            e0, e1 = int(edge_start[ci]), int(edge_start[ci + 1])
            clusters = component_clusters(members, list(zip(ea_l[e0:e1], eb_l[e0:e1], eiou_l[e0:e1], eit_l[e0:e1])),
                                          n_nodes, cluster_iou_thr)
            ncl[k] = len(clusters)
            for cluster in clusters:
                sz.append(len(cluster))
                mem.extend(cluster)
        csz, cmem = np.array(sz, dtype=np.int64), np.array(mem, dtype=np.int64)
    cstart = np.concatenate([[0], np.cumsum(csz)]).astype(np.int64)
    big = csz >= min_cluster                               # `if len(cluster) < min_cluster: continue`
    cl_comp = np.repeat(slow, ncl)[big]
    cl_box = _segment_boxes(boxes[cmem], cstart[:-1])[big] if len(csz) else np.zeros((0, 6), dtype=np.int64)
    # both kinds in discovery order: components ascending (a component is of one kind only),
    # clusters of a component in the order the cluster graph yields them
    comp_all = np.concatenate([one, cl_comp])
    src_start = np.concatenate([node_start[one], len(node_order) + cstart[:-1][big]])
    length = np.concatenate([comp_size[one], csz[big]])
    box_all = np.concatenate([comp_box[one], cl_box])
    sel = np.argsort(comp_all, kind="stable")
    comp_all, src_start, length, box_all = comp_all[sel], src_start[sel], length[sel], box_all[sel]
    mstart = np.concatenate([[0], np.cumsum(length)]).astype(np.int64)
    src = np.concatenate([node_order, cmem])
    gather = np.repeat(src_start - mstart[:-1], length) + np.arange(mstart[-1])
    return Candidates(comp_all, mstart, src[gather] if len(gather) else np.zeros(0, np.int64), box_all)


def merge_overlapping_candidates(cands, csize, ca, cb, cinter):
    """merge_overlapping per component (consensus.py:144-195,461-466): candidates of one component
    whose voted voxel sets overlap (IoU > 0.01 or > 100 voxels) become one instance; final ids
    1..n in discovery order. `csize[cid]`: voted voxels of candidate cid (1-based); (ca, cb,
    cinter): voxels claimed by both candidates of a pair. Returns (cid -> final id [n_cands + 1],
    {final id: box}).

    The reference walks the components in order and, inside one, networkx's connected components
    of the overlap graph over the candidates that kept a voxel (`if len(voted_ranges) > 0`), i.e.
    groups ordered by their smallest candidate id - candidate ids ascend with the component, so
    the final ids are the ranks of the groups' smallest ids. Boxes are min / max merges."""
    n = len(cands)
    cid_final = np.zeros(n + 1, dtype=np.int32)
    csize = np.asarray(csize, dtype=np.int64)
    live = np.flatnonzero(csize[1:n + 1] > 0)              # 0-based candidate indices
    if len(live) == 0:
        return cid_final, {}
    ca, cb, it = (np.asarray(v, dtype=np.int64) for v in (ca, cb, cinter))
    ok = (ca >= 1) & (cb >= 1) & (ca <= n) & (cb <= n) & (it > 0)
    ca, cb, it = ca[ok], cb[ok], it[ok]
    iou = it / (csize[ca] + csize[cb] - it)
    edge = ((iou > MIN_IOU) | (it > MIN_OVERLAP)) & (cands.comp[ca - 1] == cands.comp[cb - 1])
    adj = coo_matrix((np.ones(int(edge.sum()), dtype=np.int8), (ca[edge] - 1, cb[edge] - 1)), shape=(n, n))
    _, lab = connected_components(adj, directed=False)
    gmin = np.full(int(lab.max()) + 1, n, dtype=np.int64)   # smallest live candidate of every group
    np.minimum.at(gmin, lab[live], live)
    heads = np.unique(gmin[lab[live]])                      # ascending = discovery order
    fid = np.searchsorted(heads, gmin[lab[live]]) + 1
    cid_final[live + 1] = fid
    by_fid = np.argsort(fid, kind="stable")
    fb = _segment_boxes(cands.boxes[live][by_fid], np.searchsorted(fid[by_fid], np.arange(1, len(heads) + 1)))
    return cid_final, {i + 1: tuple(b) for i, b in enumerate(fb.tolist())}


def membership_tables(cands, n_nodes):
    """node (1-based) -> candidate ids (1-based) as CSR host arrays (ids ascending per node)."""
    sizes = cands.mstart[1:] - cands.mstart[:-1]
    nodes = cands.members + 1
    cids = np.repeat(np.arange(1, len(cands) + 1, dtype=np.int64), sizes)
    order = np.argsort(nodes, kind="stable")          # stable: candidate ids stay ascending per node
    memb_off = np.zeros(n_nodes + 3, dtype=np.int32)
    memb_off[1:] = np.cumsum(np.bincount(nodes, minlength=n_nodes + 2))
    memb_list = cids[order].astype(np.int32) if len(cids) else np.array([0], dtype=np.int32)
    return memb_off, memb_list


def _node_tables(trackers):
    """Nodes in tracker order / dict order (consensus.py:400-406): (n_nodes, sizes, boxes as an
    (n, 6) int64 array, per-plane label -> node LUTs as numpy arrays). Sizes come from `_b200_sizes` when the engine left them
    on the tracker, else from the run-length tables."""
    node_sizes, node_boxes, luts = [], [], []
    nid = 0
    for tr in trackers:
        labels = [int(l) for l in tr.instances.keys()]
        lut = np.zeros((max(labels) + 1) if labels else 1, dtype=np.int32)
        if labels:
            lut[np.asarray(labels, dtype=np.int64)] = np.arange(nid + 1, nid + 1 + len(labels), dtype=np.int32)
            nid += len(labels)
            known = getattr(tr, "_b200_sizes", None)
            if known is not None:
                node_sizes.extend(int(known[l]) for l in labels)
            else:
                node_sizes.extend(int(np.sum(a["runs"])) for a in tr.instances.values())
            node_boxes.append(np.asarray([a["box"] for a in tr.instances.values()], dtype=np.int64).reshape(-1, 6))
        luts.append(lut)
    boxes = np.concatenate(node_boxes) if node_boxes else np.zeros((0, 6), dtype=np.int64)
    return nid, node_sizes, boxes, luts


def tracker_nodes(trackers, dev):
    """`_node_tables` + the LUTs and the label volumes on the device."""
    nid, node_sizes, node_boxes, luts = _node_tables(trackers)
    return nid, node_sizes, node_boxes, [torch.from_numpy(l).to(dev) for l in luts], [dense_volume(tr, dev) for tr in trackers]


def tracker_node_tables(trackers):
    """`tracker_nodes` without the device volumes (host tables only)."""
    return _node_tables(trackers)


def keep_mask(final_boxes, fsize, n_final, min_size, min_extent):
    """filters.remove_small_objects / remove_pancakes (filters.py:22-56) on the tables."""
    keep = np.ones(n_final + 1, dtype=bool)
    keep[0] = False
    if n_final == 0:
        return keep
    b = np.array([final_boxes[fid] for fid in range(1, n_final + 1)], dtype=np.int64).reshape(n_final, 6)
    if min_size is not None:
        keep[1:] &= np.asarray(fsize[1:n_final + 1]) >= min_size
    if min_extent is not None:
        keep[1:] &= ((b[:, 3:] - b[:, :3]) >= min_extent).all(axis=1)
    return keep


def instances_from_ranges(ids, starts, lens, final_boxes, keep, n_final):
    """{id: {'box', 'starts', 'runs'}} from joined ranges sorted by (id, start)."""
    bounds = np.searchsorted(ids, np.arange(1, n_final + 2))
    instances = {}
    for fid in range(1, n_final + 1):
        if not keep[fid]:
            continue
        a, b = bounds[fid - 1], bounds[fid]
        instances[fid] = {"box": final_boxes[fid], "starts": starts[a:b], "runs": lens[a:b]}
    return instances


class ConsensusShard:
    """Device side of the consensus for one z-slab [z0, z0 + dz) of the volume: the three
    planes' label slabs (dz, H, W) + label -> node tables. Every method is one phase of
    `consensus_driver`; arguments and results are small host tables (picklable), so the same
    driver runs over in-process shards (single GPU) or one shard per rank (multigpu.py)."""

    def __init__(self, vols, luts, z0=0):
        vols, luts = list(vols), list(luts)
        while len(vols) < 3:
            vols.append(None)
            luts.append(None)
        first = next(v for v in vols if v is not None)
        self.dev = first.device
        self.shape = tuple(int(x) for x in first.shape)
        self.z0 = int(z0)
        self.vols, self.luts = vols, luts
        self.runs = None
        self.painted = None

    def _lut_tensors(self, luts):
        return [None if l is None else (l if isinstance(l, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(l, dtype=np.int32)).to(self.dev))
                for l in luts]

    def pairs(self, cap, luts=None):
        """Triple runs of the slab + overlap table between nodes of different planes."""
        if luts is not None:
            self.luts = luts
        self.luts = self._lut_tensors(self.luts)
        D, H, W = self.shape
        self.runs = TripleRuns(self.vols, self.luts, self.shape, flat0=self.z0 * H * W)
        return self.runs.pairs(cap)

    def stats(self, memb_off, memb_list, vote_thr, n_cands, cap):
        self.memb = (torch.from_numpy(memb_off).to(self.dev), torch.from_numpy(memb_list).to(self.dev))
        self.vote_thr = int(vote_thr)
        return self.runs.stats(self.memb[0], self.memb[1], vote_thr, n_cands, cap)

    def final_sizes(self, cid_final, n_final):
        self.cid_final = torch.from_numpy(cid_final).to(self.dev)
        return self.runs.final_sizes(self.memb[0], self.memb[1], self.vote_thr, self.cid_final, n_final)

    def paint(self, keep, on_volume_ready=None):
        """Paints the slab (kept in `self.painted`) and returns the joined ranges (ids, starts,
        lengths) of the slab, sorted by (id, start), starts as GLOBAL flat indices."""
        keep_d = torch.from_numpy(np.ascontiguousarray(keep, dtype=np.int32)).to(self.dev)
        self.runs.records(self.memb[0], self.memb[1], self.vote_thr, self.cid_final, keep_d)
        self.painted = self.runs.paint(torch.empty(self.shape, dtype=torch.int32, device=self.dev))
        if on_volume_ready is not None:
            on_volume_ready(self.painted)
        return self.runs.joined_ranges()

    def zero(self):
        self.painted = torch.zeros(self.shape, dtype=torch.int32, device=self.dev)

    @property
    def launches(self):
        return self.runs.launches if self.runs is not None else 0


def _sum_by_pair(tables):
    """[(a, b, count)] per shard -> one table with the counts of equal (a, b) summed."""
    a = np.concatenate([t[0] for t in tables]).astype(np.int64)
    b = np.concatenate([t[1] for t in tables]).astype(np.int64)
    v = np.concatenate([t[2] for t in tables]).astype(np.int64)
    if len(tables) == 1 or a.size == 0:
        return a, b, v
    key = (a << 32) | b
    uk, inv = np.unique(key, return_inverse=True)
    return uk >> 32, uk & 0xFFFFFFFF, np.bincount(inv, weights=v, minlength=len(uk)).astype(np.int64)


def join_shard_ranges(parts, n_final):
    """Per-shard joined ranges (ids ascending, starts ascending per id; shards in ascending z
    order) -> per-id (starts, lengths) over the whole volume. Ranges that meet at a shard
    boundary (the flat index runs on from the last voxel of one slab into the first voxel of the
    next) are joined, as they are when the volume is processed in one piece."""
    bounds = [np.searchsorted(ids, np.arange(1, n_final + 2)) for ids, _, _ in parts]
    out = {}
    for fid in range(1, n_final + 1):
        ss, ll = [], []
        for (ids, starts, lens), bd in zip(parts, bounds):
            a, b = bd[fid - 1], bd[fid]
            if b > a:
                if ss and ss[-1][-1] + ll[-1][-1] == starts[a]:
                    ll[-1] = ll[-1].copy()
                    ll[-1][-1] += lens[a]
                    a += 1
                    if b == a:
                        continue
                ss.append(starts[a:b])
                ll.append(lens[a:b])
        if ss:
            out[fid] = (ss[0], ll[0]) if len(ss) == 1 else (np.concatenate(ss), np.concatenate(ll))
    return out


def consensus_driver(each, n_shards, n_nodes, node_sizes, node_boxes, luts, pixel_vote_thr, cluster_iou_thr,
                     min_cluster, min_size, min_extent, mark=lambda name: None, on_volume_ready=None):
    """Host side of consensus.py:348-469 (+ the two tracker filters, inference.py:149-150) over
    `n_shards` z-slabs. `each(method, *args)` runs `ConsensusShard.method(*args)` on every shard
    and returns the list of results in ascending z order. Returns the instances dict; the painted
    slabs stay on the shards (`ConsensusShard.painted`)."""
    tables = each("pairs", _next_pow2(CAPS["pairs"] or max(1 << 16, 16 * n_nodes)), luts)
    pa, pb, inter = _sum_by_pair(tables)
    mark('triple runs + pairs kernel')
    sizes = np.array(node_sizes, dtype=np.int64)
    cands = cluster_candidates(n_nodes, node_boxes, sizes, pa, pb, inter, cluster_iou_thr, min_cluster)
    mark('host graph clustering')
    if len(cands) == 0:
        each("zero")
        return {}
    memb_off, memb_list = membership_tables(cands, n_nodes)
    res = each("stats", memb_off, memb_list, pixel_vote_thr, len(cands), _next_pow2(CAPS["votes"]))
    csize = np.sum([r[0] for r in res], axis=0)
    ca, cb, cinter = _sum_by_pair([r[1:] for r in res])
    mark('vote stats kernel')
    cid_final, final_boxes = merge_overlapping_candidates(cands, csize, ca, cb, cinter)
    n_final = len(final_boxes)
    if n_final == 0:
        each("zero")
        return {}
    fsize = np.sum(each("final_sizes", cid_final, n_final), axis=0)
    keep = keep_mask(final_boxes, fsize, n_final, min_size, min_extent)
    mark('host merge_overlapping + sizes + filters')
    parts = each("paint", keep.astype(np.int32), on_volume_ready) if on_volume_ready is not None else each("paint", keep.astype(np.int32))
    mark('records + paint + sort + join kernels')
    if n_shards == 1:
        ids, starts, lens = parts[0]
        instances = instances_from_ranges(ids, starts, lens, final_boxes, keep, n_final)
    else:
        joined = join_shard_ranges(parts, n_final)
        instances = {fid: {"box": final_boxes[fid], "starts": joined[fid][0], "runs": joined[fid][1]}
                     for fid in range(1, n_final + 1) if keep[fid] and fid in joined}
    mark('instances dict')
    return instances


def merge_objects_from_trackers(trackers, pixel_vote_thr=2, cluster_iou_thr=0.75, bypass=False,
                                min_size=None, min_extent=None, on_volume_ready=None, z_shards=1):
    """consensus.py:348-469 followed by the two tracker filters (inference.py:149-150).
    Returns (device int32 volume with the final ids painted, instances dict). `z_shards` > 1
    processes the volume as that many z-slabs through the sharded driver (the code path of the
    multi-GPU engines, exercised on one GPU by the tests)."""
    global LAST_LAUNCHES, LAST_PROFILE
    import os as _os, time as _time
    _prof_on = _os.environ.get("B200_EMPANADA_PROFILE") == "1"
    LAST_PROFILE = {}
    _t0 = [_time.perf_counter()]

    def _mark(name):
        if _prof_on:
            torch.cuda.synchronize()
            t1 = _time.perf_counter()
            LAST_PROFILE[name] = t1 - _t0[0]
            _t0[0] = t1
    LAST_LAUNCHES = 0
    dev = torch.device("cuda", torch.cuda.current_device())
    shape3d = tuple(int(s) for s in trackers[0].shape3d)
    n_votes = len(trackers)
    if n_votes > 3:
        raise _lib.B200EmpanadaError("consensus kernels take at most three planes")
    min_cluster = 1 if bypass else (n_votes // 2) + 1
    if pixel_vote_thr < min_cluster:
        cluster_iou_thr = 0

    if any(getattr(tr, "_b200_xz_wrap", False) for tr in trackers):
        import warnings
        warnings.warn("an xz instance spans the full slice width: the reference lifts such runs unsplit "
                      "(tracker.py:80-84) and counts their self-overlaps as extra votes; this consensus votes once "
                      "per plane and may miss those voxels (DESIGN.md section 5)", RuntimeWarning, stacklevel=2)
    n_nodes, node_sizes, node_boxes, luts, vols = tracker_nodes(trackers, dev)
    if n_nodes == 0:
        return torch.zeros(shape3d, dtype=torch.int32, device=dev), {}
    D = shape3d[0]
    z_shards = max(1, min(int(z_shards), D))
    if z_shards == 1:
        shards = [ConsensusShard(vols, luts)]
    else:
        cuts = [round(i * D / z_shards) for i in range(z_shards + 1)]
        shards = [ConsensusShard([v[a:b] for v in vols], luts, z0=a) for a, b in zip(cuts[:-1], cuts[1:])]

    def each(method, *args):
        return [getattr(sh, method)(*args) for sh in shards]

    hook = on_volume_ready if z_shards == 1 else None
    instances = consensus_driver(each, len(shards), n_nodes, node_sizes, node_boxes, None, pixel_vote_thr,
                                 cluster_iou_thr, min_cluster, min_size, min_extent, _mark, hook)
    LAST_LAUNCHES = sum(sh.launches for sh in shards)
    if z_shards == 1:
        return shards[0].painted, instances
    out = torch.cat([sh.painted for sh in shards])
    if on_volume_ready is not None:
        on_volume_ready(out)
    return out, instances


def merge_semantic_from_trackers(trackers, pixel_vote_thr=2, dev=None, runs_fn=None):
    """`merge_semantic_from_trackers` (consensus.py:289-346) for a stuff class: every plane holds
    at most one label, the consensus is the set of voxels that at least `pixel_vote_thr` planes
    claim (`vote_by_ranges` on the sorted ranges = maximal flat-index runs of the voted mask;
    `join_ranges` when the threshold is 1), its box the merge of the planes' boxes, its id 1.
    Returns (device int32 volume holding 1 on the voted voxels, instances dict). `dev` / `runs_fn`
    (default: the current CUDA device and the library's run extraction) exist for the host-side
    unit test of the voting logic."""
    dev = torch.device("cuda", torch.cuda.current_device()) if dev is None else dev
    runs_fn = extract_runs if runs_fn is None else runs_fn
    shape3d = tuple(int(s) for s in trackers[0].shape3d)
    boxes, vols, voters = [], [], []
    for tr in trackers:
        assert len(tr.instances.keys()) <= 1, 'Semantic classes only have 1 label!'
        for attrs in tr.instances.values():
            boxes.append(tuple(int(v) for v in attrs["box"]))
            vols.append(getattr(tr, "_b200_dense", None))
            voters.append(tr)
    if not boxes:
        return torch.zeros(shape3d, dtype=torch.int32, device=dev), {}
    box = boxes[0]
    for b in boxes[1:]:
        box = merge_boxes(box, b)
    if pixel_vote_thr > 1 and len(vols) < pixel_vote_thr:
        # `vote_by_ranges` hands back a 1-D empty array here and the reference fails on it
        # (array_utils.py:631-635, consensus.py:340)
        raise IndexError("too many indices for array: array is 1-dimensional, but 2 were indexed")
    # Votes per voxel. A plane's label volume gives one vote where it is set - except for a plane
    # whose run-length table overlaps itself (the xz row-wrap quirk of tracker.py:80-84, DESIGN.md
    # section 5; or any tracker that carries no label volume, e.g. one loaded from JSON): the
    # reference votes on the RANGES, so a voxel covered by k ranges of one plane has k votes, and a
    # range may even run past the end of the volume. Those planes are counted from their tables.
    n = int(np.prod(shape3d))
    from_tables = [getattr(tr, "_b200_dense", None) is None or getattr(tr, "_b200_xz_wrap", False) for tr in voters]
    n_ext = n
    for tr, ft in zip(voters, from_tables):
        if ft:
            for attrs in tr.instances.values():
                if len(attrs["starts"]):
                    n_ext = max(n_ext, int(np.max(np.asarray(attrs["starts"]) + np.asarray(attrs["runs"]))))
    votes = torch.zeros(n_ext, dtype=torch.int32, device=dev)
    for tr, ft, v in zip(voters, from_tables, vols):
        if not ft:
            votes[:n] += (v.reshape(-1) != 0)
            continue
        diff = torch.zeros(n_ext + 1, dtype=torch.int32, device=dev)
        for attrs in tr.instances.values():
            st = torch.from_numpy(np.ascontiguousarray(attrs["starts"], dtype=np.int64)).to(dev)
            en = st + torch.from_numpy(np.ascontiguousarray(attrs["runs"], dtype=np.int64)).to(dev)
            one = torch.ones(st.numel(), dtype=torch.int32, device=dev)
            diff.index_add_(0, st, one)
            diff.index_add_(0, en, -one)
        votes += torch.cumsum(diff, 0)[:n_ext].to(torch.int32)
    voted = (votes >= int(pixel_vote_thr)).to(torch.int32)
    # ranges = maximal flat-index runs of the voted mask (over the extended index range when a
    # table runs past the volume; the painted volume is clipped as numpy slicing clips the fill)
    _, starts, lens = runs_fn(voted.view(shape3d) if n_ext == n else voted)
    return voted[:n].view(shape3d), {1: {"box": box, "starts": starts.cpu().numpy(), "runs": lens.cpu().numpy().astype(np.int64)}}


def instance_relabel(tracker):
    """empanada_napari/inference.py:31-54."""
    out = {}
    iid = 1
    for attrs in tracker.instances.values():
        cat = np.stack([attrs["starts"], attrs["runs"]], axis=1)
        cat = cat[np.argsort(cat[:, 0], kind="stable")]
        out[iid] = {"box": attrs["box"], "starts": cat[:, 0], "runs": cat[:, 1]}
        iid += 1
    return out


def fill_volume_device(src_tracker, new_instances, dtype, to_host=None):
    """`fill_volume` (patterns.py:204-213) for the relabelled stack: tracker label -> 1..n on the
    device-resident volume; instances filtered out become 0. `to_host(vol_d, dtype)`: optional
    device->host copier (page-locked staging pool of the caller)."""
    dev = torch.device("cuda", torch.cuda.current_device())
    vol = dense_volume(src_tracker, dev).clone()
    labels = [int(l) for l in src_tracker.instances.keys()]
    lut = np.zeros((max(labels) + 1) if labels else 1, dtype=np.int32)
    for new_id, l in enumerate(labels, start=1):
        if new_id in new_instances:
            lut[l] = new_id
    call("be_lut_inplace", ptr(vol), vol.numel(), ptr(torch.from_numpy(lut).to(dev)), int(lut.shape[0]), stream_ptr())
    if to_host is not None:
        return to_host(vol, dtype)
    return vol.cpu().numpy().astype(dtype, copy=False)
