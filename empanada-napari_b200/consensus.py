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


def dense_volume(tracker, dev):
    """Device (D,H,W) int32 label volume of a tracker: the one `Engine3d.infer_on_axis` left on
    the GPU, or a rasterisation of the RLE (trackers loaded from JSON)."""
    vol = getattr(tracker, "_b200_dense", None)
    if vol is not None:
        return vol
    shape3d = tuple(int(s) for s in tracker.shape3d)
    vol = torch.zeros(shape3d, dtype=torch.int32, device=dev)
    flat = vol.view(-1)
    for label, attrs in tracker.instances.items():
        starts = torch.from_numpy(np.asarray(attrs["starts"], dtype=np.int64)).to(dev)
        runs = torch.from_numpy(np.asarray(attrs["runs"], dtype=np.int64)).to(dev)
        if starts.numel() == 0:
            continue
        total = int(runs.sum().item())
        rep = torch.repeat_interleave(starts - torch.cumsum(runs, 0) + runs, runs)
        idx = rep + torch.arange(total, device=dev)
        flat[idx] = int(label)
    return vol


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


# ------------------------------------------------------------------------- consensus
def merge_objects_from_trackers(trackers, pixel_vote_thr=2, cluster_iou_thr=0.75, bypass=False,
                                min_size=None, min_extent=None, on_volume_ready=None):
    """consensus.py:348-469 followed by the two tracker filters (inference.py:149-150).
    Returns (device int32 volume with the final ids painted, instances dict)."""
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
    LAST_LAUNCHES = 12  # hash clear/compact x2, pairs, vote stats, vote paint, hist, lut, runs x2
    dev = torch.device("cuda", torch.cuda.current_device())
    shape3d = tuple(int(s) for s in trackers[0].shape3d)
    n_vox = int(np.prod(shape3d))
    W = shape3d[2]
    n_votes = len(trackers)
    if n_votes > 3:
        raise _lib.B200EmpanadaError("consensus kernels take at most three planes")
    min_cluster = 1 if bypass else (n_votes // 2) + 1
    if pixel_vote_thr < min_cluster:
        cluster_iou_thr = 0

    # nodes in tracker order / dict order (consensus.py:400-406)
    node_sizes, node_boxes, luts, vols = [], [], [], []
    nid = 0
    for tr in trackers:
        labels = [int(l) for l in tr.instances.keys()]
        lut = np.zeros((max(labels) + 1) if labels else 1, dtype=np.int32)
        known = getattr(tr, "_b200_sizes", None)
        for l in labels:
            nid += 1
            lut[l] = nid
            node_sizes.append(int(known[l]) if known is not None else int(np.sum(tr.instances[l]["runs"])))
            node_boxes.append(tuple(int(v) for v in tr.instances[l]["box"]))
        luts.append(torch.from_numpy(lut).to(dev))
        vols.append(dense_volume(tr, dev))
    n_nodes = nid
    empty_vol = torch.zeros(shape3d, dtype=torch.int32, device=dev)
    if n_nodes == 0:
        return empty_vol, {}
    while len(vols) < 3:
        vols.append(None)
        luts.append(None)
    vargs = [ptr(v) for v in vols] + [ptr(l) for l in luts] + [int(l.numel()) if l is not None else 0 for l in luts]

    # pass 1: overlaps between instances of different planes
    cap = _next_pow2(max(1 << 16, 16 * n_nodes))
    while True:
        keys, vals = _hash_table(cap, dev)
        overflow = torch.zeros(1, dtype=torch.int32, device=dev)
        call("be_plane_pairs", *vargs, n_vox, W, ptr(keys), ptr(vals), cap, ptr(overflow), stream_ptr())
        if int(overflow.item()) == 0:
            break
        cap *= 4
    pa, pb, inter = _hash_items(keys, vals, cap, dev)
    _mark('plane pairs kernel')
    order = np.lexsort((pb, pa))
    sizes = np.array(node_sizes, dtype=np.int64)
    ea, eb, eit = pa[order] - 1, pb[order] - 1, inter[order]
    eiou = eit / (sizes[ea] + sizes[eb] - eit)          # float64, as `intersection / union`
    pos = eiou > 0
    ea, eb, eit, eiou = ea[pos], eb[pos], eit[pos], eiou[pos]
    _mark('host build nx graph')
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
    cands = []  # (component index, member node list, merged box)
    for ci in np.flatnonzero(comp_size >= min_cluster):
        members = node_order[node_start[ci]:node_start[ci + 1]]
        if comp_min_iou[ci] > cluster_iou_thr:
            # every edge survives the IoU cut: the component is one cluster and the cluster graph
            # has no edges, so create_graph_of_clusters / merge_clusters reduce to the identity
            clusters = [members.tolist()]
        else:
            eidx = edge_order[edge_start[ci]:edge_start[ci + 1]]
            sub = component_subgraph(members.tolist(), [(int(ea[k]), int(eb[k]), float(eiou[k]), int(eit[k])) for k in eidx], n_nodes)
            cg = merge_clusters(create_graph_of_clusters(sub, cluster_iou_thr))
            clusters = [list(cg.nodes[node]["cluster"]) for node in cg.nodes]
        for cluster in clusters:
            if len(cluster) < min_cluster:
                continue
            box = node_boxes[cluster[0]]
            for m in cluster[1:]:
                box = merge_boxes(box, node_boxes[m])
            cands.append((int(ci), cluster, box))
    _mark('host graph clustering')
    if not cands:
        return empty_vol, {}

    # membership lists: node (1-based) -> candidate ids (1-based)
    memb = [[] for _ in range(n_nodes + 2)]
    for cid, (_, cluster, _) in enumerate(cands, start=1):
        for m in cluster:
            memb[m + 1].append(cid)
    memb_off = np.zeros(n_nodes + 3, dtype=np.int32)
    memb_off[1:] = np.cumsum([len(m) for m in memb])
    memb_list = np.array([c for m in memb for c in m] or [0], dtype=np.int32)
    memb_off_d = torch.from_numpy(memb_off).to(dev)
    memb_list_d = torch.from_numpy(memb_list).to(dev)

    # pass 2: voxels claimed by each candidate, overlaps between candidates
    csize_d = torch.zeros(len(cands) + 1, dtype=torch.int32, device=dev)
    cap2 = 1 << 16
    while True:
        csize_d.zero_()
        keys, vals = _hash_table(cap2, dev)
        overflow = torch.zeros(1, dtype=torch.int32, device=dev)
        call("be_vote_stats", *vargs, n_vox, W, ptr(memb_off_d), ptr(memb_list_d),
             int(pixel_vote_thr), ptr(csize_d), ptr(keys), ptr(vals), cap2, ptr(overflow), stream_ptr())
        ov = int(overflow.item())
        if ov == 2:
            raise _lib.B200EmpanadaError("a voxel is claimed by more than 32 consensus clusters")
        if ov == 0:
            break
        cap2 *= 4
    csize = csize_d.cpu().numpy().astype(np.int64)
    _mark('vote stats kernel')
    ca, cb, cinter = _hash_items(keys, vals, cap2, dev)
    cpair = {(int(a), int(b)): int(v) for a, b, v in zip(ca, cb, cinter)}

    # merge_overlapping per component, final ids in discovery order
    cid_final = np.zeros(len(cands) + 1, dtype=np.int32)
    final_boxes = {}
    next_id = 1
    by_comp = {}
    for cid, (ci, _, _) in enumerate(cands, start=1):
        if csize[cid] > 0:  # `if len(voted_ranges) > 0`
            by_comp.setdefault(ci, []).append(cid)
    for ci in sorted(by_comp.keys()):
        ids = by_comp[ci]
        if len(ids) < 2:
            groups = [set(ids)]
        else:
            mg = nx.Graph()
            mg.add_nodes_from(ids)
            for a, b in combinations(ids, 2):
                it = cpair.get((min(a, b), max(a, b)), 0)
                iou = it / (csize[a] + csize[b] - it)
                if iou > MIN_IOU or it > MIN_OVERLAP:
                    mg.add_edge(a, b)
            groups = list(nx.connected_components(mg))
        for grp in groups:
            box = None
            for cid in ids:  # dict order of cluster_instances
                if cid in grp:
                    cid_final[cid] = next_id
                    box = cands[cid - 1][2] if box is None else merge_boxes(box, cands[cid - 1][2])
            final_boxes[next_id] = tuple(int(v) for v in box)
            next_id += 1
    n_final = next_id - 1
    if n_final == 0:
        return empty_vol, {}

    # pass 3: paint (max final id wins, as the reference's in-order fill) + multi-claim side list
    out = torch.empty(shape3d, dtype=torch.int32, device=dev)
    cid_final_d = torch.from_numpy(cid_final).to(dev)
    side_cap = 1 << 20
    while True:
        side_vox = torch.empty(side_cap, dtype=torch.int64, device=dev)
        side_id = torch.empty(side_cap, dtype=torch.int32, device=dev)
        side_count = torch.zeros(1, dtype=torch.int32, device=dev)
        call("be_vote_paint", *vargs, n_vox, W, ptr(memb_off_d), ptr(memb_list_d),
             int(pixel_vote_thr), ptr(cid_final_d), ptr(out), ptr(side_vox), ptr(side_id),
             side_cap, ptr(side_count), stream_ptr())
        n_side = int(side_count.item())
        if n_side <= side_cap:
            break
        side_cap = _next_pow2(n_side)
    _mark('host merge_overlapping + paint kernel')
    hist = torch.zeros(n_final + 1, dtype=torch.int32, device=dev)
    call("be_label_hist", ptr(out), n_vox, W, n_final + 1, ptr(hist), stream_ptr())
    fsize = hist.cpu().numpy().astype(np.int64)
    flat = out.view(-1)
    side_vox, side_id = side_vox[:n_side], side_id[:n_side]
    if n_side:
        hidden = side_id != flat[side_vox]  # claimed by id but painted with a larger id
        hv, hi = side_vox[hidden].cpu().numpy(), side_id[hidden].cpu().numpy()
        np.add.at(fsize, hi, 1)
    else:
        hv, hi = np.zeros(0, np.int64), np.zeros(0, np.int32)

    # filters (filters.py:22-56) on the tables
    keep = np.ones(n_final + 1, dtype=bool)
    keep[0] = False
    for fid in range(1, n_final + 1):
        b = final_boxes[fid]
        if min_size is not None and fsize[fid] < min_size:
            keep[fid] = False
        if min_extent is not None and any(s < min_extent for s in (b[3] - b[0], b[4] - b[1], b[5] - b[2])):
            keep[fid] = False
    keep_lut = np.where(keep, np.arange(n_final + 1), 0).astype(np.int32)
    if not keep.all():
        call("be_lut_inplace", ptr(out), n_vox, ptr(torch.from_numpy(keep_lut).to(dev)), n_final + 1, stream_ptr())
        if n_side:  # voxels whose painted instance was dropped fall back to a surviving claim
            sid = torch.from_numpy(keep_lut).to(dev)[side_id.long()]
            flat.scatter_reduce_(0, side_vox, sid, reduce="amax", include_self=True)

    # the painted volume is final from here on: let the caller start its device->host copy while
    # the run-length tables are extracted
    if on_volume_ready is not None:
        on_volume_ready(out)
    # instances: runs of the painted volume (+ hidden voxels of overlapped instances)
    _mark('hist + filters')
    labels, starts, lens = extract_runs(out)
    order = torch.argsort(labels.long(), stable=True)
    lab_s = labels[order].cpu().numpy()
    st_s = starts[order].cpu().numpy()
    ln_s = lens[order].long().cpu().numpy()
    bounds = np.searchsorted(lab_s, np.arange(1, n_final + 2))
    if len(hv):
        # voxels claimed by an instance but painted with another id (after the fix-up above):
        # one gather for all of them, then grouped by instance
        still = out.view(-1)[torch.from_numpy(hv).to(dev)].cpu().numpy() != hi
        hv, hi = hv[still], hi[still]
        o = np.argsort(hi, kind="stable")
        hv, hi = hv[o], hi[o]
    hb = np.searchsorted(hi, np.arange(1, n_final + 2))
    instances = {}
    for fid in range(1, n_final + 1):
        if not keep[fid]:
            continue
        a, b = bounds[fid - 1], bounds[fid]
        s, r = st_s[a:b], ln_s[a:b]
        extra = hv[hb[fid - 1]:hb[fid]]
        if len(extra):
            # join touching / overlapping ranges (array_utils.py:659-752 `_join_ranges`)
            rng = np.concatenate([np.stack([s, s + r], 1), np.stack([extra, extra + 1], 1)])
            rng = rng[np.argsort(rng[:, 0], kind="stable")]
            st_, en_ = rng[:, 0], rng[:, 1]
            reach = np.maximum.accumulate(en_)
            head = np.ones(len(st_), dtype=bool)
            head[1:] = st_[1:] > reach[:-1]
            hidx = np.flatnonzero(head)
            s = st_[hidx]
            r = np.maximum.reduceat(en_, hidx) - s
        instances[fid] = {"box": final_boxes[fid], "starts": s, "runs": r}
    _mark('runs + instances dict')
    return out, instances


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
