"""CPU restatement of the reference's orthoplane consensus (empanada/consensus.py:35-469) and
its wrappers (empanada/inference/patterns.py:154-220, empanada_napari/inference.py:56-169).
TEST INFRASTRUCTURE ONLY (see oracle/post.py header).

Third-party code the reference calls and that is NOT vendored under /root/reference: networkx
(not even listed in install_requires; call sites consensus.py:57-59,100-142,176-191,267-285,427).
Its graph containers define the iteration orders the final instance ids depend on
(connected_components order, subgraph node order), so networkx itself is used here, with the
same call sequence. Pinned against the reference run in this container:
tests/golden/consensus_*.npz.
"""
from itertools import combinations

import networkx as nx
import numpy as np

from .ranges import (box_iou_pairs, merge_boxes, merge_rles, numpy_fill_instances, rle_iou,
                     vote_by_ranges)
from .tracking import (InstanceTracker, instance_relabel, remove_pancakes, remove_small_objects)

MIN_OVERLAP = 100
MIN_IOU = 1e-2


def _avg_edge(G, c1, c2, key):
    """consensus.py:10-33."""
    w = []
    for a in c1:
        for b in c2:
            w.append(G[a][b][key] if G.has_edge(a, b) else 0)
    return sum(w) / len(w)


def create_graph_of_clusters(G, cluster_iou_thr):
    """consensus.py:35-74."""
    H = G.copy()
    for (u, v, d) in G.edges(data=True):
        if d["iou"] <= cluster_iou_thr:
            H.remove_edge(u, v)
    CG = nx.Graph()
    for i, cluster in enumerate(nx.connected_components(H)):
        CG.add_node(i, cluster=cluster)
    for n1, n2 in combinations(CG.nodes, 2):
        c1 = CG.nodes[n1]["cluster"]
        c2 = CG.nodes[n2]["cluster"]
        iw = _avg_edge(G, c1, c2, "iou")
        ow = _avg_edge(G, c1, c2, "overlap")
        if iw > MIN_IOU or ow > MIN_OVERLAP:
            CG.add_edge(n1, n2, iou=iw, overlap=ow)
    return CG


def _push(G, src, dst):
    """consensus.py:76-84."""
    G.nodes[dst]["cluster"] = G.nodes[dst]["cluster"].union(G.nodes[src]["cluster"])
    G.remove_edge(src, dst)


def merge_clusters(G):
    """consensus.py:86-142."""
    H = G.copy()
    while len(H.edges()) > 0:
        mc = sorted(H.nodes, key=lambda x: len(list(H.neighbors(x))), reverse=True)[0]
        nbrs = sorted(H.neighbors(mc), key=lambda x: len(H.nodes[x]["cluster"]), reverse=True)
        if len(H.nodes[nbrs[0]]["cluster"]) > len(H.nodes[mc]["cluster"]):
            for nb in nbrs:
                _push(H, mc, nb)
            H.remove_node(mc)
        else:
            for nb in nbrs:
                _push(H, nb, mc)
                for sn in list(H.neighbors(nb)):
                    if not H.has_edge(mc, sn):
                        H.add_edge(mc, nb, iou=H[nb][sn]["iou"])
                H.remove_node(nb)
    return H


def merge_instances(d):
    """consensus.py:144-164."""
    if len(d) < 2:
        return list(d.values())[0]
    box = starts = runs = None
    for attrs in d.values():
        if box is None:
            box, starts, runs = attrs["box"], attrs["starts"], attrs["runs"]
        else:
            box = merge_boxes(box, attrs["box"])
            starts, runs = merge_rles(starts, runs, attrs["starts"], attrs["runs"])
    return dict(box=box, starts=starts, runs=runs)


def merge_overlapping(cluster_instances):
    """consensus.py:166-195."""
    if len(cluster_instances) < 2:
        return list(cluster_instances.values())
    ids = list(cluster_instances.keys())
    g = nx.Graph()
    g.add_nodes_from(ids)
    for a, b in combinations(ids, 2):
        iou, inter = rle_iou(cluster_instances[a]["starts"], cluster_instances[a]["runs"],
                             cluster_instances[b]["starts"], cluster_instances[b]["runs"],
                             return_intersection=True)
        if iou > MIN_IOU or inter > MIN_OVERLAP:
            g.add_edge(a, b)
    out = []
    for comp in nx.connected_components(g):
        out.append(merge_instances({k: v for k, v in cluster_instances.items() if k in comp}))
    return out


def bounding_box_screening(boxes, source_indices):
    """consensus.py:197-231."""
    pairs, _, _ = box_iou_pairs(boxes)
    pairs = pairs[source_indices[pairs[:, 0]] != source_indices[pairs[:, 1]]]
    pairs = np.sort(pairs, axis=-1)
    return np.unique(pairs, axis=0)


def object_iou_graph(source_indices, labels, boxes, starts, runs):
    """consensus.py:233-287."""
    matches = bounding_box_screening(boxes, source_indices)
    g = nx.Graph()
    for n in range(len(labels)):
        g.add_node(n, box=boxes[n], starts=starts[n], runs=runs[n])
    for r1, r2 in zip(*tuple(matches.T)):
        iou, inter = rle_iou(g.nodes[r1]["starts"], g.nodes[r1]["runs"], g.nodes[r2]["starts"],
                             g.nodes[r2]["runs"], return_intersection=True)
        if iou > 0:
            g.add_edge(r1, r2, iou=iou, overlap=inter)
    return g


def merge_objects_from_trackers(trackers, pixel_vote_thr=2, cluster_iou_thr=0.75, bypass=False):
    """consensus.py:348-469."""
    n_votes = len(trackers)
    min_cluster = 1 if bypass else (n_votes // 2) + 1
    if pixel_vote_thr < min_cluster:
        cluster_iou_thr = 0
    src, labels, boxes, starts, runs = [], [], [], [], []
    for ti, tr in enumerate(trackers):
        for iid, attrs in tr.instances.items():
            src.append(ti)
            labels.append(int(iid))
            boxes.append(attrs["box"])
            starts.append(attrs["starts"])
            runs.append(attrs["runs"])
    src = np.array(src)
    labels = np.array(labels)
    boxes = np.array(boxes)
    if len(boxes) == 0:
        return {}
    graph = object_iou_graph(src, labels, boxes, starts, runs)
    instance_id = 1
    instances = {}
    for comp in nx.connected_components(graph):
        if len(comp) < min_cluster:
            continue
        cg = merge_clusters(create_graph_of_clusters(graph.subgraph(comp), cluster_iou_thr))
        cid = 1
        cinst = {}
        for node in cg.nodes:
            cluster = list(cg.nodes[node]["cluster"])
            if len(cluster) < min_cluster:
                continue
            box = graph.nodes[cluster[0]]["box"]
            for n in cluster[1:]:
                box = merge_boxes(box, graph.nodes[n]["box"])
            all_ranges = [np.stack([graph.nodes[n]["starts"],
                                    graph.nodes[n]["starts"] + graph.nodes[n]["runs"]], axis=1)
                          for n in cluster]
            voted = vote_by_ranges(all_ranges, pixel_vote_thr)
            if len(voted) > 0:
                cinst[cid] = {"box": tuple(int(x) for x in box), "starts": voted[:, 0],
                              "runs": voted[:, 1] - voted[:, 0]}
                cid += 1
        for attrs in merge_overlapping(cinst):
            instances[instance_id] = attrs
            instance_id += 1
    return instances


def merge_semantic_from_trackers(trackers, pixel_vote_thr=2):
    """consensus.py:289-346."""
    boxes, starts, runs = [], [], []
    for tr in trackers:
        assert len(tr.instances.keys()) <= 1
        for attrs in tr.instances.values():
            boxes.append(attrs["box"])
            starts.append(attrs["starts"])
            runs.append(attrs["runs"])
    if not boxes:
        return {}
    box = boxes[0]
    for b in boxes[1:]:
        box = merge_boxes(box, b)
    ranges = vote_by_ranges([np.stack([s, s + r], axis=1) for s, r in zip(starts, runs)],
                            pixel_vote_thr)
    return {1: {"box": box, "starts": ranges[:, 0], "runs": ranges[:, 1] - ranges[:, 0]}}


def get_axis_trackers_by_class(trackers, class_id):
    """patterns.py:154-166."""
    return [t for axis_trackers in trackers.values() for t in axis_trackers
            if t.class_id == class_id]


def tracker_consensus(trackers, model_config, label_divisor=1000, pixel_vote_thr=2,
                      cluster_iou_thr=0.75, allow_one_view=False, min_size=200, min_extent=4,
                      dtype=np.uint32):
    """empanada_napari/inference.py:111-169 (in-memory branch)."""
    thing_list = model_config["thing_list"]
    for class_id, class_name in model_config["class_names"].items():
        cts = get_axis_trackers_by_class(trackers, class_id)
        shape3d = cts[0].shape3d
        out = InstanceTracker(class_id, cts[0].label_divisor, shape3d, "xy")
        if class_id in thing_list:
            out.instances = merge_objects_from_trackers(cts, pixel_vote_thr, cluster_iou_thr,
                                                        allow_one_view)
            remove_small_objects(out, min_size=min_size)
            remove_pancakes(out, min_span=min_extent)
        else:
            out.instances = merge_semantic_from_trackers(cts, pixel_vote_thr)
        vol = np.zeros(shape3d, dtype=dtype)
        numpy_fill_instances(vol, out.instances)
        yield vol, class_name, out.instances


def stack_postprocessing(trackers, model_config, label_divisor=1000, min_size=200, min_extent=4,
                         dtype=np.uint32):
    """empanada_napari/inference.py:56-109 (in-memory branch)."""
    thing_list = model_config["thing_list"]
    for class_id, class_name in model_config["class_names"].items():
        ct = get_axis_trackers_by_class(trackers, class_id)[0]
        st = InstanceTracker(class_id, label_divisor, ct.shape3d, "xy")
        st.instances = instance_relabel(ct)
        if class_id in thing_list:
            remove_small_objects(st, min_size=min_size)
            remove_pancakes(st, min_span=min_extent)
        vol = np.zeros(ct.shape3d, dtype=dtype)
        numpy_fill_instances(vol, st.instances)
        yield vol, class_name, st.instances
