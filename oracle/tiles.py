"""CPU restatement of the reference's tiled 2-D inference: `Tiler` / `calculate_overlap_rle`
(empanada/inference/tile.py:8-166), `merge_objects_from_tiles` / `merge_semantic_from_tiles`
(empanada/consensus.py:471-626) and the tiled branch of `Engine2d.infer`
(empanada_napari/inference.py:283-318). The tile rectangles come from the caller (`layout`):
the reference takes them from the third-party `cztile` package (absent here), so the layout is
injected exactly as oracle/make_golden.py injects it into the reference.

Pinned on tests/golden/tiled_cases.npz (tests/test_oracle_golden.py).

TEST INFRASTRUCTURE ONLY: never imported by the product path.
"""
import networkx as nx
import numpy as np

from . import post
from .consensus import object_iou_graph
from .ranges import join_ranges, merge_boxes, merge_rles, rle_ioa, rle_voting
from .tracking import pan_seg_to_rle_seg, rle_seg_to_pan_seg


def calculate_overlap_rle(yranges, xranges, image_shape):
    """tile.py:8-52: rows inside two distinct y ranges, columns inside two distinct x ranges."""
    y = np.array(rle_voting(np.unique(np.stack(yranges, axis=0), axis=0), 2))
    x = np.array(rle_voting(np.unique(np.stack(xranges, axis=0), axis=0), 2))
    if len(y) > 0:
        row_starts = y[:, 0] * image_shape[1]
        row_runs = y[:, 1] * image_shape[1] - row_starts
    else:
        row_starts, row_runs = [], []
    if len(x) > 0:
        col_ranges = np.concatenate([x + r * image_shape[1] for r in range(image_shape[0])], axis=0)
        col_starts = col_ranges[:, 0]
        col_runs = col_ranges[:, 1] - col_starts
    else:
        col_starts, col_runs = [], []
    if len(row_starts) > 0 or len(col_starts) > 0:
        return merge_rles(row_starts, row_runs, col_starts, col_runs)
    return [], []


class Tiler:
    """tile.py:54-194 with the rectangles supplied by `layout(image_shape, (th, tw), overlap)`
    -> (yranges, xranges)."""

    def __init__(self, image_shape, tile_size, overlap_width, layout):
        if isinstance(tile_size, int):
            tile_size = (tile_size, tile_size)
        assert isinstance(overlap_width, int)
        assert len(image_shape) == 2, "Tiler only works with 2D images"
        self.image_shape = image_shape
        th, tw = min(tile_size[0], image_shape[0]), min(tile_size[1], image_shape[1])
        self.yranges, self.xranges = layout(image_shape, (th, tw), overlap_width)
        self.overlap_rle = calculate_overlap_rle(self.yranges, self.xranges, image_shape)

    def __len__(self):
        return len(self.yranges)

    def __call__(self, image, i):
        return image[slice(*self.yranges[i]), slice(*self.xranges[i])]

    def translate_rle_seg(self, rle_seg, i):
        """tile.py:126-166: boxes and run STARTS move into the image frame; run lengths stay (a
        run that wraps around a tile row end therefore leaves the tile on the right)."""
        ys, ye = self.yranges[i]
        xs, xe = self.xranges[i]
        w = xe - xs
        for labels in rle_seg.values():
            for attrs in labels.values():
                b = list(attrs["box"])
                attrs["box"] = (b[0] + ys, b[1] + xs, b[2] + ys, b[3] + xs)
                starts = attrs["starts"]
                attrs["starts"] = np.ravel_multi_index((starts // w + ys, starts % w + xs), dims=self.image_shape)
        return rle_seg


def merge_semantic_from_tiles(tiles):
    """consensus.py:471-519."""
    label_id, boxes, starts, runs = None, [], [], []
    for tile_instances in tiles:
        for iid, attrs in tile_instances.items():
            if label_id is None:
                label_id = iid
            boxes.append(attrs["box"])
            starts.append(attrs["starts"])
            runs.append(attrs["runs"])
    if len(boxes) == 0:
        return {}
    box = np.array(boxes)[0]
    for b in np.array(boxes)[1:]:
        box = merge_boxes(box, b)
    rng = join_ranges([np.stack([s, s + r], axis=1) for s, r in zip(starts, runs)])
    return {label_id: {"box": box, "starts": rng[:, 0], "runs": rng[:, 1] - rng[:, 0]}}


def merge_objects_from_tiles(tiles, overlap_rle=None):
    """consensus.py:524-626: objects of different tiles that overlap are one object; a single
    detection with more than 10 % of its pixels inside the region covered by two tiles is dropped."""
    tile_idx, labels, boxes, starts, runs = [], [], [], [], []
    for ti, tile_instances in enumerate(tiles):
        for iid, attrs in tile_instances.items():
            tile_idx.append(ti)
            labels.append(int(iid))
            boxes.append(attrs["box"])
            starts.append(attrs["starts"])
            runs.append(attrs["runs"])
    tile_idx, labels, boxes = np.array(tile_idx), np.array(labels), np.array(boxes)
    if len(boxes) == 0:
        return {}
    graph = object_iou_graph(tile_idx, labels, boxes, starts, runs)
    if overlap_rle is not None:
        ov_starts, ov_runs = overlap_rle
    instance_id = int(np.min(labels))
    instances = {}
    for cluster in nx.connected_components(graph):
        cluster = list(cluster)
        box = graph.nodes[cluster[0]]["box"]
        for n in cluster[1:]:
            box = merge_boxes(box, graph.nodes[n]["box"])
        voted = join_ranges([np.stack([graph.nodes[n]["starts"], graph.nodes[n]["starts"] + graph.nodes[n]["runs"]], axis=1)
                             for n in cluster])
        if overlap_rle is not None and len(cluster) < 2 and np.any(voted):
            if rle_ioa(ov_starts, ov_runs, voted[:, 0], voted[:, 1] - voted[:, 0]) > 0.1:
                voted = []
        if np.any(voted):
            instances[instance_id] = {"box": tuple(int(v) for v in box), "starts": voted[:, 0],
                                      "runs": voted[:, 1] - voted[:, 0]}
            instance_id += 1
    return instances


def engine2d_infer_tiled(image, heads_fn, model_config, tile_size, layout, label_divisor=1000, nms_threshold=0.1,
                         nms_kernel=3, confidence_thr=0.3, stuff_area=64, void_label=0, fine_boundaries=False,
                         semantic_only=False, inference_scale=1):
    """The tiled branch of Engine2d.infer (empanada_napari/inference.py:283-318). `heads_fn(t, x)`
    stands for the model on tile t."""
    from .transforms import resize_by_factor
    labels = model_config["labels"]
    thing_list = [] if semantic_only else model_config["thing_list"]
    norms = model_config["norms"]
    tiler = Tiler(image.shape, tile_size, min(128, int(tile_size * 0.1)), layout)
    eng = post.RenderEnginePost(thing_list, label_divisor, stuff_area, void_label, nms_threshold, nms_kernel,
                                confidence_thr, None, not fine_boundaries)
    rle_segs = []
    for t in range(len(tiler)):
        tile = tiler(image, t)
        size = tile.shape
        x = post.factor_pad(post.normalize(resize_by_factor(tile, inference_scale), norms["mean"], norms["std"]),
                            model_config["padding_factor"])
        sem_logits, ctr, off = heads_fn(t, x)
        pan = eng(post.sigmoid(sem_logits), ctr, off, size, inference_scale).astype(np.int32)
        rle_segs.append(tiler.translate_rle_seg(pan_seg_to_rle_seg(pan, labels, label_divisor, thing_list), t))
    rle_seg = {}
    for label in labels:
        if label in thing_list:
            rle_seg[label] = merge_objects_from_tiles([rs[label] for rs in rle_segs], tiler.overlap_rle)
        else:
            rle_seg[label] = merge_semantic_from_tiles([rs[label] for rs in rle_segs])
    return rle_seg_to_pan_seg(rle_seg, image.shape)
