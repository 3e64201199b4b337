"""2-D tiled inference for images larger than `tile_size`:
`empanada.inference.tile.Tiler` (tile.py:54-194), `merge_objects_from_tiles` /
`merge_semantic_from_tiles` (empanada/consensus.py:471-626) and the tiled branch of
`Engine2d.infer` (empanada_napari/inference.py:283-318).

The reference gets its tile rectangles from `cztile.AlmostEqualBorderFixedTotalAreaStrategy2D`, a
third-party package that is neither vendored in the reference nor installed here: when `cztile`
is importable it is used; otherwise `fixed_total_area_tiles_1d` restates its documented contract
(tiles of constant total size, zero border at the image edges, at least `min_border` on inner
sides, non-zero borders differing by at most one pixel). PARITY UNPINNED for that layout; everything
downstream of the layout is bit-exact against the reference (tests inject the same layout).

All tiles have the same total size, so they run as ONE batch through the per-slice kernels; the
merge needs only sparse tables: per-tile components (row runs), their overlaps with components of
neighbouring tiles inside the shared rectangles, and their pixel counts inside the overlap region.
"""
import math

import numpy as np
import torch

from . import _lib
from ._lib import call, ptr, stream_ptr


def fixed_total_area_tiles_1d(length, total, min_border):
    """[(start, size)] of the tile rectangles (interior + borders) along one axis."""
    if length <= 0:
        return []
    if total >= length:
        return [(0, length)]
    if 2 * min_border >= total:
        raise AssertionError("twice the minimum border must be smaller than the tile size")
    edge_interior = total - min_border
    inner_interior = total - 2 * min_border
    n = max(int(math.ceil((length - 2 * edge_interior) / inner_interior)), 0) + 2
    n_borders = 2 * n - 2
    excess = n * total - length - n_borders * min_border
    frac = max(excess / n_borders, 0.0)
    cum = [int(np.round(frac * i)) for i in range(n_borders + 1)]
    borders = [min_border + cum[i] - cum[i - 1] for i in range(1, n_borders + 1)]     # tile0 right, tile1 left, tile1 right, ...
    out, pos = [], 0
    for k in range(n):
        left = borders[2 * k - 1] if k > 0 else 0
        right = borders[2 * k] if k < n - 1 else 0
        out.append((pos - left, total))
        pos += total - left - right
    assert pos == length and out[-1][0] + total == length
    return out


def tile_rectangles(image_shape, tile_size, overlap_width, layout=None):
    """(yranges, xranges) of every tile, in the order the reference visits them."""
    h, w = image_shape
    th, tw = min(tile_size[0], h), min(tile_size[1], w)
    if layout is not None:
        return layout(image_shape, (th, tw), overlap_width)
    try:
        from cztile.fixed_total_area_strategy_2d import AlmostEqualBorderFixedTotalAreaStrategy2D
        from cztile.tiling_strategy import Region2D
        tiler = AlmostEqualBorderFixedTotalAreaStrategy2D(total_tile_width=tw, total_tile_height=th,
                                                          min_border_width=overlap_width)
        yr, xr = [], []
        for tile in tiler.tile_rectangle(Region2D(x=0, y=0, w=w, h=h)):
            yr.append((tile.roi.y, tile.roi.y + tile.roi.h))
            xr.append((tile.roi.x, tile.roi.x + tile.roi.w))
        return yr, xr
    except ImportError:
        ys = fixed_total_area_tiles_1d(h, th, overlap_width)
        xs = fixed_total_area_tiles_1d(w, tw, overlap_width)
        yr, xr = [], []
        for x0, sx in xs:
            for y0, sy in ys:
                yr.append((y0, y0 + sy))
                xr.append((x0, x0 + sx))
        return yr, xr


class Tiler:
    """tile.py:54-124: tile rectangles + the region covered by two or more tiles."""

    def __init__(self, image_shape, tile_size=2048, overlap_width=128, layout=None):
        if isinstance(tile_size, int):
            tile_size = (tile_size, tile_size)
        assert isinstance(overlap_width, int)
        assert len(image_shape) == 2, "Tiler only works with 2D images"
        self.image_shape = tuple(int(s) for s in image_shape)
        self.tile_size = tile_size
        self.overlap_width = overlap_width
        self.yranges, self.xranges = tile_rectangles(self.image_shape, tile_size, overlap_width, layout)
        # calculate_overlap_rle (tile.py:8-52): rows inside >= 2 DISTINCT y ranges, columns inside
        # >= 2 distinct x ranges
        h, w = self.image_shape
        self.row_overlap = self._covered_twice(self.yranges, h)
        self.col_overlap = self._covered_twice(self.xranges, w)

    @staticmethod
    def _covered_twice(ranges, n):
        cover = np.zeros(n + 1, dtype=np.int64)
        for a, b in sorted(set((int(a), int(b)) for a, b in ranges)):
            cover[a] += 1
            cover[b] -= 1
        return np.cumsum(cover)[:n] >= 2

    def __len__(self):
        return len(self.yranges)

    def overlap_mask(self):
        return (self.row_overlap[:, None] | self.col_overlap[None, :]).astype(np.float64)


def _pair_counts(a, b):
    """Overlap counts between the non-zero labels of two equal-shape int32 device images
    (be_pair_overlap on the two-image stack): (label_a, label_b, pixels) numpy arrays."""
    dev = a.device
    h, w = a.shape
    two = torch.stack([a, b]).contiguous()
    cap = 1 << 12
    while True:
        keys = torch.empty(cap, dtype=torch.int64, device=dev)
        vals = torch.empty(cap, dtype=torch.int32, device=dev)
        overflow = torch.zeros(1, dtype=torch.int32, device=dev)
        call("be_hash_clear", ptr(keys), ptr(vals), cap, stream_ptr())
        call("be_pair_overlap", ptr(two), h, w, 1, 2, ptr(keys), ptr(vals), cap, ptr(overflow), stream_ptr())
        if int(overflow.item()) == 0:
            break
        cap *= 4
    out_keys = torch.empty(cap, dtype=torch.int64, device=dev)
    out_vals = torch.empty(cap, dtype=torch.int32, device=dev)
    cursor = torch.zeros(1, dtype=torch.int32, device=dev)
    call("be_hash_compact", ptr(keys), ptr(vals), cap, ptr(out_keys), ptr(out_vals), cap, ptr(cursor), stream_ptr())
    n = int(cursor.item())
    k = out_keys[:n].cpu().numpy().view(np.uint64)
    return ((k >> np.uint64(20)) & np.uint64(0xFFFFF)).astype(np.int64), (k & np.uint64(0xFFFFF)).astype(np.int64), \
        out_vals[:n].cpu().numpy().astype(np.int64)


def _warn_if_runs_wrap(tiles):
    """The reference run-length encodes every tile over its flat indices and translates only the
    START of a run into the image frame (tile.py:126-166): a run that continues from the last
    pixel of one tile row into the first pixel of the next (an object as wide as the tile) keeps
    its length and is painted past the tile's right edge instead of into the next row. That
    quirk is not reproduced here (every tile is painted where it is); say so when it applies."""
    if tiles.shape[1] > 1 and bool(((tiles[:, :-1, -1] == tiles[:, 1:, 0]) & (tiles[:, 1:, 0] != 0)).any()):
        import warnings
        warnings.warn("an object spans the full width of a tile: the reference paints the wrapped part of such "
                      "runs outside the tile (tile.py:126-166); this result keeps it inside (DESIGN.md section 7)",
                      RuntimeWarning, stacklevel=3)


def merge_tiles(post, tiler, thing, label_base, filter_overlap=True):
    """`merge_objects_from_tiles` (thing class) / `merge_semantic_from_tiles` (stuff class) on the
    batched tile post-processor `post` (run_cc done): returns the (h, w) int32 device image of the
    merged class, i.e. `rle_seg_to_pan_seg` of the merged run-length tables (rle.py:88-118)."""
    H, W = tiler.image_shape
    n = len(tiler)
    dev = post.dev
    out = torch.zeros((H, W), dtype=torch.int32, device=dev)
    n_cc = post._n_cc_host
    if int(n_cc.sum()) == 0:
        return out
    if not thing:
        # one label per class: the union of every tile's mask
        lut = np.full((n, post.cc_cap + 1), label_base, dtype=np.int32)
        lut[:, 0] = 0
        tiles = post.relabel(lut, "xy", (n, post.h, post.w))
        _warn_if_runs_wrap(tiles)
        for t in range(n):
            (y0, y1), (x0, x1) = tiler.yranges[t], tiler.xranges[t]
            torch.maximum(out[y0:y1, x0:x1], tiles[t], out=out[y0:y1, x0:x1])
        return out
    # nodes: (tile, component) in tile order, components in raster order (= the dict order of
    # pan_seg_to_rle_seg per tile, rle.py:60-83); node id of component c of tile t = off[t] + c - 1
    off = np.concatenate([[0], np.cumsum(n_cc)]).astype(np.int64)
    n_nodes = int(off[-1])
    parent = np.arange(n_nodes)

    def find(a):
        while parent[a] != a:
            parent[a] = parent[parent[a]]
            a = parent[a]
        return a

    cc = post.cc_images(0, n)
    for i in range(n):
        (ay0, ay1), (ax0, ax1) = tiler.yranges[i], tiler.xranges[i]
        for j in range(i + 1, n):
            (by0, by1), (bx0, bx1) = tiler.yranges[j], tiler.xranges[j]
            y0, y1, x0, x1 = max(ay0, by0), min(ay1, by1), max(ax0, bx0), min(ax1, bx1)
            if y0 >= y1 or x0 >= x1 or n_cc[i] == 0 or n_cc[j] == 0:
                continue
            la, lb, _ = _pair_counts(cc[i, y0 - ay0:y1 - ay0, x0 - ax0:x1 - ax0].contiguous(),
                                     cc[j, y0 - by0:y1 - by0, x0 - bx0:x1 - bx0].contiguous())
            for a, b in zip(la.tolist(), lb.tolist()):       # intersection > 0  <=>  IoU > 0: an edge
                ra, rb = find(off[i] + a - 1), find(off[j] + b - 1)
                if ra != rb:
                    parent[max(ra, rb)] = min(ra, rb)
    roots = np.array([find(a) for a in range(n_nodes)])
    # clusters in networkx's order (first node of each component); a root is its smallest member
    is_root = roots == np.arange(n_nodes)
    comp_size = np.bincount(roots, minlength=n_nodes)
    keep = np.ones(n_nodes, dtype=bool)
    if filter_overlap:
        # single-detection clusters with more than 10 % of their pixels inside the overlap region
        # are dropped as likely false positives (consensus.py:604-616)
        area, inside = _overlap_region_counts(post, tiler, off, n_nodes)
        single = is_root & (comp_size == 1)
        keep[single & (inside / np.maximum(area, 1) > 0.1)] = False
    final = np.zeros(n_nodes, dtype=np.int64)
    ids = np.flatnonzero(is_root & keep)
    final[ids] = label_base + 1 + np.arange(len(ids))     # instance_id = min(object_labels) = base + 1, then +1 each
    node_final = np.where(keep[roots], final[roots], 0)
    lut = np.zeros((n, post.cc_cap + 1), dtype=np.int32)
    for t in range(n):
        lut[t, 1:n_cc[t] + 1] = node_final[off[t]:off[t + 1]]
    tiles = post.relabel(lut, "xy", (n, post.h, post.w))
    _warn_if_runs_wrap(cc)          # per-tile components: one run-length table each (rle.py:60-83)
    for t in range(n):
        (y0, y1), (x0, x1) = tiler.yranges[t], tiler.xranges[t]
        torch.maximum(out[y0:y1, x0:x1], tiles[t], out=out[y0:y1, x0:x1])
    return out


def _overlap_region_counts(post, tiler, off, n_nodes):
    """Per node: pixels, and pixels inside the region covered by >= 2 tiles (from the row runs)."""
    r = post.runs
    total = r["total"]
    yx = r["yx"][:total].cpu().numpy()
    x1 = r["x1"][:total].cpu().numpy().astype(np.int64)
    comp = r["cc"][:total].cpu().numpy().astype(np.int64)
    so = r["slice_off"].cpu().numpy().astype(np.int64)
    tile = np.searchsorted(so, np.arange(total), side="right") - 1
    ys = np.array([a for a, _ in tiler.yranges], dtype=np.int64)[tile]
    xs = np.array([a for a, _ in tiler.xranges], dtype=np.int64)[tile]
    gy = ys + yx[:, 0]
    gx0, gx1 = xs + yx[:, 1], xs + x1
    length = gx1 - gx0
    colcum = np.concatenate([[0], np.cumsum(tiler.col_overlap.astype(np.int64))])
    inside = np.where(tiler.row_overlap[gy], length, colcum[gx1] - colcum[gx0])
    node = off[tile] + comp - 1
    return (np.bincount(node, weights=length, minlength=n_nodes).astype(np.int64),
            np.bincount(node, weights=inside, minlength=n_nodes).astype(np.int64))
