"""Differential fuzzing of the oracle against the UNMODIFIED reference (/root/reference, through
oracle/ref_shim.py), in the build container: random small volumes, head noise and engine options
run through `Engine3d.infer_on_axis` x3 + `tracker_consensus` + `stack_postprocessing` of the
reference and through `oracle.pipeline` / `oracle.consensus`; trackers, stacks, consensus volumes
and instance tables must be identical. The fixtures under tests/golden pin the oracle on a few
cases; this widens the net (thousands of slices, many option combinations).

    python -m oracle.fuzz_reference [first_seed] [n_cases] [tiled]

Needs /root/reference, so it is not part of the test suite (the GPU box has no reference).
TEST INFRASTRUCTURE ONLY.
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402

ref_shim.install()

import empanada_napari_b200.synthetic as syn  # noqa: E402
from oracle import consensus as ocons, pipeline  # noqa: E402
from oracle.make_golden import MODEL_CONFIG, FakeModel, noisy_heads, scaled_heads, slices_of  # noqa: E402


def same_instances(a, b):
    if list(a.keys()) != list(b.keys()):
        return False
    for k in a:
        if tuple(int(v) for v in a[k]["box"]) != tuple(int(v) for v in b[k]["box"]):
            return False
        if not (np.array_equal(a[k]["starts"], b[k]["starts"]) and np.array_equal(a[k]["runs"], b[k]["runs"])):
            return False
    return True


def one_case(seed):
    import empanada_napari.inference as inf
    rng = np.random.default_rng(seed)
    shape = (int(rng.integers(8, 26)), int(rng.integers(20, 56)), int(rng.integers(20, 56)))
    ks = int(rng.choice([1, 3, 5]))
    noise = float(rng.choice([0.0, 0.3, 0.6, 1.0]))
    nms_kernel = int(rng.choice([3, 3, 5, 7]))
    conf = float(rng.choice([0.3, 0.5]))
    scale = int(rng.choice([1, 1, 1, 2]))
    semantic_only = bool(rng.random() < 0.15)
    stuff_config = bool(rng.random() < 0.1)
    vote = int(rng.choice([1, 2, 2, 2, 3]))
    allow_one = bool(rng.random() < 0.2)
    erosion = int(rng.choice([0, 0, 0, 1]))
    dilation = int(rng.choice([0, 0, 0, 1]))
    fill = bool(rng.random() < 0.15)
    min_size, min_extent = int(rng.choice([5, 20, 60])), int(rng.choice([1, 2, 3]))
    n_objects = int(rng.integers(3, 16))
    params = dict(shape=shape, ks=ks, noise=noise, nms_kernel=nms_kernel, conf=conf, scale=scale,
                  semantic_only=semantic_only, stuff_config=stuff_config, vote=vote, allow_one=allow_one,
                  erosion=erosion, dilation=dilation, fill=fill, min_size=min_size, min_extent=min_extent,
                  n_objects=n_objects)
    if min(shape) < ks:
        return "skipped", params
    cfg = dict(MODEL_CONFIG)
    if stuff_config:
        cfg["thing_list"] = []
    vol, lab, _ = syn.make_volume(shape, seed=seed, n_objects=n_objects, scale=1.0)
    ref_tr, ora_tr = {}, {}
    orig_loader = inf.load_model_to_device
    try:
        for axis_name, axis in (("xy", 0), ("xz", 1), ("yz", 2)):
            if scale > 1:
                heads = [scaled_heads(s, scale, noise, rng) for s in slices_of(lab, axis)]
            else:
                heads = [noisy_heads(s, 16, noise, rng) for s in slices_of(lab, axis)]
            fake = FakeModel(heads)
            inf.load_model_to_device = lambda url, device: fake
            eng = inf.Engine3d(cfg, inference_scale=scale, label_divisor=1000, median_kernel_size=ks,
                               nms_threshold=0.1, nms_kernel=nms_kernel, confidence_thr=conf, min_size=min_size,
                               min_extent=min_extent, use_gpu=False, save_panoptic=True, semantic_only=semantic_only,
                               label_erosion=erosion, label_dilation=dilation, fill_holes_in_segmentation=fill)
            stack, trs = eng.infer_on_axis(vol, axis_name)
            ostack, otrs = pipeline.infer_on_axis(
                vol, axis_name, lambda i, x: heads[i], cfg, median_kernel_size=ks, nms_kernel=nms_kernel,
                confidence_thr=conf, min_size=min_size, min_extent=min_extent, semantic_only=semantic_only,
                inference_scale=scale, label_erosion=erosion, label_dilation=dilation,
                fill_holes_in_segmentation=fill)
            if not (np.array_equal(stack, ostack) and same_instances(trs[0].instances, otrs[0].instances)):
                return f"plane {axis_name} differs", params
            ref_tr[axis_name], ora_tr[axis_name] = trs, otrs
    finally:
        inf.load_model_to_device = orig_loader
    kw = dict(label_divisor=1000, pixel_vote_thr=vote, cluster_iou_thr=0.75, allow_one_view=allow_one,
              min_size=min_size, min_extent=min_extent, dtype=np.int32)
    try:
        ref_c = [(v, i) for v, _, i in inf.tracker_consensus(ref_tr, None, cfg, **kw)]
        ref_err = None
    except Exception as e:      # the oracle must fail the same way
        ref_c, ref_err = None, type(e).__name__
    try:
        ora_c = [(v, i) for v, _, i in ocons.tracker_consensus(ora_tr, cfg, **kw)]
        ora_err = None
    except Exception as e:
        ora_c, ora_err = None, type(e).__name__
    if ref_err != ora_err:
        return f"consensus error behaviour differs ({ref_err} vs {ora_err})", params
    if ref_err is None:
        for (v, i), (ov, oi) in zip(ref_c, ora_c):
            if not (np.array_equal(v, ov) and same_instances(i, oi)):
                return "consensus differs", params
    kw2 = dict(label_divisor=1000, min_size=min_size, min_extent=min_extent, dtype=np.int32)
    for (v, _, i), (ov, _, oi) in zip(inf.stack_postprocessing({"xy": ref_tr["xy"]}, None, cfg, **kw2),
                                      ocons.stack_postprocessing({"xy": ora_tr["xy"]}, cfg, **kw2)):
        if not (np.array_equal(v, ov) and same_instances(i, oi)):
            return "stack_postprocessing differs", params
    return "ok" if ref_err is None else f"ok (both raise {ref_err})", params


def one_case_tiled(seed):
    """Tiled `Engine2d.infer` (empanada_napari/inference.py:283-318) of the reference against
    oracle/tiles.py; the tile layout is injected into both (cztile is absent)."""
    import types
    import empanada.inference.tile as ref_tile
    import empanada_napari.inference as inf
    from empanada_napari_b200.tiling import fixed_total_area_tiles_1d
    from oracle import tiles as otiles

    class FakeStrategy:
        def __init__(self, total_tile_width, total_tile_height, min_border_width):
            self.tw, self.th, self.b = total_tile_width, total_tile_height, min_border_width

        def tile_rectangle(self, rect):
            return [types.SimpleNamespace(roi=types.SimpleNamespace(x=x0, y=y0, w=sx, h=sy))
                    for x0, sx in fixed_total_area_tiles_1d(rect.w, self.tw, self.b)
                    for y0, sy in fixed_total_area_tiles_1d(rect.h, self.th, self.b)]

    def layout(shape, tile, overlap):      # the rectangles FakeStrategy hands to the reference, same order
        yr, xr = [], []
        for x0, sx in fixed_total_area_tiles_1d(shape[1], tile[1], overlap):
            for y0, sy in fixed_total_area_tiles_1d(shape[0], tile[0], overlap):
                yr.append((y0, y0 + sy))
                xr.append((x0, x0 + sx))
        return yr, xr

    ref_tile.AlmostEqualBorderFixedTotalAreaStrategy2D = FakeStrategy
    ref_tile.czrect = lambda x, y, w, h: types.SimpleNamespace(x=x, y=y, w=w, h=h)
    rng = np.random.default_rng(50000 + seed)
    shape = (int(rng.integers(140, 330)), int(rng.integers(140, 330)))
    tile_size = int(rng.choice([96, 128, 160]))
    semantic_only = bool(rng.random() < 0.25)
    scale = int(rng.choice([1, 1, 1, 2]))
    wide = bool(rng.random() < 0.3)        # an object wider than a tile: its runs wrap around tile rows
    params = dict(shape=shape, tile_size=tile_size, semantic_only=semantic_only, scale=scale, wide=wide)
    _, lab, _ = syn.make_volume((1,) + shape, seed=50000 + seed, n_objects=int(rng.integers(4, 16)), scale=1.6)
    lab = lab[0].copy()
    if wide:
        y0 = int(rng.integers(10, shape[0] - 30))
        lab[y0:y0 + int(rng.integers(6, 20)), :] = int(lab.max()) + 1
    img = np.clip(np.where(lab > 0, 70.0, 170.0) + rng.normal(0, 8.0, shape), 0, 255).astype(np.uint8)
    yr, xr = layout(shape, (min(tile_size, shape[0]), min(tile_size, shape[1])), min(128, int(tile_size * 0.1)))
    heads = []
    for (y0, y1), (x0, x1) in zip(yr, xr):
        tl = lab[y0:y1, x0:x1]
        sem, ctr, off = scaled_heads(tl, scale, 0.0, rng) if scale > 1 else noisy_heads(tl, 16, 0.0, rng)
        heads.append((sem, ctr, (off + rng.normal(0, 1.5, off.shape)).astype(np.float32)))
    fake = FakeModel(heads)
    orig_loader = inf.load_model_to_device
    inf.load_model_to_device = lambda url, device: fake
    try:
        eng = inf.Engine2d(MODEL_CONFIG, inference_scale=scale, label_divisor=1000, nms_threshold=0.1, nms_kernel=3,
                           confidence_thr=0.5, semantic_only=semantic_only, tile_size=tile_size, use_gpu=False)
        try:
            ref, ref_err = eng.infer(img).astype(np.int64), None
        except Exception as e:
            ref, ref_err = None, type(e).__name__
    finally:
        inf.load_model_to_device = orig_loader
    try:
        if any(s > tile_size for s in shape):
            ora = otiles.engine2d_infer_tiled(img, lambda t, x: heads[t], MODEL_CONFIG, tile_size,
                                              layout, nms_kernel=3,
                                              confidence_thr=0.5, semantic_only=semantic_only, inference_scale=scale).astype(np.int64)
        else:       # the image fits into one tile: the plain branch (inference.py:319-325)
            ora = pipeline.engine2d_infer(img, lambda t, x: heads[0], MODEL_CONFIG, nms_kernel=3, confidence_thr=0.5,
                                          semantic_only=semantic_only, inference_scale=scale).astype(np.int64)
        ora_err = None
    except Exception as e:
        ora, ora_err = None, type(e).__name__
    if ref_err != ora_err:
        return f"tiled: error behaviour differs ({ref_err} vs {ora_err})", params
    if ref_err is None and not np.array_equal(ref, ora):
        return "tiled: label image differs", params
    return ("ok" if ref_err is None else f"ok (both raise {ref_err})") + (" [wide object]" if wide else ""), params


def main():
    first = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    count = int(sys.argv[2]) if len(sys.argv) > 2 else 50
    global one_case
    if len(sys.argv) > 3 and sys.argv[3] == "tiled":
        one_case = one_case_tiled
    t0 = time.time()
    tally = {}
    devnull = open(os.devnull, "w")
    for seed in range(first, first + count):
        out, sys.stdout = sys.stdout, devnull      # the reference prints progress
        try:
            res, params = one_case(seed)
        finally:
            sys.stdout = out
        tally[res] = tally.get(res, 0) + 1
        if not res.startswith("ok") and res != "skipped":
            print("seed", seed, res, params, flush=True)
    print(f"{count} cases from seed {first} in {time.time() - t0:.0f} s:", tally)


if __name__ == "__main__":
    main()
