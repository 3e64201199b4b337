"""Generates tests/golden/*.npz by running the UNMODIFIED reference (/root/reference) through
oracle/ref_shim.py in the build container. The fixtures travel to the GPU box; the reference
does not. Run: `python -m oracle.make_golden` from the repo root.

TEST INFRASTRUCTURE ONLY.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")

from oracle import ref_shim  # noqa: E402

ref_shim.install()

import torch  # noqa: E402

import empanada_napari_b200.synthetic as syn  # noqa: E402

MODEL_CONFIG = {
    "class_names": {1: "mito"}, "labels": [1], "thing_list": [1], "model": "unused",
    "model_quantized": None, "padding_factor": 16, "norms": {"mean": 0.57571, "std": 0.12765},
}


def noisy_heads(label_slice, pad_to, noise, rng):
    sem, ctr, off = syn.analytic_heads(label_slice, pad_to=pad_to)
    if noise > 0:
        sem = sem + rng.normal(0, 2.5 * noise, size=sem.shape).astype(np.float32)
        ctr = ctr + rng.normal(0, 0.05 * noise, size=ctr.shape).astype(np.float32)
        off = off + rng.normal(0, 2.0 * noise, size=off.shape).astype(np.float32)
    return sem.astype(np.float32), ctr.astype(np.float32), off.astype(np.float32)


class FakeModel(torch.nn.Module):
    """Stands in for the TorchScript network: returns precomputed heads slice by slice
    (the engine calls the model exactly once per slice, in order)."""

    def __init__(self, heads):
        super().__init__()
        self.dummy = torch.nn.Parameter(torch.zeros(1))
        self.heads = heads
        self.i = 0

    def forward(self, x, render_steps: int = 2, interpolate_ins: bool = False):
        sem, ctr, off = self.heads[self.i]
        self.i += 1
        up = 2 ** (render_steps - 2)    # PointRend renders `render_steps` doublings from the /4 map
        assert tuple(v * up for v in x.shape[-2:]) == tuple(sem.shape[-2:]), (x.shape, sem.shape)
        assert tuple(v // 4 for v in x.shape[-2:]) == tuple(ctr.shape[-2:]), (x.shape, ctr.shape)
        return {"sem_logits": torch.from_numpy(sem)[None], "ctr_hmp": torch.from_numpy(ctr)[None, None],
                "offsets": torch.from_numpy(off)[None]}


def slices_of(lab, axis):
    return [np.take(lab, i, axis=axis) for i in range(lab.shape[axis])]


def pack_instances(prefix, instances, out):
    labels = np.array(list(instances.keys()), dtype=np.int64)
    out[prefix + "labels"] = labels
    out[prefix + "boxes"] = np.array([instances[k]["box"] for k in labels], dtype=np.int64).reshape(-1, 6)
    lens = np.array([len(instances[k]["starts"]) for k in labels], dtype=np.int64)
    out[prefix + "lens"] = lens
    out[prefix + "starts"] = (np.concatenate([instances[k]["starts"] for k in labels])
                              if len(labels) else np.zeros(0, np.int64)).astype(np.int64)
    out[prefix + "runs"] = (np.concatenate([instances[k]["runs"] for k in labels])
                            if len(labels) else np.zeros(0, np.int64)).astype(np.int64)


def gen_post_cases():
    """Per-slice post-processing vectors (pieces 2,3): reference engines/postprocess on heads."""
    from empanada.inference.engines import PanopticDeepLabRenderEngine
    from empanada.inference.postprocess import find_instance_center

    rng = np.random.default_rng(7)
    cases = []
    shapes = [(64, 64), (48, 80), (128, 96), (16, 16)]
    for ci in range(14):
        H, W = shapes[ci % len(shapes)]
        h4, w4 = H // 4, W // 4
        kind = ci % 7
        nms_kernel = [3, 7, 3, 5, 4, 3, 7][kind]
        thr = 0.1
        if kind == 0:      # analytic objects, few centres
            lab = syn.label_volume((1, H, W), syn.make_ellipsoids((1, H, W), 3, seed=ci, scale=1.0)
                                   * np.array([0, 1, 1, 1, 1, 1], np.float32) + np.array([0.4, 0, 0, 0, 0, 0], np.float32))[0]
            sem, ctr, off = noisy_heads(lab, None, 0.3, rng)
        elif kind == 1:    # random everything, many centres (> 20 -> chunked path)
            sem = rng.normal(0, 3, (1, H, W)).astype(np.float32)
            ctr = rng.uniform(0, 1, (h4, w4)).astype(np.float32)
            off = rng.normal(0, 6, (2, h4, w4)).astype(np.float32)
        elif kind == 2:    # no centres at all
            sem = rng.normal(0, 3, (1, H, W)).astype(np.float32)
            ctr = rng.uniform(0, 0.09, (h4, w4)).astype(np.float32)
            off = rng.normal(0, 6, (2, h4, w4)).astype(np.float32)
        elif kind == 3:    # plateaus + exact distance ties (integer offsets, quantised heat map)
            sem = rng.normal(1, 3, (1, H, W)).astype(np.float32)
            ctr = (rng.integers(0, 4, (h4, w4)) / 4.0).astype(np.float32)
            off = rng.integers(-8, 9, (2, h4, w4)).astype(np.float32)
        elif kind == 4:    # even NMS kernel
            sem = rng.normal(0, 3, (1, H, W)).astype(np.float32)
            ctr = rng.uniform(0, 1, (h4, w4)).astype(np.float32) ** 4
            off = rng.normal(0, 3, (2, h4, w4)).astype(np.float32)
        elif kind == 5:    # 1..20 centres (argmin path), zero offsets => symmetric ties
            sem = np.full((1, H, W), 3.0, np.float32)
            ctr = np.zeros((h4, w4), np.float32)
            for _ in range(6):
                ctr[rng.integers(0, h4), rng.integers(0, w4)] = 0.9
            off = np.zeros((2, h4, w4), np.float32)
        else:              # everything foreground, huge offsets (> 1e5 with K > 20 -> id 0)
            sem = np.full((1, H, W), 3.0, np.float32)
            ctr = rng.uniform(0, 1, (h4, w4)).astype(np.float32)
            off = rng.normal(0, 6, (2, h4, w4)).astype(np.float32)
            off[:, : h4 // 2] += 3e5
        conf = 0.5
        eng = PanopticDeepLabRenderEngine(
            torch.nn.Conv2d(1, 1, 1), thing_list=[1], label_divisor=1000, stuff_area=64, void_label=0,
            nms_threshold=thr, nms_kernel=nms_kernel, confidence_thr=conf, padding_factor=16,
            coarse_boundaries=True)
        prob = torch.sigmoid(torch.from_numpy(sem)[None])
        centers = find_instance_center(torch.from_numpy(ctr.copy())[None, None], thr, nms_kernel).numpy()
        cells = eng.get_instance_cells(torch.from_numpy(ctr.copy())[None, None], torch.from_numpy(off)[None], 1)
        pan = eng.postprocess(prob, cells)
        cases.append(dict(prob=prob[0].numpy(), ctr=ctr, off=off, nms_kernel=nms_kernel, thr=thr, conf=conf,
                          centers=centers, cells=cells[0, 0].numpy(), pan=pan[0].numpy()))
    out = {"n": len(cases)}
    for i, c in enumerate(cases):
        for k, v in c.items():
            out[f"c{i}_{k}"] = np.asarray(v)
    np.savez_compressed(os.path.join(GOLD, "post_cases.npz"), **out)
    print("post_cases", len(cases))


def gen_median():
    from empanada.inference.engines import _MedianQueue

    rng = np.random.default_rng(3)
    out = {}
    for ci, (ks, n) in enumerate([(3, 7), (5, 9), (1, 4), (7, 7), (3, 3)]):
        q = _MedianQueue(ks)
        x = rng.uniform(0, 1, (n, 1, 6, 5)).astype(np.float32)
        emitted = []
        for t in range(n):
            q.enqueue({"sem": torch.from_numpy(x[t].copy())[None], "t": t})
            o = q.get_next(keys=["sem"])
            if o is not None:
                emitted.append((o["t"], o["sem"][0].numpy().copy()))
        for o in q.end():
            emitted.append((o["t"], o["sem"][0].numpy().copy()))
        out[f"m{ci}_ks"] = ks
        out[f"m{ci}_x"] = x
        out[f"m{ci}_t"] = np.array([e[0] for e in emitted])
        out[f"m{ci}_y"] = np.stack([e[1] for e in emitted])
    out["n"] = 5
    np.savez_compressed(os.path.join(GOLD, "median_cases.npz"), **out)
    print("median done")


def scaled_heads(label_slice, scale, noise, rng):
    """Heads of the model for a slice that was down-sampled by `scale` (inference_scale): centre
    and offset maps on the /4 grid of the padded down-sampled slice, semantic logits rendered back
    to `scale` times that size (the FakeModel asserts both shapes)."""
    sem, _, _ = noisy_heads(label_slice, 16 * scale, noise, rng)
    _, ctr, off = noisy_heads(label_slice[::scale, ::scale], 16, noise, rng)
    return sem, ctr, off


def run_reference_volume(shape, seed, noise, ks, n_objects, min_size, min_extent, tag, pixel_vote_thr=2,
                         allow_one_view=False, semantic_only=False, inference_scale=1, erosion=0, dilation=0,
                         fill_holes=False, stuff_config=False):
    """Engine3d.infer_on_axis x3 + tracker_consensus + stack_postprocessing, unmodified.
    `stuff_config`: the model config lists no thing class (`thing_list: []`), so the class is a
    stuff class everywhere - one label per plane and `create_semantic_consensus` (a plain voxel
    vote, consensus.py:289-346) instead of the instance consensus."""
    import empanada_napari.inference as inf
    MODEL_CONFIG = dict(globals()["MODEL_CONFIG"])
    if stuff_config:
        MODEL_CONFIG["thing_list"] = []

    vol, lab, ell = syn.make_volume(shape, seed=seed, n_objects=n_objects, scale=1.0)
    out = {"shape": np.array(shape), "seed": seed, "noise": noise, "ks": ks, "n_objects": n_objects,
           "min_size": min_size, "min_extent": min_extent, "pixel_vote_thr": pixel_vote_thr,
           "allow_one_view": int(allow_one_view), "semantic_only": int(semantic_only),
           "inference_scale": int(inference_scale), "erosion": int(erosion), "dilation": int(dilation),
           "fill_holes": int(fill_holes), "stuff_config": int(stuff_config)}
    rng = np.random.default_rng(seed + 100)
    trackers = {}
    orig_loader = inf.load_model_to_device
    for axis_name, axis in (("xy", 0), ("xz", 1), ("yz", 2)):
        if inference_scale > 1:
            heads = [scaled_heads(s, inference_scale, noise, rng) for s in slices_of(lab, axis)]
        else:
            heads = [noisy_heads(s, 16, noise, rng) for s in slices_of(lab, axis)]
        out[f"{axis_name}_sem"] = np.stack([h[0] for h in heads]).astype(np.float16)  # +-4 +- noise: stored as f16
        out[f"{axis_name}_ctr"] = np.stack([h[1] for h in heads]).astype(np.float32)
        out[f"{axis_name}_off"] = np.stack([h[2] for h in heads]).astype(np.float32)
        # the reference must see exactly what the fixture stores
        heads = [(out[f"{axis_name}_sem"][i].astype(np.float32), h[1], h[2]) for i, h in enumerate(heads)]
        fake = FakeModel(heads)
        inf.load_model_to_device = lambda url, device: fake
        eng = inf.Engine3d(MODEL_CONFIG, inference_scale=inference_scale, label_divisor=1000, median_kernel_size=ks,
                           nms_threshold=0.1, nms_kernel=3, confidence_thr=0.5, min_size=min_size,
                           min_extent=min_extent, use_gpu=False, save_panoptic=True, semantic_only=semantic_only,
                           label_erosion=erosion, label_dilation=dilation, fill_holes_in_segmentation=fill_holes)
        stack, trs = eng.infer_on_axis(vol, axis_name)
        trackers[axis_name] = trs
        out[f"{axis_name}_stack"] = stack.astype(np.int32)
        pack_instances(f"{axis_name}_tr_", trs[0].instances, out)
    inf.load_model_to_device = orig_loader
    for v, name, instances in inf.tracker_consensus(
            trackers, None, MODEL_CONFIG, label_divisor=1000, pixel_vote_thr=pixel_vote_thr,
            cluster_iou_thr=0.75, allow_one_view=allow_one_view, min_size=min_size,
            min_extent=min_extent, dtype=np.int32):
        out["consensus_vol"] = v.astype(np.int32)
        pack_instances("consensus_", instances, out)
    for v, name, instances in inf.stack_postprocessing(
            {"xy": trackers["xy"]}, None, MODEL_CONFIG, label_divisor=1000, min_size=min_size,
            min_extent=min_extent, dtype=np.int32):
        out["stackpost_vol"] = v.astype(np.int32)
        pack_instances("stackpost_", instances, out)
    np.savez_compressed(os.path.join(GOLD, f"volume_{tag}.npz"), **out)
    print("volume", tag, {k: len(trackers[k][0].instances) for k in trackers},
          "consensus", len(out["consensus_labels"]))


def gen_post_cases_fine():
    """Per-slice vectors with `coarse_boundaries=False` (fine boundaries: full-resolution centre /
    offset maps, grouping step 1; engines.py:258-275) from the reference engine."""
    from empanada.inference.engines import PanopticDeepLabRenderEngine
    from empanada.inference.postprocess import find_instance_center

    rng = np.random.default_rng(17)
    cases = []
    for ci, (H, W) in enumerate([(64, 64), (48, 80), (32, 96), (16, 16), (80, 48), (64, 32)]):
        kind = ci % 3
        nms_kernel = [3, 7, 5][kind]
        thr, conf = 0.1, 0.5
        sem = rng.normal(0.5, 3, (1, H, W)).astype(np.float32)
        if kind == 0:      # smooth heat map with a handful of peaks, offsets towards them
            yy, xx = np.mgrid[0:H, 0:W].astype(np.float32)
            ctr = np.zeros((H, W), np.float32)
            off = np.zeros((2, H, W), np.float32)
            pts = [(rng.integers(4, H - 4), rng.integers(4, W - 4)) for _ in range(5)]
            d2 = np.stack([(yy - py) ** 2 + (xx - px) ** 2 for py, px in pts])
            near = d2.argmin(0)
            ctr = np.exp(-d2.min(0) / 50.0).astype(np.float32)
            off[0] = np.array([p[0] for p in pts], np.float32)[near] - yy
            off[1] = np.array([p[1] for p in pts], np.float32)[near] - xx
            off += rng.normal(0, 0.7, off.shape).astype(np.float32)
        elif kind == 1:    # random, many centres (> 20 -> chunked path)
            ctr = rng.uniform(0, 1, (H, W)).astype(np.float32) ** 6
            off = rng.normal(0, 5, (2, H, W)).astype(np.float32)
        else:              # plateaus and integer offsets: exact ties
            ctr = (rng.integers(0, 4, (H, W)) / 4.0).astype(np.float32) * (rng.uniform(0, 1, (H, W)) > 0.9)
            off = rng.integers(-6, 7, (2, H, W)).astype(np.float32)
        eng = PanopticDeepLabRenderEngine(
            torch.nn.Conv2d(1, 1, 1), thing_list=[1], label_divisor=1000, stuff_area=64, void_label=0,
            nms_threshold=thr, nms_kernel=nms_kernel, confidence_thr=conf, padding_factor=16,
            coarse_boundaries=False)
        prob = torch.sigmoid(torch.from_numpy(sem)[None])
        centers = find_instance_center(torch.from_numpy(ctr.copy())[None, None], thr, nms_kernel).numpy()
        cells = eng.get_instance_cells(torch.from_numpy(ctr.copy())[None, None], torch.from_numpy(off)[None], 1)
        pan = eng.postprocess(prob, cells)
        cases.append(dict(prob=prob[0].numpy(), ctr=ctr.astype(np.float32), off=off, nms_kernel=nms_kernel, thr=thr,
                          conf=conf, centers=centers, cells=cells[0, 0].numpy(), pan=pan[0].numpy()))
    out = {"n": len(cases)}
    for i, c in enumerate(cases):
        for k, v in c.items():
            out[f"c{i}_{k}"] = np.asarray(v)
    np.savez_compressed(os.path.join(GOLD, "post_cases_fine.npz"), **out)
    print("post_cases_fine", len(cases), [len(c["centers"]) for c in cases])


def gen_resize_cases():
    """`resize_by_factor` (data/utils/transforms.py:9-21) with the build container's opencv-python:
    pins oracle/transforms.py and the CUDA kernel on OpenCV's 8-bit INTER_LINEAR arithmetic."""
    import cv2
    from empanada.data.utils import resize_by_factor
    rng = np.random.default_rng(21)
    out = {"cv2_version": cv2.__version__}
    cases = [(64, 64, 2), (63, 64, 2), (64, 63, 2), (100, 75, 2), (128, 128, 4), (127, 130, 4), (33, 47, 4),
             (257, 255, 2), (96, 80, 8), (50, 50, 1), (17, 23, 16), (200, 333, 2)]
    for i, (h, w, f) in enumerate(cases):
        img = rng.integers(0, 256, (h, w), dtype=np.uint8)
        if i % 3 == 0:     # smooth content as well as noise
            yy, xx = np.mgrid[0:h, 0:w]
            img = ((np.sin(yy / 7.0) + np.cos(xx / 5.0) + 2) * 63).astype(np.uint8)
        out[f"r{i}_img"], out[f"r{i}_f"], out[f"r{i}_out"] = img, f, resize_by_factor(img, f)
    out["n"] = len(cases)
    np.savez_compressed(os.path.join(GOLD, "resize_cases.npz"), **out)
    print("resize cases", len(cases), "opencv", cv2.__version__)


def gen_tiled_cases(wide=False):
    """The tiled branch of the reference's `Engine2d.infer` (empanada_napari/inference.py:283-318:
    Tiler, per-tile engine call, pan_seg_to_rle_seg, translate_rle_seg, merge_objects_from_tiles /
    merge_semantic_from_tiles, rle_seg_to_pan_seg), unmodified. `cztile` (third party, absent) is
    replaced by a strategy object that returns the tile rectangles of
    `empanada_napari_b200.tiling.fixed_total_area_tiles_1d`, so the fixture pins everything
    downstream of the tile layout."""
    import types
    import empanada.inference.tile as ref_tile
    import empanada_napari.inference as inf
    from empanada_napari_b200.tiling import fixed_total_area_tiles_1d

    class FakeStrategy:
        def __init__(self, total_tile_width, total_tile_height, min_border_width):
            self.tw, self.th, self.b = total_tile_width, total_tile_height, min_border_width

        def tile_rectangle(self, rect):
            out = []
            for x0, sx in fixed_total_area_tiles_1d(rect.w, self.tw, self.b):
                for y0, sy in fixed_total_area_tiles_1d(rect.h, self.th, self.b):
                    out.append(types.SimpleNamespace(roi=types.SimpleNamespace(x=x0, y=y0, w=sx, h=sy)))
            return out

    ref_tile.AlmostEqualBorderFixedTotalAreaStrategy2D = FakeStrategy
    ref_tile.czrect = lambda x, y, w, h: types.SimpleNamespace(x=x, y=y, w=w, h=h)
    out = {}
    rng = np.random.default_rng(31)
    cases = [((300, 420), 128, False, 1, 14), ((260, 200), 96, False, 1, 8), ((200, 330), 128, True, 1, 10),
             ((280, 280), 128, False, 2, 9)]
    if wide:
        # an object wider than a tile: its flat-index runs wrap around tile row ends and the
        # reference translates only their starts (tile.py:126-166), painting the wrapped part
        # outside the tile. Pins the oracle's restatement of that quirk (the CUDA path does not
        # reproduce it and warns instead).
        cases = [((240, 300), 96, False, 1, 8), ((220, 260), 128, True, 1, 7), ((260, 230), 96, False, 1, 10)]
    orig_loader = inf.load_model_to_device
    for ci, (shape, tile_size, semantic_only, scale, n_obj) in enumerate(cases):
        # The reference's `_join_ranges` (array_utils.py:657-690) fails on a cluster that consists of
        # ONE run (an object clipped to a single row): seeds are tried until the reference runs.
        for seed in range(40 + 10 * ci, 40 + 10 * ci + 10):
            _, lab, _ = syn.make_volume((1,) + shape, seed=seed, n_objects=n_obj, scale=1.6)
            lab = lab[0].copy()
            if wide:
                y0 = 30 + 17 * ci
                lab[y0:y0 + 9 + 4 * ci, :] = int(lab.max()) + 1
            img = np.clip(np.where(lab > 0, 70.0, 170.0) + rng.normal(0, 8.0, shape), 0, 255).astype(np.uint8)
            tiler = ref_tile.Tiler(shape, tile_size=tile_size, overlap_width=min(128, int(tile_size * 0.1)))
            heads = []
            for (y0, y1), (x0, x1) in zip(tiler.yranges, tiler.xranges):
                tl = lab[y0:y1, x0:x1]
                sem, ctr, off = scaled_heads(tl, scale, 0.0, rng) if scale > 1 else noisy_heads(tl, 16, 0.0, rng)
                off = (off + rng.normal(0, 1.5, off.shape)).astype(np.float32)     # ragged instance borders
                heads.append((sem, ctr, off))
            fake = FakeModel(heads)
            inf.load_model_to_device = lambda url, device: fake
            eng = inf.Engine2d(MODEL_CONFIG, inference_scale=scale, label_divisor=1000, nms_threshold=0.1, nms_kernel=3,
                               confidence_thr=0.5, semantic_only=semantic_only, tile_size=tile_size, use_gpu=False)
            try:
                pan = eng.infer(img)
                break
            except Exception as e:
                print("  seed", seed, "reference failed:", type(e).__name__)
        for k, name in enumerate(("sem", "ctr", "off")):
            for t, h in enumerate(heads):
                out[f"t{ci}_{name}{t}"] = h[k]
        out[f"t{ci}_img"], out[f"t{ci}_pan"] = img, pan.astype(np.int32)
        out[f"t{ci}_meta"] = np.array([tile_size, int(semantic_only), scale, len(tiler)])
        print("tiled case", ci, shape, "tiles", len(tiler), "labels", len(np.unique(pan)) - 1)
    inf.load_model_to_device = orig_loader
    out["n"] = len(cases)
    np.savez_compressed(os.path.join(GOLD, "tiled_cases_wide.npz" if wide else "tiled_cases.npz"), **out)


def gen_model_tiny():
    """Reference PDL classes with the seeded weights of synthetic.make_pdl_state_dict(0)."""
    import yaml
    from empanada.models.quantization.panoptic_deeplab import QuantizablePanopticDeepLabPR

    cfg = yaml.safe_load(open(os.path.join(ref_shim.REF_ROOT, "empanada_napari/training/pdl_model.yaml")))
    cfg.pop("arch")
    cfg["num_classes"] = 1
    m = QuantizablePanopticDeepLabPR(**cfg, quantize=False).eval()
    m.fuse_model()
    sd = syn.make_pdl_state_dict(0)
    m.load_state_dict(sd)
    m = torch.jit.script(m)  # the deployed form
    rng = np.random.default_rng(11)
    img = rng.integers(0, 256, (96, 80), dtype=np.uint8)
    from empanada_napari.utils import Preprocessor
    x = Preprocessor(**MODEL_CONFIG["norms"])(img)["image"].unsqueeze(0)
    from empanada.inference.postprocess import factor_pad
    x = factor_pad(x, 16)
    with torch.no_grad():
        o = m(x, 2, False)
    np.savez_compressed(os.path.join(GOLD, "model_pdl_tiny.npz"), img=img, x=x.numpy(),
                        sem_logits=o["sem_logits"].numpy(), ctr_hmp=o["ctr_hmp"].numpy(),
                        offsets=o["offsets"].numpy())
    print("model tiny", {k: tuple(v.shape) for k, v in o.items()})


def gen_model_bifpn_tiny():
    """Reference PanopticBiFPN-PR classes with the seeded weights of
    synthetic.make_bifpn_state_dict(0), exported as `_train.py:59-73` does."""
    import yaml
    from empanada.models.quantization.panoptic_bifpn import QuantizablePanopticBiFPNPR

    cfg = yaml.safe_load(open(os.path.join(ref_shim.REF_ROOT, "empanada_napari/training/bifpn_model.yaml")))
    cfg.pop("arch")
    cfg["num_classes"] = 1
    m = QuantizablePanopticBiFPNPR(**cfg, quantize=False).eval()
    m.fuse_model()
    sd = syn.make_bifpn_state_dict(0)
    ref_keys = list(m.state_dict().keys())
    assert ref_keys == list(sd.keys()), [k for k in ref_keys if k not in sd][:5]
    m.load_state_dict(sd)
    m = torch.jit.script(m)  # the deployed form
    rng = np.random.default_rng(12)
    img = rng.integers(0, 256, (100, 200), dtype=np.uint8)
    from empanada_napari.utils import Preprocessor
    x = Preprocessor(**MODEL_CONFIG["norms"])(img)["image"].unsqueeze(0)
    from empanada.inference.postprocess import factor_pad
    x = factor_pad(x, 128)
    with torch.no_grad():
        o = m(x, 2, False)
    np.savez_compressed(os.path.join(GOLD, "model_bifpn_tiny.npz"), img=img, x=x.numpy(),
                        sem_logits=o["sem_logits"].numpy(), ctr_hmp=o["ctr_hmp"].numpy(),
                        offsets=o["offsets"].numpy())
    print("bifpn tiny", {k: tuple(v.shape) for k, v in o.items()})


def gen_eval_cases():
    """Reference `Evaluator` (evaluation/evaluator.py:24-122) with f1_50 / f1_75 / iou on pairs of
    tracker JSONs written by the reference InstanceTracker: ground truth = exact ellipsoid labels,
    prediction = the same labels eroded / shifted / with objects dropped and merged."""
    import json
    import tempfile
    from empanada.evaluation.evaluator import Evaluator
    from empanada.evaluation.instance_metrics import f1_50, f1_75
    from empanada.evaluation.semantic_metrics import iou
    from empanada.inference.tracker import InstanceTracker
    from empanada.array_utils import rle_encode

    def tracker_json(lab, path):
        tr = InstanceTracker(1, 1000, lab.shape, "xy")
        flat = lab.ravel()
        for l in np.unique(flat):
            if l == 0:
                continue
            idx = np.flatnonzero(flat == l)
            starts, runs = rle_encode(idx)
            zz, yy, xx = np.unravel_index(idx, lab.shape)
            tr.instances[int(l)] = {"box": (int(zz.min()), int(yy.min()), int(xx.min()), int(zz.max()) + 1,
                                            int(yy.max()) + 1, int(xx.max()) + 1), "starts": starts, "runs": runs}
        tr.finished = True
        tr.write_to_json(path)

    cases = []
    ev = Evaluator(semantic_metrics={"iou": iou}, instance_metrics={"f1_50": f1_50, "f1_75": f1_75})
    with tempfile.TemporaryDirectory() as tmp:
        for ci, (seed, shape, mode) in enumerate([(3, (24, 40, 36), "same"), (4, (28, 44, 40), "shift"),
                                                  (5, (20, 48, 48), "drop"), (6, (24, 40, 40), "merge")]):
            _, lab, _ = syn.make_volume(shape, seed=seed, scale=1.0)
            pred = lab.copy()
            if mode == "shift":
                pred = np.roll(pred, 2, axis=2)
                pred[:, :, :2] = 0
            elif mode == "drop":
                ids = np.unique(pred)[1:]
                pred[np.isin(pred, ids[::3])] = 0
            elif mode == "merge":
                ids = np.unique(pred)[1:]
                pred[pred == ids[1]] = ids[0]
                pred[:, ::2, :][pred[:, ::2, :] == ids[2]] = 0
            gp, pp = os.path.join(tmp, f"gt{ci}.json"), os.path.join(tmp, f"pr{ci}.json")
            tracker_json(lab, gp)
            tracker_json(pred, pp)
            res = ev(gp, pp)
            cases.append({"gt": json.load(open(gp)), "pred": json.load(open(pp)),
                          "results": {k: float(v) for k, v in res.items()}})
    with open(os.path.join(GOLD, "eval_cases.json"), "w") as f:
        json.dump(cases, f)
    print("eval cases:", [c["results"] for c in cases])


if __name__ == "__main__":
    os.makedirs(GOLD, exist_ok=True)
    which = sys.argv[1:] or ["post", "post_fine", "median", "model", "bifpn", "volumes", "resize", "volumes2", "tiled", "morph",
                             "stuff", "tiled_wide"]
    if "stuff" in which:
        run_reference_volume((24, 42, 38), seed=15, noise=0.5, ks=3, n_objects=10, min_size=20, min_extent=2,
                             tag="stuff_class", stuff_config=True)
        run_reference_volume((20, 36, 44), seed=16, noise=0.3, ks=1, n_objects=8, min_size=20, min_extent=2,
                             tag="stuff_class_vote3", stuff_config=True, pixel_vote_thr=3)
    if "resize" in which:
        gen_resize_cases()
    if "tiled" in which:
        gen_tiled_cases()
    if "tiled_wide" in which:
        gen_tiled_cases(wide=True)
    if "volumes2" in which:
        run_reference_volume((26, 44, 40), seed=5, noise=0.5, ks=3, n_objects=10, min_size=20, min_extent=2,
                             tag="semantic_only", semantic_only=True)
        run_reference_volume((28, 50, 46), seed=6, noise=0.4, ks=3, n_objects=10, min_size=20, min_extent=2,
                             tag="scale2", inference_scale=2)
        run_reference_volume((32, 70, 61), seed=7, noise=0.3, ks=1, n_objects=8, min_size=20, min_extent=2,
                             tag="scale4_semantic", inference_scale=4, semantic_only=True)
    if "morph" in which:
        run_reference_volume((26, 44, 40), seed=8, noise=0.5, ks=3, n_objects=12, min_size=20, min_extent=2,
                             tag="erode1", erosion=1)
        run_reference_volume((24, 40, 36), seed=9, noise=0.5, ks=3, n_objects=12, min_size=20, min_extent=2,
                             tag="dilate2_fill", dilation=2, fill_holes=True)
        run_reference_volume((22, 48, 40), seed=10, noise=0.7, ks=1, n_objects=10, min_size=10, min_extent=2,
                             tag="erode1_dilate1_fill", erosion=1, dilation=1, fill_holes=True)
    if "eval" in which:
        gen_eval_cases()
    if "post" in which:
        gen_post_cases()
    if "post_fine" in which:
        gen_post_cases_fine()
    if "median" in which:
        gen_median()
    if "model" in which:
        gen_model_tiny()
    if "bifpn" in which:
        gen_model_bifpn_tiny()
    if "volumes" in which:
        run_reference_volume((24, 40, 48), seed=1, noise=0.0, ks=3, n_objects=10, min_size=50, min_extent=3, tag="clean")
        run_reference_volume((32, 36, 44), seed=2, noise=0.6, ks=3, n_objects=14, min_size=20, min_extent=2, tag="noisy")
        run_reference_volume((20, 33, 30), seed=3, noise=0.3, ks=5, n_objects=8, min_size=10, min_extent=2, tag="ks5_odd",
                             pixel_vote_thr=1, allow_one_view=True)
