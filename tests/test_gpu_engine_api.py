"""-m gpu: the `empanada.inference.engines` object API on the CUDA kernels
(empanada-napari_b200/engines.py <- /root/reference/empanada/inference/engines.py:223-394):
per-slice engines called the way the reference's own orchestration calls them, compared bit for
bit with the reference-generated fixtures and with the oracle's RenderEnginePost."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _engine(cls, heads, **kw):
    """Engine whose model returns the given (sem_logits (1,H,W), ctr (h4,w4), off (2,h4,w4))
    numpy heads, slice after slice."""
    import torch
    from empanada_napari_b200.model import SyntheticHeadsModel
    dev = torch.device("cuda:0")
    it = iter(heads)

    def heads_fn(axis, s0, s1):
        sem, ctr, off = next(it)
        return (torch.from_numpy(np.ascontiguousarray(sem)).to(dev), torch.from_numpy(np.ascontiguousarray(ctr[None])).to(dev),
                torch.from_numpy(np.ascontiguousarray(off[None])).to(dev))
    return cls(SyntheticHeadsModel(heads_fn), **kw)


@pytest.mark.parametrize("fixture,coarse", [("post_cases.npz", True), ("post_cases_fine.npz", False)])
def test_engine_pieces_match_reference_fixture(fixture, coarse):
    """get_instance_cells + postprocess (engines.py:258-298) on the reference's own vectors."""
    import torch
    from empanada_napari_b200.engines import PanopticDeepLabRenderEngine
    z = np.load(os.path.join(GOLDEN, fixture))
    for i in range(int(z["n"])):
        eng = _engine(PanopticDeepLabRenderEngine, [], thing_list=[1], label_divisor=1000, stuff_area=64,
                      void_label=0, nms_threshold=float(z[f"c{i}_thr"]), nms_kernel=int(z[f"c{i}_nms_kernel"]),
                      confidence_thr=float(z[f"c{i}_conf"]), padding_factor=16, coarse_boundaries=coarse)
        ctr = torch.from_numpy(z[f"c{i}_ctr"])[None, None].cuda()
        off = torch.from_numpy(z[f"c{i}_off"])[None].cuda()
        prob = torch.from_numpy(z[f"c{i}_prob"])[None].cuda()
        cells = eng.get_instance_cells(ctr, off, 1)
        assert cells.dtype == torch.float32 and tuple(cells.shape) == (1, 1) + z[f"c{i}_cells"].shape
        assert np.array_equal(cells[0, 0].cpu().numpy(), z[f"c{i}_cells"]), i
        pan = eng.postprocess(prob, cells)
        assert pan.dtype == torch.int64 and np.array_equal(pan[0].cpu().numpy(), z[f"c{i}_pan"]), i


@pytest.mark.parametrize("upsampling,coarse,thing_list", [(1, True, [1]), (2, True, [1]), (4, False, [1]),
                                                          (1, True, []), (2, False, [])])
def test_render_engine_call_vs_oracle(upsampling, coarse, thing_list):
    """engine(image, size, upsampling) -> pan_seg (1, h, w) int64 (engines.py:300-325), with extra
    render steps / nearest upsampling of the cells for upsampling > 1 and stuff pasting for
    thing_list = []."""
    import torch
    import empanada_napari_b200.synthetic as syn
    from empanada_napari_b200.engines import PanopticDeepLabRenderEngine
    from oracle import post
    rng = np.random.default_rng(3 + upsampling)
    h, w = 75, 100                         # size of the ORIGINAL image
    dh, dw = -(-h // upsampling), -(-w // upsampling)
    Hd, Wd = dh + (16 - dh % 16) % 16, dw + (16 - dw % 16) % 16
    _, lab, _ = syn.make_volume((1, Hd * upsampling, Wd * upsampling), seed=4, n_objects=9, scale=float(upsampling))
    sem = np.where(lab[0] > 0, 4.0, -4.0).astype(np.float32)[None] + rng.normal(0, 2.0, (1, Hd * upsampling, Wd * upsampling)).astype(np.float32)
    g = 4 if coarse else 1
    _, ctr, off = syn.analytic_heads(lab[0][::upsampling, ::upsampling], step=g)
    off = (off + rng.normal(0, 0.5, off.shape)).astype(np.float32)
    kw = dict(thing_list=thing_list, label_divisor=1000, stuff_area=64, void_label=0, nms_threshold=0.1, nms_kernel=3,
              confidence_thr=0.5, padding_factor=16, coarse_boundaries=coarse)
    eng = _engine(PanopticDeepLabRenderEngine, [(sem, ctr, off)], **kw)
    if not coarse:    # the engine up-samples the /4 heads itself (interpolate_ins); feed it /4 maps
        from empanada_napari_b200.inference import upsample_instance_heads
        _, c4, o4 = syn.analytic_heads(lab[0][::upsampling, ::upsampling], step=4)
        eng = _engine(PanopticDeepLabRenderEngine, [(sem, c4, o4)], **kw)
        cu, ou = upsample_instance_heads(torch.from_numpy(c4[None]).cuda(), torch.from_numpy(o4[None]).cuda())
        ctr, off = cu[0].cpu().numpy(), ou[0].cpu().numpy()
    image = torch.zeros((1, 1, dh, dw), dtype=torch.float32)
    pan = eng(image, (h, w), upsampling=upsampling)
    assert pan.dtype == torch.int64 and tuple(pan.shape) == (1, h, w) and pan.is_cuda
    okw = dict(kw)
    okw.pop("padding_factor")
    want = post.RenderEnginePost(median_kernel_size=None, **okw)(post.sigmoid(sem), ctr, off, (h, w), upsampling)
    assert np.array_equal(pan[0].cpu().numpy(), want)
    assert len(np.unique(want)) > (3 if thing_list else 1)


def test_render_engine3d_queue_protocol_vs_oracle():
    """PanopticDeepLabRenderEngine3d: None while the median queue builds, filtered slices after,
    `end()` for the tail, `reset()`, and `ks` / `mid_idx` mutated as Engine3d.update_params does
    (engines.py:327-394, empanada_napari/inference.py:439-455)."""
    import torch
    import empanada_napari_b200.synthetic as syn
    from empanada_napari_b200.engines import PanopticDeepLabRenderEngine3d
    from oracle import post
    rng = np.random.default_rng(9)
    _, lab, _ = syn.make_volume((9, 64, 80), seed=12, n_objects=8, scale=1.0)
    heads = []
    for i in range(9):
        sem, ctr, off = syn.analytic_heads(lab[i], pad_to=16)
        heads.append(((sem + rng.normal(0, 3.0, sem.shape)).astype(np.float32), ctr, off))
    kw = dict(thing_list=[1], label_divisor=1000, stuff_area=64, void_label=0, nms_threshold=0.1, nms_kernel=3,
              confidence_thr=0.5, coarse_boundaries=True)
    for ks in (3, 5):
        eng = _engine(PanopticDeepLabRenderEngine3d, heads, median_kernel_size=3, padding_factor=16, **kw)
        eng.ks, eng.mid_idx = ks, (ks - 1) // 2        # as update_params does, then reset()
        eng.reset()
        oracle = post.RenderEnginePost(median_kernel_size=ks, **kw)
        image = torch.zeros((1, 1, 64, 80), dtype=torch.float32)
        got, want = [], []
        for i in range(9):
            got.append(eng(image, (64, 80)))
            want.append(oracle(post.sigmoid(heads[i][0]), heads[i][1], heads[i][2], (64, 80)))
        got += eng.end()
        want += oracle.end()
        assert len(got) == len(want) == 9 + (ks - 1) // 2
        for a, b in zip(got, want):
            assert (a is None) == (b is None)
            if a is not None:
                assert a.dtype == torch.int64 and np.array_equal(a[0].cpu().numpy(), b)
        assert sum(a is None for a in got) == (ks - 1) // 2


def test_tiled_inference_matches_reference_fixture():
    """`Engine2d.infer` with tile_size > 0 on images larger than the tile (inference.py:283-318):
    tests/golden/tiled_cases.npz holds the unmodified reference's output for the same tile layout
    (oracle/make_golden.py gen_tiled_cases) - thing class with merging across tiles and the
    overlap-region false-positive filter, semantic-only, and inference_scale 2."""
    import torch
    from empanada_napari_b200.inference import Engine2d
    from empanada_napari_b200.model import SyntheticHeadsModel
    z = np.load(os.path.join(GOLDEN, "tiled_cases.npz"))
    dev = torch.device("cuda:0")
    for ci in range(int(z["n"])):
        tile_size, semantic_only, scale, n_tiles = (int(v) for v in z[f"t{ci}_meta"])
        sem = torch.from_numpy(np.stack([z[f"t{ci}_sem{t}"][0] for t in range(n_tiles)])).to(dev)
        ctr = torch.from_numpy(np.stack([z[f"t{ci}_ctr{t}"] for t in range(n_tiles)])).to(dev)
        off = torch.from_numpy(np.stack([z[f"t{ci}_off{t}"] for t in range(n_tiles)])).to(dev)
        cfg = {"class_names": {1: "mito"}, "labels": [1], "thing_list": [1], "padding_factor": 16,
               "norms": {"mean": 0.57571, "std": 0.12765},
               "model": SyntheticHeadsModel(lambda a, s0, s1: (sem[s0:s1], ctr[s0:s1], off[s0:s1]))}
        eng = Engine2d(cfg, inference_scale=scale, label_divisor=1000, nms_threshold=0.1, nms_kernel=3,
                       confidence_thr=0.5, semantic_only=bool(semantic_only), tile_size=tile_size)
        img = z[f"t{ci}_img"]
        out = eng.infer(img)
        assert eng.last_stats["tiles"] == n_tiles
        assert out.dtype == np.int32 and out.shape == img.shape
        assert np.array_equal(out, z[f"t{ci}_pan"]), ci
