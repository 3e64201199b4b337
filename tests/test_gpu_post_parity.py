"""-m gpu: the CUDA post-processing / tracking / consensus path against the CPU oracle on the
same seeded inputs, and against the golden fixtures generated from the reference itself.
Bit-exact (integer label maps, RLE tables)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, assert_instances_equal, unpack_instances

pytestmark = pytest.mark.gpu

MODEL_CONFIG = {"class_names": {1: "mito"}, "labels": [1], "thing_list": [1], "padding_factor": 16,
                "norms": {"mean": 0.57571, "std": 0.12765}, "model": None}


def _engine(heads_by_axis, **kw):
    import torch
    from empanada_napari_b200.inference import Engine3d
    from empanada_napari_b200.model import SyntheticHeadsModel

    def heads_fn(axis, s0, s1):
        sem, ctr, off = heads_by_axis[axis]
        dev = torch.device("cuda:0")
        return (torch.from_numpy(np.ascontiguousarray(sem[s0:s1, 0])).to(dev),
                torch.from_numpy(np.ascontiguousarray(ctr[s0:s1])).to(dev),
                torch.from_numpy(np.ascontiguousarray(off[s0:s1])).to(dev))

    cfg = dict(MODEL_CONFIG)
    cfg["model"] = SyntheticHeadsModel(heads_fn)
    return Engine3d(cfg, **kw), cfg


def test_post_cases_golden():
    """per-slice kernels vs reference-generated vectors (centres, cells, pan_seg)."""
    import torch
    from empanada_napari_b200.postproc import PlanePost
    z = np.load(os.path.join(GOLDEN, "post_cases.npz"))
    dev = torch.device("cuda:0")
    for i in range(int(z["n"])):
        ctr, off, prob = z[f"c{i}_ctr"], z[f"c{i}_off"], z[f"c{i}_prob"]
        H, W = prob.shape[-2:]
        post = PlanePost(1, H, W, H, W, ks=1, thing_class=1, label_divisor=1000,
                         nms_threshold=float(z[f"c{i}_thr"]), nms_kernel=int(z[f"c{i}_nms_kernel"]),
                         confidence_thr=float(z[f"c{i}_conf"]), device=dev)
        post.push_heads(torch.from_numpy(prob).to(dev), torch.from_numpy(ctr[None]).to(dev),
                        torch.from_numpy(off[None]).to(dev), is_prob=True)
        post.finish_heads()
        k = int(post.center_counts[0].item())
        packed = post.centers[0, :k].cpu().numpy()
        centers = np.stack([packed >> 16, packed & 0xFFFF], axis=1).reshape(-1, 2)
        assert np.array_equal(centers, z[f"c{i}_centers"].reshape(-1, 2)), i
        cells4 = post.cells4[0].cpu().numpy()
        assert np.array_equal(np.repeat(np.repeat(cells4, 4, 0), 4, 1), z[f"c{i}_cells"].astype(np.int32)), i
        pan = post.pan_batch(0, 1)[0].cpu().numpy()
        assert np.array_equal(pan, z[f"c{i}_pan"]), i


def test_median_golden():
    import torch
    from empanada_napari_b200.postproc import PlanePost
    z = np.load(os.path.join(GOLDEN, "median_cases.npz"))
    dev = torch.device("cuda:0")
    for i in range(int(z["n"])):
        ks, x, ts, ys = int(z[f"m{i}_ks"]), z[f"m{i}_x"], z[f"m{i}_t"], z[f"m{i}_y"]
        n, _, H, W = x.shape
        if n < ks:
            continue
        for bs in (1, 2, n):
            post = PlanePost(n, H, W, H, W, ks=ks, thing_class=1, label_divisor=1000, device=dev,
                             keep_prob=True, scale=1)
            for s0 in range(0, n, bs):
                s1 = min(n, s0 + bs)
                post.push_heads(torch.from_numpy(x[s0:s1, 0]).to(dev),
                                torch.zeros((s1 - s0, H, W), device=dev),
                                torch.zeros((s1 - s0, 2, H, W), device=dev), is_prob=True)
            post.finish_heads()
            got = post.prob.cpu().numpy()
            assert ts.tolist() == list(range(n))
            assert np.array_equal(got, ys[:, 0]), (i, bs)
            assert np.array_equal(post.hard.cpu().numpy(), (ys[:, 0] >= np.float32(0.5)).astype(np.uint8))


@pytest.mark.parametrize("tag", ["clean", "noisy", "ks5_odd"])
def test_volume_golden(tag):
    """Engine3d.infer_on_axis x3 + tracker_consensus + stack_postprocessing vs the reference."""
    import empanada_napari_b200.synthetic as syn
    from empanada_napari_b200.inference import stack_postprocessing, tracker_consensus
    z = np.load(os.path.join(GOLDEN, f"volume_{tag}.npz"))
    shape = tuple(int(v) for v in z["shape"])
    vol, _, _ = syn.make_volume(shape, seed=int(z["seed"]), n_objects=int(z["n_objects"]), scale=1.0)
    heads = {a: (z[f"{n}_sem"].astype(np.float32), z[f"{n}_ctr"], z[f"{n}_off"])
             for a, n in enumerate(("xy", "xz", "yz"))}
    eng, cfg = _engine(heads, median_kernel_size=int(z["ks"]), nms_kernel=3, confidence_thr=0.5,
                       min_size=int(z["min_size"]), min_extent=int(z["min_extent"]),
                       save_panoptic=True, batch_size=5)
    trackers = {}
    for axis_name in ("xy", "xz", "yz"):
        stack, trs = eng.infer_on_axis(vol, axis_name)
        assert_instances_equal(trs[0].instances, unpack_instances(z, f"{axis_name}_tr_"))
        assert stack.dtype == np.int32 and np.array_equal(stack, z[f"{axis_name}_stack"])
        trackers[axis_name] = trs
    for v, name, inst in tracker_consensus(
            trackers, None, cfg, pixel_vote_thr=int(z["pixel_vote_thr"]),
            allow_one_view=bool(z["allow_one_view"]), min_size=int(z["min_size"]),
            min_extent=int(z["min_extent"]), dtype=np.int32):
        assert_instances_equal(inst, unpack_instances(z, "consensus_"))
        assert np.array_equal(v, z["consensus_vol"])
    for v, name, inst in stack_postprocessing(
            {"xy": trackers["xy"]}, None, cfg, min_size=int(z["min_size"]),
            min_extent=int(z["min_extent"]), dtype=np.int32):
        assert_instances_equal(inst, unpack_instances(z, "stackpost_"))
        assert np.array_equal(v, z["stackpost_vol"])


@pytest.mark.parametrize("seed,shape,noise,ks", [(11, (40, 64, 72), 0.0, 3), (12, (48, 56, 50), 0.5, 3),
                                                 (13, (36, 70, 41), 0.8, 5), (14, (64, 64, 64), 0.3, 1)])
def test_volume_vs_oracle(seed, shape, noise, ks):
    """Larger random volumes: CUDA path vs the pinned CPU oracle, bit-exact."""
    import empanada_napari_b200.synthetic as syn
    from empanada_napari_b200.inference import tracker_consensus
    from oracle import consensus as ocons, pipeline
    vol, lab, _ = syn.make_volume(shape, seed=seed, scale=1.0)
    rng = np.random.default_rng(seed)
    heads = {}
    for axis in range(3):
        hs = []
        for i in range(shape[axis]):
            sem, ctr, off = syn.analytic_heads(np.take(lab, i, axis=axis), pad_to=16)
            sem = sem + rng.normal(0, 2.5 * noise, sem.shape).astype(np.float32)
            ctr = ctr + rng.normal(0, 0.05 * noise, ctr.shape).astype(np.float32)
            off = off + rng.normal(0, 2.0 * noise, off.shape).astype(np.float32)
            hs.append((sem.astype(np.float32), ctr.astype(np.float32), off.astype(np.float32)))
        heads[axis] = (np.stack([h[0] for h in hs]), np.stack([h[1] for h in hs]), np.stack([h[2] for h in hs]))
    eng, cfg = _engine(heads, median_kernel_size=ks, nms_kernel=3, confidence_thr=0.5, min_size=30,
                       min_extent=3, save_panoptic=True, batch_size=7)
    got, want = {}, {}
    for a, axis_name in enumerate(("xy", "xz", "yz")):
        stack, trs = eng.infer_on_axis(vol, axis_name)
        sem, ctr, off = heads[a]
        ostack, otrs = pipeline.infer_on_axis(vol, axis_name, lambda i, x: (sem[i], ctr[i], off[i]), cfg,
                                              median_kernel_size=ks, nms_kernel=3, confidence_thr=0.5,
                                              min_size=30, min_extent=3)
        assert_instances_equal(trs[0].instances, otrs[0].instances)
        assert np.array_equal(stack, ostack)
        got[axis_name], want[axis_name] = trs, otrs
    for (v, _, inst), (ov, _, oinst) in zip(
            tracker_consensus(got, None, cfg, pixel_vote_thr=2, min_size=30, min_extent=3, dtype=np.int32),
            ocons.tracker_consensus(want, cfg, pixel_vote_thr=2, min_size=30, min_extent=3, dtype=np.int32)):
        assert_instances_equal(inst, oinst)
        assert np.array_equal(v, ov)
        assert len(inst) > 0


def test_deferred_replay_and_lazy_rle_equal_eager():
    """The overlapped (worker-thread) matcher replay + lazily extracted per-plane RLE give the
    same trackers and the same consensus as the eager path and the oracle."""
    import empanada_napari_b200.synthetic as syn
    from empanada_napari_b200.inference import tracker_consensus
    from empanada_napari_b200.tracking import PendingTracker
    from oracle import consensus as ocons, pipeline
    shape, seed, ks = (44, 60, 52), 21, 3
    vol, lab, _ = syn.make_volume(shape, seed=seed, scale=1.0)
    heads = {}
    for axis in range(3):
        hs = [syn.analytic_heads(np.take(lab, i, axis=axis), pad_to=16) for i in range(shape[axis])]
        heads[axis] = tuple(np.stack([h[k] for h in hs]).astype(np.float32) for k in range(3))
    kw = dict(median_kernel_size=ks, nms_kernel=3, confidence_thr=0.5, min_size=30, min_extent=3, batch_size=6)
    eng, cfg = _engine(heads, save_panoptic=False, **kw)
    got, want = {}, {}
    for axis_name in ("xy", "xz", "yz"):      # all three planes first: replays overlap the next plane
        stack, got[axis_name] = eng.infer_on_axis(vol, axis_name)
        assert stack is None and isinstance(got[axis_name][0], PendingTracker)
    for a, axis_name in enumerate(("xy", "xz", "yz")):
        sem, ctr, off = heads[a]
        _, want[axis_name] = pipeline.infer_on_axis(vol, axis_name, lambda i, x: (sem[i], ctr[i], off[i]), cfg,
                                                    median_kernel_size=ks, nms_kernel=3, confidence_thr=0.5,
                                                    min_size=30, min_extent=3)
    for (v, _, inst), (ov, _, oinst) in zip(
            tracker_consensus(got, None, cfg, pixel_vote_thr=2, min_size=30, min_extent=3, dtype=np.int32),
            ocons.tracker_consensus(want, cfg, pixel_vote_thr=2, min_size=30, min_extent=3, dtype=np.int32)):
        assert_instances_equal(inst, oinst)
        assert np.array_equal(v, ov)
    for axis_name in ("xy", "xz", "yz"):      # RLE materialised on first access, after consensus
        assert_instances_equal(got[axis_name][0].instances, want[axis_name][0].instances)


def _heads_from_labels(lab, shape):
    import empanada_napari_b200.synthetic as syn
    heads = {}
    for axis in range(3):
        hs = [syn.analytic_heads(np.take(lab, i, axis=axis), pad_to=16) for i in range(shape[axis])]
        heads[axis] = tuple(np.stack([h[k] for h in hs]).astype(np.float32) for k in range(3))
    return heads


def test_empty_volume_all_background():
    """Edge case: no foreground anywhere -> empty trackers, zero volumes, empty consensus."""
    from empanada_napari_b200.inference import stack_postprocessing, tracker_consensus
    shape = (12, 40, 33)
    lab = np.zeros(shape, dtype=np.int32)
    vol = np.full(shape, 170, dtype=np.uint8)
    eng, cfg = _engine(_heads_from_labels(lab, shape), median_kernel_size=3, nms_kernel=3, confidence_thr=0.5,
                       min_size=10, min_extent=2, save_panoptic=True, batch_size=5)
    trackers = {}
    for axis_name in ("xy", "xz", "yz"):
        stack, trs = eng.infer_on_axis(vol, axis_name)
        assert stack.shape == shape and stack.dtype == np.int32 and not stack.any()
        assert trs[0].instances == {}
        trackers[axis_name] = trs
    for v, name, inst in tracker_consensus(trackers, None, cfg, pixel_vote_thr=2, min_size=10, min_extent=2, dtype=np.uint32):
        assert inst == {} and v.shape == shape and v.dtype == np.uint32 and not v.any()
    for v, name, inst in stack_postprocessing({"xy": trackers["xy"]}, None, cfg, min_size=10, min_extent=2, dtype=np.uint32):
        assert inst == {} and not v.any()


def test_stack_shorter_than_median_kernel_raises():
    """The reference's median queue never emits for stacks shorter than the kernel; here the
    engine refuses loudly instead of returning an empty stack."""
    lab = np.zeros((2, 32, 32), dtype=np.int32)
    eng, cfg = _engine(_heads_from_labels(lab, lab.shape), median_kernel_size=3, confidence_thr=0.5, batch_size=2)
    with pytest.raises(ValueError):
        eng.infer_on_axis(np.zeros(lab.shape, dtype=np.uint8), "xy")


def test_float_and_non_uint8_inputs_rejected():
    lab = np.zeros((4, 32, 32), dtype=np.int32)
    eng, cfg = _engine(_heads_from_labels(lab, lab.shape), median_kernel_size=1, confidence_thr=0.5)
    with pytest.raises(Exception, match="float"):
        eng.infer_on_axis(np.zeros(lab.shape, dtype=np.float32), "xy")
    with pytest.raises(NotImplementedError):
        eng.infer_on_axis(np.zeros(lab.shape, dtype=np.uint16), "xy")


def test_rle_volume_round_trip_properties_256():
    """Size-independent properties on a 256^3 job (the benchmark's synthetic generator): every
    tracker's RLE decodes to exactly its dense label volume; consensus ids are 1..n; run lengths
    sum to the label histogram; the painted consensus volume re-encodes to the same runs."""
    import torch
    import bench
    from empanada_napari_b200 import consensus
    from empanada_napari_b200.inference import Engine3d, tracker_consensus
    from empanada_napari_b200.model import SyntheticHeadsModel
    S = 256
    dev = torch.device("cuda:0")
    vol_d, lab_d, n_obj = bench.synth_on_device(S, dev)
    heads = {a: bench.analytic_heads_on_device(lab_d, a, n_obj) for a in range(3)}
    cfg = dict(bench.MODEL_CONFIG)
    cfg["model"] = SyntheticHeadsModel(lambda a, s0, s1: tuple(t[s0:s1] for t in heads[a]))
    eng = Engine3d(cfg, median_kernel_size=3, nms_kernel=3, confidence_thr=0.5, min_size=500, min_extent=5, batch_size=16)
    trackers = {name: eng.infer_on_axis(vol_d, name)[1] for name in ("xy", "xz", "yz")}

    def decode(instances, n):
        out = np.zeros(n, dtype=np.int32)
        for label, a in instances.items():
            for s, r in zip(a["starts"], a["runs"]):
                out[s:s + r] = label
        return out

    for name, trs in trackers.items():
        tr = trs[0]
        dense = tr._b200_dense.cpu().numpy()
        assert len(tr.instances) > 5
        if name != "xz":   # xz runs may wrap a 2-D row (tracker.py:80-84)
            assert np.array_equal(decode(tr.instances, S ** 3).reshape(dense.shape), dense), name
        hist = np.bincount(dense.ravel())
        for label, a in tr.instances.items():
            assert int(a["runs"].sum()) == int(hist[label]) == tr._b200_sizes[label], (name, label)
    for v, _, inst in tracker_consensus(trackers, None, cfg, pixel_vote_thr=2, min_size=500, min_extent=5, dtype=np.int32):
        assert list(inst.keys()) == sorted(inst.keys()) and len(inst) > 5
        painted = decode(inst, S ** 3)          # later ids overwrite earlier ones, as fill_volume does
        assert np.array_equal(painted.reshape(v.shape), v)
        lab2, st2, ln2 = consensus.extract_runs(torch.from_numpy(v.copy()).to(dev))
        assert int(ln2.sum().item()) == int((v > 0).sum())
        for label, a in inst.items():
            z0, y0, x0, z1, y1, x1 = a["box"]
            zz, yy, xx = np.nonzero(v == label)
            if len(zz):
                assert z0 <= zz.min() and zz.max() < z1 and y0 <= yy.min() and yy.max() < y1 and x0 <= xx.min() and xx.max() < x1


def test_up4_align_corners_vs_torch():
    """be_up4 == F.interpolate(scale_factor=4, mode='bilinear', align_corners=True) (fp32; the
    reference applies it inside the model when interpolate_ins=True)."""
    import torch
    from empanada_napari_b200 import _lib
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(3)
    for planes, h, w in ((3, 13, 17), (2, 32, 64), (1, 1, 5)):
        x = torch.randn(planes, h, w, generator=g)
        out = torch.empty(planes, 4 * h, 4 * w, device=dev)
        _lib.call("be_up4", _lib.ptr(x.to(dev)), planes, h, w, _lib.ptr(out), _lib.stream_ptr())
        ref = torch.nn.functional.interpolate(x[None], scale_factor=4.0, mode="bilinear", align_corners=True)[0]
        assert float((out.cpu() - ref).abs().max()) <= 2e-6 * max(1.0, float(ref.abs().max()))


def test_fine_boundaries_vs_oracle():
    """`fine_boundaries=True` (interpolate_ins, grouping step 1): bit-exact against the oracle fed
    with the same full-resolution centre / offset maps."""
    import torch
    import empanada_napari_b200.synthetic as syn
    from empanada_napari_b200.inference import tracker_consensus, upsample_instance_heads
    from oracle import consensus as ocons, pipeline
    shape, seed, ks = (28, 48, 40), 41, 3
    vol, lab, _ = syn.make_volume(shape, seed=seed, scale=1.0)
    heads = _heads_from_labels(lab, shape)
    dev = torch.device("cuda:0")
    full = {}
    for a in range(3):
        c, o = upsample_instance_heads(torch.from_numpy(heads[a][1]).to(dev), torch.from_numpy(heads[a][2]).to(dev))
        full[a] = (heads[a][0], c.cpu().numpy(), o.cpu().numpy())
    kw = dict(median_kernel_size=ks, nms_kernel=3, confidence_thr=0.5, min_size=30, min_extent=3, batch_size=5)
    eng, cfg = _engine(heads, save_panoptic=True, fine_boundaries=True, **kw)
    got, want = {}, {}
    for a, axis_name in enumerate(("xy", "xz", "yz")):
        stack, got[axis_name] = eng.infer_on_axis(vol, axis_name)
        sem, ctr, off = full[a]
        ostack, want[axis_name] = pipeline.infer_on_axis(
            vol, axis_name, lambda i, x: (sem[i], ctr[i], off[i]), cfg, median_kernel_size=ks, nms_kernel=3,
            confidence_thr=0.5, min_size=30, min_extent=3, fine_boundaries=True)
        assert_instances_equal(got[axis_name][0].instances, want[axis_name][0].instances)
        assert np.array_equal(stack, ostack)
    for (v, _, inst), (ov, _, oinst) in zip(
            tracker_consensus(got, None, cfg, pixel_vote_thr=2, min_size=30, min_extent=3, dtype=np.int32),
            ocons.tracker_consensus(want, cfg, pixel_vote_thr=2, min_size=30, min_extent=3, dtype=np.int32)):
        assert_instances_equal(inst, oinst)
        assert np.array_equal(v, ov)
        assert len(inst) > 0
