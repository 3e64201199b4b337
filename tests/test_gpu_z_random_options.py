"""-m gpu: random combinations of the engine options on random noisy volumes, CUDA path vs the
CPU oracle, bit-exact. Same generator family as oracle/fuzz_reference.py (which runs the
unmodified reference against the oracle). The file name sorts after the other GPU suites: this is
the long random sweep and runs last."""
import numpy as np
import pytest

from conftest import assert_instances_equal
from test_gpu_post_parity import MODEL_CONFIG, _engine  # noqa: F401


def random_engine_case(seed):
    """A random small volume + noisy head maps + a random combination of engine options (the
    same generator family as oracle/fuzz_reference.py, which runs the unmodified reference against
    the oracle on it). Returns (volume, heads per axis, options)."""
    import empanada_napari_b200.synthetic as syn
    rng = np.random.default_rng(7000 + seed)
    shape = (int(rng.integers(8, 22)), int(rng.integers(20, 50)), int(rng.integers(20, 50)))
    opt = dict(ks=int(rng.choice([1, 3, 5])), noise=float(rng.choice([0.0, 0.3, 0.6, 1.0])),
               nms_kernel=int(rng.choice([3, 3, 5, 7])), conf=float(rng.choice([0.3, 0.5])),
               scale=int(rng.choice([1, 1, 1, 2])), semantic_only=bool(rng.random() < 0.15),
               stuff_config=bool(rng.random() < 0.1), vote=int(rng.choice([1, 2, 2, 2, 3])),
               allow_one=bool(rng.random() < 0.2), erosion=int(rng.choice([0, 0, 0, 1])),
               dilation=int(rng.choice([0, 0, 0, 1])), fill=bool(rng.random() < 0.15),
               min_size=int(rng.choice([5, 20, 60])), min_extent=int(rng.choice([1, 2, 3])),
               batch=int(rng.choice([3, 5, 8])))
    vol, lab, _ = syn.make_volume(shape, seed=7000 + seed, n_objects=int(rng.integers(3, 16)), scale=1.0)
    sc, noise = opt["scale"], opt["noise"]

    def noisy(label_slice, pad_to):
        sem, ctr, off = syn.analytic_heads(label_slice, pad_to=pad_to)
        if noise > 0:
            sem = sem + rng.normal(0, 2.5 * noise, size=sem.shape)
            ctr = ctr + rng.normal(0, 0.05 * noise, size=ctr.shape)
            off = off + rng.normal(0, 2.0 * noise, size=off.shape)
        return sem.astype(np.float32), ctr.astype(np.float32), off.astype(np.float32)

    heads = {}
    for axis in range(3):
        hs = []
        for i in range(shape[axis]):
            sl = np.take(lab, i, axis=axis)
            if sc > 1:   # the network sees the down-sampled slice; PointRend renders back up
                sem = noisy(sl, 16 * sc)[0]
                _, ctr, off = noisy(sl[::sc, ::sc], 16)
            else:
                sem, ctr, off = noisy(sl, 16)
            hs.append((sem, ctr, off))
        heads[axis] = tuple(np.stack([h[k] for h in hs]) for k in range(3))
    return vol, heads, opt


def _consensus_or_error(gen):
    try:
        return [(v, i) for v, _, i in gen], None
    except Exception as e:       # the reference's own failure modes must be reproduced
        return None, type(e).__name__


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(8))
def test_random_option_combinations_vs_oracle(seed):
    """Random combinations of the engine options (median kernel, NMS kernel, thresholds,
    inference_scale, semantic-only, stuff config, vote threshold, one-view consensus, tracker
    morphology, filters, slice batch) on random noisy volumes: CUDA path vs oracle, bit-exact."""
    from empanada_napari_b200.inference import stack_postprocessing, tracker_consensus
    from oracle import consensus as ocons, pipeline
    vol, heads, o = random_engine_case(seed)
    if min(vol.shape) < o["ks"]:
        pytest.skip("stack shorter than the median kernel")
    eng, cfg = _engine(heads, config={"thing_list": []} if o["stuff_config"] else None,
                       median_kernel_size=o["ks"], nms_kernel=o["nms_kernel"], confidence_thr=o["conf"],
                       min_size=o["min_size"], min_extent=o["min_extent"], save_panoptic=True, batch_size=o["batch"],
                       semantic_only=o["semantic_only"], inference_scale=o["scale"], label_erosion=o["erosion"],
                       label_dilation=o["dilation"], fill_holes_in_segmentation=o["fill"])
    got, want = {}, {}
    for a, axis_name in enumerate(("xy", "xz", "yz")):
        sem, ctr, off = heads[a]
        stack, trs = eng.infer_on_axis(vol, axis_name)
        ostack, otrs = pipeline.infer_on_axis(
            vol, axis_name, lambda i, x: (sem[i], ctr[i], off[i]), cfg, median_kernel_size=o["ks"],
            nms_kernel=o["nms_kernel"], confidence_thr=o["conf"], min_size=o["min_size"], min_extent=o["min_extent"],
            semantic_only=o["semantic_only"], inference_scale=o["scale"], label_erosion=o["erosion"],
            label_dilation=o["dilation"], fill_holes_in_segmentation=o["fill"])
        assert_instances_equal(trs[0].instances, otrs[0].instances)
        assert np.array_equal(stack, ostack), (axis_name, o)
        got[axis_name], want[axis_name] = trs, otrs
    kw = dict(pixel_vote_thr=o["vote"], allow_one_view=o["allow_one"], min_size=o["min_size"],
              min_extent=o["min_extent"], dtype=np.int32)
    res, err = _consensus_or_error(tracker_consensus(got, None, cfg, **kw))
    ores, oerr = _consensus_or_error(ocons.tracker_consensus(want, cfg, **kw))
    assert err == oerr, (err, oerr, o)
    if err is None:
        for (v, inst), (ov, oinst) in zip(res, ores):
            assert_instances_equal(inst, oinst)
            assert np.array_equal(v, ov), o
    kw2 = dict(min_size=o["min_size"], min_extent=o["min_extent"], dtype=np.int32)
    for (v, _, inst), (ov, _, oinst) in zip(stack_postprocessing({"xy": got["xy"]}, None, cfg, **kw2),
                                            ocons.stack_postprocessing({"xy": want["xy"]}, cfg, **kw2)):
        assert_instances_equal(inst, oinst)
        assert np.array_equal(v, ov), o
