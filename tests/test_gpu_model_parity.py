"""-m gpu: the bf16 tcgen05 forward pass against the fp32 PyTorch oracle (oracle/model.py, itself
pinned bit-exactly against the reference classes). Tolerances are stated per tensor: relative
L2 error <= 3e-2 for intermediate features and /4-resolution heads (bf16 activations through
~55 layers); full-resolution PointRend logits are compared with a mismatch budget because which
8192 points get re-predicted depends on top-k tie/near-tie order."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

NORMS = {"mean": 0.57571, "std": 0.12765}


def rel_l2(a, b):
    return float(np.linalg.norm(a.astype(np.float64) - b.astype(np.float64)) / (np.linalg.norm(b.astype(np.float64)) + 1e-12))


@pytest.mark.parametrize("h,w,B", [(96, 80, 2), (120, 200, 3), (256, 256, 1)])
def test_pdl_forward_vs_oracle(h, w, B):
    import torch
    import empanada_napari_b200.synthetic as syn
    from empanada_napari_b200.pdl import PDLModel
    from oracle import model as omodel, post
    sd = syn.make_pdl_state_dict(0)
    dev = torch.device("cuda:0")
    m = PDLModel(sd, dev)
    rng = np.random.default_rng(5)
    vol = rng.integers(0, 256, (B, h, w), dtype=np.uint8)
    vol_d = torch.from_numpy(vol).to(dev)
    sem, ctr, off = m.forward_slices(vol_d, 0, 0, B, NORMS, 16)
    torch.cuda.synchronize()
    plan = m.last_plan
    x = np.stack([post.factor_pad(post.normalize(vol[i], NORMS["mean"], NORMS["std"]), 16) for i in range(B)])[:, None]
    ref = omodel.pdl_forward(sd, torch.from_numpy(x), 2, False)
    nhwc = lambda t: t.float().cpu().numpy().transpose(0, 3, 1, 2)
    errs = {
        "p2": rel_l2(nhwc(plan.p2), ref["p2"].numpy()),
        "p5": rel_l2(nhwc(plan.p5), ref["p5"].numpy()),
        "semantic_x": rel_l2(nhwc(plan.semantic_x), ref["semantic_x"].numpy()),
        "instance_x": rel_l2(nhwc(plan.instance_x), ref["instance_x"].numpy()),
        "coarse": rel_l2(plan.coarse.cpu().numpy(), ref["coarse_logits"].numpy()),
        "ctr": rel_l2(ctr.cpu().numpy(), ref["ctr_hmp"].numpy()[:, 0]),
        "off": rel_l2(off.cpu().numpy(), ref["offsets"].numpy()),
    }
    print(errs)
    for k, v in errs.items():
        assert v <= 3e-2, (k, v, errs)
    got = sem.cpu().numpy()
    want = ref["sem_logits"].numpy()[:, 0]
    scale = float(np.abs(want).mean())
    bad = np.abs(got - want) > 0.1 * scale + 0.05 * np.abs(want)
    frac = float(bad.mean())
    print("sem_logits rel_l2", rel_l2(got, want), "mismatch fraction", frac)
    assert frac <= 0.03, frac


def test_engine2d_end_to_end_agreement():
    """Engine2d.infer with the real forward pass vs the oracle pipeline fed by the oracle model:
    instance maps agree (pixel agreement >= 0.99 on the foreground mask)."""
    import torch
    import empanada_napari_b200.synthetic as syn
    from empanada_napari_b200.inference import Engine2d
    from oracle import model as omodel, pipeline
    sd = syn.make_pdl_state_dict(0)
    cfg = {"class_names": {1: "mito"}, "labels": [1], "thing_list": [1], "padding_factor": 16,
           "norms": NORMS, "model": sd}
    eng = Engine2d(cfg, confidence_thr=0.5, nms_threshold=0.1, nms_kernel=3)
    img, _, _ = syn.make_volume((1, 200, 184), seed=3, scale=1.0)
    img = img[0]
    out = eng.infer(img)
    assert out.shape == img.shape and out.dtype == np.int32

    def heads_fn(i, x):
        o = omodel.pdl_forward(sd, torch.from_numpy(x[None, None]), 2, False)
        return o["sem_logits"][0].numpy(), o["ctr_hmp"][0, 0].numpy(), o["offsets"][0].numpy()

    want = pipeline.engine2d_infer(img, heads_fn, cfg, confidence_thr=0.5, nms_threshold=0.1, nms_kernel=3)
    fg_agree = float(((out > 0) == (want > 0)).mean())
    print("foreground agreement", fg_agree, "objects", len(np.unique(out)) - 1, len(np.unique(want)) - 1)
    assert fg_agree >= 0.99


@pytest.mark.parametrize("h,w,B", [(100, 200, 2), (256, 128, 1)])
def test_bifpn_forward_vs_oracle(h, w, B):
    """PanopticBiFPN-PointRend (MitoNet_v1_mini architecture, padding factor 128) on the bf16
    tcgen05 path vs the fp32 oracle restatement (pinned on tests/golden/model_bifpn_tiny.npz).
    Same tolerances as the PanopticDeepLab test."""
    import torch
    import empanada_napari_b200.synthetic as syn
    from empanada_napari_b200.bifpn import BiFPNModel
    from oracle import model as omodel, post
    sd = syn.make_bifpn_state_dict(0)
    dev = torch.device("cuda:0")
    m = BiFPNModel(sd, dev)
    rng = np.random.default_rng(6)
    vol = rng.integers(0, 256, (B, h, w), dtype=np.uint8)
    sem, ctr, off = m.forward_slices(torch.from_numpy(vol).to(dev), 0, 0, B, NORMS, 128)
    torch.cuda.synchronize()
    plan = m.last_plan
    x = np.stack([post.factor_pad(post.normalize(vol[i], NORMS["mean"], NORMS["std"]), 128) for i in range(B)])[:, None]
    ref = omodel.bifpn_forward(sd, torch.from_numpy(x), 2, False)
    nhwc = lambda t: t.float().cpu().numpy().transpose(0, 3, 1, 2)
    D = plan.W.D
    errs = {
        "p2": rel_l2(nhwc(plan.p2), ref["p2"].numpy()),
        "p5": rel_l2(nhwc(plan.p5), ref["p5"].numpy()),
        "p2_resampled": rel_l2(nhwc(plan.p2f[..., D:]), ref["p2f"].numpy()),
        "semantic_x": rel_l2(nhwc(plan.semantic_x), ref["semantic_x"].numpy()),
        "instance_x": rel_l2(nhwc(plan.instance_x), ref["instance_x"].numpy()),
        "coarse": rel_l2(plan.coarse.cpu().numpy(), ref["coarse_logits"].numpy()),
        "ctr": rel_l2(ctr.cpu().numpy(), ref["ctr_hmp"].numpy()[:, 0]),
        "off": rel_l2(off.cpu().numpy(), ref["offsets"].numpy()),
    }
    print(errs)
    for k, v in errs.items():
        assert v <= 3e-2, (k, v, errs)
    got = sem.cpu().numpy()
    want = ref["sem_logits"].numpy()[:, 0]
    scale = float(np.abs(want).mean())
    bad = np.abs(got - want) > 0.1 * scale + 0.05 * np.abs(want)
    frac = float(bad.mean())
    print("sem_logits rel_l2", rel_l2(got, want), "mismatch fraction", frac)
    assert frac <= 0.03, frac


def test_engine2d_bifpn_mini_config():
    """BASELINE config C1: MitoNet_v1_mini-class 2D inference (padding factor 128, nms_kernel 7)
    through Engine2d vs the oracle pipeline fed by the oracle BiFPN."""
    import torch
    import empanada_napari_b200.synthetic as syn
    from empanada_napari_b200.inference import Engine2d
    from oracle import model as omodel, pipeline
    sd = syn.make_bifpn_state_dict(0)
    cfg = {"class_names": {1: "mito"}, "labels": [1], "thing_list": [1], "padding_factor": 128,
           "norms": NORMS, "model": sd}
    eng = Engine2d(cfg, confidence_thr=0.5, nms_threshold=0.1, nms_kernel=7)
    img, _, _ = syn.make_volume((1, 200, 184), seed=3, scale=1.0)
    img = img[0]
    out = eng.infer(img)
    assert out.shape == img.shape and out.dtype == np.int32

    def heads_fn(i, x):
        o = omodel.bifpn_forward(sd, torch.from_numpy(x[None, None]), 2, False)
        return o["sem_logits"][0].numpy(), o["ctr_hmp"][0, 0].numpy(), o["offsets"][0].numpy()

    want = pipeline.engine2d_infer(img, heads_fn, cfg, confidence_thr=0.5, nms_threshold=0.1, nms_kernel=7)
    fg_agree = float(((out > 0) == (want > 0)).mean())
    print("foreground agreement", fg_agree, "objects", len(np.unique(out)) - 1, len(np.unique(want)) - 1)
    assert fg_agree >= 0.99


@pytest.mark.parametrize("net", ["pdl", "bifpn"])
def test_graph_replay_equals_plain_replay(net):
    """Small launch lists are replayed as one CUDA graph after the first call; every replay must
    give exactly the bits of the plain launch sequence (different slices of the volume per call)."""
    import torch
    import empanada_napari_b200.synthetic as syn
    from empanada_napari_b200 import pdl as pdl_mod
    from empanada_napari_b200.bifpn import BiFPNModel
    from empanada_napari_b200.pdl import PDLModel
    dev = torch.device("cuda:0")
    sd = syn.make_pdl_state_dict(0) if net == "pdl" else syn.make_bifpn_state_dict(0)
    cls, pf = (PDLModel, 16) if net == "pdl" else (BiFPNModel, 128)
    vol = torch.randint(0, 256, (6, 128, 128), dtype=torch.uint8, device=dev, generator=torch.Generator(device=dev).manual_seed(1))
    outs = {}
    for use in (True, False):
        pdl_mod.USE_GRAPHS = use
        m = cls(sd, dev)
        res = []
        for s0 in (0, 2, 4, 2):
            sem, ctr, off = m.forward_slices(vol, 0, s0, s0 + 2, NORMS, pf)
            res.append((sem.clone(), ctr.clone(), off.clone()))
        torch.cuda.synchronize()
        outs[use] = res
    pdl_mod.USE_GRAPHS = True
    for a, b in zip(outs[True], outs[False]):
        for x, y in zip(a, b):
            assert torch.equal(x, y)
    # and replays of the same slices agree with each other
    for x, y in zip(outs[True][1], outs[True][3]):
        assert torch.equal(x, y)


def test_engine3d_end_to_end_agreement():
    """Whole 3-D path with the network's OWN heads (no substitution): bf16 tcgen05 forward +
    CUDA post-processing + consensus vs the fp32 oracle model + oracle pipeline on the same
    volume. Voxel agreement of the consensus foreground >= 0.99 and instance F1 (IoU >= 0.5
    matching) >= 0.9 when there are instances to match."""
    import torch
    import empanada_napari_b200.synthetic as syn
    from empanada_napari_b200.inference import Engine3d, tracker_consensus
    from oracle import consensus as ocons, model as omodel, pipeline
    sd = syn.make_pdl_state_dict(0)
    cfg = {"class_names": {1: "mito"}, "labels": [1], "thing_list": [1], "padding_factor": 16,
           "norms": NORMS, "model": sd}
    shape = (20, 48, 40)
    vol, _, _ = syn.make_volume(shape, seed=9, scale=1.0)
    kw = dict(median_kernel_size=3, nms_kernel=3, confidence_thr=0.5, min_size=8, min_extent=1)
    eng = Engine3d(cfg, batch_size=4, **kw)
    got = {ax: eng.infer_on_axis(vol, ax)[1] for ax in ("xy", "xz", "yz")}

    def heads_fn(i, x):
        o = omodel.pdl_forward(sd, torch.from_numpy(x[None, None]), 2, False)
        return o["sem_logits"][0].numpy(), o["ctr_hmp"][0, 0].numpy(), o["offsets"][0].numpy()

    want = {ax: pipeline.infer_on_axis(vol, ax, heads_fn, cfg, save_panoptic=False, **kw)[1] for ax in ("xy", "xz", "yz")}
    (v, _, inst), = list(tracker_consensus(got, None, cfg, pixel_vote_thr=2, min_size=8, min_extent=1, dtype=np.int32))
    (ov, _, oinst), = list(ocons.tracker_consensus(want, cfg, pixel_vote_thr=2, min_size=8, min_extent=1, dtype=np.int32))
    fg = float(((v > 0) == (ov > 0)).mean())
    # instance matching at IoU >= 0.5
    tp = 0
    for l in np.unique(ov)[1:]:
        m = ov == l
        cand, cnt = np.unique(v[m], return_counts=True)
        for c, n in zip(cand, cnt):
            if c > 0 and n / float((m | (v == c)).sum()) >= 0.5:
                tp += 1
                break
    n_got, n_want = len(np.unique(v)) - 1, len(np.unique(ov)) - 1
    f1 = 2 * tp / max(1, n_got + n_want)
    print("foreground agreement", fg, "instances", n_got, n_want, "matched", tp, "F1", f1)
    assert fg >= 0.99
    if min(n_got, n_want) >= 3:
        assert f1 >= 0.9
