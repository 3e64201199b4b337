"""`__graft_entry__.smoke()`: one small invocation of the whole hot path on cuda:0 (network
forward on the tcgen05 kernels, post-processing, tracking, consensus) checked against the CPU
oracle. The oracle is imported here only as the checker."""
import numpy as np


def smoke():
    import torch
    from . import synthetic as syn
    from .inference import Engine3d, tracker_consensus
    from .model import SyntheticHeadsModel
    from .pdl import PDLModel
    from oracle import consensus as ocons, model as omodel, pipeline, post

    if not torch.cuda.is_available():
        raise RuntimeError("smoke() needs a CUDA device")
    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    norms = {"mean": 0.57571, "std": 0.12765}
    cfg = {"class_names": {1: "mito"}, "labels": [1], "thing_list": [1], "padding_factor": 16,
           "norms": norms, "model": None}
    shape = (24, 48, 40)
    vol, lab, _ = syn.make_volume(shape, seed=5, n_objects=8, scale=1.0)

    # (1) network forward vs the fp32 oracle
    sd = syn.make_pdl_state_dict(0)
    pdl = PDLModel(sd, dev)
    vol_d = torch.from_numpy(vol).to(dev)
    sem, ctr, off = pdl.forward_slices(vol_d, 0, 0, 2, norms, 16)
    x = np.stack([post.factor_pad(post.normalize(vol[i], norms["mean"], norms["std"]), 16) for i in range(2)])[:, None]
    ref = omodel.pdl_forward(sd, torch.from_numpy(x), 2, False)
    for got, want, name in ((ctr.cpu().numpy(), ref["ctr_hmp"].numpy()[:, 0], "ctr_hmp"),
                            (off.cpu().numpy(), ref["offsets"].numpy(), "offsets")):
        err = np.linalg.norm(got - want) / (np.linalg.norm(want) + 1e-12)
        if not err < 3e-2:
            raise AssertionError(f"forward mismatch on {name}: rel L2 {err}")

    # (1b) PanopticBiFPN-PointRend (MitoNet_v1_mini architecture, padding factor 128)
    from .bifpn import BiFPNModel
    sdb = syn.make_bifpn_state_dict(0)
    bif = BiFPNModel(sdb, dev)
    _, ctr_b, off_b = bif.forward_slices(vol_d, 0, 0, 1, norms, 128)
    xb = post.factor_pad(post.normalize(vol[0], norms["mean"], norms["std"]), 128)[None, None]
    refb = omodel.bifpn_forward(sdb, torch.from_numpy(xb), 2, False)
    for got, want, name in ((ctr_b.cpu().numpy(), refb["ctr_hmp"].numpy()[:, 0], "bifpn ctr_hmp"),
                            (off_b.cpu().numpy(), refb["offsets"].numpy(), "bifpn offsets")):
        err = np.linalg.norm(got - want) / (np.linalg.norm(want) + 1e-12)
        if not err < 3e-2:
            raise AssertionError(f"forward mismatch on {name}: rel L2 {err}")

    # (2) post-processing + tracking + consensus vs the oracle, bit exact
    heads = {}
    for a in range(3):
        hs = [syn.analytic_heads(np.take(lab, i, axis=a), pad_to=16) for i in range(shape[a])]
        heads[a] = tuple(np.stack([h[j] for h in hs]) for j in range(3))

    def heads_fn(axis, s0, s1):
        s, c, o = heads[axis]
        return (torch.from_numpy(np.ascontiguousarray(s[s0:s1, 0])).to(dev),
                torch.from_numpy(np.ascontiguousarray(c[s0:s1])).to(dev),
                torch.from_numpy(np.ascontiguousarray(o[s0:s1])).to(dev))

    cfg_g = dict(cfg)
    cfg_g["model"] = SyntheticHeadsModel(heads_fn, inner=pdl)
    eng = Engine3d(cfg_g, median_kernel_size=3, confidence_thr=0.5, min_size=20, min_extent=2, batch_size=8)
    got, want = {}, {}
    for a, name in enumerate(("xy", "xz", "yz")):
        _, got[name] = eng.infer_on_axis(vol, name)
        s, c, o = heads[a]
        _, want[name] = pipeline.infer_on_axis(vol, name, lambda i, x: (s[i], c[i], o[i]), cfg,
                                               median_kernel_size=3, confidence_thr=0.5, min_size=20,
                                               min_extent=2, save_panoptic=False)
        if list(got[name][0].instances.keys()) != list(want[name][0].instances.keys()):
            raise AssertionError(f"tracker labels differ on plane {name}")
    (v, _, inst), = list(tracker_consensus(got, None, cfg_g, min_size=20, min_extent=2, dtype=np.int32))
    (ov, _, oinst), = list(ocons.tracker_consensus(want, cfg, min_size=20, min_extent=2, dtype=np.int32))
    if not np.array_equal(v, ov) or list(inst.keys()) != list(oinst.keys()):
        raise AssertionError("consensus volume differs from the oracle")
    print(f"smoke ok: forward within tolerance, {len(inst)} consensus instances bit-exact vs oracle")
