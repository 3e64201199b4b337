"""Single-process check of the widget-facing `MultiGPUEngine3d` (needs >= 2 GPUs): constructed
with the reference's keywords, it must return the same trackers and label volumes as the
single-GPU `Engine3d`, and `tracker_consensus` on them the same consensus.
    python tools/check_multigpu_front.py [world]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    import empanada_napari_b200.synthetic as syn
    from empanada_napari_b200.inference import Engine3d, tracker_consensus
    from empanada_napari_b200.model import HostHeadsModel
    from empanada_napari_b200.multigpu import MultiGPUEngine3d
    world = int(sys.argv[1]) if len(sys.argv) > 1 else min(4, torch.cuda.device_count())
    shape = (50, 70, 61)
    vol, lab, _ = syn.make_volume(shape, seed=31, scale=1.0)
    heads = {}
    for axis in range(3):
        hs = [syn.analytic_heads(np.take(lab, i, axis=axis), pad_to=16) for i in range(shape[axis])]
        heads[axis] = (np.stack([h[0][0] for h in hs]), np.stack([h[1] for h in hs]), np.stack([h[2] for h in hs]))
    cfg = {"class_names": {1: "mito"}, "labels": [1], "thing_list": [1], "padding_factor": 16,
           "norms": {"mean": 0.57571, "std": 0.12765}, "model": HostHeadsModel(heads)}
    kw = dict(median_kernel_size=3, nms_kernel=3, confidence_thr=0.5, min_size=30, min_extent=3)
    meng = MultiGPUEngine3d(cfg, save_panoptic=False, world_size=world, batch_size=4, **kw)
    seng = Engine3d(cfg, save_panoptic=True, batch_size=4, **kw)
    ok = True
    got, ref, rstacks = {}, {}, {}
    # the widget's orthoplane sequence (_volume_inference.py:331-346): three planes, instance
    # counts read in between, then tracker_consensus on the dictionary of trackers
    for name in ("xy", "xz", "yz"):
        stack, got[name] = meng.infer_on_axis(vol, name)
        n_inst = len(got[name][0].instances.keys())
        rstacks[name], ref[name] = seng.infer_on_axis(vol, name)
        a, b = got[name][0], ref[name][0]
        same = (stack is None and list(a.instances.keys()) == list(b.instances.keys())
                and all(tuple(a.instances[k]["box"]) == tuple(b.instances[k]["box"]) for k in a.instances))
        print(name, "instances", n_inst, "tables equal", same, flush=True)
        ok &= bool(same)
    outs = []
    for trs in (got, ref):      # `got`: label volumes still sharded -> every GPU votes on its z-slab
        for v, _, inst in tracker_consensus(trs, None, cfg, pixel_vote_thr=2, min_size=30, min_extent=3, dtype=np.int32):
            outs.append((v.copy(), inst))
    same = (np.array_equal(outs[0][0], outs[1][0]) and list(outs[0][1].keys()) == list(outs[1][1].keys())
            and all(np.array_equal(outs[0][1][k]["starts"], outs[1][1][k]["starts"])
                    and np.array_equal(outs[0][1][k]["runs"], outs[1][1][k]["runs"]) for k in outs[0][1]))
    print("consensus instances", len(outs[0][1]), "equal", same, flush=True)
    ok &= bool(same) and len(outs[0][1]) > 0
    # per-plane run-length tables are fetched from the ranks on first use
    for name in ("xy", "xz", "yz"):
        a, b = got[name][0], ref[name][0]
        same = all(np.array_equal(a.instances[k]["starts"], b.instances[k]["starts"])
                   and np.array_equal(a.instances[k]["runs"], b.instances[k]["runs"]) for k in a.instances)
        print(name, "lazily fetched RLE equal", same, flush=True)
        ok &= bool(same)
    # stacks on request (save_panoptic)
    meng.save_panoptic = True
    stack, _ = meng.infer_on_axis(vol, "yz")
    same = stack is not None and stack.dtype == np.int32 and np.array_equal(stack, rstacks["yz"])
    print("yz stack equal", same, flush=True)
    ok &= bool(same)
    # a second, different volume of the same shape through the same engine (re-upload + broadcast)
    vol2 = np.ascontiguousarray(vol[::-1])
    _, t2 = meng.infer_on_axis(vol2, "xy")
    print("second volume instances", len(t2[0].instances))
    meng.close()
    print("MULTIGPU_FRONT", "PASS" if ok else "FAIL")
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
