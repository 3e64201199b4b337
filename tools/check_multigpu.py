"""torchrun helper (N >= 2 GPUs): DistributedEngine3d + tracker_consensus on rank 0 must equal the
single-GPU Engine3d result bit for bit (trackers' boxes / sizes, dense volumes, consensus RLE).
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29533 tools/check_multigpu.py"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import empanada_napari_b200.synthetic as syn  # noqa: E402
from empanada_napari_b200 import multigpu  # noqa: E402
from empanada_napari_b200.inference import Engine3d, tracker_consensus  # noqa: E402
from empanada_napari_b200.model import SyntheticHeadsModel  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
shape = (50, 70, 61)
vol, lab, _ = syn.make_volume(shape, seed=31, scale=1.0)
heads = {}
for axis in range(3):
    hs = [syn.analytic_heads(np.take(lab, i, axis=axis), pad_to=16) for i in range(shape[axis])]
    heads[axis] = tuple(torch.from_numpy(np.stack([h[k] for h in hs]).astype(np.float32)).to(dev) for k in range(3))
pdl = None
if os.environ.get("CHECK_WITH_NETWORK", "1") == "1":   # run the real network too (results unused)
    from empanada_napari_b200.pdl import PDLModel
    pdl = PDLModel(syn.make_pdl_state_dict(0), dev)
cfg = {"class_names": {1: "mito"}, "labels": [1], "thing_list": [1], "padding_factor": 16,
       "norms": {"mean": 0.57571, "std": 0.12765}}
cfg["model"] = SyntheticHeadsModel(lambda a, s0, s1: (heads[a][0][s0:s1, 0], heads[a][1][s0:s1], heads[a][2][s0:s1]), inner=pdl)
kw = dict(median_kernel_size=3, nms_kernel=3, confidence_thr=0.5, min_size=30, min_extent=3, batch_size=4)
engine_cls = {"sharded": multigpu.ShardedEngine3d, "gather": multigpu.DistributedEngine3d}[os.environ.get("CHECK_ENGINE", "sharded")]
if os.environ.get("CHECK_FINE") == "1":
    kw["fine_boundaries"] = True
if os.environ.get("CHECK_KS"):
    kw["median_kernel_size"] = int(os.environ["CHECK_KS"])
deng = engine_cls(cfg, **kw)
trackers = {}
for name in ("xy", "xz", "yz"):
    _, trackers[name] = deng.infer_on_axis(vol, name)
trackers = deng.finalize(trackers)
ok = True
if rank == 0:
    seng = Engine3d(cfg, **kw)
    ref = {name: seng.infer_on_axis(vol, name)[1] for name in ("xy", "xz", "yz")}
    for name in ("xy", "xz", "yz"):
        a, b = trackers[name][0], ref[name][0]
        same = (list(a.instances.keys()) == list(b.instances.keys())
                and all(tuple(a.instances[k]["box"]) == tuple(b.instances[k]["box"]) for k in a.instances)
                and a._b200_sizes == b._b200_sizes and bool(torch.equal(a._b200_dense, b._b200_dense)))
        print(name, "instances", len(a.instances), "equal", same)
        ok &= same
    outs = []
    for trs in (trackers, ref):
        for v, _, inst in tracker_consensus(trs, None, cfg, pixel_vote_thr=2, min_size=30, min_extent=3, dtype=np.int32):
            outs.append((v.copy(), inst))
    same = (np.array_equal(outs[0][0], outs[1][0]) and list(outs[0][1].keys()) == list(outs[1][1].keys())
            and all(np.array_equal(outs[0][1][k]["starts"], outs[1][1][k]["starts"])
                    and np.array_equal(outs[0][1][k]["runs"], outs[1][1][k]["runs"]) for k in outs[0][1]))
    print("consensus instances", len(outs[0][1]), "equal", same)
    ok &= same and len(outs[0][1]) > 0
# the z-slab sharded consensus (every rank votes on its own slab) against the same reference
if engine_cls is multigpu.ShardedEngine3d:
    # gather_dense=False: planes stay sharded and are completed behind the next plane's forward pass
    deng2 = multigpu.ShardedEngine3d(cfg, gather_dense=False, **kw)
    trackers = {}
    for name in ("xy", "xz", "yz"):
        _, trackers[name] = deng2.infer_on_axis(vol, name)
    trackers = deng2.finalize(trackers)
    if rank == 0:
        for name in ("xy", "xz", "yz"):
            a, b = trackers[name][0], ref[name][0]
            same = (list(a.instances.keys()) == list(b.instances.keys()) and a._b200_sizes == b._b200_sizes
                    and all(tuple(a.instances[k]["box"]) == tuple(b.instances[k]["box"]) for k in a.instances))
            print(name, "sharded tables equal", same)
            ok &= bool(same)
    v, _, inst = deng2.sharded_consensus(trackers, cfg, pixel_vote_thr=2, min_size=30, min_extent=3)
    if rank == 0:
        rv, rinst = outs[1]
        same = (np.array_equal(v.cpu().numpy(), rv) and list(inst.keys()) == list(rinst.keys())
                and all(np.array_equal(inst[k]["starts"], rinst[k]["starts"]) and np.array_equal(inst[k]["runs"], rinst[k]["runs"])
                        and tuple(inst[k]["box"]) == tuple(rinst[k]["box"]) for k in inst))
        print("sharded consensus instances", len(inst), "equal", same)
        ok &= bool(same)
flag = torch.tensor([1 if ok else 0], device=dev)
dist.broadcast(flag, 0)
dist.destroy_process_group()
if rank == 0:
    print("MULTIGPU_CHECK", engine_cls.__name__, "PASS" if ok else "FAIL")
sys.exit(0 if int(flag.item()) == 1 else 1)
