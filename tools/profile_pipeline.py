"""Phase breakdown of one orthoplane job (wall clock with synchronisation). Diagnostic only."""
import os, sys, json, time
os.environ["B200_EMPANADA_PROFILE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
import empanada_napari_b200.synthetic as syn
from empanada_napari_b200 import consensus
from empanada_napari_b200.inference import Engine3d, tracker_consensus
from empanada_napari_b200.model import SyntheticHeadsModel
from empanada_napari_b200.pdl import PDLModel
S = int(sys.argv[1]) if len(sys.argv) > 1 else 512
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
dev = torch.device("cuda:0")
t0 = time.perf_counter()
vol_d, lab_d, n_obj = bench.synth_on_device(S, dev)
heads = {a: bench.analytic_heads_on_device(lab_d, a, n_obj) for a in range(3)}
torch.cuda.synchronize(); print("setup s", time.perf_counter() - t0)
pdl = PDLModel(syn.make_pdl_state_dict(0), dev)
cfg = dict(bench.MODEL_CONFIG)
cfg["model"] = SyntheticHeadsModel(lambda a, s0, s1: tuple(t[s0:s1] for t in heads[a]), inner=pdl)
eng = Engine3d(cfg, median_kernel_size=3, nms_kernel=3, confidence_thr=0.5, min_size=500, min_extent=5, batch_size=B)
for it in range(2):
    trackers = {}
    tot = {}
    t0 = time.perf_counter()
    for name in ("xy", "xz", "yz"):
        _, trackers[name] = eng.infer_on_axis(vol_d, name)
        for k, v in eng.last_profile.items():
            tot[k] = tot.get(k, 0) + v
        print(name, {k: round(v, 3) for k, v in eng.last_profile.items()}, "instances", len(trackers[name][0].instances),
              "runs", sum(len(i["runs"]) for i in trackers[name][0].instances.values()))
    t1 = time.perf_counter()
    for vol, cname, inst in tracker_consensus(trackers, None, cfg, pixel_vote_thr=2, min_size=500, min_extent=5, dtype=np.int32, to_host=False):
        pass
    torch.cuda.synchronize(); t2 = time.perf_counter()
    print("planes total", round(t1 - t0, 3), {k: round(v, 3) for k, v in tot.items()})
    print("consensus", round(t2 - t1, 3), {k: round(v, 3) for k, v in consensus.LAST_PROFILE.items()}, "instances", len(inst))
