"""ncu helper: tracker_consensus of a S^3 orthoplane job inside a cudaProfiler window."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
import empanada_napari_b200.synthetic as syn
from empanada_napari_b200.inference import Engine3d, tracker_consensus
from empanada_napari_b200.model import SyntheticHeadsModel

S = int(sys.argv[1]) if len(sys.argv) > 1 else 512
dev = torch.device("cuda:0")
vol_d, lab_d, n_obj = bench.synth_on_device(S, dev)
heads = {a: bench.analytic_heads_on_device(lab_d, a, n_obj) for a in range(3)}
cfg = dict(bench.MODEL_CONFIG)
cfg["model"] = SyntheticHeadsModel(lambda a, s0, s1: tuple(t[s0:s1] for t in heads[a]), inner=None)
eng = Engine3d(cfg, median_kernel_size=3, nms_kernel=3, confidence_thr=0.5, min_size=500, min_extent=5, batch_size=16)
trackers = {name: eng.infer_on_axis(vol_d, name)[1] for name in ("xy", "xz", "yz")}
for rep in range(2):
    if rep == 1:
        torch.cuda.synchronize(); torch.cuda.profiler.start()
    for vol, cname, inst in tracker_consensus(trackers, None, cfg, pixel_vote_thr=2, min_size=500, min_extent=5, dtype=np.int32, to_host=False):
        pass
    if rep == 1:
        torch.cuda.synchronize(); torch.cuda.profiler.stop()
print("consensus instances", len(inst))
