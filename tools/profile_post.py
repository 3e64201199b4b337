"""ncu helper: post-processing kernels of one plane (N slices of SxS) inside a cudaProfiler window."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from empanada_napari_b200.postproc import PlanePost
from empanada_napari_b200 import tracking
S = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
N = int(sys.argv[2]) if len(sys.argv) > 2 else 32
dev = torch.device("cuda:0")
# a slab of the 1024^3 synthetic volume: same object density as the benchmark
import empanada_napari_b200.synthetic as syn
ell = syn.make_ellipsoids((S, S, S), seed=0)
lab = torch.zeros((N, S, S), dtype=torch.int32, device=dev)
ar = torch.arange(S, device=dev, dtype=torch.float32)
for i, (cz, cy, cx, rz, ry, rx) in enumerate(ell.tolist(), start=1):
    cz -= S // 2 - N // 2
    z0, z1 = max(0, int(cz - rz)), min(N, int(cz + rz) + 2)
    y0, y1 = max(0, int(cy - ry)), min(S, int(cy + ry) + 2)
    x0, x1 = max(0, int(cx - rx)), min(S, int(cx + rx) + 2)
    if z0 >= z1 or y0 >= y1 or x0 >= x1:
        continue
    m = (((ar[z0:z1] - cz) / rz) ** 2)[:, None, None] + (((ar[y0:y1] - cy) / ry) ** 2)[None, :, None] + (((ar[x0:x1] - cx) / rx) ** 2)[None, None, :] <= 1.0
    lab[z0:z1, y0:y1, x0:x1][m] = i
sem, ctr, off = bench.analytic_heads_on_device(lab, 0, len(ell))
for rep in range(2):
    if rep == 1:
        torch.cuda.synchronize(); torch.cuda.profiler.start()
    post = PlanePost(N, S, S, S, S, ks=3, thing_class=1, label_divisor=1000, confidence_thr=0.5, device=dev)
    for s0 in range(0, N, 8):
        post.push_heads(sem[s0:s0 + 8], ctr[s0:s0 + 8], off[s0:s0 + 8])
    post.finish_heads()
    post.run_cc()
    lut, labels, sizes, boxes = post.replay("xy")
    dense = post.relabel(lut, "xy", (N, S, S))
    inst = post.tracker_instances("xy", (N, S, S), lut, labels, boxes, dense)
    if rep == 1:
        torch.cuda.synchronize(); torch.cuda.profiler.stop()
print("instances", len(inst), "n_cc mean", float(post.n_cc.float().mean()))
