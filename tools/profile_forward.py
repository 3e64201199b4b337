"""Per-op device times of the forward launch list (CUDA events between ops). Diagnostic only."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import empanada_napari_b200.synthetic as syn
from empanada_napari_b200.pdl import PDLModel

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
S = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
dev = torch.device("cuda:0")
m = PDLModel(syn.make_pdl_state_dict(0), dev)
vol = torch.randint(0, 256, (B, S, S), dtype=torch.uint8, device=dev)
norms = {"mean": 0.57571, "std": 0.12765}
for _ in range(3):
    m.forward_slices(vol, 0, 0, B, norms, 16)
torch.cuda.synchronize()
plan = m.last_plan
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    m.forward_slices(vol, 0, 0, B, norms, 16)
e1.record(); torch.cuda.synchronize()
total = e0.elapsed_time(e1) / 5
ms = plan.run_timed(vol, (S * S, S, 1), 0)
ms = plan.run_timed(vol, (S * S, S, 1), 0)
agg = {}
for (kind, fl), t in zip(plan.op_info, ms):
    a = agg.setdefault(kind, [0.0, 0.0, 0]); a[0] += t; a[1] += fl; a[2] += 1
print(f"B={B} {S}x{S}: forward {total:.3f} ms/batch = {total/B:.3f} ms/slice; sum of per-op {ms.sum():.3f} ms")
for k, (t, fl, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"  {k:16s} n={n:3d} {t:8.3f} ms  {fl/1e9:9.1f} GF  {fl/t*1e-9 if t>0 else 0:8.1f} TF/s")
convs = [(t, fl, i) for i, ((kind, fl), t) in enumerate(zip(plan.op_info, ms)) if kind == "conv"]
print("all ops (ideal = max(bytes/6.5TB/s, flops/1400TF/s)):")
gap = 0.0
for i, ((kind, fl), t) in enumerate(zip(plan.op_info, ms)):
    d = plan.op_desc[i]
    ideal = 0.0
    if kind == "conv":
        mb = float(d.split("bytes=")[1][:-2])
        ideal = max(mb * 1e6 / 6.5e12, fl / 1.4e15) * 1e3
        gap += t - ideal
    print(f"   op{i:3d} {t:7.3f} ms ideal {ideal:6.3f} {fl/1e9:8.1f} GF {fl/t*1e-9 if t > 0 else 0:8.1f} TF/s  {d}")
print("conv gap to ideal (ms/batch):", gap)
tot_fl = sum(fl for _, fl, _ in convs)
print(f"conv total {sum(t for t,_,_ in convs):.3f} ms, {tot_fl/1e9:.1f} GF/batch = {tot_fl/B/1e9:.2f} GF/slice; whole-forward {tot_fl/total*1e-9:.1f} TF/s")
