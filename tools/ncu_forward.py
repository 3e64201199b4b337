"""ncu helper: exactly one forward launch list (B slices of SxS) inside a cudaProfiler window."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import empanada_napari_b200.synthetic as syn
from empanada_napari_b200.pdl import PDLModel

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
S = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
dev = torch.device("cuda:0")
m = PDLModel(syn.make_pdl_state_dict(0), dev)
vol = torch.randint(0, 256, (B, S, S), dtype=torch.uint8, device=dev)
norms = {"mean": 0.57571, "std": 0.12765}
for _ in range(3):
    m.forward_slices(vol, 0, 0, B, norms, 16)
torch.cuda.synchronize()
torch.cuda.profiler.start()
m.forward_slices(vol, 0, 0, B, norms, 16)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ops", len(m.last_plan.op_info), "launches", m.last_plan.launches)
for i, ((k, fl), d) in enumerate(zip(m.last_plan.op_info, m.last_plan.op_desc)):
    print(i, k, fl, d)
