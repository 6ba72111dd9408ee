"""Measurement helper: the shipped batch size (intrinsic_cat.yaml:32, 4096 rays) -- eager and CUDA-graph step times."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, ctypes as C
import inf_b200
from inf_b200 import trainer, _lib
import bench
dev = torch.device("cuda")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
faces = bench.torus_faces(224, dev)
table = inf_b200.EigenTable.from_dense(bench.make_table(224 * 224, 1023, dev), device=dev)
torch.manual_seed(0)
m = inf_b200.make_model(bench.MODEL_CFG).to(dev)
o = trainer.FusedFieldAdam(m, lr=1e-4)
g = trainer.GraphedTrainStep(m, o)
v, b, t = bench.make_rays(faces, B, dev, seed=7)
sb = {"eigenfunction_table": table, "vertex_idxs_of_hit_faces": v, "barycentric_coords": b, "expected_rgbs": t}
lib = _lib.load()
def timed(fn, n=100):
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n
for _ in range(5): trainer.fused_train_step_async(m, o, sb)
lib.inf_profile_enable(1)
ms = timed(lambda: trainer.fused_train_step_async(m, o, sb))
arr = (C.c_float * 6)(); n = C.c_int(); lib.inf_profile_read_stages(arr, C.byref(n)); lib.inf_profile_enable(0)
for _ in range(4): g.step_async(sb)
msg = timed(lambda: g.step_async(sb))
print(f"B={B}: eager {ms:.4f} ms  graph {msg:.4f} ms  ({B / msg / 1e3:.1f} M rays/s)  stages {dict(zip(bench.STAGES, [round(x, 4) for x in arr]))}")
