"""Measurement helper: the optimiser tail of field_reduce_kernel on its own (inf_adam_pack_step: Adam + operand refresh from
a given flat gradient) against the full reduction stage of a training step."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, ctypes as C
import inf_b200
from inf_b200 import trainer, _lib
import bench
dev = torch.device("cuda")
torch.manual_seed(0)
m = inf_b200.make_model(bench.MODEL_CFG).to(dev)
o = trainer.FusedFieldAdam(m, lr=1e-4)
flat = torch.randn(o.flat.numel(), device=dev) * 1e-3
def timed(fn, n=200):
    for _ in range(5): fn()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n * 1e3
step_dev = o.step_dev if hasattr(o, "step_dev") else torch.ones(1, dtype=torch.int32, device=dev)
fn = lambda: torch.ops.inf_b200.field_adam_pack(flat, o.packed, m.cfg.key(), o.flat, o.exp_avg, o.exp_avg_sq, 1e-4, 0.9, 0.999, 1e-8, step_dev)
print(f"adam tail only (back-to-back launches): {timed(fn):.2f} us per call")
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    for _ in range(10): fn()
print(f"adam tail only (graph of 10): {timed(g.replay, 50) / 10:.2f} us per call")
