"""Measurement helper: held-out PSNR of the 2000-step protocols on both evaluation paths (what tests/test_gpu_parity.py asserts
against the reference ensemble), printed instead of asserted."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import json
from inf_b200 import ops
import test_gpu_parity as T
ens = json.load(open(os.path.join(ROOT, "tests", "golden", "psnr_ensemble.json")))
for name, fused in (("cat_2000", True), ("view_2000", True), ("view_2000", False)):
    lo, hi = min(e["psnr"] for e in ens[name]), max(e["psnr"] for e in ens[name])
    for path in ("ray", "vertex"):
        ops.DEFAULT_PATH = path
        losses, psnr = T._run_protocol(name, fused_adam=fused)
        print(f"{name} fused_adam={fused} path={path}: PSNR {psnr:.3f} dB (reference ensemble {lo:.3f} .. {hi:.3f}), final loss {losses[-1]:.5f}", flush=True)
