"""Measurement helper: host -> host inference (FieldRenderer.predict_hits_host) of a stream of rays at several chunk sizes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import time, torch
import inf_b200
from inf_b200.renderer import FieldRenderer
import bench
dev = torch.device("cuda")
faces = bench.torus_faces(224, dev)
table = inf_b200.EigenTable.from_dense(bench.make_table(224 * 224, 1023, dev), device=dev)
torch.manual_seed(0)
m = inf_b200.make_model(bench.MODEL_CFG).to(dev).eval()
N = 1 << 22                                              # ~ two 1080p frames of hits
v, b, _ = bench.make_rays(faces, N, dev, seed=3)
hv, hb = v.to(torch.int32).cpu().pin_memory(), b.cpu().pin_memory()
out = torch.empty((N, 3), dtype=torch.float32).pin_memory()
r = FieldRenderer(m, table, H=1, W=1, device=dev)
for chunk in (1 << 17, 1 << 18, 1 << 19, 1 << 20):
    for _ in range(3):
        r.predict_hits_host(hv, hb, chunk=chunk, out=out)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    reps = 5
    for _ in range(reps):
        r.predict_hits_host(hv, hb, chunk=chunk, out=out)
    dt = (time.perf_counter() - t0) / reps
    print(f"chunk {chunk:8d}: {dt * 1e3:7.3f} ms per {N} rays = {N / dt / 1e6:8.1f} M rays/s host to host "
          f"({N * 36 / dt / 1e9:.1f} GB/s up+down)")
