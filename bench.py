#!/usr/bin/env python
"""Benchmark of the intrinsic-neural-field hot path (BASELINE.json metric: rays/sec, training fwd+bwd).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B] [--mode train|infer]

A "step" is one training pass of the hot path over one batch of synthetic rays: fused gather ->
interpolate -> MLP forward, L1 loss, backward (all parameter gradients), [N>1: one NCCL allreduce of
the flat gradient buffer], torch.optim.Adam step and operand re-pack.  Workload = BASELINE.json
configs[1] (cat texture reconstruction model: k=1023 eigenfunctions, 6x128 MLP, skip 3, L1, lr 1e-4) on
a synthetic 224x224 grid-torus mesh (V=50 176, F=100 352), fp32 eigenfunction table (205 MB > L2),
`--batch` rays per GPU per step (default 131 072; the shipped batch_size 4096 is reported alongside).

Prints ONE JSON line (rank 0).  `--impl reference` times the CPU restatement of the reference path
(oracle/, kind "port": /root/reference is not present on the GPU box) on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

# per-ray algorithmic work, k=1023 H=128 L=6 (BASELINE.md section 3 / SURVEY.md section 8d)
K_EIG, HID, LAYERS, SKIP = 1023, 128, 6, 3
GRID = 224                       # grid-torus: V = 224*224 = 50 176 vertices, F = 100 352 faces
LR = 1e-4                        # intrinsic_cat.yaml:33


def flop_per_ray(train: bool) -> float:
    mac_fwd = 2 * K_EIG * HID + (LAYERS - 2) * HID * HID + HID * 3
    mac = mac_fwd + (mac_fwd + (LAYERS - 2) * HID * HID + HID * 3 if train else 0)
    return 2.0 * mac


def bytes_per_ray(train: bool) -> float:
    # 3 fp32 table rows + i64 vids + bary + rgb out (+ target when training); gather counted once per step
    return 3 * K_EIG * 4 + 24 + 12 + 12 + (12 if train else 0)


def ncu_traffic(kernel_prefix: str):
    """DRAM bytes per launch of `kernel_prefix` from the newest committed ncu summary (profiles/r*_kernels.txt), or None."""
    import glob
    import re
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_kernels.txt")), reverse=True):
        blocks = open(path).read().split("---")
        for b in blocks:
            if ("Kernel Name" in b) and (kernel_prefix in b):
                rd = re.search(r"dram__bytes_read.sum\s+([0-9.]+) \[(\w+)\]", b)
                wr = re.search(r"dram__bytes_write.sum\s+([0-9.]+) \[(\w+)\]", b)
                if rd and wr:
                    unit = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
                    return float(rd.group(1)) * unit[rd.group(2)] + float(wr.group(1)) * unit[wr.group(2)], os.path.basename(path)
    return None, None


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"], "bf16_tflops_sustained": d["bf16_tflops_sustained"],
                "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


# ------------------------------------------------------------------------------------------------
# synthetic workload
# ------------------------------------------------------------------------------------------------
def torus_faces(n: int, device) -> torch.Tensor:
    i, j = torch.meshgrid(torch.arange(n, device=device), torch.arange(n, device=device), indexing="ij")
    v00 = i * n + j
    v01 = i * n + (j + 1) % n
    v10 = ((i + 1) % n) * n + j
    v11 = ((i + 1) % n) * n + (j + 1) % n
    return torch.cat([torch.stack([v00, v10, v01], -1).reshape(-1, 3), torch.stack([v01, v10, v11], -1).reshape(-1, 3)])


def make_table(V: int, k: int, device, seed: int = 0) -> torch.Tensor:
    g = torch.Generator(device=device).manual_seed(seed)
    t = torch.randn(V, k, device=device, generator=g)
    return t / (t.max(0, keepdim=True).values - t.min(0, keepdim=True).values)   # 'standard' rescale, mesh.py:99-102


def make_rays(faces: torch.Tensor, B: int, device, seed: int, coherent: bool = False):
    """Random hit faces (worst-case locality; what a shuffled training batch looks like) or, with `coherent`, a scanline
    pattern: consecutive rays walk over neighbouring faces, ~8 rays per face (render_video.py:122-134 style frames)."""
    g = torch.Generator(device=device).manual_seed(seed)
    if coherent:
        start = int(torch.randint(0, faces.shape[0], (1,), device=device, generator=g))
        fid = (start + torch.arange(B, device=device) // 8) % faces.shape[0]
    else:
        fid = torch.randint(0, faces.shape[0], (B,), device=device, generator=g)
    vids = faces[fid].to(torch.int64).contiguous()
    u = torch.rand(B, 3, device=device, generator=g)
    bary = (u / u.sum(-1, keepdim=True)).float().contiguous()
    target = torch.rand(B, 3, device=device, generator=g)
    return vids, bary, target


class ClockSampler:
    """nvidia-smi clock / throttle-reason sampler running during the timed region."""

    def __init__(self, index: int):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the oracle's restatement of the reference path on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_reference_rate(train: bool, steps: int, warmup: int, batch: int, V: int = GRID * GRID):
    """rays/s of get_k_eigenfunc_vec_vals -> TextureField -> L1 -> backward -> Adam (trainer.py:71-84) on the CPU."""
    from oracle import field_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    dev = torch.device("cpu")
    faces = torus_faces(GRID, dev)
    table = make_table(V, K_EIG, dev)
    params = O.make_params(K_EIG, HID, LAYERS, SKIP, seed=0)
    q = {n: t.clone().requires_grad_(train) for n, t in params.items()}
    opt = torch.optim.Adam(list(q.values()), lr=LR) if train else None
    times = []
    for s in range(warmup + steps):
        vids, bary, target = make_rays(faces, batch, dev, seed=100 + s)
        t0 = time.perf_counter()
        if train:
            pred = O.field_forward(q, table, vids, bary, num_layers=LAYERS, skip=SKIP)
            loss = O.l1_loss(pred, target)
            opt.zero_grad(set_to_none=True)
            loss.backward()
            opt.step()
            loss.item()
        else:
            with torch.no_grad():
                O.field_forward(q, table, vids, bary, num_layers=LAYERS, skip=SKIP)
        dt = time.perf_counter() - t0
        if s >= warmup:
            times.append(dt)
    total = sum(times)
    return batch * len(times) / total, total / len(times) * 1e3


def torch_gpu_context_rate(train: bool, dev, batch: int = 131072, steps: int = 5):
    """Context only: the reference's PyTorch fp32 path (oracle restatement: materialised gather,
    cuBLAS SGEMMs with TF32 off, autograd, Adam) on the SAME B200, same workload.  Not the baseline, not the product."""
    from oracle import field_oracle as O
    torch.backends.cuda.matmul.allow_tf32 = False
    faces = torus_faces(GRID, dev)
    table = make_table(GRID * GRID, K_EIG, dev)
    params = {n: t.to(dev).requires_grad_(train) for n, t in O.make_params(K_EIG, HID, LAYERS, SKIP, seed=0).items()}
    opt = torch.optim.Adam(list(params.values()), lr=LR) if train else None
    vids, bary, target = make_rays(faces, batch, dev, seed=5)

    def one():
        if train:
            pred = O.field_forward(params, table, vids, bary, num_layers=LAYERS, skip=SKIP)
            loss = O.l1_loss(pred, target)
            opt.zero_grad(set_to_none=True)
            loss.backward()
            opt.step()
        else:
            with torch.no_grad():
                O.field_forward(params, table, vids, bary, num_layers=LAYERS, skip=SKIP)
    for _ in range(2):
        one()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        one()
    e1.record()
    torch.cuda.synchronize()
    return batch * steps / (e0.elapsed_time(e1) / 1e3)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    train = args.mode == "train"
    batch = 16384
    rate, ms = cpu_reference_rate(train, args.steps, args.warmup, batch)
    cores = os.cpu_count() or 1
    out = {
        "impl": "reference", "metric": "rays_per_sec_train_fwd_bwd" if train else "rays_per_sec_inference",
        "value": rate, "unit": "rays/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": workload_config(args, batch_override=batch),
        "cpu_baseline": {"value": rate, "unit": "rays/s", "cores": cores, "kind": "port",
                         "sample": f"{args.steps} steps x {batch} rays of the same workload (oracle/ restatement of the "
                                   "reference PyTorch fp32 path, torch CPU, all host threads)"},
        "e2e": {"value": rate, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out), flush=True)


def workload_config(args, batch_override=None):
    b = batch_override if batch_override is not None else args.batch
    cfg = {"workload": f"cat_texture_recon_{args.mode}_step_torus{GRID}x{GRID}_V{GRID * GRID}_k{K_EIG}_L{LAYERS}x{HID}_skip{SKIP}",
           "rays_per_gpu_per_step": b, "table": f"{args.table_dtype} (V x 1024 padded)",
           "parallelism": f"ray-sharded dp{args.gpus}",
           "timing": "inputs larger than L2 (205 MB table + per-step stash); 4 rotating ray batches of random hit faces"}
    if args.mode == "train":
        cfg.update(optimizer="Adam lr 1e-4 (fused into the reduction kernel)", loss="L1")
    return cfg


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def _say(msg):
    if os.environ.get("INF_BENCH_VERBOSE"):
        print(f"[bench r{os.environ.get('RANK', '0')}] {msg}", file=sys.stderr, flush=True)


def run_ours(args):
    import torch.distributed as dist
    import inf_b200
    from inf_b200 import _lib, trainer, parallel

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the B200 path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    train = args.mode == "train"
    B = args.batch
    V = GRID * GRID

    faces = torus_faces(GRID, dev)
    table_dense = make_table(V, K_EIG, dev)
    table = inf_b200.EigenTable.from_dense(table_dense, device=dev, dtype=args.table_dtype, path=args.path)
    del table_dense
    torch.manual_seed(0)
    model = inf_b200.make_model(dict(k=K_EIG, num_layers=LAYERS, mlp_hidden_dim=HID, skip_layer_idx=SKIP)).to(dev)
    # config.py:108 Adam(lr): the update + operand refresh run inside the gradient-reduction kernel (FusedFieldAdam)
    optim = trainer.FusedFieldAdam(model, lr=LR)
    graphed = trainer.GraphedTrainStep(model, optim)
    NB = 4
    dev_batches, host_parts = [], []
    for i in range(NB):
        vids, bary, target = make_rays(faces, B, dev, seed=1000 * rank + i)
        dev_batches.append({"eigenfunction_table": table, "vertex_idxs_of_hit_faces": vids, "barycentric_coords": bary,
                            "expected_rgbs": target})
        # on the wire the vertex ids are int32, exactly as the reference stores them on disk (mesh.py:513-516)
        host_parts.append((vids.to(torch.int32).cpu(), bary.cpu(), target.cpu()))

    def step_resident(i):
        b = dev_batches[i % NB]
        if train:
            return trainer.fused_train_step_async(model, optim, b)[0]
        with torch.no_grad():
            return model(b)

    # end-to-end arm: every step's gather inputs start in pinned HOST memory and are copied inside the timed region
    # (HostRayLoader: copy stream, one batch ahead); the step's result is read back to the host every step
    # (training: loss.item() as trainer.py:84; inference: the (B,3) RGB as renderer.py:117).
    from inf_b200.ray_dataloader import HostRayLoader
    from inf_b200.renderer import FieldRenderer
    n_e2e = min(args.steps + 6, 24)      # host batches kept in pinned memory; longer runs cycle through them
    hv = torch.cat([host_parts[i % NB][0] for i in range(n_e2e)]).pin_memory()
    hb = torch.cat([host_parts[i % NB][1] for i in range(n_e2e)]).pin_memory()
    ht = torch.cat([host_parts[i % NB][2] for i in range(n_e2e)]).pin_memory() if train else None
    if train:
        e2e_loader = HostRayLoader(table, hv, hb, ht, batch_size=B, device=dev)
    else:
        e2e_renderer = FieldRenderer(model, table, H=1, W=1, device=dev)
        e2e_out = torch.empty((n_e2e * B, 3), dtype=torch.float32).pin_memory()

    def run_e2e(steps):
        done = 0
        if train:
            while done < steps:
                for b in e2e_loader:                             # one pass over the pinned host split
                    if done >= steps:
                        break
                    graphed.step(b)                              # one graph launch; loss.item(): D2H read every step
                    done += 1
        else:
            # a stream of frames of B rays each through ONE renderer call per pass over the host split: the copies of
            # frame i+1 (in) and i-1 (out) overlap the kernels of frame i; a call returns when the last RGB is in host memory
            while done < steps:
                n = min(steps - done, n_e2e)
                e2e_renderer.predict_hits_host(hv[:n * B], hb[:n * B], chunk=B, out=e2e_out[:n * B])
                done += n

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    _say("setup done")
    lib = _lib.load()
    sampler = ClockSampler(local)     # 100 ms samples from here to the end of the end-to-end loop (warm-up + both timed regions)
    if rank == 0:
        sampler.start()
    for i in range(max(args.warmup, 3)):
        step_resident(i)
    _say("warmup done")
    lib.inf_profile_enable(1)
    ms_total = timed(step_resident, args.steps)
    import ctypes as C
    km, kw, kr, kn = C.c_float(), C.c_float(), C.c_float(), C.c_int()
    lib.inf_profile_read(C.byref(km), C.byref(kw), C.byref(kr), C.byref(kn))
    stage_arr = (C.c_float * _lib.PROFILE_STAGES)()
    lib.inf_profile_read_stages(stage_arr, C.byref(kn))
    stages = dict(zip(["incidence_lists", "vertex_project", "rays_fwd_bwd", "vertex_gather", "wgrad", "reduce_adam"],
                      [float(x) for x in stage_arr]))
    lib.inf_profile_enable(0)
    assert inf_b200.kernel_error_word(dev) == 0, "kernel watchdog fired"

    _say("resident timing done")
    if args.skip_e2e:
        sampler.stop() if rank == 0 else None
        print(json.dumps({"profiling_run": True, "ms_per_step": ms_total / args.steps, "rays_per_sec": B * args.steps / (ms_total / 1e3),
                          "step_kernels_ms": stages}), flush=True)
        if world > 1:
            dist.destroy_process_group()
        return
    run_e2e(6 if train else 2)       # per staging slot: one eager step, one graph capture, then replays
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run_e2e(args.steps)
    e1.record()
    barrier()
    t_e2e = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    ms_e2e = t_e2e.item()
    _say("e2e timing done")
    clocks = sampler.stop() if rank == 0 else None

    # rendering-style coherent rays: neighbouring pixels hit the same / adjacent faces
    coherent = None
    if (not train) and rank == 0 and world == 1:
        cb = []
        for i in range(NB):
            v, b_, _t = make_rays(faces, B, dev, seed=77 + i, coherent=True)
            cb.append({"eigenfunction_table": table, "vertex_idxs_of_hit_faces": v, "barycentric_coords": b_})
        with torch.no_grad():
            for i in range(3):
                model(cb[i % NB])
            ms_c = timed(lambda i: model(cb[i % NB]), args.steps)
        coherent = {"rays_per_sec": B * args.steps / (ms_c / 1e3), "ms_per_step": ms_c / args.steps,
                    "pattern": "scanline over neighbouring faces, 8 rays per face"}

    # shipped batch size (intrinsic_cat.yaml:32) for context, same model/table
    small = None
    if train and rank == 0 and world == 1:
        vids, bary, target = make_rays(faces, 4096, dev, seed=7)
        sb = {"eigenfunction_table": table, "vertex_idxs_of_hit_faces": vids, "barycentric_coords": bary, "expected_rgbs": target}
        for _ in range(5):
            trainer.fused_train_step_async(model, optim, sb)
        ms_small = timed(lambda i: trainer.fused_train_step_async(model, optim, sb), 50)
        for _ in range(4):
            graphed.step_async(sb)
        ms_small_g = timed(lambda i: graphed.step_async(sb), 50)
        small = {"rays_per_step": 4096, "ms_per_step": ms_small / 50, "rays_per_sec": 4096 * 50 / (ms_small / 1e3),
                 "cuda_graph_ms_per_step": ms_small_g / 50, "cuda_graph_rays_per_sec": 4096 * 50 / (ms_small_g / 1e3)}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    pk = peaks()
    rate = B * world * args.steps / (ms_total / 1e3)
    rate_e2e = B * world * args.steps / (ms_e2e / 1e3)
    main_ms = km.value if kn.value else ms_total / args.steps
    achieved = B * bytes_per_ray(train) / (main_ms / 1e3) / 1e9
    h2d = B * (12 + 12 + (12 if train else 0))
    # DRAM bytes of one launch of the dominant kernel from the committed ncu capture of this workload (same B, same table)
    traffic, traffic_src = (ncu_traffic("field_main_kernel<%d, float>" % (1 if train else 0))
                            if (args.table_dtype == "f32" and B == 131072 and GRID == 224) else (None, None))
    out = {
        "metric": "rays_per_sec_train_fwd_bwd" if train else "rays_per_sec_inference",
        "value": rate, "unit": "rays/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f16 operands / f32 accumulate", "data": "synthetic", "config": workload_config(args),
        "clocks": clocks,
        "e2e": {"value": rate_e2e, "unit": "rays/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4 if train else B * 12,
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": (3 if train else 1) * args.steps,
        "roofline": {"bound": "hbm", "kernel": "field_main_kernel<train>" if train else "field_main_kernel<infer>",
                     "achieved": achieved, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": achieved / pk["hbm_gbs"],
                     "traffic": traffic, "traffic_source": traffic_src, "peak_source": pk["source"], "kernel_ms": main_ms,
                     # rows shared by rays are served from L2 (the 205 MB table is partly resident), so the algorithmic
                     # rate can exceed the DRAM rate -- and the copy peak; the DRAM side of the same launch:
                     "dram_gbs": (traffic / (main_ms / 1e3) / 1e9) if traffic else None,
                     "algorithmic_bytes_per_ray": bytes_per_ray(train),
                     "step_kernels_ms": stages,
                     "tensor_frac_of_sustained_bf16": rate / world * flop_per_ray(train) / (pk["bf16_tflops_sustained"] * 1e12)},
    }
    if small is not None:
        out["config_batch_4096"] = small
    if coherent is not None:
        out["coherent_rays"] = coherent
    if not args.no_cpu_baseline and world == 1:
        cr, cms = cpu_reference_rate(train, steps=6, warmup=1, batch=16384)
        out["cpu_baseline"] = {"value": cr, "unit": "rays/s", "cores": os.cpu_count() or 1, "kind": "port",
                               "sample": "6 steps x 16384 rays of the same workload (oracle/ restatement, torch CPU, all threads)"}
        try:
            out["context_reference_torch_fp32_on_this_gpu"] = {
                "value": torch_gpu_context_rate(train, dev), "unit": "rays/s",
                "note": "oracle restatement of the reference PyTorch path on the same B200 (TF32 off); context, not the baseline"}
        except Exception as exc:      # context only: never fail the bench line
            out["context_reference_torch_fp32_on_this_gpu"] = {"value": None, "note": f"not measured: {type(exc).__name__}"}
    print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="train", choices=["train", "infer"])
    ap.add_argument("--batch", type=int, default=131072, help="rays per GPU per step")
    ap.add_argument("--table-dtype", default="f32", choices=["f32", "f16"])
    ap.add_argument("--path", default="auto", choices=["auto", "ray", "vertex"],
                    help="evaluation of the two K = k contractions: per ray, per vertex, or the library's rule (InfDesc.path)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--skip-e2e", action="store_true", help="profiling runs: resident steps only (the JSON line is not a bench result)")
    ap.add_argument("--grid", type=int, default=224, help="torus grid side: V = grid^2 vertices (default: config 2)")
    args = ap.parse_args()
    global GRID
    GRID = args.grid
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
