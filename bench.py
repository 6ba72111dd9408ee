#!/usr/bin/env python
"""Benchmark of the intrinsic-neural-field hot path (BASELINE.json metric: rays/sec, training fwd+bwd and inference).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B] [--mode train|infer]

A "step" is one pass of the hot path over one batch of synthetic rays.  Training (`--mode train`, the default): gather ->
interpolate -> MLP forward, L1 loss, backward (all parameter gradients), [N>1: one allreduce of the flat gradient buffer],
Adam step and operand refresh (config.py:108, trainer.py:71-84).  Inference (`--mode infer`): gather -> interpolate -> MLP
-> RGB (renderer.py:113-118).  Workload = BASELINE.json configs[1] (cat texture reconstruction model: k=1023
eigenfunctions, 6x128 MLP, skip 3, L1, lr 1e-4) on a synthetic 224x224 grid-torus mesh (V=50 176, F=100 352), fp32
eigenfunction table (205 MB > L2), `--batch` rays per GPU per step (default 131 072), random hit faces.

Prints ONE JSON line (rank 0).  The default line carries, besides the headline arm: the other mode (`inference` /
`training`), the shipped batch size 4096, config 4 (400k-row table, the gather-bandwidth-bound case), the first-layer
tensor-pipe fraction, the CPU baseline and the reference's PyTorch path on the same GPU (context).
`--impl reference` times the reference's own CPU implementation (oracle/_ref staged from /root/reference by
oracle/make_ref.py; the oracle port if that is absent) on the host cores, at the same rays per step.
"""
import argparse
import ctypes as C
import glob
import json
import os
import re
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
K_PAD = 1024
GRID = 224                       # grid-torus: V = 224*224 = 50 176 vertices, F = 100 352 faces
LR = 1e-4                        # intrinsic_cat.yaml:33
MODEL_CFG = dict(k=K_EIG, num_layers=LAYERS, mlp_hidden_dim=HID, skip_layer_idx=SKIP)
STAGES = ["incidence_lists", "vertex_project", "rays_fwd_bwd", "vertex_gather", "wgrad", "reduce_adam"]


def flop_per_ray(train: bool) -> float:
    mac_fwd = 2 * K_EIG * HID + (LAYERS - 2) * HID * HID + HID * 3
    mac = mac_fwd + (mac_fwd + (LAYERS - 2) * HID * HID + HID * 3 if train else 0)
    return 2.0 * mac


def bytes_per_ray(train: bool) -> float:
    # 3 fp32 table rows + i64 vids + bary + rgb out (+ target when training); gather counted once per step
    return 3 * K_EIG * 4 + 24 + 12 + 12 + (12 if train else 0)


def ncu_kernels():
    """{kernel name: (dram bytes per launch, duration us, tensor pipe %)} from the newest committed ncu summary."""
    unit = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_kernels.txt")), reverse=True):
        out = {}
        for b in open(path).read().split("---"):
            name = re.search(r"Kernel Name\s+(.+?) \[", b)
            rd = re.search(r"dram__bytes_read.sum\s+([0-9.]+) \[(\w+)\]", b)
            wr = re.search(r"dram__bytes_write.sum\s+([0-9.]+) \[(\w+)\]", b)
            tp = re.search(r"sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active\s+([0-9.]+)", b)
            if name and rd and wr:
                out.setdefault(name.group(1).strip(), (float(rd.group(1)) * unit[rd.group(2)] + float(wr.group(1)) * unit[wr.group(2)],
                                                       float(tp.group(1)) if tp else None))
        if out:
            return out, os.path.basename(path)
    return {}, None


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
# reference arm / CPU baseline: the reference's own CPU implementation on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_reference_rate(train: bool, steps: int, warmup: int, batch: int, V: int = None):
    """rays/s of get_k_eigenfunc_vec_vals -> TextureField -> L1 -> backward -> Adam (trainer.py:71-84) on the CPU, with the
    unmodified reference modules staged under oracle/_ref ("reference"); the oracle's restatement ("port") if absent."""
    from oracle import ref_path
    torch.set_num_threads(os.cpu_count() or 1)
    dev = torch.device("cpu")
    V = GRID * GRID if V is None else V
    faces = torus_faces(GRID, dev)
    table = make_table(V, K_EIG, dev)
    kind = "reference" if ref_path.available() else "port"
    if kind == "reference":
        torch.manual_seed(0)
        ref = ref_path.ReferenceStep(dict(MODEL_CFG, feature_strategy="efuncs"), table, LR, train=train)
        step = ref.step
    else:
        from oracle import field_oracle as O
        q = {n: t.clone().requires_grad_(train) for n, t in O.make_params(K_EIG, HID, LAYERS, SKIP, seed=0).items()}
        opt = torch.optim.Adam(list(q.values()), lr=LR) if train else None

        def step(vids, bary, target=None):
            if not train:
                with torch.no_grad():
                    return O.field_forward(q, table, vids, bary, num_layers=LAYERS, skip=SKIP)
            pred = O.field_forward(q, table, vids, bary, num_layers=LAYERS, skip=SKIP)
            loss = O.l1_loss(pred, target)
            opt.zero_grad(set_to_none=True)
            loss.backward()
            opt.step()
            return loss.item(), pred
    times = []
    for s in range(warmup + steps):
        vids, bary, target = make_rays(faces, batch, dev, seed=100 + s)
        t0 = time.perf_counter()
        step(vids, bary, target if train else None)
        dt = time.perf_counter() - t0
        if s >= warmup:
            times.append(dt)
    total = sum(times)
    return batch * len(times) / total, total / len(times) * 1e3, kind


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
    batch = args.batch                      # the same rays per step as the GPU arm
    rate, ms, kind = cpu_reference_rate(train, args.steps, args.warmup, batch)
    cores = os.cpu_count() or 1
    out = {
        "impl": "reference", "metric": "rays_per_sec_train_fwd_bwd" if train else "rays_per_sec_inference",
        "value": rate, "unit": "rays/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": workload_config(args),
        "cpu_baseline": {"value": rate, "unit": "rays/s", "cores": cores, "kind": kind,
                         "sample": f"{args.steps} steps x {batch} rays of the same workload: "
                                   + ("the reference's own get_k_eigenfunc_vec_vals + make_model + F.l1_loss + Adam (oracle/_ref), "
                                      if kind == "reference" else "oracle/ restatement of the reference PyTorch fp32 path, ")
                                   + "torch CPU, all host threads"},
        "e2e": {"value": rate, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out), flush=True)


def workload_config(args, batch_override=None):
    b = batch_override if batch_override is not None else args.batch
    cfg = {"workload": f"cat_texture_recon_{args.mode}_step_torus{GRID}x{GRID}_V{GRID * GRID}_k{K_EIG}_L{LAYERS}x{HID}_skip{SKIP}",
           "rays_per_gpu_per_step": b, "table": f"{args.table_dtype} (V x 1024 padded)",
           "parallelism": f"ray-sharded dp{args.gpus}",
           "timing": "inputs larger than L2 (205 MB table + its 103 MB operand image + per-step activation stash); "
                     "4 rotating ray batches of random hit faces"}
    if args.mode == "train":
        cfg.update(optimizer="Adam lr 1e-4 (fused into the reduction kernel)", loss="L1")
    return cfg


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def _say(msg):
    if os.environ.get("INF_BENCH_VERBOSE"):
        print(f"[bench r{os.environ.get('RANK', '0')}] {msg}", file=sys.stderr, flush=True)


class Timer:
    """CUDA-event timing of K calls bracketed by barrier + synchronize on both sides, max over ranks."""

    def __init__(self, dev, world):
        self.dev, self.world = dev, world

    def barrier(self):
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def __call__(self, fn, steps):
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        self.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=self.dev)
        if self.world > 1:
            import torch.distributed as dist
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()


def read_stages(lib):
    from inf_b200 import _lib
    arr = (C.c_float * _lib.PROFILE_STAGES)()
    n = C.c_int()
    lib.inf_profile_read_stages(arr, C.byref(n))
    return dict(zip(STAGES, [float(x) for x in arr])), n.value


def kernel_rooflines(stages, train, vertex, B, V, pk, ncu):
    """Per-kernel rooflines of one step: what each kernel executes against the roof that binds it."""
    tf = pk["bf16_tflops_sustained"] * 1e12
    ntv = (V + 127) // 128
    ker = []

    def add(name, ms, flops=None, bytes_=None, bound=None, ncu_name=None, note=None):
        if ms <= 0:
            return
        e = {"kernel": name, "ms": ms}
        if flops is not None:
            e.update(tflops=flops / (ms / 1e3) / 1e12, tensor_frac_of_sustained_bf16=flops / (ms / 1e3) / tf)
        if bytes_ is not None:
            e.update(gbs=bytes_ / (ms / 1e3) / 1e9, hbm_frac=bytes_ / (ms / 1e3) / 1e9 / pk["hbm_gbs"], bytes=bytes_)
        t = ncu.get(f"{ncu_name or name}<{1 if train else 0}") or ncu.get(ncu_name or name)
        if t is not None:
            e.update(ncu_dram_bytes=t[0], ncu_tensor_pipe_pct=t[1])
        e["bound"] = bound
        if note:
            e["note"] = note
        ker.append(e)
    hid = (LAYERS - 2) * HID * HID * 2.0
    if vertex:
        if train or stages["vertex_project"] > 0.01:   # inference of a frozen model: P is projected once, outside the step
            add("vertex_project_kernel", stages["vertex_project"], flops=2.0 * ntv * 128 * K_PAD * 256,
                bytes_=ntv * 128 * (K_PAD * 2 + 512), bound="tensor",
                note="P = E [W0;Wy]^T once per vertex: the first layer + skip branch of every ray; bytes = table image in, P out")
        add("field_rays_kernel", stages["rays_fwd_bwd"], flops=B * hid * (2 if train else 1),
            bytes_=B * (3 * 512 + 36 + 12 + (12 + 10 * 256 + 32 if train else 0)), bound="latency",
            note="per-ray chain; bytes = 3 P rows + inputs + RGB (+ activation stash out); neither roof binds, see DESIGN.md")
        if train:
            add("vertex_gather_kernel", stages["vertex_gather"], bytes_=B * 3 * 512 + ntv * 65536, bound="l2",
                note="G[v] = sum bary * dz: 3 stashed 512 B rows per ray, G image out")
            add("field_wgrad_kernel", stages["wgrad"], flops=B * hid + 2.0 * ntv * 128 * K_PAD * 256,
                bytes_=B * (10 * 256 + 32) + ntv * 128 * (K_PAD * 2 + 512), bound="hbm",
                note="hidden dW from the stash (K = rays) + [dW0;dWy] = G^T E (K = vertices); bytes = stash + table image + G")
    else:
        fwd = 2.0 * (2 * K_EIG * HID + (LAYERS - 2) * HID * HID + 3 * HID)
        add("field_main_kernel", stages["rays_fwd_bwd"], flops=B * (fwd + (hid + 6.0 * HID if train else 0)),
            bytes_=B * bytes_per_ray(train), bound="hbm", note="per-ray gather of 3 table rows fused into GEMM1 (+ dgrad chain)")
        if train:
            add("field_wgrad_kernel", stages["wgrad"], flops=B * fwd, bytes_=B * 4864, bound="hbm",
                note="all weight gradients from the 4.75 KB/ray stash")
    if train:
        add("field_reduce_lean_kernel", stages["reduce_adam"], bound="latency",
            note="partial sums + Adam + operand refresh (+ the N-GPU gradient exchange); the stage time includes the ~5 us launch gap")
    return ker


def run_ours(args):
    import torch.distributed as dist
    import inf_b200
    from inf_b200 import _lib, trainer
    from inf_b200.ray_dataloader import HostRayLoader
    from inf_b200.renderer import FieldRenderer

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the B200 path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    lib = _lib.load()
    debug = lib.inf_debug_flags()
    if debug & (7 | 0x300):
        raise RuntimeError(f"INF_DEBUG / INF_REDUCE_DBG flags {debug:#x}: measurement switches that skip work and produce wrong "
                           "gradients are set -- refusing to benchmark")
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    headline_train = args.mode == "train"
    B = args.batch
    V = GRID * GRID
    timed = Timer(dev, world)

    faces = torus_faces(GRID, dev)
    table_dense = make_table(V, K_EIG, dev)
    table = inf_b200.EigenTable.from_dense(table_dense, device=dev, dtype=args.table_dtype, path=args.path)
    del table_dense
    torch.manual_seed(0)
    model = inf_b200.make_model(MODEL_CFG).to(dev)
    # config.py:108 Adam(lr): the update + operand refresh run inside the gradient-reduction kernel (FusedFieldAdam)
    optim = trainer.FusedFieldAdam(model, lr=LR)
    # N > 1: the ranks' gradients are summed inside the reduction kernel over NVLink peer memory (INF_PEER_EXCHANGE=0 or
    # --nccl: one NCCL allreduce of the flat gradient + a separate Adam launch instead)
    exchange = optim.enable_peer_exchange() if (world > 1 and not args.nccl) else None
    graphed = trainer.GraphedTrainStep(model, optim)
    NB = 4
    dev_batches, host_parts = [], []
    for i in range(NB):
        vids, bary, target = make_rays(faces, B, dev, seed=1000 * rank + i)
        dev_batches.append({"eigenfunction_table": table, "vertex_idxs_of_hit_faces": vids, "barycentric_coords": bary,
                            "expected_rgbs": target})
        # on the wire the vertex ids are int32, exactly as the reference stores them on disk (mesh.py:513-516)
        host_parts.append((vids.to(torch.int32).cpu(), bary.cpu(), target.cpu()))
    d0 = inf_b200.ops._desc(model.cfg, B, V, path=inf_b200.ops.PATHS[args.path])
    t0 = _lib.InfTensors(table_image=table.image.data_ptr() if table.image is not None else None)
    vertex = bool(lib.inf_uses_vertex_path(C.byref(d0), C.byref(t0)))

    from inf_b200.renderer import GraphedForward
    gforward = GraphedForward(model)

    # resident inputs: one CUDA-graph launch per step (GraphedTrainStep / GraphedForward: the kernels of a step take
    # 0.1-0.4 ms, the eager Python path into them ~0.2 ms); the eager calls are used for the per-stage device times
    def train_step(i):
        return graphed.step_async(dev_batches[i % NB])[0]

    def train_step_eager(i):
        return trainer.fused_train_step_async(model, optim, dev_batches[i % NB])[0]

    def infer_step(i):
        return gforward(dev_batches[i % NB])

    def infer_step_eager(i):
        with torch.no_grad():
            return model(dev_batches[i % NB])

    # end-to-end arms: every step's gather inputs start in pinned HOST memory and are copied inside the timed region
    # (HostRayLoader: copy stream, one batch ahead); the step's result is read back to the host every step
    # (training: loss.item() as trainer.py:84; inference: the (B,3) RGB as renderer.py:117).
    n_e2e = min(args.steps + 6, 24)      # host batches kept in pinned memory; longer runs cycle through them
    hv = torch.cat([host_parts[i % NB][0] for i in range(n_e2e)]).pin_memory()
    hb = torch.cat([host_parts[i % NB][1] for i in range(n_e2e)]).pin_memory()
    ht = torch.cat([host_parts[i % NB][2] for i in range(n_e2e)]).pin_memory()
    e2e_loader = HostRayLoader(table, hv, hb, ht, batch_size=B, device=dev)
    e2e_renderer = FieldRenderer(model, table, H=1, W=1, device=dev)
    e2e_out = torch.empty((n_e2e * B, 3), dtype=torch.float32).pin_memory()

    def train_e2e(steps):
        done = 0
        while done < steps:
            for b in e2e_loader:                             # one pass over the pinned host split
                if done >= steps:
                    break
                graphed.step_logged(b)                       # one graph launch; the loss is read back (D2H) every step,
                done += 1                                    # one step behind the launch (no device drain per step)
        graphed.flush_logged()                               # the last step's loss: every step's result has reached the host

    def infer_e2e(steps):
        # a stream of frames of B rays each through ONE renderer call per pass over the host split: the copies of
        # frame i+1 (in) and i-1 (out) overlap the kernels of frame i; a call returns when the last RGB is in host memory
        done = 0
        while done < steps:
            n = min(steps - done, n_e2e)
            e2e_renderer.predict_hits_host(hv[:n * B], hb[:n * B], out=e2e_out[:n * B])   # the renderer's own chunking (2^18 rays)
            done += n

    def measure(train):
        """(resident ms/step, stage ms, launches/step, e2e ms/step) of one mode."""
        step, eager, e2e = (train_step, train_step_eager, train_e2e) if train else (infer_step, infer_step_eager, infer_e2e)
        model.train(train)
        for i in range(max(args.warmup, 3, 2 * NB)):     # every resident batch: one eager call, one capture
            step(i)
        ms = timed(step, args.steps)
        # per-stage device times and the launch count from eager calls (a graph replay does not pass through the library)
        for i in range(3):
            eager(i)
        lib.inf_profile_enable(1)
        l0 = lib.inf_launch_count()
        timed(eager, min(args.steps, 20))
        launches = (lib.inf_launch_count() - l0) / min(args.steps, 20)
        stages, _ = read_stages(lib)
        lib.inf_profile_enable(0)
        assert inf_b200.kernel_error_word(dev) == 0, "kernel watchdog fired"
        if args.skip_e2e:
            return ms / args.steps, stages, launches, None
        # per staging slot (2 training, 3 inference): one eager step, one graph capture, then replays; an inference call cuts the
        # host stream into the renderer's 2^18-ray chunks, so three passes over the pinned split warm every slot's graph
        if train:
            e2e(6)
        else:
            for _ in range(3):
                e2e(n_e2e)
        ms_e2e = timed(lambda i: e2e(args.steps), 1)
        return ms / args.steps, stages, launches, ms_e2e / args.steps

    _say("setup done")
    sampler = ClockSampler(local)     # 100 ms samples over warm-up + every timed region
    if rank == 0:
        sampler.start()
    res = {}
    # both halves of BASELINE.json's metric at every N (inference shards chunks / frames with no collective)
    modes = [headline_train] if (args.skip_e2e or args.headline_only) else [headline_train, not headline_train]
    for tr in modes:
        res[tr] = measure(tr)
        _say(f"mode train={tr} done")
    clocks = sampler.stop() if rank == 0 else None
    ranks_identical = None
    if world > 1:
        # identical Adam steps on every rank: the weights must still agree bit for bit after all the steps above
        mine = optim.flat.detach().clone()
        allw = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allw, mine)
        ranks_identical = all(torch.equal(allw[0], w) for w in allw)
        assert ranks_identical, "rank weights diverged"

    if args.skip_e2e:
        ms, stages, launches, _ = res[headline_train]
        if rank == 0:
            print(json.dumps({"profiling_run": True, "ms_per_step": ms, "rays_per_sec": B * world / (ms / 1e3), "vertex_path": vertex,
                              "launches_per_step": launches, "step_kernels_ms": stages}), flush=True)
        if world > 1:
            dist.destroy_process_group()
        return

    proj_ms = None
    if vertex and rank == 0 and world == 1:
        from inf_b200 import ops as _ops
        model.eval()
        with torch.no_grad():
            packed_now, pbuf = model.prepare(dev_batches[0])
            if pbuf is not None:
                proj_ms = timed(lambda i: _ops.project_vertices_raw(model.cfg, table, packed_now, out=pbuf), 10) / 10
    extras = {}
    if rank == 0 and world == 1 and not args.headline_only:
        model.eval()
        # rendering-style coherent rays: neighbouring pixels hit the same / adjacent faces
        cb = []
        for i in range(NB):
            v, b_, _t = make_rays(faces, B, dev, seed=77 + i, coherent=True)
            cb.append({"eigenfunction_table": table, "vertex_idxs_of_hit_faces": v, "barycentric_coords": b_})
        with torch.no_grad():
            for i in range(3):
                model(cb[i % NB])
            ms_c = timed(lambda i: model(cb[i % NB]), args.steps)
        extras["coherent_rays_inference"] = {"rays_per_sec": B * args.steps / (ms_c / 1e3), "ms_per_step": ms_c / args.steps,
                                             "pattern": "scanline over neighbouring faces, 8 rays per face"}
        del cb
        # shipped batch size (intrinsic_cat.yaml:32), same model/table: eager launches and one CUDA-graph launch per step
        model.train(True)
        vids, bary, target = make_rays(faces, 4096, dev, seed=7)
        sb = {"eigenfunction_table": table, "vertex_idxs_of_hit_faces": vids, "barycentric_coords": bary, "expected_rgbs": target}
        for _ in range(5):
            trainer.fused_train_step_async(model, optim, sb)
        ms_small = timed(lambda i: trainer.fused_train_step_async(model, optim, sb), 50)
        for _ in range(4):
            graphed.step_async(sb)
        ms_small_g = timed(lambda i: graphed.step_async(sb), 50)
        extras["config_batch_4096"] = {"rays_per_step": 4096, "ms_per_step": ms_small / 50, "rays_per_sec": 4096 * 50 / (ms_small / 1e3),
                                       "cuda_graph_ms_per_step": ms_small_g / 50,
                                       "cuda_graph_rays_per_sec": 4096 * 50 / (ms_small_g / 1e3)}
        # config 3 (human, view-dependence branch: intrinsic angle -> Fourier features -> 2-layer directional MLP), same table
        try:
            import types
            normals = torch.nn.functional.normalize(torch.randn(faces.shape[0], 3, device=dev), dim=-1)
            vcfg = dict(MODEL_CFG, view_dependence=dict(strategy="intrinsic", in_dim_view_dir=1, include_view_dir=True, embed_size=4,
                                                        bottleneck_vec_dim=128, directional_hidden_dim=64))
            mv = inf_b200.make_model(vcfg, mesh=types.SimpleNamespace(face_normals=normals.cpu().numpy())).to(dev)
            ov = trainer.FusedFieldAdam(mv, lr=LR)
            bv = []
            for i in range(NB):
                gvi = torch.Generator(device=dev).manual_seed(300 + i)
                fid = torch.randint(0, faces.shape[0], (B,), device=dev, generator=gvi)
                u = torch.rand(B, 3, device=dev, generator=gvi)
                dd = torch.nn.functional.normalize(torch.randn(B, 3, device=dev, generator=gvi), dim=-1)
                dd = torch.where(((-dd * normals[fid]).sum(-1) < 0)[:, None], -dd, dd)
                bv.append({"eigenfunction_table": table, "vertex_idxs_of_hit_faces": faces[fid].to(torch.int64).contiguous(),
                           "barycentric_coords": (u / u.sum(-1, keepdim=True)).contiguous(),
                           "expected_rgbs": torch.rand(B, 3, device=dev, generator=gvi), "unit_ray_dirs": dd.contiguous(),
                           "hit_face_idxs": fid})
            gv = trainer.GraphedTrainStep(mv, ov)
            mv.train(True)
            for i in range(3 * NB):
                gv.step_async(bv[i % NB])
            ms_vt = timed(lambda i: gv.step_async(bv[i % NB])[0], args.steps)
            mv.eval()
            gfv = GraphedForward(mv)
            with torch.no_grad():
                for i in range(3 * NB):
                    gfv(bv[i % NB])
                ms_vi = timed(lambda i: gfv(bv[i % NB]), args.steps)
            extras["config3_view_dependent"] = {
                "rays_per_step": B, "view_dependence": vcfg["view_dependence"],
                "training": {"rays_per_sec": B * args.steps / (ms_vt / 1e3), "ms_per_step": ms_vt / args.steps},
                "inference_frozen_weights": {"rays_per_sec": B * args.steps / (ms_vi / 1e3), "ms_per_step": ms_vi / args.steps}}
            del mv, ov, bv, gv, gfv
        except Exception as exc:       # an extra: never fail the bench line
            extras["config3_view_dependent"] = {"error": f"{type(exc).__name__}: {exc}"}
        # config 4 (human_cloud400k): 400k-row table, 1.64 GB, every gathered row comes from DRAM -- the gather-bandwidth case
        try:
            g4 = 632
            V4 = g4 * g4
            t4 = inf_b200.EigenTable.from_dense(make_table(V4, K_EIG, dev, seed=4), device=dev, image=False)
            f4 = torus_faces(g4, dev)
            m4 = inf_b200.make_model(MODEL_CFG).to(dev)
            o4 = trainer.FusedFieldAdam(m4, lr=LR)
            b4 = []
            for i in range(NB):
                v, b_, t_ = make_rays(f4, B, dev, seed=40 + i)
                b4.append({"eigenfunction_table": t4, "vertex_idxs_of_hit_faces": v, "barycentric_coords": b_, "expected_rgbs": t_})
            out4 = {"V": V4, "table_bytes": V4 * K_PAD * 4, "rays_per_step": B}
            for tr in (False, True):
                m4.train(tr)

                def s4(i, tr=tr):
                    if tr:
                        return trainer.fused_train_step_async(m4, o4, b4[i % NB])[0]
                    with torch.no_grad():
                        return m4(b4[i % NB])
                for i in range(3):
                    s4(i)
                lib.inf_profile_enable(1)
                ms4 = timed(s4, args.steps)
                st4, _ = read_stages(lib)
                lib.inf_profile_enable(0)
                key = "training" if tr else "inference"
                out4[key] = {"rays_per_sec": B * args.steps / (ms4 / 1e3), "ms_per_step": ms4 / args.steps,
                             "gather_kernel_ms": st4["rays_fwd_bwd"],
                             "achieved_hbm_gbs": B * bytes_per_ray(tr) / (st4["rays_fwd_bwd"] / 1e3) / 1e9,
                             "frac_of_hbm_peak": B * bytes_per_ray(tr) / (st4["rays_fwd_bwd"] / 1e3) / 1e9 / peaks()["hbm_gbs"]}
            # the same 400k-row table rendered by a FROZEN model: the vertices are projected once (204 MB of P, streamed
            # from DRAM: it does not fit L2) and every step gathers 1.5 KB of P per ray instead of 12 KB of table rows
            t4.build_image()
            m4.eval()
            with torch.no_grad():
                for i in range(3):
                    m4(b4[i % NB])
                g4f = GraphedForward(m4)
                for i in range(3 * NB):
                    g4f(b4[i % NB])
                ms4f = timed(lambda i: g4f(b4[i % NB]), args.steps)
            out4["inference_frozen_weights"] = {"rays_per_sec": B * args.steps / (ms4f / 1e3), "ms_per_step": ms4f / args.steps,
                                                "note": "vertex projection computed once per (weights, table), per-ray kernel only per step"}
            extras["config4_gather_bound"] = out4
            del t4, m4, o4, b4, g4f
            torch.cuda.empty_cache()
        except Exception as exc:       # an extra: never fail the bench line
            extras["config4_gather_bound"] = {"error": f"{type(exc).__name__}: {exc}"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    pk = peaks()
    ncu, ncu_src = ncu_kernels()
    # "void inf::field_rays_kernel<1, 1>(inf::FieldParams)" -> "field_rays_kernel<1" (the first template argument is the
    # train / inference flavour of the templated kernels) and the bare name, first capture wins
    ncu_named = {}
    for k, v in ncu.items():
        base = k.replace("void ", "").replace("inf::", "").split("(")[0]
        flav = base.split(",")[0].rstrip(">")
        ncu_named.setdefault(flav, v)
        ncu_named.setdefault(base.split("<")[0], v)
    ncu = ncu_named

    def block(train):
        ms, stages, launches, ms_e2e = res[train]
        kernels_ms = sum(stages.values())
        rate = B * world / (ms / 1e3)
        h2d = B * (12 + 12 + (12 if train else 0))
        ker = kernel_rooflines(stages, train, vertex, B, V, pk, ncu)
        traffic = sum(k["ncu_dram_bytes"] for k in ker if "ncu_dram_bytes" in k) or None
        achieved = B * bytes_per_ray(train) / (kernels_ms / 1e3) / 1e9
        first = next((k for k in ker if k["kernel"] in ("vertex_project_kernel", "field_main_kernel")), None)
        out = {
            "value": rate, "unit": "rays/s", "ms_per_step": ms,
            "e2e": {"value": B * world / (ms_e2e / 1e3), "unit": "rays/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": 4 if train else B * 12, "ms_per_step": ms_e2e},
            "gpu_launches": int(round(launches * args.steps)), "launches_per_step": launches,
            "roofline": {
                "bound": "hbm",
                "kernel": ("whole step: the path's work is split over %d kernels (per vertex + per ray), timed with CUDA events "
                           "on the launching stream" % round(launches)) if vertex else "field_main_kernel + field_wgrad_kernel + field_reduce_kernel",
                "achieved": achieved, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": achieved / pk["hbm_gbs"],
                "traffic": traffic, "traffic_source": ncu_src, "peak_source": pk["source"], "kernel_ms": kernels_ms,
                "algorithmic_bytes_per_ray": bytes_per_ray(train),
                "note": ("achieved = SURVEY 8(d) algorithmic bytes (3 fp32 table rows per ray, counted once) / device time of the step's "
                         "kernels.  The vertex path does not move those bytes: it reads the table image once per contraction "
                         "(2 x 103 MB per training step) and gathers 1.5 KB of P per ray from L2, so frac compares the path's "
                         "throughput with what a perfect per-ray gather could do, not this implementation's DRAM traffic "
                         "(traffic: ncu).") if vertex else "rows shared by rays are served from L2, so frac can exceed the DRAM rate",
                "step_kernels_ms": stages, "kernels": ker,
                "tensor_frac_of_sustained_bf16": rate / world * flop_per_ray(train) / (pk["bf16_tflops_sustained"] * 1e12)},
            "first_layer_tensor_frac": None if first is None else {
                "kernel": first["kernel"], "ms": first["ms"],
                "executed_tflops": first.get("tflops"), "frac_of_sustained_bf16": first.get("tensor_frac_of_sustained_bf16"),
                "per_ray_equivalent_tflops": B * 4.0 * K_EIG * HID / (first["ms"] / 1e3) / 1e12,
                "note": "executed = FLOPs the kernel really performs (vertex path: 2 V k 256); per_ray_equivalent = the first-layer "
                        "FLOPs of the reference formulation (2 B k 256) / the same time"},
        }
        return out

    head = block(headline_train)
    out = {
        "metric": "rays_per_sec_train_fwd_bwd" if headline_train else "rays_per_sec_inference",
        "value": head["value"], "unit": "rays/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f16 operands / f32 accumulate", "data": "synthetic", "config": workload_config(args),
        "clocks": clocks, "e2e": head["e2e"], "gpu_launches": head["gpu_launches"], "roofline": head["roofline"],
        "first_layer_tensor_frac": head["first_layer_tensor_frac"],
        "evaluation_path": "vertex (P = E [W0;Wy]^T per vertex, rays gather P)" if vertex else "ray (3 table rows gathered per ray)",
        "gradient_exchange": None if world == 1 else ("peer memory over NVLink inside field_reduce_kernel (rank-ordered sum, Adam fused)"
                                                      if exchange is not None else "NCCL allreduce of the flat gradient + Adam launch"),
        "rank_weights_bit_identical": ranks_identical,
        "inf_debug": debug,
    }
    if (not headline_train) in res:
        other = block(not headline_train)
        out["inference" if headline_train else "training"] = other
    inf_block = out.get("inference") if headline_train else out
    if vertex and inf_block is not None and proj_ms is not None:
        # a renderer evaluates a FROZEN model (model.eval(), renderer.py:36-146): the vertex projection is computed once per
        # (weights, table) and every chunk / frame launches the per-ray kernel only
        inf_block["vertex_projection"] = {"cached": True, "once_ms": proj_ms, "steps_to_amortise_to_1pct": int(100 * proj_ms / max(inf_block["ms_per_step"], 1e-9)) + 1,
                                          "ms_per_step_if_projected_every_step": inf_block["ms_per_step"] + proj_ms,
                                          "note": "eval mode: P = E [W0;Wy]^T is projected once per (weights, table) before the timed "
                                                  "region, as for every frame of render_video.py:117-134; training projects every step"}
    out.update(extras)
    if not args.no_cpu_baseline and world == 1:
        cr, cms, kind = cpu_reference_rate(headline_train, steps=3, warmup=1, batch=B)
        out["cpu_baseline"] = {"value": cr, "unit": "rays/s", "cores": os.cpu_count() or 1, "kind": kind,
                               "sample": f"3 steps x {B} rays of the same workload ("
                                         + ("the reference's own modules from oracle/_ref" if kind == "reference" else "oracle/ restatement")
                                         + ", torch CPU, all host threads)"}
        try:
            out["context_reference_torch_fp32_on_this_gpu"] = {
                "value": torch_gpu_context_rate(headline_train, dev), "unit": "rays/s",
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
    ap.add_argument("--nccl", action="store_true", help="N > 1: NCCL allreduce + separate Adam launch instead of the peer-memory exchange")
    ap.add_argument("--headline-only", action="store_true", help="only the headline mode (no other mode, no extras)")
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
