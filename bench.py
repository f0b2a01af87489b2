#!/usr/bin/env python
"""bench.py -- ConstraintModule forward+backward samples/sec (BASELINE.json metric) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload cfg5] [--batch B]

One "step" = one forward + one backward of the layer over one batch of B samples per GPU.  The default
workload is BASELINE.json's configs[4] ("Mixed L+Q+SOC+LMI, dim=32, batch=262144 sharded across 8xB200"):
every GPU owns a 32768-sample shard (weak scaling; 8 GPUs = the named 262144 batch), no data-path
collective.  Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for every field.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "ConstraintModule fwd+bwd samples/sec"
POOL = 16  # rotating buffer sets; their total size exceeds the 126 MB L2 for the default workload


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg5", choices=["cfg2", "cfg3", "cfg4", "cfg5"])
    ap.add_argument("--batch", type=int, default=0, help="samples per GPU (default: the workload's named batch)")
    ap.add_argument("--no-extra", action="store_true", help="skip the secondary configs / sweeps")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the CPU oracle timing (profiling runs)")
    ap.add_argument("--tm", type=int, default=0)
    ap.add_argument("--lanes", type=int, default=0)
    return ap.parse_args()


def workload_desc(name, shp, batch, world):
    fam = []
    if shp["m"]:
        fam.append(f"{shp['m']} linear rows")
    if shp["eta"]:
        fam.append(f"{shp['eta']} ellipsoids")
    if shp["mu"]:
        fam.append(f"{shp['mu']} cones (r_M={shp['r_M']})")
    if shp["r"]:
        fam.append(f"one {shp['r']}x{shp['r']} LMI")
    return f"{name}: dim={shp['k']}, " + " + ".join(fam) + f", batch {batch}/GPU x {world} GPU"


# ----------------------------------------------------------------------------- clocks sampler
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.path = tempfile.NamedTemporaryFile(prefix="clocks_", suffix=".csv", delete=False).name
        self.proc = None
        self.gpu_index = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu_index), f"--query-gpu={self.QUERY}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.path)
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(np.max(mx)), reasons=sorted(reasons), samples=len(sm))
        return out


# ----------------------------------------------------------------------------- the reference arm (CPU)
def cpu_oracle_rate(spec_name, sample, repeats, warmup, threads):
    """samples/sec of the op-for-op torch restatement of the reference (oracle port) on the host cores."""
    from oracle.rayen_oracle import OracleSet, TorchOracle
    from rayen_b200 import synthetic
    torch.set_num_threads(threads)
    spec = synthetic.config_spec(spec_name)
    cs = synthetic.build_constraints(spec)
    orc = TorchOracle(OracleSet.from_constraints(cs), torch.float32)
    v, gy = synthetic.sample_inputs(sample, cs.n, cs.k)
    times = []
    for it in range(warmup + repeats):
        t0 = time.perf_counter()
        orc.forward_backward(v, gy)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    return sample / float(np.mean(times)), float(np.mean(times)), float(np.min(times))


def run_reference(args, rank, world):
    """--impl reference: the reference's own PyTorch op sequence (oracle port; the reference is Python and
    cannot travel to the GPU box) on the host cores, same workload, bounded sample per step."""
    if rank != 0:
        return
    from rayen_b200 import synthetic
    shp = synthetic.CONFIG_SHAPES[args.workload]
    batch = args.batch or shp["batch"]
    cores = os.cpu_count() or 1
    sample = min(batch, 2048)
    steps, warmup = max(1, args.steps), max(1, min(args.warmup, 3))
    # keep the whole run bounded: ~0.3 s per 2048-sample step of cfg5 on 8 cores
    steps = min(steps, 40)
    rate, mean_t, _ = cpu_oracle_rate(args.workload, sample, steps, warmup, cores)
    line = {
        "impl": "reference", "metric": METRIC, "value": rate * 1.0, "unit": "samples/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": mean_t * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_desc(args.workload, shp, batch, world), "sample_per_step": sample},
        "cpu_baseline": {"value": rate, "unit": "samples/s", "cores": cores, "kind": "port",
                         "sample": f"{sample} samples of the workload per step, torch {torch.__version__} CPU fp32, "
                                   f"{cores} threads"},
        "e2e": {"value": rate, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- the B200 arm
class DeviceBench:
    """Pre-allocated buffer pool + direct C-ABI launches on the current stream (no allocation in the loop)."""

    def __init__(self, layer, batch, device, seed=1, pool=POOL):
        from rayen_b200 import _cabi, synthetic
        self.lib = _cabi.lib()
        self.cabi = _cabi
        self.layer, self.B, self.device = layer, batch, device
        self.n, self.k = layer.n, layer.k
        self.plan = layer._device_plan(device)
        self.sets = []
        self.pool = pool
        self.want_grad = 1   # forward leaves d kappa/du of the LMI-bound samples for backward (training step)
        for i in range(pool):
            v, gy = synthetic.sample_inputs(batch, self.n, self.k, seed_v=seed + i, seed_g=7 + i)
            self.sets.append(dict(
                v=v.to(device), gy=gy.to(device),
                y=torch.empty((batch, self.k), device=device), gv=torch.empty((batch, self.n), device=device),
                kappa=torch.empty((batch,), device=device), active=torch.empty((batch,), dtype=torch.int32, device=device),
                ws=torch.empty((max(self.plan.workspace_bytes(batch), 16),), dtype=torch.uint8, device=device)))
        self.stream = ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)

    def forward(self, s, stage=3):
        rc = self.lib.rayen_forward_stage_f32(self.plan.handle, s["v"].data_ptr(), self.n, s["y"].data_ptr(),
                                              s["kappa"].data_ptr(), s["active"].data_ptr(), self.B, 0, self.want_grad,
                                              stage, s["ws"].data_ptr(), self.stream)
        self.cabi.check(rc, "rayen_forward_stage_f32")

    def backward(self, s, stage=3):
        rc = self.lib.rayen_backward_stage_f32(self.plan.handle, s["v"].data_ptr(), self.n, s["gy"].data_ptr(),
                                               s["kappa"].data_ptr(), s["active"].data_ptr(), s["gv"].data_ptr(),
                                               self.n, self.B, 0, self.want_grad, stage, s["ws"].data_ptr(),
                                               self.stream)
        self.cabi.check(rc, "rayen_backward_stage_f32")

    def step(self, i):
        s = self.sets[i % self.pool]
        self.forward(s)
        self.backward(s)

    def time_loop(self, fn, steps, warmup):
        for i in range(warmup):
            fn(i)
        torch.cuda.synchronize(self.device)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(warmup + i)
        e1.record()
        torch.cuda.synchronize(self.device)
        return e0.elapsed_time(e1) / steps  # ms per call


def module_step_fn(layer, bench):
    """The user-facing path: nn.Module forward + autograd backward on device-resident tensors."""
    xs = [s["v"].clone().requires_grad_(True) for s in bench.sets]

    def fn(i):
        x = xs[i % bench.pool]
        x.grad = None
        y = layer(x.unsqueeze(2))
        y.backward(bench.sets[i % bench.pool]["gy"].view(bench.B, bench.k, 1))
    return fn


def e2e_step_fn(layer, bench, device):
    host = []
    for i in range(4):
        s = bench.sets[i]
        host.append(dict(v=s["v"].cpu().pin_memory(), gy=s["gy"].cpu().pin_memory(),
                         y=torch.empty((bench.B, bench.k)).pin_memory(), gv=torch.empty((bench.B, bench.n)).pin_memory()))

    def fn(i):
        h = host[i % 4]
        layer.forward_backward_host(h["v"], h["gy"], h["y"], h["gv"], device=device)
    return fn, host


def e2e_pipelined_ms(layer, host, device, steps, warmup, in_flight=2):
    """Same step, submitted without blocking with `in_flight` steps in the air (double-buffered input pipeline):
    wall-clock per step from the first submit to the last wait."""
    def run(count):
        for i in range(count):
            h = host[i % 4]
            if i >= in_flight:
                layer.host_wait((i - in_flight) % 4, device=device)
            layer.forward_backward_host(h["v"], h["gy"], h["y"], h["gv"], device=device, slot=i % 4)
        for i in range(max(0, count - in_flight), count):
            layer.host_wait(i % 4, device=device)
    run(warmup)
    torch.cuda.synchronize(device)
    t0 = time.perf_counter()
    run(steps)
    return (time.perf_counter() - t0) / steps * 1e3


def run_b200(args, rank, local_rank, world):
    from rayen_b200 import _cabi, synthetic
    from rayen_b200.constraint_module import ConstraintModule
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py --impl b200 needs a CUDA device (there is no CPU fallback)")
    _cabi.lib()
    device = torch.device("cuda", local_rank)
    torch.cuda.set_device(device)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        # NCCL prints its version banner to stdout (C level) at communicator creation: keep stdout for the JSON line
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group(backend="nccl", device_id=device)
            dist.barrier()
            torch.cuda.synchronize(device)
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)

    shp = synthetic.CONFIG_SHAPES[args.workload]
    batch = args.batch or shp["batch"]
    spec = synthetic.config_spec(args.workload)
    cs = synthetic.build_constraints(spec)
    layer = ConstraintModule(cs, create_map=False).to(device)
    if args.tm or args.lanes:
        layer.set_tuning(args.tm, args.lanes, device=device)
    bench = DeviceBench(layer, batch, device, seed=1 + 100 * rank)
    n, k = layer.n, layer.k
    steps, warmup = args.steps, max(args.warmup, 3)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(device)

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- headline: device-resident fwd+bwd through the C ABI (rayen_forward_f32 + rayen_backward_f32), the
    # boundary the Python drop-in binds; K steps back to back between two CUDA events on the launching stream
    clocks = ClockSampler(local_rank)
    barrier()
    for i in range(warmup):
        bench.step(i)
    barrier()
    clocks.start()
    launches0 = _cabi.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        bench.step(warmup + i)
    e1.record()
    barrier()
    launches = _cabi.launch_count() - launches0
    direct_ms = max_over_ranks(e0.elapsed_time(e1) / steps)
    # the same through nn.Module.forward + autograd backward (adds PyTorch's eager/autograd overhead per step)
    mod_fn = module_step_fn(layer, bench)
    ms_module = max_over_ranks(bench.time_loop(mod_fn, steps, warmup))
    # ... and replayed from CUDA graphs (one captured fwd+bwd per buffer set): launch overhead removed
    graph_ms = None
    try:
        graphs = []
        side = torch.cuda.Stream(device)
        with torch.cuda.stream(side):
            for s in bench.sets:
                gph = torch.cuda.CUDAGraph()
                bench.stream = ctypes.c_void_p(side.cuda_stream)
                with torch.cuda.graph(gph, stream=side):
                    bench.forward(s)
                    bench.backward(s)
                graphs.append(gph)
        bench.stream = ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)
        torch.cuda.synchronize(device)
        graph_ms = max_over_ranks(bench.time_loop(lambda i: graphs[i % POOL].replay(), steps, warmup))
    except Exception as exc:  # noqa: BLE001 - the graph variant is informational only
        bench.stream = ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)
        graph_err = repr(exc)[:200]
    t_extra = time.time()
    while time.time() - t_extra < 0.5:
        bench.step(0)
    torch.cuda.synchronize(device)
    clock_info = clocks.stop()

    # ---- e2e: host buffers through the public API (H2D + fwd + bwd + D2H every step)
    e2e_fn, host = e2e_step_fn(layer, bench, device)
    barrier()
    e2e_ms = bench.time_loop(e2e_fn, max(5, steps // 2), 3)
    barrier()
    e2e_ms = max_over_ranks(e2e_ms)
    e2e_pipe_ms = max_over_ranks(e2e_pipelined_ms(layer, host, device, max(8, steps // 2), 4))
    barrier()

    # the link this box gives the step: one 64 MB pinned copy each way, alone (explains e2e, which moves
    # batch*4*(n+k) bytes in each direction per step)
    link = {}
    try:
        hbuf = torch.empty(64 << 20, dtype=torch.uint8).pin_memory()
        dbuf = torch.empty(64 << 20, dtype=torch.uint8, device=device)
        for label, (src, dst) in (("h2d_gbs", (hbuf, dbuf)), ("d2h_gbs", (dbuf, hbuf))):
            dst.copy_(src, non_blocking=True)
            torch.cuda.synchronize(device)
            c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            c0.record()
            for _ in range(4):
                dst.copy_(src, non_blocking=True)
            c1.record()
            torch.cuda.synchronize(device)
            link[label] = 4 * (64 << 20) / (c0.elapsed_time(c1) * 1e-3) / 1e9
        del hbuf, dbuf
    except Exception as exc:  # noqa: BLE001 - informational
        link["error"] = repr(exc)[:120]

    if rank != 0:
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return

    value = world * batch / (direct_ms * 1e-3)
    line = {
        "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": world, "steps": steps, "warmup": warmup,
        "ms_per_step": direct_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": workload_desc(args.workload, shp, batch, world), "batch_per_gpu": batch,
                   "global_batch": batch * world, "n": n, "k": k,
                   "l2": f"inputs larger than L2: rotating pool of {POOL} buffer sets, "
                         f"{POOL * batch * 4 * (3 * n + 2 * k) / 1e6:.0f} MB of algorithmic traffic per cycle",
                   "path": "rayen_forward_f32 + rayen_backward_f32 (C ABI) on device-resident buffers, plain launches"},
        "clocks": {"sm_mhz": clock_info["sm_mhz"], "sm_max_mhz": clock_info["sm_max_mhz"],
                   "reasons": clock_info["reasons"], "samples": clock_info["samples"]},
        "e2e": {"value": world * batch / (e2e_ms * 1e-3), "unit": "samples/s",
                "h2d_bytes_per_step": batch * 4 * (n + k), "d2h_bytes_per_step": batch * 4 * (n + k),
                "ms_per_step": e2e_ms, "link": link,
                "pipelined": {"value": world * batch / (e2e_pipe_ms * 1e-3), "ms_per_step": e2e_pipe_ms, "steps_in_flight": 2,
                              "path": "forward_backward_host(slot=...) + host_wait: the next step's copy-in overlaps this "
                                      "step's kernels and copy-out; wall clock, every step's H2D and D2H inside"},
                "copy_floor_ms": (batch * 4 * (n + k) / 1e6 / max(min(link.get("h2d_gbs", 0.0), link.get("d2h_gbs", 0.0)), 1e-9)
                                  if "h2d_gbs" in link and "d2h_gbs" in link else None),
                "path": "ConstraintModule.forward_backward_host -> rayen_forward_backward_host_f32 "
                                               "(pinned host buffers; copy-in, kernels and copy-out on three streams, "
                                               "synchronous per step)"},
        "gpu_launches": int(launches),
        "module_autograd": {"value": world * batch / (ms_module * 1e-3), "ms_per_step": ms_module,
                            "path": "nn.Module forward + autograd backward (PyTorch eager overhead included)"},
        "cuda_graph": ({"value": world * batch / (graph_ms * 1e-3), "ms_per_step": graph_ms,
                        "path": "the C-ABI fwd+bwd of each buffer set captured once, replayed"} if graph_ms else None),
    }

    # ---- per-kernel durations (CUDA events on the launching stream) and the roofline of the dominant one
    has_lmi = shp["r"] > 0
    # Single-stage launches must see the state the previous stage leaves behind (kappa/active and the work
    # lists), so: full forward everywhere, then the two backward stages, then the two forward stages.
    for i in range(POOL):
        bench.forward(bench.sets[i])
    durs = {}
    durs["lqs_backward_kernel"] = bench.time_loop(lambda i: bench.backward(bench.sets[i % POOL], 1), steps, POOL)
    if has_lmi:
        durs["lmi_backward_kernel"] = bench.time_loop(lambda i: bench.backward(bench.sets[i % POOL], 2), steps, POOL)
    durs["lqs_forward_kernel"] = bench.time_loop(lambda i: bench.forward(bench.sets[i % POOL], 1), steps, POOL)
    if has_lmi:
        durs["lmi_forward_kernel"] = bench.time_loop(lambda i: bench.forward(bench.sets[i % POOL], 2), steps, POOL)
    kernel_names = {"lqs_forward_kernel": "lqs_tc_forward_kernel (tcgen05)" if n >= 16 else "lqs_forward_kernel (FP32 pipe)",
                    "lmi_forward_kernel": "lmi_forward_kernel<WITH_GRAD> (FP32-pipe contraction; the step carries gradient work)"}
    dominant = max(durs, key=durs.get)
    fwd_bytes, bwd_bytes = batch * 4 * (n + k), batch * 4 * (2 * n + k)
    alg_bytes = fwd_bytes if "forward" in dominant else bwd_bytes
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    achieved = alg_bytes / (durs[dominant] * 1e-3) / 1e9
    # DRAM traffic of the dominant kernel per launch, from the committed `ncu --set full` capture of this workload
    traffic, traffic_src = None, None
    prof = os.path.join(ROOT, "profiles", "r01_lmi_forward.md")
    if dominant == "lmi_forward_kernel" and args.workload == "cfg5" and batch == 32768 and os.path.isfile(prof):
        unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        tot = 0.0
        for row in open(prof):
            f = [x.strip() for x in row.split("|")]
            if len(f) > 3 and f[1] in ("dram__bytes_read.sum", "dram__bytes_write.sum") and f[2] in unit:
                tot += float(f[3]) * unit[f[2]]
        traffic, traffic_src = tot, "profiles/r01_lmi_forward.md (dram__bytes_read.sum + dram__bytes_write.sum, one launch)"
    step_bytes = fwd_bytes + bwd_bytes
    line["roofline"] = {
        "bound": "hbm", "kernel": dominant, "achieved": achieved, "peak": peak, "unit": "GB/s",
        "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
        "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": durs[dominant],
        "kernel_ms_all": durs, "kernel_names": kernel_names,
        "kernel_share_of_step": durs[dominant] / sum(durs.values()),
        "whole_step": {"algorithmic_bytes": step_bytes, "achieved": step_bytes / (direct_ms * 1e-3) / 1e9,
                       "frac": step_bytes / (direct_ms * 1e-3) / 1e9 / peak},
        "note": "the named shapes are FP32-issue/LSU bound, not HBM bound (DESIGN.md, Roofline); the HBM fraction "
                "is reported as the contract asks",
    }

    # ---- LMI forward without gradient work (inference): FP32-pipe contraction vs the tcgen05 contraction, forced
    if has_lmi:
        lmi_modes = {}
        for label, mode in (("fp32_pipe", 0), ("tcgen05", 1)):
            layer.set_lmi_tensor_cores(mode, device=device)
            bench.want_grad = 0
            lmi_modes[label + "_fwd_nograd_ms"] = bench.time_loop(lambda i: bench.forward(bench.sets[i % POOL]), steps, 3)
        layer.set_lmi_tensor_cores(None, device=device)
        bench.want_grad = 1
        line["lmi_contraction"] = dict(lmi_modes, note="whole forward (LQS + LMI kernels), want_grad=0; automatic policy: "
                                       "tcgen05 for K=32 dense launches without gradient work, FP32 pipe otherwise")

    # ---- feasibility of the outputs (fp64 residuals of every constraint on a sub-sample)
    from oracle.rayen_oracle import OracleSet, max_violation
    s0 = bench.sets[0]
    bench.forward(s0)
    torch.cuda.synchronize(device)
    sub = min(batch, 4096)
    line["max_violation"] = max_violation(OracleSet.from_constraints(cs), s0["y"][:sub].cpu().numpy(),
                                          spec["A1"], spec["b1"], spec["A2"], spec["b2"])
    # ... and of the FULL batch on the GPU (rayen_violation_f32, float32 residuals)
    line["max_violation_gpu_full_batch"] = float(layer.violation(s0["y"]).max())
    act = s0["active"].cpu().numpy() >> 24
    line["active_family_hist"] = {name: int(c) for name, c in zip(["none", "linear", "quad", "soc", "lmi"],
                                                                  np.bincount(act, minlength=5))}

    # ---- CPU baseline (oracle port) on this box's host cores, bounded sample; rank 0, N == 1 only
    if world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        sample = min(batch, 4096)
        rate, mean_t, best_t = cpu_oracle_rate(args.workload, sample, 8, 1, cores)
        line["cpu_baseline"] = {"value": rate, "unit": "samples/s", "cores": cores, "kind": "port",
                                "sample": f"{sample} samples of the same workload, mean of 8 passes after 1 warm-up, "
                                          f"torch {torch.__version__} CPU fp32, {cores} threads, {mean_t * 1e3:.0f} ms/pass"}

    # ---- the other BASELINE.json configs at their named batch, and a large-batch point (not bench lines)
    if not args.no_extra and world == 1:
        extra = {}
        for name in ("cfg2", "cfg3", "cfg4", "cfg5"):
            eshp = synthetic.CONFIG_SHAPES[name]
            for eb in sorted({eshp["batch"], 262144}):
                if name == args.workload and eb == batch:
                    continue
                ecs = synthetic.build_constraints(synthetic.config_spec(name))
                elayer = ConstraintModule(ecs, create_map=False).to(device)
                per_set = eb * 4 * (3 * elayer.n + 2 * elayer.k)
                eb_bench = DeviceBench(elayer, eb, device, pool=max(2, min(POOL, int(300e6 // per_set) + 1)))
                ms = eb_bench.time_loop(eb_bench.step, 30, 5)
                fwd_ms = eb_bench.time_loop(lambda i: eb_bench.forward(eb_bench.sets[i % eb_bench.pool]), 30, 5)
                abytes = eb * 4 * (3 * elayer.n + 2 * elayer.k)
                extra[f"{name}_B{eb}"] = {"fwd_bwd_samples_per_s": eb / (ms * 1e-3), "ms_per_step": ms, "fwd_ms": fwd_ms,
                                          "hbm_frac": abytes / (ms * 1e-3) / 1e9 / peak}
                del eb_bench, elayer
                torch.cuda.empty_cache()
        # wide sets (n > 32, wide.cuh): dim 64 (256 rows + 4 ellipsoids + 4 cones) and dim 256 (1024 rows)
        for wname, (wk, wm, weta, wmu, wrm, wb) in (("wide_n64", (64, 256, 4, 4, 32, 65536)),
                                                    ("wide_n256", (256, 1024, 0, 0, 0, 8192))):
            ecs = synthetic.build_constraints(synthetic.wide_spec(wk, wm, weta, wmu, wrm, 0, seed=1))
            elayer = ConstraintModule(ecs, create_map=False).to(device)
            per_set = wb * 4 * (3 * elayer.n + 2 * elayer.k)
            eb_bench = DeviceBench(elayer, wb, device, pool=max(2, min(POOL, int(300e6 // per_set) + 1)))
            ms = eb_bench.time_loop(eb_bench.step, 20, 5)
            fwd_ms = eb_bench.time_loop(lambda i: eb_bench.forward(eb_bench.sets[i % eb_bench.pool]), 20, 5)
            extra[f"{wname}_B{wb}"] = {"fwd_bwd_samples_per_s": wb / (ms * 1e-3), "ms_per_step": ms, "fwd_ms": fwd_ms,
                                       "hbm_frac": per_set / (ms * 1e-3) / 1e9 / peak}
            del eb_bench, elayer
            torch.cuda.empty_cache()
        line["extra"] = extra

    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    run_b200(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
