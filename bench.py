#!/usr/bin/env python
"""bench.py -- ConstraintModule forward+backward samples/sec (BASELINE.json metric) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload cfg5] [--batch B]
                    [--strong] [--gather] [--no-extra] [--no-cpu-baseline]

One "step" = one forward + one backward of the layer over one batch of B samples per GPU.  The default
workload is BASELINE.json's configs[4] ("Mixed L+Q+SOC+LMI, dim=32, batch=262144 sharded across 8xB200"):
every GPU owns a 32768-sample shard (weak scaling; 8 GPUs = the named 262144 batch), no data-path
collective.  `--strong` keeps the global batch at 262144 and splits it over the ranks.  `--gather` adds the optional
exchange step (NCCL all-gather of y, reduce-scatter of its gradient).  Prints ONE JSON line (rank 0).  See DESIGN.md
"Measurement" for every field.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "ConstraintModule fwd+bwd samples/sec"
POOL = 16  # rotating buffer sets; their total size exceeds the 126 MB L2 for the default workload
STRONG_GLOBAL_BATCH = 262144


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg5", choices=["cfg2", "cfg3", "cfg4", "cfg5"])
    ap.add_argument("--batch", type=int, default=0, help="samples per GPU (default: the workload's named batch)")
    ap.add_argument("--strong", action="store_true", help="strong scaling: global batch fixed at 262144, split over the ranks")
    ap.add_argument("--gather", action="store_true", help="also time the optional all-gather of y (+ reduce-scatter of g_y)")
    ap.add_argument("--no-extra", action="store_true", help="skip the secondary configs / sweeps")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the CPU reference timing (profiling runs)")
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


def flops_per_sample(shp):
    """Algorithmic flops of one forward in the z-space formulation: linear 2mn, quadratic eta(2n^2+4n), SOC
    mu(2 r_M n + 2n + 20), LMI contraction 2 n r^2 + eigen-solve 4/3 r^3, scale step 2kn; backward ~ one more pass."""
    n = k = shp["k"]
    r = shp["r"]
    parts = {"linear": 2.0 * shp["m"] * n, "quadratic": shp["eta"] * (2.0 * n * n + 4 * n),
             "soc": shp["mu"] * (2.0 * shp["r_M"] * n + 2 * n + 20), "lmi_contraction": 2.0 * n * r * r,
             "lmi_eigen": 4.0 / 3.0 * r ** 3, "scale": 2.0 * k * n}
    return parts


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
class CpuArm:
    """The reference's own implementation of the path on the host cores: the UNMODIFIED reference package (from
    oracle/_ref/, a verbatim git-ignored copy made by oracle/make_ref.py; kind "reference") when it is there, else the
    op-for-op torch restatement in oracle/rayen_oracle.py (kind "port").  One pass = y = layer(v); (y*g_y).sum().backward()."""

    def __init__(self, workload, threads):
        from rayen_b200 import synthetic
        from oracle import reference_loader
        torch.set_num_threads(threads)
        self.spec = synthetic.config_spec(workload)
        self.kind = "port"
        self.n = self.k = None
        if reference_loader.reference_available():
            try:
                ref = reference_loader.load_reference()
                prev = torch.get_default_dtype()
                torch.set_default_dtype(torch.float32)
                try:
                    cs = synthetic.build_constraints(self.spec, module=ref.constraints)
                    self.layer = ref.constraint_module.ConstraintModule(cs, method="RAYEN", create_map=False)
                finally:
                    torch.set_default_dtype(prev)
                self.n, self.k = cs.n, cs.k
                self.kind = "reference"
            except Exception as exc:  # noqa: BLE001 - fall back to the port, say why
                self.why_port = repr(exc)[:160]
        if self.kind == "port":
            from oracle.rayen_oracle import OracleSet, TorchOracle
            cs = synthetic.build_constraints(self.spec)
            self.orc = TorchOracle(OracleSet.from_constraints(cs), torch.float32)
            self.n, self.k = cs.n, cs.k

    def one_pass(self, v, gy):
        if self.kind == "reference":
            x = v.unsqueeze(2).clone().requires_grad_(True)
            y = self.layer(x)
            (y * gy.unsqueeze(2)).sum().backward()
            return y, x.grad
        return self.orc.forward_backward(v, gy)

    def time(self, sample, steps, warmup):
        from rayen_b200 import synthetic
        v, gy = synthetic.sample_inputs(sample, self.n, self.k)
        times = []
        for it in range(warmup + steps):
            t0 = time.perf_counter()
            self.one_pass(v, gy)
            dt = time.perf_counter() - t0
            if it >= warmup:
                times.append(dt)
        return sample / float(np.mean(times)), float(np.mean(times))

    def describe(self, sample, cores, per_pass_s):
        what = ("the unmodified reference (rayen.constraint_module.ConstraintModule, method='RAYEN', from oracle/_ref)"
                if self.kind == "reference" else "the oracle's op-for-op torch restatement of the reference")
        return (f"{sample} samples of the workload per pass, {what}, torch {torch.__version__} CPU fp32, {cores} threads, "
                f"{per_pass_s * 1e3:.0f} ms/pass")


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path on this box's host cores, same workload; one
    step = one forward+backward pass over a sample of the batch sized so that the whole run stays within minutes."""
    if rank != 0:
        return
    from rayen_b200 import synthetic
    shp = synthetic.CONFIG_SHAPES[args.workload]
    batch = args.batch or (STRONG_GLOBAL_BATCH // world if args.strong else shp["batch"])
    cores = os.cpu_count() or 1
    arm = CpuArm(args.workload, cores)
    steps, warmup = max(1, args.steps), max(1, args.warmup)
    probe_rate, _ = arm.time(min(batch, 1024), 1, 1)
    budget_s = 150.0
    sample = int(min(batch, max(1024, probe_rate * budget_s / (steps + warmup) // 1024 * 1024)))
    rate, mean_t = arm.time(sample, steps, warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": "samples/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": mean_t * 1e3, "higher_is_better": True,
        "scaling": "strong" if args.strong else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_desc(args.workload, shp, batch, world), "batch_per_gpu": batch,
                   "sample_per_step": sample, "same_batch_as_b200_arm": sample == batch},
        "cpu_baseline": {"value": rate, "unit": "samples/s", "cores": cores, "kind": arm.kind,
                         "sample": arm.describe(sample, cores, mean_t)},
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
                ws=torch.zeros((max(self.plan.workspace_bytes(batch), 16),), dtype=torch.uint8, device=device)))
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

    def time_loop_median(self, fn, steps, warmup, groups=5):
        """Median over `groups` timed loops: a secondary figure that one host-side hiccup (the clocks sampler's
        nvidia-smi poll holding a driver lock, a GC pause) must not decide."""
        vals = [self.time_loop(fn, steps, warmup if g == 0 else 1) for g in range(groups)]
        return float(np.median(vals))

    def counters(self, s):
        """(LMI work-list length, fail-list length) the last forward on this buffer set left in its workspace."""
        c = s["ws"][:12].view(torch.int32).cpu().numpy()
        return int(c[0]), int(c[2])


def module_step_fn(layer, bench):
    """The user-facing path: nn.Module forward + autograd backward on device-resident tensors."""
    xs = [s["v"].clone().requires_grad_(True) for s in bench.sets]

    def fn(i):
        x = xs[i % bench.pool]
        x.grad = None
        y = layer(x.unsqueeze(2))
        y.backward(bench.sets[i % bench.pool]["gy"].view(bench.B, bench.k, 1))
    return fn


def module_graph_steps(layer, bench):
    """The same user-facing step (nn.Module forward + autograd backward, input gradient included) captured whole into one
    CUDA graph per buffer set (rayen_b200.graphed.GraphedStep, loss = <y, g_y>): replay costs the kernels, not the eager
    dispatch and the autograd engine."""
    from rayen_b200.graphed import GraphedStep
    steps = []
    for s in bench.sets:
        gy = s["gy"].view(bench.B, bench.k, 1)
        steps.append(GraphedStep(layer, lambda y, t: (y * t).sum(), None, s["v"].unsqueeze(2), gy, input_grad=True, warmup=2))
    return steps


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


def measured_traffic(kernel):
    """DRAM bytes per launch of `kernel` from the committed ncu capture of this round (profiles/r02_traffic.json, written
    by scripts/summarize_profiles.py from `ncu --set full`: dram__bytes_read.sum + dram__bytes_write.sum), or None."""
    path = os.path.join(ROOT, "profiles", "r02_traffic.json")
    if not os.path.isfile(path):
        return None, None
    try:
        table = json.load(open(path))
    except ValueError:
        return None, None
    for name, entry in table.items():
        if name == kernel:
            return float(entry["dram_bytes"]), f"profiles/r02_traffic.json ({entry.get('source', 'ncu --set full')})"
    return None, None


def run_b200(args, rank, local_rank, world):
    from rayen_b200 import _cabi, synthetic
    from rayen_b200.constraint_module import ConstraintModule
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py --impl b200 needs a CUDA device (there is no CPU fallback)")
    lib = _cabi.lib()
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
    batch = args.batch or (STRONG_GLOBAL_BATCH // world if args.strong else shp["batch"])
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
    # boundary the Python drop-in binds.  One step = the forward and the backward call of one buffer set, captured once
    # per buffer set into a CUDA graph and replayed: K replays back to back between two CUDA events on the launching
    # stream.  (The kernels take 10-70 us; issued call by call the loop is bound by the host's launch rate on slower
    # boxes -- `direct_launch` below is that number.)  A capture failure falls back to direct launches, loudly.
    clocks = ClockSampler(local_rank)
    barrier()
    for i in range(warmup):
        bench.step(i)
    barrier()
    graphs, launches_per_step = [], None
    try:
        side = torch.cuda.Stream(device)
        with torch.cuda.stream(side):
            for s in bench.sets:
                gph = torch.cuda.CUDAGraph()
                bench.stream = ctypes.c_void_p(side.cuda_stream)
                l0 = _cabi.launch_count()
                with torch.cuda.graph(gph, stream=side):
                    bench.forward(s)
                    bench.backward(s)
                launches_per_step = _cabi.launch_count() - l0
                graphs.append(gph)
    except Exception as exc:  # noqa: BLE001
        print(f"bench.py: CUDA-graph capture failed ({exc}); timing direct launches", file=sys.stderr)
        graphs = []
    bench.stream = ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)
    torch.cuda.synchronize(device)
    step_fn = (lambda i: graphs[i % POOL].replay()) if graphs else bench.step
    for i in range(warmup):
        step_fn(i)
    barrier()
    clocks.start()
    launches0 = _cabi.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        step_fn(warmup + i)
    e1.record()
    barrier()
    launches = (launches_per_step * steps) if graphs else (_cabi.launch_count() - launches0)
    headline_ms = max_over_ranks(e0.elapsed_time(e1) / steps)
    # the same K steps issued call by call (no graph)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        bench.step(warmup + i)
    e1.record()
    barrier()
    direct_launch_ms = max_over_ranks(e0.elapsed_time(e1) / steps)
    direct_ms = headline_ms
    graph_ms = headline_ms if graphs else None
    # the same through nn.Module.forward + autograd backward (adds PyTorch's eager/autograd overhead per step)
    mod_fn = module_step_fn(layer, bench)
    ms_module = max_over_ranks(bench.time_loop_median(mod_fn, steps, warmup))
    ms_module_graph = None
    try:
        gsteps = module_graph_steps(layer, bench)
        ms_module_graph = max_over_ranks(bench.time_loop_median(lambda i: gsteps[i % POOL].graph.replay(), steps, warmup))
        del gsteps
    except Exception as exc:  # noqa: BLE001 - informational
        print(f"bench.py: GraphedStep capture failed ({exc})", file=sys.stderr)
    # launch floor: a chain of as many empty kernels as one step launches
    per_step_launches = max(1, int(round(launches / max(steps, 1))))
    floor_ms = bench.time_loop_median(lambda i: lib.rayen_launch_empty(per_step_launches, bench.stream), 200, 20, groups=3)
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

    # ---- optional exchange step: all-gather of y over NCCL, reduce-scatter of the gradient
    gather = None
    if args.gather and dist is not None:
        from rayen_b200 import sharding
        y_local = bench.sets[0]["y"].clone().requires_grad_(True)
        g_full = torch.randn(batch * world, k, device=device)

        def gather_fwd(i):
            sharding.all_gather_outputs(y_local.detach())

        def gather_both(i):
            y_local.grad = None
            sharding.all_gather_outputs(y_local).backward(g_full)
        barrier()
        fwd_ms = max_over_ranks(bench.time_loop_median(gather_fwd, 30, 5))
        barrier()
        both_ms = max_over_ranks(bench.time_loop_median(gather_both, 30, 5))
        out_bytes = batch * world * k * 4
        gather = {"all_gather_ms": fwd_ms, "all_gather_plus_reduce_scatter_ms": both_ms,
                  "bytes_in_per_rank": batch * k * 4, "bytes_out_per_rank": out_bytes,
                  "all_gather_bus_gbs": out_bytes * (world - 1) / world / (fwd_ms * 1e-3) / 1e9,
                  "step_with_gather_ms": direct_ms + both_ms,
                  "path": "rayen_b200.sharding.all_gather_outputs: dist.all_gather_into_tensor forward, "
                          "dist.reduce_scatter_tensor backward (NCCL)"}
        # the same exchange through this library's own kernel over peer memory (rayen_gather_push_f32: P2P stores, or one
        # multimem.st per 16 bytes through the NVSwitch multicast mapping), and fused into the forward kernels' epilogue
        try:
            pg = sharding.PeerGather(batch, k, device=device)
            yl = bench.sets[0]["y"].contiguous()
            vq = bench.sets[0]["v"]
            barrier()
            gather["push_p2p_ms"] = max_over_ranks(bench.time_loop_median(lambda i: pg.all_gather(yl, use_multicast=False), 30, 5))
            gather["has_multicast"] = bool(pg.has_multicast)
            if pg.has_multicast:
                barrier()
                gather["push_multicast_ms"] = max_over_ranks(bench.time_loop_median(lambda i: pg.all_gather(yl), 30, 5))
                with torch.no_grad():
                    barrier()
                    gather["forward_fused_epilogue_ms"] = max_over_ranks(
                        bench.time_loop_median(lambda i: sharding.forward_gathered(layer, vq, pg), 30, 5))
                    barrier()
                    gather["forward_then_push_multicast_ms"] = max_over_ranks(
                        bench.time_loop_median(lambda i: pg.all_gather(layer(vq.unsqueeze(2))[:, :, 0]), 30, 5))
                    barrier()
                    gather["forward_then_nccl_all_gather_ms"] = max_over_ranks(
                        bench.time_loop_median(lambda i: sharding.all_gather_outputs(layer(vq.unsqueeze(2))[:, :, 0]), 30, 5))
            best = min(v_ for k_, v_ in gather.items() if k_.startswith("push_") and k_.endswith("_ms"))
            gather["push_bus_gbs"] = out_bytes * (world - 1) / world / (best * 1e-3) / 1e9
            gather["peer_path"] = ("rayen_b200.sharding.PeerGather (torch symmetric memory for the handle exchange and the barrier; "
                                   "rayen_gather_push_f32 moves the bytes) / sharding.forward_gathered (y stored through the multicast "
                                   "mapping by the forward kernels themselves)")
        except Exception as exc:  # noqa: BLE001 - symmetric memory may be unavailable on a box; NCCL numbers stand
            gather["peer_path_error"] = repr(exc)[:300]
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

    # ... and what the step's copies alone cost when both directions run at once, as they do in the step: H2D of v and
    # g_y on one stream, D2H of y and g_v on another, no kernels (the floor of ANY host-buffer implementation on this box)
    try:
        h = host[0]
        s0 = bench.sets[0]
        sin, sout = torch.cuda.Stream(device), torch.cuda.Stream(device)

        def copies_only(i):
            cur = torch.cuda.current_stream(device)
            sin.wait_stream(cur)
            sout.wait_stream(cur)
            with torch.cuda.stream(sin):
                s0["v"].copy_(h["v"], non_blocking=True)
                s0["gy"].copy_(h["gy"], non_blocking=True)
            with torch.cuda.stream(sout):
                h["y"].copy_(s0["y"], non_blocking=True)
                h["gv"].copy_(s0["gv"], non_blocking=True)
            cur.wait_stream(sin)
            cur.wait_stream(sout)
        link["step_copies_both_directions_ms"] = bench.time_loop_median(copies_only, 20, 5, groups=3)
    except Exception as exc:  # noqa: BLE001 - informational
        link["copies_error"] = repr(exc)[:120]

    if rank != 0:
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return

    value = world * batch / (direct_ms * 1e-3)
    line = {
        "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": world, "steps": steps, "warmup": warmup,
        "ms_per_step": direct_ms, "higher_is_better": True, "scaling": "strong" if args.strong else "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_desc(args.workload, shp, batch, world), "batch_per_gpu": batch,
                   "global_batch": batch * world, "n": n, "k": k,
                   "l2": f"inputs larger than L2: rotating pool of {POOL} buffer sets, "
                         f"{POOL * batch * 4 * (3 * n + 2 * k) / 1e6:.0f} MB of algorithmic traffic per cycle",
                   "path": "rayen_forward_f32 + rayen_backward_f32 (C ABI; called here through their *_stage_* twins with "
                           "stage_mask = 3, which launch exactly the same kernels) on device-resident buffers; one step = the two "
                           "calls of one buffer set, captured once into a CUDA graph and replayed (direct_launch: issued call by call)"},
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
                "copies_only_ms": link.get("step_copies_both_directions_ms"),
                "path": "ConstraintModule.forward_backward_host -> rayen_forward_backward_host_f32 "
                        "(pinned host buffers; copy-in, kernels and copy-out on three streams, replayed from a CUDA graph, "
                        "synchronous per step); copy_floor_ms = bytes of one direction / the slower direction's solo rate, "
                        "copies_only_ms = the step's four copies with both directions busy and no kernels"},
        "gpu_launches": int(launches),
        "launch_floor_ms": floor_ms,
        "module_autograd": {"value": world * batch / (ms_module * 1e-3), "ms_per_step": ms_module,
                            "path": "nn.Module forward + autograd backward (PyTorch eager overhead included); median of 5 loops"},
        "module_cuda_graph": ({"value": world * batch / (ms_module_graph * 1e-3), "ms_per_step": ms_module_graph,
                               "path": "the same nn.Module forward + autograd backward captured whole into one CUDA graph "
                                       "(rayen_b200.graphed.GraphedStep) and replayed"} if ms_module_graph else None),
        "direct_launch": {"value": world * batch / (direct_launch_ms * 1e-3), "ms_per_step": direct_launch_ms,
                          "path": "the same K steps issued call by call (cudaLaunchKernel per kernel, no graph): host-launch-rate bound on slow hosts"},
        "cuda_graph": ({"value": world * batch / (graph_ms * 1e-3), "ms_per_step": graph_ms,
                        "path": "= the headline: the C-ABI fwd+bwd of each buffer set captured once, replayed"}
                       if graph_ms else None),
    }
    if gather is not None:
        line["gather"] = gather

    # ---- per-kernel durations (CUDA events on the launching stream) and the roofline of the dominant one
    has_lmi = shp["r"] > 0
    # Single-stage launches must see the state the previous stage leaves behind (kappa/active and the work
    # lists), so: full forward everywhere, then the two backward stages, then the two forward stages.
    for i in range(POOL):
        bench.forward(bench.sets[i])
    torch.cuda.synchronize(device)
    list_len, fail_len = bench.counters(bench.sets[0]) if has_lmi else (0, 0)
    durs = {}
    durs["lqs_backward_kernel"] = bench.time_loop(lambda i: bench.backward(bench.sets[i % POOL], 1), steps, POOL)
    durs["lqs_forward_kernel"] = bench.time_loop(lambda i: bench.forward(bench.sets[i % POOL], 1), steps, POOL)
    if has_lmi:
        durs["lmi_forward_kernel"] = bench.time_loop(lambda i: bench.forward(bench.sets[i % POOL], 2), steps, POOL)
    warp_path = has_lmi and os.environ.get("RAYEN_LMI_WARP", "2") != "0" and n >= 1 and shp["r"] > 8 and (shp["m"] or shp["eta"] or shp["mu"])
    kernel_names = {
        "lqs_forward_kernel": "lqs_tc_forward_kernel (tcgen05 3xTF32 GEMM)" if n >= 16 else "lqs_forward_kernel (FP32 pipe)",
        "lmi_forward_kernel": ("lmi_forward_warp_kernel<WITH_GRAD> x2 (definiteness filter + one-warp-per-matrix solver on the work "
                               "list, then the same kernel on the fail list)" if warp_path else
                               "lmi_forward_kernel<WITH_GRAD> (8 lanes per matrix, FP32-pipe contraction)"),
        "lqs_backward_kernel": "lqs_backward_kernel (closed-form backward, all families)"}
    dominant = max(durs, key=durs.get)
    fwd_bytes_s, bwd_bytes_s = 4 * (n + k), 4 * (2 * n + k)     # algorithmic bytes per sample
    units = {"lqs_forward_kernel": batch, "lqs_backward_kernel": batch,
             "lmi_forward_kernel": list_len if (has_lmi and shp["m"] + shp["eta"] + shp["mu"] > 0) else batch}
    per_unit = bwd_bytes_s if "backward" in dominant else fwd_bytes_s
    alg_bytes = units[dominant] * per_unit
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(peaks_path):
        peaks = json.load(open(peaks_path))
        peak, peak_src = float(peaks["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        tensor_peak = float(peaks.get("bf16_tflops", 0.0)) or None
    else:
        peak, peak_src, tensor_peak = 6650.0, "fallback (B200_PROFILING.md)", None
    achieved = alg_bytes / (durs[dominant] * 1e-3) / 1e9
    traffic, traffic_src = measured_traffic(dominant)
    step_bytes = batch * (fwd_bytes_s + bwd_bytes_s)
    fl = flops_per_sample(shp)
    fwd_flops = sum(fl.values())
    # FP32 FMA peak of this GPU: 148 SMs x 128 lanes x 2 flop x SM clock; TF32 tensor peak taken as half the measured bf16
    sm_ghz = (clock_info["sm_mhz"] or 1965.0) / 1e3
    fp32_peak_tflops = 148 * 128 * 2 * sm_ghz / 1e3
    lqs_flops = batch * (fl["linear"] + fl["quadratic"] + fl["soc"] + fl["scale"])
    compute = {
        "flops_per_sample_forward": fl, "flops_per_sample_forward_total": fwd_flops,
        "whole_step_tflops": batch * 2 * fwd_flops / (direct_ms * 1e-3) / 1e12,
        "fp32_fma_peak_tflops": fp32_peak_tflops,
        "whole_step_frac_of_fp32_peak": batch * 2 * fwd_flops / (direct_ms * 1e-3) / 1e12 / fp32_peak_tflops,
        "lqs_forward": {"algorithmic_tflops": lqs_flops / (durs["lqs_forward_kernel"] * 1e-3) / 1e12,
                        "note": "on the tensor pipe the kernel issues 3 TF32 MMAs per product (3xTF32) over zero-padded "
                                "128-row panels: issued flops are ~3.5x the algorithmic ones",
                        "tf32_peak_tflops_assumed": (tensor_peak / 2 if tensor_peak else None)},
    }
    if has_lmi:
        lmi_units = units["lmi_forward_kernel"]
        compute["lmi_forward"] = {
            "units_processed": lmi_units,
            "algorithmic_tflops": lmi_units * (fl["lmi_contraction"] + fl["lmi_eigen"]) / (durs["lmi_forward_kernel"] * 1e-3) / 1e12,
            "tensor_pipe": "none: the contraction of the surviving samples runs on the FP32 pipe, register-tiled 4 samples per "
                           "F~ word (lmi_warp.cuh); the tcgen05 contraction (lmi_tc.cuh) serves dense LMI-only inference launches"}
    line["roofline"] = {
        "bound": "hbm", "kernel": dominant, "achieved": achieved, "peak": peak, "unit": "GB/s",
        "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
        "algorithmic_bytes_per_unit": per_unit, "units_processed": units[dominant],
        "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": durs[dominant],
        "kernel_ms_all": durs, "kernel_names": kernel_names,
        "kernel_share_of_step": durs[dominant] / sum(durs.values()),
        "whole_step": {"algorithmic_bytes": step_bytes, "achieved": step_bytes / (direct_ms * 1e-3) / 1e9,
                       "frac": step_bytes / (direct_ms * 1e-3) / 1e9 / peak},
        "compute": compute,
        "launch_floor_ms": floor_ms,
        "note": "the named shapes are FP32-issue / latency bound, not HBM bound (DESIGN.md, Roofline); the HBM fraction "
                "is reported as the contract asks, on the units the launch really processes",
    }
    if has_lmi:
        line["prune"] = {"batch": batch, "work_list": list_len, "prune_rate": 1.0 - list_len / batch,
                         "failed_filter_beyond_in_kernel_budget": fail_len,
                         "note": "work_list = samples the Wolkowicz-Styan bound could not settle; they go through the LDL' "
                                 "definiteness filter, the ones that fail it through the eigen-solver"}

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

    # ---- CPU baseline on this box's host cores, bounded sample; rank 0, N == 1 only
    if world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        arm = CpuArm(args.workload, cores)
        sample = min(batch, 8192)
        rate, mean_t = arm.time(sample, 4, 1)
        line["cpu_baseline"] = {"value": rate, "unit": "samples/s", "cores": cores, "kind": arm.kind,
                                "sample": arm.describe(sample, cores, mean_t) + ", mean of 4 passes after 1 warm-up"}

    # ---- the other BASELINE.json configs at their named batch, large-batch points, data-dependence of the headline
    if not args.no_extra and world == 1:
        extra = {}

        def run_extra(label, ecs, eb, prune=True, steps_=30):
            elayer = ConstraintModule(ecs, create_map=False).to(device)
            if not prune:
                elayer.set_pruning(False, device=device)
            per_set = eb * 4 * (3 * elayer.n + 2 * elayer.k)
            eb_bench = DeviceBench(elayer, eb, device, pool=max(2, min(POOL, int(300e6 // per_set) + 1)))
            ms = eb_bench.time_loop(eb_bench.step, steps_, 5)
            fwd_ms = eb_bench.time_loop(lambda i: eb_bench.forward(eb_bench.sets[i % eb_bench.pool]), steps_, 5)
            entry = {"fwd_bwd_samples_per_s": eb / (ms * 1e-3), "ms_per_step": ms, "fwd_ms": fwd_ms,
                     "hbm_frac": per_set / (ms * 1e-3) / 1e9 / peak}
            if ecs.lmic is not None and (ecs.lc is not None or ecs.qcs or ecs.socs) and not elayer._packed.fields.get("lmi_big"):
                eb_bench.forward(eb_bench.sets[0])
                torch.cuda.synchronize(device)
                ll, fl_ = eb_bench.counters(eb_bench.sets[0])
                entry["work_list"] = ll if prune else eb
                entry["fail_list"] = fl_
            extra[label] = entry
            del eb_bench, elayer
            torch.cuda.empty_cache()

        for name in ("cfg2", "cfg3", "cfg4", "cfg5"):
            eshp = synthetic.CONFIG_SHAPES[name]
            ecs = synthetic.build_constraints(synthetic.config_spec(name))
            for eb in sorted({eshp["batch"], 262144, 1 << 20} | ({1 << 22} if name in ("cfg2", "cfg3") else set())):
                if name == args.workload and eb == batch:
                    continue
                if name == "cfg4" and eb > 262144:
                    continue           # dense 32x32 eigen-solves: 8 ms per 2^20, nothing new to learn
                run_extra(f"{name}_B{eb}", ecs, eb, steps_=30 if eb <= 262144 else 10)
        # how much the headline depends on the data: the same set with pruning off (every sample goes through the
        # filter), and the "loose" variant (rows relaxed 4x: every family binds, the LMI for ~13 % of the samples)
        ecs = synthetic.build_constraints(synthetic.config_spec("cfg5"))
        run_extra("cfg5_B32768_noprune", ecs, 32768, prune=False)
        lspec = synthetic.config_spec("cfg5")
        lspec["b1"] = lspec["b1"] * 4.0
        lcs = synthetic.build_constraints(lspec)
        run_extra("cfg5_loose_B32768", lcs, 32768)
        run_extra("cfg5_loose_B32768_noprune", lcs, 32768, prune=False)
        # wide sets (n > 32, wide.cuh): dim 64 (256 rows + 4 ellipsoids + 4 cones) and dim 256 (1024 rows)
        for wname, (wk, wm, weta, wmu, wrm, wb) in (("wide_n64", (64, 256, 4, 4, 32, 65536)),
                                                    ("wide_n256", (256, 1024, 0, 0, 0, 8192))):
            ecs = synthetic.build_constraints(synthetic.wide_spec(wk, wm, weta, wmu, wrm, 0, seed=1))
            run_extra(f"{wname}_B{wb}", ecs, wb, steps_=20)
        # LMIs beyond the register-resident kernels (lmi_big.cuh): a 64 x 64 LMI in dim 32 next to 64 rows,
        # the reference sweep's r_F = 100, k = 100 point (LMI only, 2000 samples), and a wide set (dim 64) with a 33 x 33 LMI
        bspec = synthetic.random_spec(k=32, m=64, r=64, seed=5)
        bspec["b1"] = bspec["b1"] * 3.0
        run_extra("lmi_r64_n32_B8192", synthetic.build_constraints(bspec), 8192, steps_=10)
        run_extra("lmi_r100_n100_B2000", synthetic.build_constraints(synthetic.random_spec(k=100, r=100, seed=6)), 2000, steps_=10)
        run_extra("wide_n64_lmi_r33_B8192", synthetic.build_constraints(synthetic.wide_spec(64, 128, 2, 2, 16, 0, seed=7, r=33)),
                  8192, steps_=10)
        # the widest linear point of the reference sweep that fits one block of constants: 100 rows in dimension 10000
        import numpy as _np
        _rng = _np.random.default_rng(10)
        wspec = dict(A1=_rng.uniform(-1.0, 1.0, size=(100, 10000)), b1=_rng.uniform(0.1, 1.0, size=(100, 1)), A2=None, b2=None,
                     qcs=[], socs=[], lmi=None, y0=_np.zeros((10000, 1)))
        run_extra("wide_n10000_rows100_B2000", synthetic.build_constraints(wspec), 2000, steps_=5)
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
