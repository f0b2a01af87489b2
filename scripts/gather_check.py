"""The exchange step of the multi-GPU path on hardware: NCCL all-gather against this library's own kernel
over peer memory (multicast / P2P) and against the all-gather fused into the forward kernels' epilogue.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 \
        scripts/gather_check.py [--out gpurun_out/r02_gather_N.json]

Checks (every rank): the three gathered tensors are bit-identical; the gradient through the fused path equals the one
through NCCL all-gather + reduce-scatter.  Timings: CUDA events on the launching stream, max over ranks."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from rayen_b200 import synthetic, sharding
from rayen_b200.constraint_module import ConstraintModule

ap = argparse.ArgumentParser()
ap.add_argument("--out", default="")
ap.add_argument("--batch", type=int, default=32768)
args = ap.parse_args()
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
B = args.batch
cs = synthetic.build_constraints(synthetic.config_spec("cfg5"))
layer = ConstraintModule(cs, create_map=False).to(dev)
v, gy = synthetic.sample_inputs(B, cs.n, cs.k, seed_v=1 + 100 * rank)
v = v.to(dev)
torch.manual_seed(3)
g_full = torch.randn(B * world, cs.k, device=dev)          # the same on every rank (replicated loss)


def max_over_ranks(x):
    t = torch.tensor([x], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def timed(fn, iters=30, warm=5):
    for _ in range(warm):
        fn()
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return max_over_ranks(e0.elapsed_time(e1) / iters)


out = {"world": world, "batch_per_rank": B, "k": cs.k, "bytes_in_per_rank": B * cs.k * 4, "bytes_out_per_rank": B * world * cs.k * 4}
# reference path: forward, NCCL all-gather, backward through reduce-scatter
x = v.clone().requires_grad_(True)
y_local = layer(x.unsqueeze(2))[:, :, 0]
full_ref = sharding.all_gather_outputs(y_local)
(full_ref * g_full).sum().backward()
gv_ref = x.grad.clone()
pg = sharding.PeerGather(B, cs.k, device=dev)
out["has_multicast"] = bool(pg.has_multicast)
yl = y_local.detach().contiguous()
full_p2p = pg.all_gather(yl, use_multicast=False).clone()
ok = {"p2p": bool(torch.equal(full_p2p, full_ref.detach()))}
if pg.has_multicast:
    full_mc = pg.all_gather(yl, use_multicast=True).clone()
    ok["multicast"] = bool(torch.equal(full_mc, full_ref.detach()))
    x2 = v.clone().requires_grad_(True)
    full_fused = sharding.forward_gathered(layer, x2, pg)
    ok["fused_forward"] = bool(torch.equal(full_fused.detach(), full_ref.detach()))
    (full_fused * g_full).sum().backward()
    ok["fused_backward"] = bool(torch.equal(x2.grad, gv_ref))
flags = torch.tensor([int(all(ok.values()))], device=dev)
dist.all_reduce(flags, op=dist.ReduceOp.MIN)
out["checks_this_rank"] = ok
out["ok_all_ranks"] = bool(flags.item())

# timings
with torch.no_grad():
    xg = v.unsqueeze(2)
    out["forward_ms"] = timed(lambda: layer(xg))
    out["nccl_all_gather_ms"] = timed(lambda: sharding.all_gather_outputs(yl))
    out["forward_plus_nccl_all_gather_ms"] = timed(lambda: sharding.all_gather_outputs(layer(xg)[:, :, 0]))
    out["push_p2p_ms"] = timed(lambda: pg.all_gather(yl, use_multicast=False))
    out["forward_plus_push_p2p_ms"] = timed(lambda: pg.all_gather(layer(xg)[:, :, 0], use_multicast=False))
    if pg.has_multicast:
        out["push_multicast_ms"] = timed(lambda: pg.all_gather(yl, use_multicast=True))
        out["forward_plus_push_multicast_ms"] = timed(lambda: pg.all_gather(layer(xg)[:, :, 0], use_multicast=True))
        out["forward_fused_epilogue_ms"] = timed(lambda: sharding.forward_gathered(layer, v, pg))
    out["barrier_only_ms"] = timed(lambda: pg.barrier())
for key in [k_ for k_ in out if k_.endswith("_ms") and ("push" in k_ or "nccl_all_gather_ms" == k_)]:
    out[key.replace("_ms", "_bus_gbs")] = out["bytes_out_per_rank"] * (world - 1) / world / (out[key] * 1e-3) / 1e9
if rank == 0:
    print(json.dumps(out), flush=True)
    if args.out:
        os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
        json.dump(out, open(args.out, "w"), indent=1)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if out["ok_all_ranks"] else 1)
