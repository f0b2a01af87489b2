"""Fused mapper (rayen_forward_mapped_f32) vs nn.Linear + layer: forward and forward+backward time through the module."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rayen_b200 import synthetic
from rayen_b200.constraint_module import ConstraintModule
dev = torch.device("cuda", 0)


def timed(fn, steps=30, warmup=5):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps * 1e3


for name, batch, in_dim in (("cfg1", 500, 64), ("cfg2", 4096, 64), ("cfg3", 16384, 64), ("cfg5", 32768, 64), ("cfg5", 262144, 64)):
    cs = synthetic.build_constraints(synthetic.config_spec(name))
    layer = ConstraintModule(cs, input_dim=in_dim, create_map=True).to(dev)
    x = torch.rand(batch, in_dim, 1, device=dev) * 2 - 1
    gy = torch.randn(batch, cs.k, 1, device=dev)
    res = {}
    for fused in (True, False):
        layer.fuse_mapper = fused
        with torch.no_grad():
            g = torch.cuda.CUDAGraph()
            s = torch.cuda.Stream()
            with torch.cuda.stream(s):
                layer(x)
                torch.cuda.synchronize()
                with torch.cuda.graph(g, stream=s):
                    y = layer(x)
            res[("fused" if fused else "unfused") + "_fwd_graph_us"] = round(timed(g.replay), 1)
        xi = x.clone().requires_grad_(True)

        def step():
            layer.zero_grad(set_to_none=True)
            xi.grad = None
            layer(xi).backward(gy)
        res[("fused" if fused else "unfused") + "_fwd_bwd_eager_us"] = round(timed(step), 1)
    print(name, "B", batch, "input_dim", in_dim, res, flush=True)
