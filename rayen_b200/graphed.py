"""Whole-step CUDA-graph capture for models that end in a ``ConstraintModule``.

At the named batch sizes the layer's kernels take 10-70 us, less than what PyTorch's eager dispatch and autograd
engine spend per step on the host (B200: 0.14 ms per ``model(x); loss.backward()`` against 0.07 ms of GPU work at
cfg5).  The layer is capture-safe -- its C entry points only launch on the current stream, never synchronise and never
allocate -- so the usual recipe applies: run forward, loss, backward and the optimizer step once under
``torch.cuda.graph`` on static input buffers, then replay.  ``GraphedStep`` is that recipe in one object; it is the
reference's training step (examples/main.py:135-171: ``y = model(x); loss = ...; loss.backward(); optimizer.step()``)
with one ``cudaGraphLaunch`` per step.
"""
import torch


class GraphedStep:
    """``step = GraphedStep(model, loss_fn, optimizer, example_inputs, example_targets)`` then ``loss = step(x, t)``.

    ``loss_fn(y, *targets)`` must return a scalar tensor.  ``optimizer`` may be None (forward + backward only; gradients
    are left in ``.grad`` of the parameters and, when ``input_grad=True``, in ``step.input_grad``).  Inputs of other
    shapes than the examples need their own GraphedStep (a CUDA graph is shape-specific).
    """

    def __init__(self, model, loss_fn, optimizer, example_inputs, example_targets=(), input_grad=False, warmup=3):
        if not isinstance(example_inputs, (tuple, list)):
            example_inputs = (example_inputs,)
        if not isinstance(example_targets, (tuple, list)):
            example_targets = (example_targets,)
        if not all(t.is_cuda for t in example_inputs):
            raise RuntimeError("GraphedStep needs CUDA tensors")
        self.model, self.loss_fn, self.optimizer = model, loss_fn, optimizer
        self.inputs = [t.detach().clone().requires_grad_(input_grad) for t in example_inputs]
        self.targets = [t.detach().clone() for t in example_targets]
        self.input_grad_enabled = input_grad
        # warm-up on a side stream (allocator, lazy plan upload, cuBLAS handles) as torch.cuda.graphs asks for
        side = torch.cuda.Stream(device=self.inputs[0].device)
        side.wait_stream(torch.cuda.current_stream(self.inputs[0].device))
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self._eager_step(set_to_none=True)
        torch.cuda.current_stream(self.inputs[0].device).wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        if optimizer is not None:
            optimizer.zero_grad(set_to_none=True)
        with torch.cuda.graph(self.graph):
            self.loss = self._eager_step(set_to_none=False, zero=False)
        self.input_grad = [t.grad for t in self.inputs] if input_grad else None

    def _eager_step(self, set_to_none, zero=True):
        if zero:
            if self.optimizer is not None:
                self.optimizer.zero_grad(set_to_none=set_to_none)
            else:
                for p in self.model.parameters():
                    p.grad = None
            for t in self.inputs:
                t.grad = None
        y = self.model(*self.inputs)
        loss = self.loss_fn(y, *self.targets)
        loss.backward()
        if self.optimizer is not None:
            self.optimizer.step()
        return loss.detach()

    def __call__(self, inputs, targets=()):
        if not isinstance(inputs, (tuple, list)):
            inputs = (inputs,)
        if not isinstance(targets, (tuple, list)):
            targets = (targets,)
        with torch.no_grad():
            for dst, src in zip(self.inputs, inputs):
                dst.copy_(src)
            for dst, src in zip(self.targets, targets):
                dst.copy_(src)
        self.graph.replay()
        return self.loss
