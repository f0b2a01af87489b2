"""``ConstraintModule``: drop-in for the reference layer, running on hand-written sm_100a kernels.

Mirrors ``rayen/constraint_module.py`` of leggedrobotics/rayen @ 2f007f7c for the methods on the
ray-shooting path:

* constructor signature, attributes (``cs k n method mapper dim_after_map``) and the registered buffer
  names / shapes / dtypes of reference constraint_module.py:18-74 and :99-122, so reference
  ``state_dict``s load;
* ``forward(x)``: flatten -> ``mapper`` -> ray shooting -> ``[B, k, 1]`` (reference :520-533);
* ``getDimAfterMap gety0 getyFromz getzFromy`` (reference :506-518);
* ``method='RAYEN'`` (reference :468-474) and ``method='RAYEN_old'`` (reference :460-466).

What is different underneath: the chain of torch ops in ``computeKappa`` (reference :351-458) and the
autograd graph it records are replaced by ``rayen_forward_f32`` / ``rayen_backward_f32`` of
``librayen_b200.so`` (include/rayen_b200.h), called through ctypes on the current CUDA stream with a
``torch.autograd.Function`` that saves only ``(v, kappa, active)``.  The other methods of the reference
(UU, Bar, PP, UP, DC3) are comparison baselines outside this package's scope and raise
``NotImplementedError``.  There is no CPU path: a CPU input raises ``RuntimeError``.
"""
import ctypes
import warnings

import numpy as np
import torch
import torch.nn as nn

from . import _cabi, plan as plan_mod, utils

try:  # raw handle of the current stream without building a torch.cuda.Stream object (11 us -> < 1 us per call)
    _raw_stream = torch._C._cuda_getCurrentRawStream
except AttributeError:  # pragma: no cover - older / newer torch without the private helper
    def _raw_stream(index):
        return torch.cuda.current_stream(index).cuda_stream

_SUPPORTED = ("RAYEN", "RAYEN_old")
# B200, scripts/mapper_compare.py (input_dim 64, forward replayed from a CUDA graph): fused 6.3 vs 10.4 us at B = 500,
# 10.3 vs 10.4 us at B = 4096, 37 vs 23 us at B = 16384 (the in-kernel mapper is per-thread FP32 FMAs; cuBLAS wins
# once the GEMM is big enough to matter)
_FUSE_MAPPER_MAX_BATCH = 4096
_BASELINES = ("UU", "Bar", "PP", "UP", "DC3")


def _as_buffer(array_like):
    """Same conversion as the reference's ``torch.Tensor(np_array)``: the default dtype at ctor time."""
    arr = np.asarray(array_like, dtype=np.float64)
    return torch.tensor(arr, dtype=torch.get_default_dtype())


class _RayShoot(torch.autograd.Function):
    """q:[B, n(+1)] -> y:[B, k] through the C ABI; backward is the closed form.

    The Python around the two C calls is kept minimal (one scratch allocation per call, cached handles): at
    the named batch sizes the kernels take 0.1-0.2 ms, so interpreter overhead is what a training loop sees.
    """

    @staticmethod
    def forward(ctx, q, module):
        if not q.is_cuda:
            raise RuntimeError("rayen_b200.ConstraintModule runs on CUDA (sm_100a) only; move the model and its "
                               "inputs to a B200 device. There is no CPU path.")
        v = q.detach()
        if v.dtype != torch.float32:
            v = v.float()
        if v.stride(1) != 1 or v.stride(0) < v.shape[1]:
            v = v.contiguous()
        B, cols = v.shape
        device = v.device
        st = module._launch_state(device)
        k = module.k
        ws_words = st.ws_words(B)
        # one scratch tensor: kappa [B] | active [B] (int32 bits) | work lists; y is what autograd sees
        aux = torch.empty((2 * B + ws_words,), dtype=torch.float32, device=device)
        y = torch.empty((B, k), dtype=torch.float32, device=device)
        base = aux.data_ptr()
        switch = torch.cuda.current_device() != st.index
        if switch:
            prev = torch.cuda.current_device()
            torch.cuda.set_device(st.index)
        try:
            want_grad = 1 if ctx.needs_input_grad[0] else 0
            rc = st.forward(st.handle, v.data_ptr(), v.stride(0) if B > 0 else cols, y.data_ptr(), base, base + 4 * B, B,
                            module._mode, want_grad, base + 8 * B, _raw_stream(st.index))
        finally:
            if switch:
                torch.cuda.set_device(prev)
        if rc != 0:
            _cabi.check(rc, "rayen_forward_f32")
        ctx.module, ctx.in_dtype, ctx.aux, ctx.have_dkappa, ctx.state = module, q.dtype, aux, want_grad, st
        ctx.save_for_backward(v)
        module._last_aux = (aux, B)
        return y if q.dtype == torch.float32 else y.to(q.dtype)

    @staticmethod
    def backward(ctx, gy):
        (v,) = ctx.saved_tensors
        module, aux = ctx.module, ctx.aux
        gy = gy.detach()
        if gy.dtype != torch.float32:
            gy = gy.float()
        if not gy.is_contiguous():
            gy = gy.contiguous()
        B, cols = v.shape
        device = v.device
        st = ctx.state   # the plan the forward ran on (a load_state_dict in between builds new plans)
        gv = torch.empty((B, cols), dtype=torch.float32, device=device)
        base = aux.data_ptr()
        switch = torch.cuda.current_device() != st.index
        if switch:
            prev = torch.cuda.current_device()
            torch.cuda.set_device(st.index)
        try:
            rc = st.backward(st.handle, v.data_ptr(), v.stride(0) if B > 0 else cols, gy.data_ptr(), base, base + 4 * B,
                             gv.data_ptr(), cols, B, module._mode, ctx.have_dkappa, base + 8 * B, _raw_stream(st.index))
        finally:
            if switch:
                torch.cuda.set_device(prev)
        if rc != 0:
            _cabi.check(rc, "rayen_backward_f32")
        return (gv if ctx.in_dtype == torch.float32 else gv.to(ctx.in_dtype)), None


class _MappedRayShoot(torch.autograd.Function):
    """x:[B, input_dim], weight:[n, input_dim], bias:[n] -> y:[B, k] with the mapper fused into the forward kernel
    (``rayen_forward_mapped_f32``): the linear/quadratic/SOC kernel computes ``q = x W' + b`` itself, so
    ``q`` is written once (backward needs it) instead of making the round trip of a separate GEMM launch.
    Backward: ``g_q`` from ``rayen_backward_f32``, then the three products of ``nn.Linear``'s own backward."""

    @staticmethod
    def forward(ctx, x, weight, bias, module):
        B, in_dim = x.shape
        device = x.device
        st = module._launch_state(device)
        n, k = module.n, module.k
        ws_words = st.ws_words(B)
        aux = torch.empty((2 * B + ws_words,), dtype=torch.float32, device=device)
        q = torch.empty((B, n), dtype=torch.float32, device=device)
        y = torch.empty((B, k), dtype=torch.float32, device=device)
        base = aux.data_ptr()
        want_grad = 1 if any(ctx.needs_input_grad[:3]) else 0
        with torch.cuda.device(device):
            rc = st.forward_mapped(st.handle, x.data_ptr(), x.stride(0) if B > 0 else in_dim, in_dim, weight.data_ptr(),
                                   weight.stride(0), bias.data_ptr() if bias is not None else None, q.data_ptr(),
                                   y.data_ptr(), base, base + 4 * B, B, want_grad, base + 8 * B,
                                   torch.cuda.current_stream(device).cuda_stream)
        if rc != 0:
            _cabi.check(rc, "rayen_forward_mapped_f32")
        ctx.module, ctx.aux, ctx.have_dkappa, ctx.has_bias, ctx.state = module, aux, want_grad, bias is not None, st
        ctx.save_for_backward(x, weight, q)
        module._last_aux = (aux, B)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, weight, q = ctx.saved_tensors
        module, aux = ctx.module, ctx.aux
        gy = gy.detach()
        if gy.dtype != torch.float32:
            gy = gy.float()
        if not gy.is_contiguous():
            gy = gy.contiguous()
        B, n = q.shape
        device = q.device
        st = ctx.state
        gq = torch.empty((B, n), dtype=torch.float32, device=device)
        base = aux.data_ptr()
        with torch.cuda.device(device):
            rc = st.backward(st.handle, q.data_ptr(), n, gy.data_ptr(), base, base + 4 * B, gq.data_ptr(), n, B,
                             module._mode, ctx.have_dkappa, base + 8 * B, torch.cuda.current_stream(device).cuda_stream)
        if rc != 0:
            _cabi.check(rc, "rayen_backward_f32")
        gx = gq @ weight if ctx.needs_input_grad[0] else None
        gw = gq.t() @ x if ctx.needs_input_grad[1] else None
        gb = gq.sum(0) if (ctx.has_bias and ctx.needs_input_grad[2]) else None
        return gx, gw, gb, None


class _LaunchState:
    """Per-(module, device) cache of everything the hot path needs: plan handle, C entry points, sizes."""

    def __init__(self, dev_plan):
        lib = _cabi.lib()
        self.plan = dev_plan
        self.handle = dev_plan.handle
        self.index = dev_plan.device_index
        self.forward = lib.rayen_forward_f32
        self.forward_mapped = lib.rayen_forward_mapped_f32
        self.backward = lib.rayen_backward_f32
        self._ws = {}

    def ws_words(self, batch):
        w = self._ws.get(batch)
        if w is None:
            w = (self.plan.workspace_bytes(batch) + 3) // 4 + 4
            self._ws[batch] = w
        return w


class ConstraintModule(nn.Module):
    def __init__(self, cs, input_dim=None, method="RAYEN", create_map=True, args_DC3=None):
        super().__init__()
        self.method = method
        if method in _BASELINES:
            raise NotImplementedError(
                f"method='{method}' is one of the reference's comparison baselines (UU/Bar/PP/UP/DC3); "
                "rayen_b200 implements the ray-shooting methods 'RAYEN' and 'RAYEN_old' only")
        if method not in _SUPPORTED:
            raise NotImplementedError
        self.cs = cs
        self.k = cs.k  # dimension of the ambient space
        self.n = cs.n  # dimension of the embedded space
        self._mode = _cabi.MODE_RAYEN if method == "RAYEN" else _cabi.MODE_RAYEN_OLD

        # ---- buffers, same names / shapes as the reference (constraint_module.py:38-74)
        D = cs.A_p / ((cs.b_p - cs.A_p @ cs.z0) @ np.ones((1, cs.n)))
        all_P, all_q, all_r = utils.getAllPqrFromQcs(cs.qcs)
        all_M, all_s, all_c, all_d = utils.getAllMscdFromSocs(cs.socs)
        if cs.has_lmi_constraints:
            all_F = [np.array(F, dtype=np.float64) for F in cs.lmic.all_F]
            H = all_F[-1] + sum(cs.y0[i, 0] * all_F[i] for i in range(cs.lmic.dim()))
            Hinv = np.linalg.inv(H)
            self.register_buffer("mHinv", _as_buffer(-Hinv))
            self.register_buffer("L", _as_buffer(np.linalg.cholesky(Hinv)))
            # the reference accumulates H in place into its copy of all_F[-1] before registering it
            all_F = all_F[:-1] + [H]
        else:
            all_F = []
        self.register_buffer("D", _as_buffer(D))
        for name, seq in (("all_P", all_P), ("all_q", all_q), ("all_r", all_r), ("all_M", all_M),
                          ("all_s", all_s), ("all_c", all_c), ("all_d", all_d), ("all_F", all_F)):
            self.register_buffer(name, _as_buffer(np.array(seq)) if len(seq) else torch.zeros(0))
        for name in ("A_p", "b_p", "yp", "NA_E", "z0", "y0"):
            self.register_buffer(name, _as_buffer(getattr(cs, name)))
        if cs.has_quadratic_constraints:  # reference constraint_module.py:99-122
            y0 = np.asarray(cs.y0, dtype=np.float64)
            all_delta, all_phi = [], []
            for P, q, r in zip(all_P, all_q, all_r):
                w = y0.T @ P + q.T
                level = (0.5 * y0.T @ P @ y0 + q.T @ y0 + r).item()
                sigma = 2.0 * level
                all_phi.append(-w / sigma)
                all_delta.append((w.T @ w - 2.0 * level * P) / sigma ** 2)
            self.register_buffer("all_delta", _as_buffer(np.array(all_delta)))
            self.register_buffer("all_phi", _as_buffer(np.array(all_phi)))

        self.dim_after_map = self.n if method == "RAYEN" else self.n + 1
        if create_map:
            utils.verify(input_dim is not None, "input_dim needs to be provided")
            self.mapper = nn.Linear(input_dim, self.dim_after_map)
        else:
            self.mapper = nn.Sequential()  # mapper does nothing

        # fold the mapper into the forward kernel when the input layout allows it: True / False force it, "auto" (the
        # default) does it for batches in the launch-bound regime, where it was measured faster;
        # the kernel that hosts it runs unless the set is an LMI alone (same rule as the C side's has_lqs)
        self.fuse_mapper = "auto"
        self._lqs_kernel_runs = bool(cs.has_quadratic_constraints or cs.has_soc_constraints or np.any(np.asarray(D) != 0)
                                     or not cs.has_lmi_constraints)

        # host-side packed plan (float64 math, float32 block) and its per-GPU uploads
        self._packed = plan_mod.build_plan_from_constraints(cs)
        self._plans = {}
        self._states = {}
        self._last_aux = None

    # ------------------------------------------------------------------ plan management
    def _device_plan(self, device):
        index = device.index if device.index is not None else torch.cuda.current_device()
        dev_plan = self._plans.get(index)
        if dev_plan is None:
            dev_plan = _cabi.DevicePlan(self._packed, index)
            self._plans[index] = dev_plan
        return dev_plan

    def __getstate__(self):
        # ctypes handles / device plans are per process: drop them when the module is pickled (torch.save(model))
        state = self.__dict__.copy()
        for key in ("_plans", "_states"):
            state[key] = {}
        state["_last_aux"] = None
        state.pop("_host_ws", None)
        return state

    def _launch_state(self, device):
        st = self._states.get(device)
        if st is None:
            st = _LaunchState(self._device_plan(device))
            self._states[device] = st
        return st

    _BUFFER_NAMES = ("D", "all_P", "all_q", "all_r", "all_M", "all_s", "all_c", "all_d", "all_F", "A_p", "b_p", "yp",
                     "NA_E", "z0", "y0")

    def _buffer_snapshot(self):
        return {name: getattr(self, name).detach().cpu().clone() for name in self._BUFFER_NAMES}

    def _load_from_state_dict(self, state_dict, prefix, *args, **kwargs):
        before = self._buffer_snapshot()
        super()._load_from_state_dict(state_dict, prefix, *args, **kwargs)
        after = self._buffer_snapshot()
        # a state dict that carries this module's own constants (checkpoint / resume of the same set, EarlyStopping's
        # reload) changes nothing: keep the plan packed from the float64 sources and the live device plans
        if all(before[k].shape == after[k].shape and torch.equal(before[k], after[k]) for k in before):
            return
        self._repack_from_buffers()

    def _repack_from_buffers(self):
        """Rebuild the device plan from the registered buffers (after a ``load_state_dict`` that changed them)."""
        g = lambda name: getattr(self, name).detach().cpu().double().numpy()
        qcs = [(P, q, r) for P, q, r in zip(g("all_P"), g("all_q"), g("all_r"))] if self.all_P.numel() else []
        socs = [(M, s, c, d) for M, s, c, d in zip(g("all_M"), g("all_s"), g("all_c"), g("all_d"))] \
            if self.all_M.numel() else []
        lmi = None
        if self.all_F.numel():
            allF = g("all_F")
            y0 = g("y0")
            constant = allF[-1] - np.einsum("a,aij->ij", y0[:, 0], allF[:-1])  # undo the H accumulation
            lmi = [F for F in allF[:-1]] + [constant]
        # the loaded buffers may describe another set than self.cs: the violation checker then uses the polyhedron
        # encoded by A_p, b_p and N (lin_rows=None) instead of the original rows
        self._packed = plan_mod.build_plan(g("A_p"), g("b_p"), g("NA_E"), g("yp"), g("z0"), qcs, socs, lmi,
                                           lin_rows=None)
        self._lqs_kernel_runs = bool(qcs or socs or np.any(g("D") != 0) or lmi is None)
        # the old device plans are not destroyed here: an autograd context of a forward that has not run its backward
        # yet keeps its own reference (ctx.state) and must find the plan its kappa / active / workspace belong to;
        # DevicePlan.__del__ frees each plan with its last reference
        self._plans = {}
        self._states = {}

    def set_tuning(self, samples_per_thread=0, lanes_per_sample=0, device=None):
        """Override the launch geometry of the linear/quadratic/SOC kernel (0 = automatic)."""
        device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self._device_plan(device).set_tuning(samples_per_thread, lanes_per_sample)

    def forward_backward_host(self, v_host, gy_host, y_host=None, gv_host=None, device=None, slot=None):
        """End-to-end step on HOST buffers through ``rayen_forward_backward_host_f32``: copies ``v_host`` [B,n]
        and ``gy_host`` [B,k] (float32, ideally pinned) to the GPU, runs forward + backward, copies ``y`` [B,k]
        and ``g_v`` [B,n] back and synchronises.  method='RAYEN' only.  Returns (y_host, gv_host).

        ``slot=0..3`` queues the step without blocking (``rayen_forward_backward_host_submit_f32``): several steps can be
        in flight, each in its own slot with its own host buffers; ``host_wait(slot)`` blocks until that step's outputs are
        complete, and a slot may be reused only after it has been waited for."""
        utils.verify(self._mode == _cabi.MODE_RAYEN, "forward_backward_host supports method='RAYEN'")
        device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        for t in (v_host, gy_host):
            utils.verify(t.dtype == torch.float32 and t.is_contiguous() and not t.is_cuda, "host float32 contiguous tensors expected")
        B = v_host.shape[0]
        utils.verify(tuple(v_host.shape) == (B, self.n) and tuple(gy_host.shape) == (B, self.k), "bad shapes")
        if y_host is None:
            y_host = torch.empty((B, self.k), dtype=torch.float32).pin_memory()
        if gv_host is None:
            gv_host = torch.empty((B, self.n), dtype=torch.float32).pin_memory()
        lib = _cabi.lib()
        dev_plan = self._device_plan(device)
        need = lib.rayen_host_workspace_bytes(dev_plan.handle, B)
        pool = self.__dict__.setdefault("_host_ws", {})
        key = (device.index, 0 if slot is None else int(slot))
        ws = pool.get(key)
        if ws is None or ws.numel() < need:
            ws = torch.empty((max(int(need), 256),), dtype=torch.uint8, device=device)
            pool[key] = ws
        with torch.cuda.device(device):
            stream = ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)
            if slot is None:
                rc = lib.rayen_forward_backward_host_f32(dev_plan.handle, v_host.data_ptr(), gy_host.data_ptr(),
                                                         y_host.data_ptr(), gv_host.data_ptr(), B, ws.data_ptr(), stream)
            else:
                rc = lib.rayen_forward_backward_host_submit_f32(dev_plan.handle, v_host.data_ptr(), gy_host.data_ptr(),
                                                                y_host.data_ptr(), gv_host.data_ptr(), B, ws.data_ptr(),
                                                                stream, int(slot))
        _cabi.check(rc, "rayen_forward_backward_host_f32" if slot is None else "rayen_forward_backward_host_submit_f32")
        return y_host, gv_host

    def host_wait(self, slot, device=None):
        """Block until the step submitted in ``slot`` by ``forward_backward_host(..., slot=slot)`` is complete."""
        device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        _cabi.check(_cabi.lib().rayen_forward_backward_host_wait(self._device_plan(device).handle, int(slot)),
                    "rayen_forward_backward_host_wait")

    def set_pruning(self, enabled=True, device=None):
        """LMI pruning on/off (results are identical; see include/rayen_b200.h)."""
        device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self._device_plan(device).set_pruning(enabled)

    def set_tensor_cores(self, enabled=True, device=None):
        """Linear/quadratic/SOC forward on tcgen05 tensor cores (default) or on the FP32 pipe."""
        device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self._device_plan(device).set_tensor_cores(enabled)

    def set_lmi_tensor_cores(self, mode=None, device=None):
        """LMI contraction ``sum_a u_a F_a`` as a tcgen05 GEMM inside the LMI kernel (True / 1: wherever the LMI is
        larger than 8x8), on the FP32 pipe (False / 0), or by the measured policy (None / 2, the default)."""
        device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self._device_plan(device).set_lmi_tensor_cores(mode)

    def set_lmi_filter(self, mode=None, device=None):
        """LMI forward behind another family's kappa: definiteness filter + one-warp-per-matrix solver (True / 1), the
        8-lanes-per-matrix kernels (False / 0), or automatic (None / 2, the default).  Results are identical."""
        device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self._device_plan(device).set_lmi_filter(mode)

    def violation(self, y):
        """Max constraint residual of every sample of ``y`` ([B, k] or [B, k, 1], CUDA) against the original
        constraints, computed on the GPU by ``rayen_violation_f32`` (<= 0 means feasible)."""
        if not y.is_cuda:
            raise RuntimeError("ConstraintModule.violation needs a CUDA tensor")
        yy = y.detach().reshape(y.shape[0], -1).float().contiguous()
        utils.verify(yy.shape[1] == self.k, f"expected {self.k} values per sample")
        out = torch.empty((yy.shape[0],), dtype=torch.float32, device=yy.device)
        plan = self._device_plan(yy.device)
        with torch.cuda.device(yy.device):
            rc = _cabi.lib().rayen_violation_f32(plan.handle, yy.data_ptr(), yy.stride(0) if yy.shape[0] else self.k,
                                                 out.data_ptr(), yy.shape[0],
                                                 ctypes.c_void_p(torch.cuda.current_stream(yy.device).cuda_stream))
        _cabi.check(rc, "rayen_violation_f32")
        return out

    def last_kappa_and_active(self):
        """(kappa[B], active[B]) of the most recent forward: active = family << 24 | constraint index."""
        if self._last_aux is None:
            return None
        aux, B = self._last_aux
        return aux[:B], aux[B:2 * B].view(torch.int32)

    # ------------------------------------------------------------------ reference helper methods
    def getDimAfterMap(self):
        return self.dim_after_map

    def gety0(self):
        return self.getyFromz(self.z0)

    def getyFromz(self, z):
        return self.NA_E @ z + self.yp

    def getzFromy(self, y):
        return self.NA_E.T @ (y - self.yp)

    # ------------------------------------------------------------------ forward
    def _can_fuse_mapper(self, x2d):
        """The mapper is folded into the forward kernel when the layout allows it (see rayen_forward_mapped_f32)."""
        m = self.mapper
        want = self.fuse_mapper if isinstance(self.fuse_mapper, bool) else x2d.shape[0] <= _FUSE_MAPPER_MAX_BATCH
        # anything the kernel cannot take as it is (a width that is not the mapper's, tensors on different devices or of
        # other dtypes, unaligned rows) goes through nn.Linear, which raises the usual shape / device errors
        return (want and isinstance(m, nn.Linear) and self._mode == _cabi.MODE_RAYEN and x2d.is_cuda
                and x2d.dtype == torch.float32 and m.weight.dtype == torch.float32 and x2d.shape[0] > 0
                and x2d.shape[1] == m.in_features and m.out_features == self.dim_after_map
                and m.weight.device == x2d.device
                and (m.bias is None or (m.bias.device == x2d.device and m.bias.dtype == torch.float32
                                        and m.bias.is_contiguous()))
                and x2d.shape[1] % 4 == 0 and x2d.stride(1) == 1 and x2d.stride(0) % 4 == 0
                and x2d.data_ptr() % 16 == 0 and m.weight.is_contiguous() and m.weight.data_ptr() % 16 == 0
                and self._lqs_kernel_runs and not self._packed.fields.get("wide"))

    def forward(self, x):
        # x: [B, numel_input_mapper, 1] (anything that views to [B, -1]), as in the reference
        x2d = x.view(x.size(0), x[0].numel() if x.size(0) else int(np.prod(x.shape[1:])))
        if self._can_fuse_mapper(x2d):
            return _MappedRayShoot.apply(x2d, self.mapper.weight, self.mapper.bias, self).unsqueeze(2)
        q = self.mapper(x2d)
        utils.verify(q.shape[1] == self.dim_after_map,
                     f"the layer expects {self.dim_after_map} values per sample, got {q.shape[1]}")
        if q.dtype == torch.float64 and not getattr(self, "_warned_f64", False):
            warnings.warn("rayen_b200 computes in float32; float64 inputs are cast down and the result cast back")
            self._warned_f64 = True
        y = _RayShoot.apply(q, self)
        return y.unsqueeze(2)
