"""ctypes binding of ``librayen_b200.so`` (the C ABI declared in ``include/rayen_b200.h``).

There is no fallback: if the shared library is missing or cannot be loaded, every entry point raises
``RuntimeError`` telling the user to build it (``python -c "import __graft_entry__ as g; g.build()"``).
"""
import ctypes
import os
import subprocess
import threading

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# RAYEN_B200_LIB points development scripts at an instrumented build of the same sources (scripts/lmi_trace.py)
LIB_PATH = os.environ.get("RAYEN_B200_LIB") or os.path.join(CSRC, "librayen_b200.so")
SOURCES = ["rayen_b200.cu", "lqs.cuh", "lqs_tc.cuh", "lmi.cuh", "lmi_tc.cuh", "lmi_warp.cuh", "lmi_big.cuh", "lmi_big_tc.cuh", "viol.cuh", "wide.cuh",
           "common.cuh"]
HEADER = os.path.join(os.path.dirname(HERE), "include", "rayen_b200.h")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "--shared", "-Xcompiler", "-fPIC"]

ABI_VERSION = 15
MODE_RAYEN, MODE_RAYEN_OLD = 0, 1
FAM_NONE, FAM_LINEAR, FAM_QUAD, FAM_SOC, FAM_LMI = 0, 1, 2, 3, 4


class RayenPlanDesc(ctypes.Structure):
    _fields_ = [(name, ctypes.c_int32) for name in (
        "abi_version", "n", "k", "np", "k_pad", "m", "m_pad", "n_quad", "n_soc", "lmi_r", "lmi_rp",
        "n_is_identity", "lin_chunk_stride", "quad_stride", "soc_stride", "lmi_prune", "tc_panels", "tc_kp",
        "viol_in", "viol_eq", "lmitc_panels", "wide", "lmi_big", "lmib_p4", "lmibt_panels", "lmibt_slices")] + [("lmi_bound_margin", ctypes.c_float)] + [
        (name, ctypes.c_int64) for name in (
            "off_lin", "off_quad", "off_soc", "off_nmat", "off_y0", "off_bound", "off_lmi", "off_tc", "off_viol",
            "off_lmineg", "off_lmitc", "off_wide", "off_lmiw", "off_lmib", "off_lminegb", "off_lmibt", "blob_words")] + [
        ("blob", ctypes.POINTER(ctypes.c_float))]


class RayenKernelInfo(ctypes.Structure):
    _fields_ = [(name, ctypes.c_int32) for name in (
        "regs_lqs_fwd", "regs_lqs_bwd", "regs_lmi_fwd", "regs_lmi_bwd", "smem_lqs_bytes", "smem_lmi_bytes",
        "sm_count", "reserved")]


# every symbol include/rayen_b200.h declares: name -> (restype, argtypes)
_P = ctypes.c_void_p
SYMBOLS = {
    "rayen_abi_version": (ctypes.c_int, []),
    "rayen_last_error": (ctypes.c_char_p, []),
    "rayen_plan_create": (ctypes.c_int, [ctypes.POINTER(RayenPlanDesc), ctypes.c_int, ctypes.POINTER(_P)]),
    "rayen_plan_destroy": (None, [_P]),
    "rayen_plan_set_tuning": (ctypes.c_int, [_P, ctypes.c_int, ctypes.c_int]),
    "rayen_plan_set_pruning": (ctypes.c_int, [_P, ctypes.c_int]),
    "rayen_plan_set_tensor_cores": (ctypes.c_int, [_P, ctypes.c_int]),
    "rayen_plan_set_coalesced_output": (ctypes.c_int, [_P, ctypes.c_int]),
    "rayen_plan_set_lmi_tensor_cores": (ctypes.c_int, [_P, ctypes.c_int]),
    "rayen_plan_set_lmi_filter": (ctypes.c_int, [_P, ctypes.c_int]),
    "rayen_workspace_bytes": (ctypes.c_int64, [_P, ctypes.c_int64]),
    "rayen_forward_f32": (ctypes.c_int, [_P, _P, ctypes.c_int64, _P, _P, _P, ctypes.c_int64, ctypes.c_int,
                                         ctypes.c_int, _P, _P]),
    "rayen_forward_mapped_f32": (ctypes.c_int, [_P, _P, ctypes.c_int64, ctypes.c_int32, _P, ctypes.c_int64, _P, _P, _P,
                                                _P, _P, ctypes.c_int64, ctypes.c_int, _P, _P]),
    "rayen_backward_f32": (ctypes.c_int, [_P, _P, ctypes.c_int64, _P, _P, _P, _P, ctypes.c_int64, ctypes.c_int64,
                                          ctypes.c_int, ctypes.c_int, _P, _P]),
    "rayen_forward_stage_f32": (ctypes.c_int, [_P, _P, ctypes.c_int64, _P, _P, _P, ctypes.c_int64, ctypes.c_int,
                                               ctypes.c_int, ctypes.c_int, _P, _P]),
    "rayen_backward_stage_f32": (ctypes.c_int, [_P, _P, ctypes.c_int64, _P, _P, _P, _P, ctypes.c_int64,
                                                ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_int, _P, _P]),
    "rayen_host_workspace_bytes": (ctypes.c_int64, [_P, ctypes.c_int64]),
    "rayen_forward_backward_host_f32": (ctypes.c_int, [_P, _P, _P, _P, _P, ctypes.c_int64, _P, _P]),
    "rayen_forward_backward_host_submit_f32": (ctypes.c_int, [_P, _P, _P, _P, _P, ctypes.c_int64, _P, _P, ctypes.c_int]),
    "rayen_forward_backward_host_wait": (ctypes.c_int, [_P, ctypes.c_int]),
    "rayen_violation_f32": (ctypes.c_int, [_P, _P, ctypes.c_int64, _P, ctypes.c_int64, _P]),
    "rayen_gather_push_f32": (ctypes.c_int, [_P, ctypes.c_int64, ctypes.c_int32, ctypes.POINTER(_P), ctypes.c_int32, _P,
                                             ctypes.c_int64, _P]),
    "rayen_launch_count": (ctypes.c_int64, []),
    "rayen_launch_empty": (ctypes.c_int, [ctypes.c_int, _P]),
    "rayen_plan_kernel_info": (ctypes.c_int, [_P, ctypes.POINTER(RayenKernelInfo)]),
}

_lib = None
_lock = threading.Lock()


def needs_rebuild():
    if not os.path.isfile(LIB_PATH):
        return True
    built = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, s) for s in SOURCES] + [HEADER]
    return any(os.path.getmtime(d) > built for d in deps if os.path.isfile(d))


def build(force=False, verbose=False):
    """Compile the CUDA library in-tree for sm_100a with nvcc (cross-compiles without a GPU)."""
    if not force and not needs_rebuild():
        return LIB_PATH
    cmd = ["nvcc"] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH, "rayen_b200.cu"]
    proc = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + proc.stdout + proc.stderr)
    if verbose:
        print(proc.stderr)
    return LIB_PATH


def lib():
    """The loaded library (loads on first use).  Raises loudly when it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.isfile(LIB_PATH):
            raise RuntimeError(
                f"rayen_b200: the CUDA extension {LIB_PATH} is not built. Build it with "
                "`python -c \"import __graft_entry__ as g; g.build()\"` (needs nvcc). "
                "There is no CPU fallback.")
        handle = ctypes.CDLL(LIB_PATH)
        for name, (restype, argtypes) in SYMBOLS.items():
            fn = getattr(handle, name)
            fn.restype = restype
            fn.argtypes = argtypes
        got = handle.rayen_abi_version()
        if got != ABI_VERSION:
            raise RuntimeError(f"rayen_b200: {LIB_PATH} has ABI {got}, the Python side expects {ABI_VERSION}; rebuild")
        _lib = handle
    return _lib


def check(rc, what):
    if rc != 0:
        msg = lib().rayen_last_error().decode("utf-8", "replace")
        kind = "CUDA error" if rc > 0 else "error"
        raise RuntimeError(f"rayen_b200: {what} failed with {kind} {rc}: {msg}")


class DevicePlan:
    """Owns one ``rayen_plan_t`` (the constant block of a feasible set on one GPU)."""

    def __init__(self, packed_plan, device_index):
        self.packed = packed_plan
        self.device_index = int(device_index)
        self._handle = _P()
        desc = packed_plan.desc()
        check(lib().rayen_plan_create(ctypes.byref(desc), self.device_index, ctypes.byref(self._handle)),
              "rayen_plan_create")

    @property
    def handle(self):
        return self._handle

    def set_tuning(self, samples_per_thread=0, lanes_per_sample=0):
        check(lib().rayen_plan_set_tuning(self._handle, samples_per_thread, lanes_per_sample), "rayen_plan_set_tuning")

    def set_pruning(self, enabled=True):
        check(lib().rayen_plan_set_pruning(self._handle, 1 if enabled else 0), "rayen_plan_set_pruning")

    def set_tensor_cores(self, enabled=True):
        check(lib().rayen_plan_set_tensor_cores(self._handle, 1 if enabled else 0), "rayen_plan_set_tensor_cores")

    def set_coalesced_output(self, enabled=True):
        check(lib().rayen_plan_set_coalesced_output(self._handle, 1 if enabled else 0), "rayen_plan_set_coalesced_output")

    def set_lmi_tensor_cores(self, mode):
        """0 / False: never, 1 / True: wherever available, 2 / None: automatic."""
        mode = 2 if mode is None else int(mode)
        check(lib().rayen_plan_set_lmi_tensor_cores(self._handle, mode), "rayen_plan_set_lmi_tensor_cores")

    def set_lmi_filter(self, mode):
        """0 / False: never, 1 / True: wherever available, 2 / None: automatic."""
        mode = 2 if mode is None else int(mode)
        check(lib().rayen_plan_set_lmi_filter(self._handle, mode), "rayen_plan_set_lmi_filter")

    def workspace_bytes(self, batch):
        return int(lib().rayen_workspace_bytes(self._handle, int(batch)))

    def kernel_info(self):
        info = RayenKernelInfo()
        check(lib().rayen_plan_kernel_info(self._handle, ctypes.byref(info)), "rayen_plan_kernel_info")
        return {name: getattr(info, name) for name, _ in RayenKernelInfo._fields_ if name != "reserved"}

    def close(self):
        if self._handle:
            lib().rayen_plan_destroy(self._handle)
            self._handle = _P()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def launch_count():
    return int(lib().rayen_launch_count())
