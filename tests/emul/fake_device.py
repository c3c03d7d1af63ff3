"""TEST INFRASTRUCTURE ONLY.  Lets the CPU suite run the product's Python layer -- the drop-in packages,
h3dgs.pipeline, h3dgs.dist, h3dgs.graphstep -- unchanged on CPU tensors, with libh3dgs_emu.so (the kernels
compiled against the SIMT emulator) standing in for libh3dgs.so: ctypes loads the emulation build, the
"is this a device tensor" checks of the shims answer yes, and the handful of torch.cuda stream / event calls
become no-ops.  CUDA graphs and NCCL are out of reach (GraphedStep runs with capture=False, collectives on gloo)."""
import contextlib
import ctypes as C
from unittest import mock

import torch


class _Stream:
    cuda_stream = 0

    def __init__(self, *a, **k):
        pass

    def wait_event(self, *_a):
        pass

    def wait_stream(self, *_a):
        pass


class _Event:
    def __init__(self, *a, **k):
        pass

    def record(self, *_a):
        pass


@contextlib.contextmanager
def _null(*_a, **_k):
    yield


@contextlib.contextmanager
def cpu_as_device(so_path):
    from h3dgs import _lib
    import diff_gaussian_rasterization._C as rc
    import gaussian_hierarchy._C as gc
    emu = _lib.bind(C.CDLL(so_path))
    patches = [
        mock.patch.object(_lib, "_lib", emu),
        mock.patch.object(rc, "_on_device", lambda t: True),
        mock.patch.object(gc, "_on_device", lambda t: True),
        mock.patch.object(torch.cuda, "current_stream", lambda *a, **k: _Stream()),
        mock.patch.object(torch.cuda, "Stream", _Stream),
        mock.patch.object(torch.cuda, "Event", _Event),
        mock.patch.object(torch.cuda, "stream", _null),
        mock.patch.object(torch.cuda, "device", _null),
        mock.patch.object(torch.cuda, "synchronize", lambda *a, **k: None),
        mock.patch.object(torch.cuda, "is_current_stream_capturing", lambda: False),
    ]
    with contextlib.ExitStack() as st:
        for p in patches:
            st.enter_context(p)
        yield emu


def _is_cuda_dev(d):
    return (isinstance(d, str) and d.startswith("cuda")) or (isinstance(d, torch.device) and d.type == "cuda")


@contextlib.contextmanager
def cuda_names_mean_cpu():
    """device="cuda" / .cuda() / .to("cuda") in test code land on the CPU: lets the GPU test files themselves run
    against the emulation build (H3DGS_EMULATE=1, tests/conftest.py)."""
    names = ["tensor", "zeros", "ones", "empty", "full", "rand", "randn", "arange", "as_tensor", "zeros_like", "ones_like",
             "empty_like", "full_like", "rand_like", "randn_like", "linspace", "eye", "randint", "randperm", "range"]
    saved = {n: getattr(torch, n) for n in names}
    saved_to, saved_cuda = torch.Tensor.to, torch.Tensor.cuda

    def wrap(fn):
        def f(*a, **k):
            if _is_cuda_dev(k.get("device")):
                k["device"] = "cpu"
            return fn(*a, **k)
        return f

    def to(self, *a, **k):
        a = tuple("cpu" if _is_cuda_dev(x) else x for x in a)
        if _is_cuda_dev(k.get("device")):
            k["device"] = "cpu"
        return saved_to(self, *a, **k)
    try:
        for n in names:
            setattr(torch, n, wrap(saved[n]))
        torch.Tensor.to = to
        torch.Tensor.cuda = lambda self, *a, **k: self
        yield
    finally:
        for n in names:
            setattr(torch, n, saved[n])
        torch.Tensor.to, torch.Tensor.cuda = saved_to, saved_cuda
