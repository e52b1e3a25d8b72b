"""ctypes binding of libd3feat_b200.so (the C ABI of include/d3feat_b200.h).

There is no CPU fallback and no second backend: if the shared library is missing or a call fails the
error is raised here. PyTorch tensors are only containers for device memory; every call is enqueued on
torch's current CUDA stream.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# D3F_LIB: alternative build of the same library (kernel-tuning experiments); never a different backend
LIB_PATH = os.environ.get("D3F_LIB") or os.path.join(_HERE, "libd3feat_b200.so")

# every symbol include/d3feat_b200.h declares: (name, restype, argtypes)
_P, _I, _F, _Z, _LL = C.c_void_p, C.c_int, C.c_float, C.c_size_t, C.c_longlong
SYMBOLS = [
    ("d3f_version", _I, []),
    ("d3f_last_error", C.c_char_p, []),
    ("d3f_launch_count", _LL, []),
    ("d3f_bbox", _I, [_P, _I, _P, _P]),
    ("d3f_grid_subsample_workspace_bytes", _Z, [_I, _I]),
    ("d3f_grid_subsample", _I, [_P, _P, _I, _I, _F, _P, _I, _P, _I, _P, _P, _P, _P, _P, _P, _P, _Z, _P]),
    ("d3f_radius_neighbors_workspace_bytes", _Z, [_I, _I, _F, _P]),
    ("d3f_radius_neighbors_build", _I, [_P, _P, _I, _I, _F, _P, _P, _Z, _P]),
    ("d3f_radius_neighbors_count", _I, [_P, _P, _I, _P, _P, _I, _I, _F, _P, _P, _P, _P, _P]),
    ("d3f_radius_neighbors_fill", _I, [_P, _P, _I, _P, _P, _I, _I, _F, _P, _P, _I, _I, _P, _P]),
    ("d3f_kpconv_workspace_bytes", _Z, [_I, _I, _I, _I, _I, _I]),
    ("d3f_pyramid_workspace_bytes", _Z, [_I, _P, _P, _P]),
    ("d3f_pyramid_build", _I, [_P, _P, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _Z, _P, _P, _P, _P]),
    ("d3f_packed_weight_floats", _Z, [_I, _I]),
    ("d3f_pack_weight", _I, [_P, _I, _I, _P, _P]),
    ("d3f_radius_neighbors_order", _I, [_P, _I, _I, _F, _P, _P, _P]),
    ("d3f_kpconv_forward", _I, [_P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _F, _I, _I, _I, _P, _P, _P, _F, _P,
                                _P, _Z, _P, _P, _P]),
    ("d3f_kpconv_deform_forward", _I, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _F, _I, _I, _P, _P,
                                       _P, _F, _P, _P, _Z, _P, _P, _P]),
    ("d3f_unary_forward", _I, [_P, _P, _P, _I, _I, _I, _P, _P, _P, _P, _F, _P, _P, _P]),
    ("d3f_ind_max_pool_workspace_bytes", _Z, [_I]),
    ("d3f_ind_max_pool", _I, [_P, _P, _I, _I, _I, _I, _P, _P, _Z, _P, _P, _P]),
    ("d3f_closest_pool", _I, [_P, _P, _I, _I, _I, _I, _P, _P, _P, _P]),
    ("d3f_l2_normalize", _I, [_P, _I, _I, _F, _P, _P, _P]),
    ("d3f_unary_pair_forward", _I, [_P, _I, _P, _I, _P, _I, _I, _P, C.c_float, _P, _P, _P]),
    ("d3f_detection_scores_workspace_bytes", _Z, [_I, _I]),
    ("d3f_detection_scores", _I, [_P, _P, _P, _I, _I, _I, _I, _P, _P, _Z, _P, _P]),
    ("d3f_affine_leaky", _I, [_P, _I, _I, _P, _P, _P, _F, _P, _P, _P]),
]

_lib = None


class D3FError(RuntimeError):
    pass


def lib():
    """Load the library (once). Raises if it has not been built: there is no fallback path."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise D3FError(
                "libd3feat_b200.so is missing at %s -- build it with `python -c 'import __graft_entry__ as g; "
                "g.build()'` (nvcc, sm_100a). d3feat_b200 has no CPU or PyTorch fallback." % LIB_PATH)
        l = C.CDLL(LIB_PATH)
        for name, res, args in SYMBOLS:
            fn = getattr(l, name)  # AttributeError if the .so does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(rc, what):
    if rc != 0:
        msg = lib().d3f_last_error().decode("utf-8", "replace")
        if rc == -1:
            raise ValueError("%s: %s" % (what, msg))
        raise D3FError("%s failed (%d): %s" % (what, rc, msg))


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    """Device (or host) address of a contiguous tensor; None -> NULL."""
    if t is None:
        return None
    assert t.is_contiguous(), "tensor must be contiguous"
    return C.c_void_p(t.data_ptr())


def workspace(nbytes, device):
    return torch.empty((max(int(nbytes), 256),), dtype=torch.uint8, device=device)


def f32(t, device):
    if not torch.is_tensor(t):
        t = torch.as_tensor(t, dtype=torch.float32)
    return t.to(device=device, dtype=torch.float32).contiguous()


def i32(t, device):
    if not torch.is_tensor(t):
        t = torch.as_tensor(t, dtype=torch.int32)
    return t.to(device=device, dtype=torch.int32).contiguous()


def launch_count():
    return int(lib().d3f_launch_count())
