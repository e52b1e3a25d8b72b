"""Mirror of the CPython extension cpp_wrappers/cpp_subsampling (module `grid_subsampling`).

Reference: grid_subsampling.compute(points, features=None, classes=None, sampleDl=0.1,
method='barycenters', verbose=0)  -- wrapper.cpp:58-286, format "O|$OOfsi" (:70-76).
Same keyword-only arguments, the same validation messages (RuntimeError like the extension), the same
return convention (points, or a tuple (points[, features][, classes])). numpy in -> numpy out, CUDA torch
tensors in -> CUDA torch tensors out; in both cases the work runs on the GPU (no CPU path).

Cells are returned in ascending cell key (the reference returns std::unordered_map iteration order).
`classes` follow the reference literally: the largest label present in each cell.
"""
import numpy as np
import torch

from . import tf_custom_ops as _ops


def compute(points, *, features=None, classes=None, sampleDl=0.1, method="barycenters", verbose=0):
    if method not in ("barycenters", "voxelcenters"):      # wrapper.cpp:86-90 (validated, then ignored)
        raise RuntimeError('Error parsing method. Valid method names are "barycenters" and "voxelcenters" ')
    as_numpy = not torch.is_tensor(points)
    dev = torch.device("cuda", torch.cuda.current_device()) if as_numpy else points.device
    try:
        pts = torch.as_tensor(np.ascontiguousarray(points, dtype=np.float32)) if as_numpy else points
        pts = pts.to(device=dev, dtype=torch.float32)
    except (TypeError, ValueError):
        raise RuntimeError("Error converting input points to numpy arrays of type float32")
    if pts.dim() != 2 or pts.shape[1] != 3:                # wrapper.cpp:135-142
        raise RuntimeError("Wrong dimensions : points.shape is not (N, 3)")
    N = pts.shape[0]
    f = c = None
    if features is not None:
        try:
            f = torch.as_tensor(np.ascontiguousarray(features, dtype=np.float32)) if not torch.is_tensor(features) else features
            f = f.to(device=dev, dtype=torch.float32)
        except (TypeError, ValueError):
            raise RuntimeError("Error converting input features to numpy arrays of type float32")
        if f.dim() != 2 or f.shape[0] != N:                # wrapper.cpp:143-150, 186-193
            raise RuntimeError("Wrong dimensions : features.shape is not (N, d)")
    if classes is not None:
        try:
            c = torch.as_tensor(np.ascontiguousarray(classes, dtype=np.int32)) if not torch.is_tensor(classes) else classes
            c = c.to(device=dev, dtype=torch.int32)
        except (TypeError, ValueError):
            raise RuntimeError("Error converting input classes to numpy arrays of type int32")
        if c.dim() > 2 or c.shape[0] != N:                 # wrapper.cpp:152-159, 194-201
            raise RuntimeError("Wrong dimensions : classes.shape is not (N,) or (N, d)")
    if verbose > 0:
        print("Computing cloud pyramid with support points: ")
    nb = torch.tensor([N], dtype=torch.int32, device=dev)
    res = _ops._subsample(pts.contiguous(), nb, float(sampleDl), features=f, classes=c)
    if res[0].shape[0] < 1:                                # wrapper.cpp:225-229
        raise RuntimeError("Error")
    out = [res[0]] + res[2:]
    if as_numpy:
        out = [o.cpu().numpy() for o in out]
    return out[0] if len(out) == 1 else tuple(out)
