"""Host-side formats either side of the hot path: the session's parameters.txt, PLY point
clouds, and the per-fragment arrays the reference's offline evaluation consumes.

  * load_config      -- utils/config.py:Config.load (parameters.txt -> attributes; only the keys the path reads)
  * read_ply_points  -- utils/ply.py:read_ply, vertex x/y/z of ascii / binary PLY files
  * select_keypoints -- utils/tester.py:209-213 (3DMatch: all points, ascending score) and :283-290 (KITTI: top-k)
  * write_fragment   -- utils/tester.py:226-228: descriptors/<scene>/cloud_bin_N.D3Feat.npy,
                        keypoints/<scene>/cloud_bin_N.npy, scores/<scene>/cloud_bin_N.npy, the layout
                        geometric_registration/evaluate.py:39-50 reads back (get_keypts / get_desc / get_scores)
"""
import os

import numpy as np

from .synth import Config

_INT_KEYS = ("num_layers", "first_features_dim", "in_features_dim", "in_points_dim", "num_kernel_points", "num_classes",
             "use_batch_norm", "modulated", "input_threads", "batch_num")
_FLOAT_KEYS = ("first_subsampling_dl", "density_parameter", "KP_extent", "batch_norm_momentum", "in_radius")
_STR_KEYS = ("dataset", "fixed_kernel_points", "KP_influence", "convolution_mode")


def load_config(path):
    """path: a results/Log_*/ directory or its parameters.txt."""
    if os.path.isdir(path):
        path = os.path.join(path, "parameters.txt")
    kw = {}
    with open(path) as fh:
        for line in fh:
            line = line.strip()
            if not line or line.startswith("#") or " = " not in line:
                continue
            key, val = line.split(" = ", 1)
            if key == "architecture":
                kw[key] = val.split()
            elif key in _INT_KEYS:
                kw[key] = int(val)
            elif key in _FLOAT_KEYS:
                kw[key] = float(val)
            elif key in _STR_KEYS:
                kw[key] = val
    if "architecture" not in kw:
        raise ValueError("%s: no architecture line" % path)
    for k in ("use_batch_norm", "modulated"):
        if k in kw:
            kw[k] = bool(kw[k])
    return Config(**kw)


_PLY_TYPES = {"char": "i1", "int8": "i1", "uchar": "u1", "uint8": "u1", "short": "i2", "int16": "i2", "ushort": "u2",
              "uint16": "u2", "int": "i4", "int32": "i4", "uint": "u4", "uint32": "u4", "float": "f4",
              "float32": "f4", "double": "f8", "float64": "f8"}


def read_ply_points(path):
    """float32 [N,3] vertex positions of a PLY file (ascii, binary little- or big-endian)."""
    with open(path, "rb") as fh:
        if fh.readline().strip() != b"ply":
            raise ValueError("%s: not a PLY file" % path)
        fmt = None
        n_vertex = None
        props = []
        in_vertex = False
        while True:
            line = fh.readline()
            if not line:
                raise ValueError("%s: truncated PLY header" % path)
            tok = line.decode("ascii", "replace").split()
            if not tok:
                continue
            if tok[0] == "format":
                fmt = tok[1]
            elif tok[0] == "element":
                in_vertex = tok[1] == "vertex"
                if in_vertex:
                    n_vertex = int(tok[2])
            elif tok[0] == "property" and in_vertex:
                if tok[1] == "list":
                    raise ValueError("%s: list property on vertices" % path)
                props.append((tok[2], _PLY_TYPES[tok[1]]))
            elif tok[0] == "end_header":
                break
        if n_vertex is None or fmt is None:
            raise ValueError("%s: PLY header without format / vertex element" % path)
        if fmt == "ascii":
            data = np.loadtxt(fh, max_rows=n_vertex, ndmin=2)
            cols = [p[0] for p in props]
            return np.ascontiguousarray(data[:, [cols.index(a) for a in "xyz"]], np.float32)
        order = "<" if fmt == "binary_little_endian" else ">"
        rec = np.fromfile(fh, dtype=[(n, order + t) for n, t in props], count=n_vertex)
    return np.ascontiguousarray(np.stack([rec["x"], rec["y"], rec["z"]], 1), np.float32)


def select_keypoints(scores, num_keypts=None):
    """Indices of the detected keypoints of ONE cloud. scores: [N,1] or [N].
    num_keypts=None: every point in ascending score order (what the 3DMatch tester dumps; the evaluation takes the
    last 250 rows, evaluate.py:47-50). num_keypts=k: the k highest scores, ascending (the KITTI tester)."""
    s = np.asarray(scores).reshape(-1, 1)
    order = np.argsort(s, axis=0).reshape(-1)          # same call as the reference: identical tie order
    return order if num_keypts is None else order[-num_keypts:]


def write_fragment(root, scene, num_frag, points, descriptors, scores, num_keypts=None):
    """Writes the three arrays of one fragment, rows sorted by detection score. Returns the paths."""
    ids = select_keypoints(scores, num_keypts)
    paths = []
    for sub, name, arr in (("descriptors", "cloud_bin_{}.D3Feat".format(num_frag), np.asarray(descriptors)[ids]),
                           ("keypoints", "cloud_bin_{}".format(num_frag), np.asarray(points)[ids]),
                           ("scores", "cloud_bin_{}".format(num_frag), np.asarray(scores)[ids])):
        d = os.path.join(root, sub, scene)
        os.makedirs(d, exist_ok=True)
        p = os.path.join(d, name + ".npy")
        np.save(p, arr.astype(np.float32))
        paths.append(p)
    return paths
