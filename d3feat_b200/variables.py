"""Minimal stand-in for TF1 variable scopes: the reference blocks create their weights with
`weight_variable(shape)` / `tf.Variable(..., name='kernel_points')` under nested `tf.variable_scope`s and, at
test time, a Saver restores them by name (utils/tester.py:143-162). Here the restored checkpoint is a dict
{scoped name: array}; the mirrored blocks look their parameters up under the same names."""
import contextlib

import numpy as np
import torch

_scope = []
_store = None


class ParamStore:
    """Device-resident parameters keyed by the reference's variable names (e.g.
    'layer_1/resnetb_0/conv2/weights', '.../kernel_points', '.../batch_normalization/gamma')."""

    def __init__(self, params, device):
        self.device = torch.device(device)
        self.t = {}
        for k, v in params.items():
            a = np.ascontiguousarray(np.asarray(v, dtype=np.float32))
            self.t[k] = torch.from_numpy(a).to(self.device)
        self._bn = {}

    def __contains__(self, name):
        return name in self.t

    def __len__(self):
        return len(self.t)

    def get(self, name):
        try:
            return self.t[name]
        except KeyError:
            raise KeyError("variable '%s' is not in the parameter store" % name)

    def bn_affine(self, scope, eps=1e-6):
        """Inference batch norm folded to y = x*scale + shift (models/network_blocks.py:149-160 with moving
        statistics): scale = gamma / sqrt(var + eps), shift = beta - mean * scale. Folded in float64."""
        if scope not in self._bn:
            pre = scope + "/batch_normalization/"
            g, b, m, v = (self.t[pre + k].double() for k in ("gamma", "beta", "moving_mean", "moving_variance"))
            scale = g / torch.sqrt(v + eps)
            shift = b - m * scale
            self._bn[scope] = (scale.float().contiguous(), shift.float().contiguous())
        return self._bn[scope]


@contextlib.contextmanager
def use_params(store):
    global _store
    prev, _store = _store, store
    try:
        yield store
    finally:
        _store = prev


@contextlib.contextmanager
def variable_scope(name):
    _scope.append(name)
    try:
        yield
    finally:
        _scope.pop()


def current_scope():
    return "/".join(_scope)


def current_store():
    return _store


def scoped(name):
    s = current_scope()
    return s + "/" + name if s else name
