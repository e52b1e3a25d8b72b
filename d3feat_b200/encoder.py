"""User-facing driver of the hot path: stacked point-cloud fragments in, per-level encoder features (and,
optionally, the 32-d descriptors of the D3Feat decoder) out.

Mirrors how the reference is driven at test time (utils/tester.py:177-233: one sess.run per batch that
executes the tf.data pyramid on the CPU and the network on the device); here both halves run on the GPU:

    enc = KPFCNN(config, params, neighborhood_limits)
    out = enc(points_host_or_cuda, lengths)           # pyramid + encoder (+ decoder)
"""
import numpy as np
import torch

from . import network_blocks as nb
from . import pyramid
from .variables import ParamStore, use_params


class KPFCNN:
    def __init__(self, config, params, neighborhood_limits, device="cuda"):
        self.config = config
        self.device = torch.device(device)
        self.store = params if isinstance(params, ParamStore) else ParamStore(params, self.device)
        self.limits = [int(x) for x in neighborhood_limits]
        self.has_decoder = any("upsample" in b for b in config.architecture)

    def build_inputs(self, stacked_points, stacked_lengths, features=None, bbox=None):
        pts = stacked_points
        if not torch.is_tensor(pts):
            pts = torch.as_tensor(np.ascontiguousarray(pts, np.float32))
        pts = pts.to(self.device, non_blocking=True)
        lens = stacked_lengths
        if not torch.is_tensor(lens):
            lens = torch.as_tensor(np.ascontiguousarray(lens, np.int32))
        lens = lens.to(self.device, non_blocking=True)
        inputs = pyramid.descriptor_input(self.config, pts, lens, self.limits, bbox=bbox)
        if features is None:
            # the 3DMatch generator feeds a constant-one feature (datasets/ThreeDMatch.py:316)
            features = torch.ones((pts.shape[0], self.config.in_features_dim), dtype=torch.float32, device=self.device)
        elif not torch.is_tensor(features):
            features = torch.as_tensor(np.ascontiguousarray(features, np.float32)).to(self.device)
        inputs["features"] = features
        return inputs

    def encode(self, inputs):
        """assemble_CNN_blocks on prepared inputs -> list F of per-level features."""
        with use_params(self.store):
            return nb.assemble_CNN_blocks(inputs, self.config, 1.0)

    def describe(self, inputs, F):
        with use_params(self.store):
            return nb.assemble_FCNN_decoder(inputs, self.config, F, 1.0)

    def __call__(self, stacked_points, stacked_lengths, features=None, bbox=None, decoder=None):
        inputs = self.build_inputs(stacked_points, stacked_lengths, features, bbox)
        F = self.encode(inputs)
        use_dec = self.has_decoder if decoder is None else decoder
        desc = self.describe(inputs, F) if use_dec else None
        return dict(inputs=inputs, F=F, descriptors=desc)
