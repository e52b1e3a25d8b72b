"""Input pyramid of the KPFCNN encoder -- mirror of Dataset.tf_descriptor_input (datasets/common.py:1301-1413)
with big_neighborhood_filter (:399-406) and calibrate_neighbors (:572-673), on the GPU.

The reference runs this loop on the CPU inside tf.data (13 radius searches + 4 grid subsamplings per batch).
Here every level builds ONE hash grid over its points and reuses it for the three searches that share those
supports and that radius (conv_l, pool_l, and up_{l-1}), i.e. 5 grid builds instead of 13; the neighbour
matrices are produced directly at the calibrated width (the reference computes the full width and slices),
so the only device->host reads are the number of cells after each subsampling and one bbox up front.
"""
import os

import numpy as np
import torch

from . import _lib
from . import tf_custom_ops as ops


def _level_radii(config):
    """(conv radius, subsample dl, pool radius, upsample radius, has_pool) per level, exactly as the loop of
    tf_descriptor_input derives them (:1312-1396)."""
    r_normal = config.first_subsampling_dl * config.KP_extent * 2.5
    arch = list(config.architecture)
    layer_blocks = []
    levels = []
    for block_i, block in enumerate(arch):
        if "global" in block or "upsample" in block:
            break
        if not ("pool" in block or "strided" in block):
            layer_blocks.append(block)
            if block_i < len(arch) - 1 and "upsample" not in arch[block_i + 1]:
                continue
        lv = {}
        if layer_blocks:
            if np.any(["deformable" in b for b in layer_blocks[:-1]]):
                lv["conv_r"] = r_normal * config.density_parameter / (config.KP_extent * 2.5)
            else:
                lv["conv_r"] = r_normal
        else:
            lv["conv_r"] = None
        if "pool" in block or "strided" in block:
            lv["dl"] = 2 * r_normal / (config.KP_extent * 2.5)
            if "deformable" in block:
                lv["pool_r"] = r_normal * config.density_parameter / (config.KP_extent * 2.5)
            else:
                lv["pool_r"] = r_normal
            lv["up_r"] = 2 * lv["pool_r"]
        else:
            lv["dl"] = None
        levels.append(lv)
        r_normal *= 2
        layer_blocks = []
    return levels


def descriptor_input(config, stacked_points, stacked_lengths, neighborhood_limits, bbox=None):
    """Returns the dict the blocks consume: points[L], neighbors[L], pools[L], upsamples[L], lengths[L]
    (placeholders of the reference's shapes at the last level, :1374-1377).

    neighborhood_limits: per-level column caps (Dataset.neighborhood_limits). Neighbour matrices are emitted
    with exactly `limit` columns, padded with the shadow index; when the true maximum count is below the
    limit the extra columns are all-shadow and do not change any downstream result.
    """
    dev = stacked_points.device
    pts = _lib.f32(stacked_points, dev)
    lens = _lib.i32(stacked_lengths, dev)
    if bbox is None:
        bbox = ops.host_bbox(pts)
    levels = _level_radii(config)
    out = dict(points=[], neighbors=[], pools=[], upsamples=[], lengths=[], orders=[])
    grids = {}

    def grid_for(level, supports, sb, radius):
        key = (level, float(np.float32(radius)))
        if key not in grids:
            grids[key] = ops.NeighborGrid(supports, sb, radius, bbox)
        return grids[key]

    level_pts = [pts]
    level_len = [lens]
    for li, lv in enumerate(levels):
        p, b = level_pts[li], level_len[li]
        lim = int(neighborhood_limits[li])
        if lv["conv_r"] is not None:
            g = grid_for(li, p, b, lv["conv_r"])
            conv_i = g.fill(p, b, lim, p.shape[0])
            # level 0 arrives in the caller's (arbitrary) order: hand the gather kernels the grid's cell order as
            # visiting order. Deeper levels are already emitted in cell order by the subsampling.
            # (measured on B200: no gain -- the gathers are L2-latency bound, not L1-locality bound -- so the hint
            # is off unless D3F_QUERY_ORDER=1; profiles/r1_notes.md)
            use_order = li == 0 and os.environ.get("D3F_QUERY_ORDER", "0") == "1"
            order = g.order() if use_order else torch.zeros((0,), dtype=torch.int32, device=dev)
        else:
            conv_i = torch.zeros((0, 1), dtype=torch.int32, device=dev)
            order = torch.zeros((0,), dtype=torch.int32, device=dev)
        if lv["dl"] is not None:
            pool_p, pool_b = ops.batch_grid_subsampling(p, b, lv["dl"], bbox=bbox)
            pool_i = grid_for(li, p, b, lv["pool_r"]).fill(pool_p, pool_b, lim, p.shape[0])
            up_i = grid_for(li + 1, pool_p, pool_b, lv["up_r"]).fill(p, b, lim, pool_p.shape[0])
            level_pts.append(pool_p)
            level_len.append(pool_b)
        else:
            pool_i = torch.zeros((0, 1), dtype=torch.int32, device=dev)
            up_i = torch.zeros((0, 1), dtype=torch.int32, device=dev)
        out["points"].append(p)
        out["neighbors"].append(conv_i)
        out["pools"].append(pool_i)
        out["upsamples"].append(up_i)
        out["lengths"].append(b)
        out["orders"].append(order)
    return out


def flat_inputs(inputs, stacked_features):
    """The positional list models/KPFCNN_model.py:86-121 unpacks: points + neighbors + pools + upsamples +
    [features] (the batch-weight / batch-index tensors are training-only)."""
    return inputs["points"] + inputs["neighbors"] + inputs["pools"] + inputs["upsamples"] + [stacked_features]


def calibrate_neighbors(config, clouds, keep_ratio=0.8, device="cuda"):
    """Column caps per level: the smallest count c such that at least keep_ratio of the neighbourhoods have
    <= c neighbours -- datasets/common.py:572-673 (histogram of conv-neighbour counts, cumulative sum,
    percentile). `clouds` is an iterable of float32[N,3] arrays (one cloud each)."""
    levels = _level_radii(config)
    hist_n = int(np.ceil(4 / 3 * np.pi * (config.density_parameter + 1) ** 3))
    hists = np.zeros((len(levels), hist_n), np.int64)
    for cloud in clouds:
        p = torch.as_tensor(np.ascontiguousarray(cloud, np.float32)).to(device)
        b = torch.tensor([p.shape[0]], dtype=torch.int32, device=device)
        bbox = ops.host_bbox(p)
        for li, lv in enumerate(levels):
            if lv["conv_r"] is not None:
                g = ops.NeighborGrid(p, b, lv["conv_r"], bbox)
                counts, _ = g.count(p, b)
                c = np.bincount(counts.cpu().numpy(), minlength=hist_n)[:hist_n]   # counts >= hist_n are dropped (:640)
                hists[li] += c
            if lv["dl"] is None:
                break
            p, b = ops.batch_grid_subsampling(p, b, lv["dl"], bbox=bbox)
    cumsum = np.cumsum(hists.T, axis=0)
    percentiles = np.sum(cumsum < (keep_ratio * cumsum[hist_n - 1, :]), axis=0)
    return [int(x) for x in percentiles]
