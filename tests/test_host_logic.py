"""CPU: host-side logic (workload generator, parameter naming, pyramid radii, oracle pyramid)."""
import numpy as np
import pytest

from d3feat_b200 import synth
from d3feat_b200 import pyramid
from oracle import native as on
from oracle import kpconv_np as ok


def test_room_fragment_is_seeded_and_exact_size():
    a = synth.room_fragment(5, 4000)
    b = synth.room_fragment(5, 4000)
    assert a.shape == (4000, 3) and a.dtype == np.float32
    assert np.array_equal(a, b)
    assert not np.array_equal(a, synth.room_fragment(6, 4000))
    # voxelised at 0.03: no two points share a cell
    cells = np.floor(a / 0.03).astype(np.int64)
    assert np.unique(cells, axis=0).shape[0] == 4000


def test_level_radii_follow_tf_descriptor_input():
    cfg = synth.Config()
    lv = pyramid._level_radii(cfg)
    assert len(lv) == 5
    for l, d in enumerate(lv):
        assert d["conv_r"] == pytest.approx(0.075 * 2 ** l)
        if l < 4:
            assert d["dl"] == pytest.approx(0.06 * 2 ** l)
            assert d["pool_r"] == pytest.approx(0.075 * 2 ** l)
            assert d["up_r"] == pytest.approx(0.15 * 2 ** l)
        else:
            assert d["dl"] is None
    cfgd = synth.Config(architecture=synth.ARCH_KITTI_DEFORM, first_subsampling_dl=0.3)
    lvd = pyramid._level_radii(cfgd)
    # layer 3 ends with a deformable strided block: pool radius is dl*density (common.py:1361-1364)
    assert lvd[3]["pool_r"] == pytest.approx(0.75 * 8 * 2)
    # layer 4 holds a single deformable block: layer_blocks[:-1] is empty, so the radius is NOT doubled
    # (quirk of common.py:1340 kept as is)
    assert lvd[4]["conv_r"] == pytest.approx(0.75 * 16)


def test_param_names_and_shapes_match_reference_scopes():
    cfg = synth.Config()
    p = synth.make_params(cfg, 0)
    assert p["layer_0/simple_0/weights"].shape == (15, 1, 64)
    assert p["layer_0/resnetb_1/conv2/weights"].shape == (15, 32, 32)
    assert p["layer_0/resnetb_1/shortcut/weights"].shape == (64, 128)
    assert "layer_0/resnetb_strided_2/shortcut/weights" not in p        # dims already match (:604)
    assert p["layer_1/resnetb_0/conv2/weights"].shape == (15, 64, 64)   # cf. results_kitti layer_1_resnetb_0_conv2.npy
    assert p["layer_4/resnetb_0/conv2/weights"].shape == (15, 512, 512)
    assert p["layer_4/resnetb_0/conv3/weights"].shape == (512, 2048)
    assert p["uplayer_3/unary_0/weights"].shape == (3072, 512)
    assert p["uplayer_0/last_unary_1/weights"].shape == (64, 32)
    assert p["layer_2/resnetb_0/conv2/kernel_points"].shape == (15, 3)
    w = p["layer_1/resnetb_0/conv1/weights"]
    assert np.allclose(w, np.round(w * 1000) / 1000) and np.abs(w).max() <= 2 * np.sqrt(2 / w.shape[-1]) + 1e-3
    kp = p["layer_0/simple_0/kernel_points"]
    assert np.linalg.norm(kp[0]) < 0.01 and np.allclose(np.linalg.norm(kp[1:], axis=1), 1.5 * 0.03, rtol=0.1)


def test_oracle_pyramid_and_encoder_shapes():
    cfg = synth.Config(architecture=synth.ARCH_ENCODER[:6], first_features_dim=16)
    pts = np.concatenate([synth.room_fragment(0, 1500), synth.room_fragment(1, 1200)], 0)
    lens = np.array([1500, 1200], np.int32)
    inputs = ok.descriptor_input_pyramid(cfg, pts, lens, [25, 25, 25], on.port_batch_neighbors,
                                         on.port_batch_subsampling)
    assert len(inputs["points"]) == 3
    assert inputs["neighbors"][0].shape == (2700, 25)
    assert inputs["pools"][0].shape[0] == inputs["points"][1].shape[0]
    assert inputs["upsamples"][0].shape[0] == 2700
    assert inputs["pools"][2].shape == (0, 1)
    # per-cloud isolation: no neighbour index crosses the cloud boundary
    nb = inputs["neighbors"][0]
    assert nb[:1500][nb[:1500] < 2700].max() < 1500 and nb[1500:].min() >= 1500
    inputs["features"] = np.ones((2700, 1), np.float32)
    params = synth.make_params(cfg, 1)
    F = ok.EncoderOracle(cfg, params, np.float32).encoder(inputs)
    assert [f.shape[1] for f in F] == [32, 64, 128]
    assert F[2].shape[0] == inputs["points"][2].shape[0]
    assert all(np.isfinite(f).all() for f in F)


def test_detection_score_restatement_against_scalar_loop():
    """oracle/kpconv_np.detection_scores (vectorised) vs a literal per-point evaluation of models/D3Feat.py:67-115
    for the reference's own batch shape (two clouds)."""
    from oracle import kpconv_np as ok
    rng = np.random.default_rng(0)
    lengths = [40, 30]
    N, D, H = 70, 8, 6
    x = rng.normal(size=(N, D))
    x[5] = 0.0
    nb = np.concatenate([rng.integers(0, 40, (40, H)), rng.integers(40, 70, (30, H))]).astype(np.int64)
    nb[rng.uniform(size=nb.shape) < 0.3] = N
    got = ok.detection_scores(x, nb, lengths)
    scale = np.concatenate([np.full(40, x[:40].max() + 1e-6), np.full(30, x[40:].max() + 1e-6)])
    xs = np.concatenate([x / scale[:, None], np.zeros((1, D))])
    for i in range(N):
        rows = xs[nb[i]]
        num = max(int(np.count_nonzero(rows.sum(axis=1))), 1)
        local = np.log(1.0 + np.exp(xs[i] - rows.sum(axis=0) / num))
        depth = xs[i] / (1e-6 + xs[i].max())
        assert abs(got[i, 0] - (local * depth).max()) < 1e-12


def test_pyramid_buffers_and_spec_on_cpu():
    """PyramidBuffers (the pre-allocated ring slots of BatchPipeline) and the level radii for the KITTI session
    (results_kitti/Log_11011605/parameters.txt:30: first_subsampling_dl = 0.3 -> conv radius 0.75, doubling per level)."""
    import torch
    from d3feat_b200 import pyramid as pyr, synth
    cfg = synth.Config(first_subsampling_dl=0.3)
    spec, levels = pyr.make_spec(cfg, [40, 41, 42, 43, 44])
    assert spec.n_levels == 5
    assert [round(float(spec.conv_radius[l]), 4) for l in range(5)] == [0.75, 1.5, 3.0, 6.0, 12.0]
    assert [round(float(spec.sub_dl[l]), 4) for l in range(4)] == [0.6, 1.2, 2.4, 4.8] and spec.sub_dl[4] < 0
    assert [round(float(spec.up_radius[l]), 4) for l in range(4)] == [1.5, 3.0, 6.0, 12.0]
    assert [int(spec.limit[l]) for l in range(5)] == [40, 41, 42, 43, 44]
    buf = pyr.PyramidBuffers(cfg, [40, 41, 42, 43, 44], capacity=1000, n_clouds=2, device=torch.device("cpu"))
    assert buf.pts[0] is None and buf.pts[1].shape == (1000, 3) and buf.len[3].shape == (2,)
    assert [b.shape[1] for b in buf.nb] == [40, 41, 42, 43, 44]
    assert buf.pool[4] is None and buf.up[4] is None and buf.pool[0].shape == (1000, 40)
    assert buf.fits(1000, 2, [40, 41, 42, 43, 44]) and not buf.fits(1001, 2, [40, 41, 42, 43, 44])
    assert not buf.fits(10, 3, [40, 41, 42, 43, 44]) and not buf.fits(10, 2, [40] * 5)
    w1 = buf.workspace(1000)
    assert buf.workspace(900) is w1 and buf.workspace(5000).numel() >= 5000
    # deformable blocks: the reference widens the CONV radius only when a deformable block precedes the last block of
    # the level (`layer_blocks[:-1]`, datasets/common.py:1340); in this architecture that never happens, so only the
    # pooling / upsampling radii of the deformable strided block grow by density_parameter / (KP_extent * 2.5)
    cfgd = synth.Config(architecture=synth.ARCH_KITTI_DEFORM, first_subsampling_dl=0.3)
    specd, _ = pyr.make_spec(cfgd, [40] * 5)
    assert [round(float(specd.conv_radius[l]), 3) for l in range(5)] == [0.75, 1.5, 3.0, 6.0, 12.0]
    assert [round(float(specd.pool_radius[l]), 3) for l in range(4)] == [0.75, 1.5, 3.0, 12.0]
    assert round(float(specd.up_radius[3]), 3) == 24.0


def test_shape_buckets_and_per_level_capacities_on_cpu():
    """Host side of the sync-free form (encoder.GraphPipeline): capacities of a shape bucket and the per-level buffer
    shapes of a slot. A level's pool matrix has the rows of the NEXT level and the columns of this level's cap; the
    upsample matrix the rows of this level (datasets/common.py:1358-1372)."""
    import torch
    from d3feat_b200 import pyramid as pyr, synth
    sizes = [240000, 60336, 15430, 4177, 1204]
    caps = pyr.bucket_capacities(sizes)
    assert caps == [270080, 68096, 17664, 4864, 1536]
    assert all(c % 256 == 0 and c >= 1.125 * n for c, n in zip(caps, sizes))
    assert pyr.bucket_capacities([0, 1]) == [256, 256]                 # empty levels still get a launchable buffer
    assert pyr.bucket_capacities(sizes, slack=1.0, quantum=128) == [240128, 60416, 15616, 4352, 1280]
    cfg = synth.Config(architecture=synth.ARCH_ENCODER)
    limits = [40, 39, 38, 37, 36]
    small = [1000, 300, 90, 30, 10]
    buf = pyr.PyramidBuffers(cfg, limits, capacity=small, n_clouds=3, device=torch.device("cpu"),
                             bbox=np.array([0, 0, 0, 1, 1, 1], np.float32))
    assert buf.caps == small and buf.capacity == 1000
    assert [tuple(b.shape) for b in buf.nb] == [(1000, 40), (300, 39), (90, 38), (30, 37), (10, 36)]
    assert [tuple(buf.pool[l].shape) for l in range(4)] == [(300, 40), (90, 39), (30, 38), (10, 37)]
    assert [tuple(buf.up[l].shape) for l in range(4)] == [(1000, 40), (300, 39), (90, 38), (30, 37)]
    assert [tuple(buf.pts[l].shape) for l in range(1, 5)] == [(300, 3), (90, 3), (30, 3), (10, 3)]
    assert tuple(buf.points0.shape) == (1000, 3) and tuple(buf.lengths0.shape) == (3,)
    assert buf.features0.shape == (1000, cfg.in_features_dim) and float(buf.features0.min()) == 1.0
    assert buf.counts.dtype == torch.int32 and int(buf.counts.abs().sum()) == 0 and int(buf.status[0]) == 0
    assert buf.bbox.dtype == np.float32 and buf.bbox.shape == (6,)
    assert buf.fits(1000, 3, limits) and not buf.fits(1000, 2, limits)
    with pytest.raises(AssertionError):
        pyr.PyramidBuffers(cfg, limits, capacity=[10, 5], n_clouds=1, device=torch.device("cpu"))
