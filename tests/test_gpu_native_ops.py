"""GPU parity: hash-grid radius neighbours and grid subsampling vs the golden vectors / the C oracle.
Bar: bit-exact indices and barycenters (after the canonical ordering described in DESIGN.md)."""
import numpy as np
import pytest
import torch

from oracle import native as on

pytestmark = pytest.mark.gpu


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def t(a, dev, dtype=None):
    x = torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    return x.to(dtype) if dtype is not None else x


def test_subsampling_matches_golden_demo(cuda, golden):
    from d3feat_b200 import tf_custom_ops as ops
    g = golden("subsampling_demo.npz")
    p, b = ops.batch_grid_subsampling(t(g["points"], cuda), t(g["lengths"], cuda), float(g["dl"]))
    p, b = p.cpu().numpy(), b.cpu().numpy()
    assert np.array_equal(b, g["sub_lengths"])
    # canonical order == the oracle's order (ascending cell key): bit-exact, row for row
    op, ob = on.port_batch_subsampling(g["points"], g["lengths"], float(g["dl"]))
    assert np.array_equal(bits(p), bits(op))
    # and the same point set as the reference (reference order = unordered_map iteration order)
    o = 0
    for n in b:
        a_, _ = on.sort_rows(p[o:o + n])
        b_, _ = on.sort_rows(g["sub_points"][o:o + n])
        assert np.array_equal(bits(a_), bits(b_))
        o += n


def test_cpp_subsampling_compute_features_classes(cuda, golden):
    from d3feat_b200 import cpp_subsampling
    g = golden("subsampling_demo.npz")
    p, f, c = cpp_subsampling.compute(g["w_points"], features=g["w_features"], classes=g["w_classes"],
                                      sampleDl=float(g["w_dl"]), verbose=0)
    assert isinstance(p, np.ndarray) and p.dtype == np.float32 and c.dtype == np.int32
    op, of, oc = on.port_grid_subsample(g["w_points"], g["w_features"], g["w_classes"], sampleDl=float(g["w_dl"]))
    assert np.array_equal(bits(p), bits(op)) and np.array_equal(bits(f), bits(of)) and np.array_equal(c, oc)
    ours = np.concatenate([bits(p), bits(f), c.astype(np.uint32)], 1)
    ref = np.concatenate([bits(g["w_sub_points"]), bits(g["w_sub_features"]), g["w_sub_classes"].astype(np.uint32)], 1)
    assert np.array_equal(on.sort_rows(ours)[0], on.sort_rows(ref)[0])
    # return convention and validation of wrapper.cpp
    only = cpp_subsampling.compute(g["w_points"], sampleDl=0.1)
    assert isinstance(only, np.ndarray) and only.shape[1] == 3
    with pytest.raises(RuntimeError):
        cpp_subsampling.compute(g["w_points"], sampleDl=0.1, method="median")
    with pytest.raises(RuntimeError):
        cpp_subsampling.compute(g["w_points"][:, :2], sampleDl=0.1)
    with pytest.raises(RuntimeError):
        cpp_subsampling.compute(g["w_points"], features=g["w_features"][:10], sampleDl=0.1)


@pytest.mark.parametrize("halfwarp", ["1", "0"])
@pytest.mark.parametrize("name", ["demo", "frag", "lattice"])
def test_neighbors_match_golden(cuda, golden, name, halfwarp, monkeypatch):
    """Both query kernels (two queries per warp = the default, one query per warp = D3F_NB_HALFWARP=0) against the
    reference's rows: real scan data, stacked fragments, and the lattice whose exact d2 ties force the exact sort."""
    from d3feat_b200 import tf_custom_ops as ops
    monkeypatch.setenv("D3F_NB_HALFWARP", halfwarp)
    g = golden("neighbors_demo.npz" if name == "demo" else "synthetic.npz")
    if name == "demo":
        P, L, ref, r = g["points"], g["lengths"], g["neighbors"], float(g["radius"])
    elif name == "frag":
        P, L, ref, r = g["frag"], g["frag_lengths"], g["frag_neighbors"], 0.075
    else:
        P, ref, r = g["lattice"], g["lattice_neighbors"], 0.075
        L = np.array([P.shape[0]], np.int32)
    out = ops.batch_ordered_neighbors(t(P, cuda), t(P, cuda), t(L, cuda), t(L, cuda), r).cpu().numpy()
    canon, _ = on.canonicalize_neighbors(ref, P, P, P.shape[0])
    assert out.shape == ref.shape and out.dtype == np.int32
    assert np.array_equal(out, canon)


def test_pool_upsample_and_ordered_neighbors_match_golden(cuda, golden):
    from d3feat_b200 import tf_custom_ops as ops
    g = golden("neighbors_demo.npz")
    q, qb, s, sb = g["pool_points"], g["pool_lengths"], g["points"], g["lengths"]
    r = float(g["radius"])
    out = ops.tf_batch_neighbors(t(q, cuda), t(s, cuda), t(qb, cuda), t(sb, cuda), r).cpu().numpy()
    assert np.array_equal(out, on.canonicalize_neighbors(g["pool_neighbors"], q, s, s.shape[0])[0])
    out = ops.tf_batch_neighbors(t(s, cuda), t(q, cuda), t(sb, cuda), t(qb, cuda), 2 * r).cpu().numpy()
    assert np.array_equal(out, on.canonicalize_neighbors(g["up_neighbors"], s, q, q.shape[0])[0])
    out = ops.ordered_neighbors(t(g["ord_points"], cuda), t(g["ord_points"], cuda), float(g["ord_radius"])).cpu().numpy()
    assert np.array_equal(out, on.canonicalize_neighbors(g["ord_neighbors"], g["ord_points"], g["ord_points"], -1)[0])


def test_neighbors_vs_oracle_stacked_fragments(cuda):
    """Config-#2 shaped: three stacked ~8k fragments, conv / pool / upsample searches, vs the exhaustive oracle."""
    from d3feat_b200 import synth, tf_custom_ops as ops
    clouds = [synth.room_fragment(10 + i, n) for i, n in enumerate((8000, 6500, 7200))]
    P = np.concatenate(clouds, 0)
    L = np.array([c.shape[0] for c in clouds], np.int32)
    tp, tl = t(P, cuda), t(L, cuda)
    out = ops.batch_ordered_neighbors(tp, tp, tl, tl, 0.075).cpu().numpy()
    ref = on.port_batch_neighbors(P, P, L, L, 0.075)
    assert np.array_equal(out, ref)
    sp, sb = ops.batch_grid_subsampling(tp, tl, 0.06)
    rp, rb = on.port_batch_subsampling(P, L, 0.06)
    assert np.array_equal(sb.cpu().numpy(), rb) and np.array_equal(bits(sp.cpu().numpy()), bits(rp))
    pool = ops.batch_ordered_neighbors(sp, tp, sb, tl, 0.075).cpu().numpy()
    assert np.array_equal(pool, on.port_batch_neighbors(rp, P, rb, L, 0.075))
    up = ops.batch_ordered_neighbors(tp, sp, tl, sb, 0.15).cpu().numpy()
    assert np.array_equal(up, on.port_batch_neighbors(P, rp, L, rb, 0.15))
    # capped single-phase form == the reference's post-hoc column slice (datasets/common.py:399-406)
    capped = ops.batch_ordered_neighbors(tp, tp, tl, tl, 0.075, max_cols=20).cpu().numpy()
    assert np.array_equal(capped, ref[:, :20])


def test_edge_cases(cuda):
    from d3feat_b200 import tf_custom_ops as ops
    # single point; strict '<' at exactly r; empty query set; cloud of identical points
    P = torch.zeros((1, 3), device=cuda)
    one = torch.tensor([1], dtype=torch.int32, device=cuda)
    assert ops.batch_ordered_neighbors(P, P, one, one, 0.1).cpu().tolist() == [[0]]
    Q = torch.tensor([[0., 0., 0.]], device=cuda)
    S = torch.tensor([[0.5, 0., 0.], [0.25, 0., 0.]], device=cuda)
    two = torch.tensor([2], dtype=torch.int32, device=cuda)
    assert ops.batch_ordered_neighbors(Q, S, one, two, 0.5).cpu().tolist() == [[1]]
    empty = torch.zeros((0, 3), device=cuda)
    zero = torch.tensor([0], dtype=torch.int32, device=cuda)
    out = ops.batch_ordered_neighbors(empty, S, zero, two, 0.5)
    assert out.shape[0] == 0
    same = torch.full((70, 3), 0.3, device=cuda)
    n70 = torch.tensor([70], dtype=torch.int32, device=cuda)
    out = ops.batch_ordered_neighbors(same, same, n70, n70, 0.05).cpu().numpy()
    assert out.shape == (70, 70) and np.array_equal(out, np.tile(np.arange(70), (70, 1)))   # ties -> index order
    p, b = ops.batch_grid_subsampling(same, n70, 0.1)
    assert p.shape == (1, 3) and b.cpu().tolist() == [1]
    # a cloud stacked next to an empty one (the reference's `if` instead of `while`, neighbors.cpp:272, breaks here)
    L = torch.tensor([0, 70], dtype=torch.int32, device=cuda)
    out2 = ops.batch_ordered_neighbors(same, same, L, L, 0.05).cpu().numpy()
    assert np.array_equal(out2, out)


@pytest.mark.parametrize("halfwarp", ["1", "0"])
def test_dense_rows_take_the_generic_path(cuda, halfwarp, monkeypatch):
    """> 512 hits per query (shared-memory list overflows): the re-scan path must give the same rows."""
    from d3feat_b200 import tf_custom_ops as ops
    monkeypatch.setenv("D3F_NB_HALFWARP", halfwarp)
    rng = np.random.default_rng(3)
    P = rng.uniform(0, 0.2, (1500, 3)).astype(np.float32)
    L = np.array([1500], np.int32)
    out = ops.batch_ordered_neighbors(t(P, cuda), t(P, cuda), t(L, cuda), t(L, cuda), 0.15, max_cols=64).cpu().numpy()
    ref = on.port_batch_neighbors(P, P, L, L, 0.15, max_cols=64)
    assert (on.port_batch_neighbors(P, P, L, L, 0.15, return_counts=True)[1] > 512).any()
    assert np.array_equal(out, ref)


def test_full_size_properties_1m(cuda):
    """Config #5 size (1 M raw points): size-independent properties instead of the O(N^2) oracle."""
    from d3feat_b200 import synth, tf_custom_ops as ops
    P = synth.surface_cloud(0, 1000000)
    tp = t(P, cuda)
    n = torch.tensor([P.shape[0]], dtype=torch.int32, device=cuda)
    sp, sb = ops.batch_grid_subsampling(tp, n, 0.03)
    M = sp.shape[0]
    assert int(sb.item()) == M and 0 < M < P.shape[0]
    # idempotence of the cell partition: every barycenter lies in the bbox; re-subsampling the barycenters at the
    # same dl cannot create more cells than points
    sp2, sb2 = ops.batch_grid_subsampling(sp, sb, 0.03)
    assert sp2.shape[0] <= M
    # number of occupied cells agrees with an independent numpy count of the reference cell keys
    mn = P.min(0)
    org = np.floor(mn * np.float32(1 / np.float32(0.03))) * np.float32(0.03)
    cells = np.floor((P - org) / np.float32(0.03)).astype(np.int64)
    assert np.unique(cells, axis=0).shape[0] == M
    # neighbours of the subsampled cloud: rows sorted by distance, self first, symmetric membership
    m = torch.tensor([M], dtype=torch.int32, device=cuda)
    nb = ops.batch_ordered_neighbors(sp, sp, m, m, 0.075)
    nbc = nb.cpu().numpy()
    spc = sp.cpu().numpy()
    assert np.array_equal(nbc[:, 0], np.arange(M))
    valid = nbc < M
    d2 = on.sqdist_f32(spc[:, None, :], spc[np.where(valid, nbc, 0)])
    d2 = np.where(valid, d2, np.float32(1e30))      # finite pad: inf - inf would be NaN in the diff
    assert np.all(np.diff(d2, axis=1) >= 0)
    assert np.all(d2[valid] < np.float32(0.075) * np.float32(0.075))
    # symmetry on a sample of rows: j in N(i)  =>  i in N(j)
    sel = np.arange(0, M, max(M // 2000, 1))[:2000]
    J = nbc[sel]
    back = (nbc[np.where(J < M, J, 0)] == sel[:, None, None]).any(-1)
    assert np.all(back[J < M])
