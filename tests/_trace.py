"""Test infrastructure: record every fused op the GPU encoder launches (inputs and output, device tensors) so that
the numpy restatement can be evaluated on a SAMPLE OF ROWS of each op at sizes where the full float64 restatement is
too slow (BASELINE configs[1]/[2]/[3] at their real sizes). Each op is independent per output row, so checking
sampled rows against the oracle on the op's real inputs is an oracle comparison, not a path-vs-path one.
"""
import contextlib

import numpy as np

from oracle import kpconv_np as ok


class Trace:
    def __init__(self):
        self.records = []


@contextlib.contextmanager
def record_ops():
    from d3feat_b200 import convolution_ops as co
    from d3feat_b200 import network_blocks as nb
    tr = Trace()
    orig = dict(kp=co.KPConv_ops, kd=co.KPConv_deform_ops, un=co.unary_convolution, up=co.unary_pair_convolution,
                mp=nb.ind_max_pool)

    def kp(q, s, idx, f, Kp, W, extent, infl, mode, *, epilogue=None, bias=None, query_order=None, **rows):
        out = orig["kp"](q, s, idx, f, Kp, W, extent, infl, mode, epilogue=epilogue, bias=bias, query_order=query_order,
                         **rows)
        tr.records.append(dict(op="kpconv", q=q, s=s, idx=idx, f=f, Kp=Kp, W=W, extent=extent, infl=infl, mode=mode,
                               epilogue=epilogue, bias=bias, out=out))
        return out

    def kd(q, s, idx, f, Kp, off, mod, W, extent, infl, mode, *, epilogue=None, query_order=None, **rows):
        out = orig["kd"](q, s, idx, f, Kp, off, mod, W, extent, infl, mode, epilogue=epilogue, query_order=query_order,
                         **rows)
        tr.records.append(dict(op="kpconv_deform", q=q, s=s, idx=idx, f=f, Kp=Kp, off=off, mod=mod, W=W, extent=extent,
                               infl=infl, mode=mode, epilogue=epilogue, out=out))
        return out

    def un(x, w, *, epilogue=None, residual=None, rows=None):
        out = orig["un"](x, w, epilogue=epilogue, residual=residual, rows=rows)
        tr.records.append(dict(op="unary", x=x, w=w, epilogue=epilogue, residual=residual, out=out))
        return out

    def up(x1, w1, a1, x2, w2, a2, alpha, *, rows=None):
        out = orig["up"](x1, w1, a1, x2, w2, a2, alpha, rows=rows)
        tr.records.append(dict(op="unary_pair", x1=x1, w1=w1, a1=a1, x2=x2, w2=w2, a2=a2, alpha=alpha, out=out))
        return out

    def mp(x, inds, **rows):
        out = orig["mp"](x, inds, **rows)
        tr.records.append(dict(op="max_pool", x=x, inds=inds, out=out))
        return out

    co.KPConv_ops, co.KPConv_deform_ops, co.unary_convolution, co.unary_pair_convolution = kp, kd, un, up
    nb.ind_max_pool = mp
    try:
        yield tr
    finally:
        co.KPConv_ops, co.KPConv_deform_ops, co.unary_convolution, co.unary_pair_convolution = (
            orig["kp"], orig["kd"], orig["un"], orig["up"])
        nb.ind_max_pool = orig["mp"]


def _np(t):
    return None if t is None else t.detach().cpu().numpy()


def _epi(y, epilogue, residual=None):
    if epilogue is not None:
        scale, shift, alpha = epilogue
        if scale is not None:
            y = y * _np(scale).astype(np.float64) + _np(shift).astype(np.float64)
    else:
        alpha = None
    if residual is not None:
        y = y + residual
    if alpha is not None:
        y = np.where(y > 0, y, alpha * y)
    return y


def check_sampled_rows(trace, n_rows, rng, rtol, min_kpconv=None):
    """For every recorded op: float64 restatement on `n_rows` sampled output rows (all rows when the op has fewer) vs
    the GPU output rows. The tolerance is max-norm relative PER TENSOR: the denominator is the max |value| of the op's
    full GPU output (the quantity test_gpu_kpconv.py normalises by), the numerator the max error over the sample.
    Returns a list of (op, shape, err) for the report."""
    report = []
    n_kp = 0
    for r in trace.records:
        out = _np(r["out"])
        N = out.shape[0]
        rows = np.arange(N) if N <= n_rows else np.sort(rng.choice(N, n_rows, replace=False))
        denom = max(float(np.abs(out).max()), 1e-30)
        if r["op"] == "kpconv":
            ref = ok.kpconv_ops(_np(r["q"])[rows], _np(r["s"]), _np(r["idx"])[rows], _np(r["f"]), _np(r["Kp"]),
                                _np(r["W"]), r["extent"], r["infl"], r["mode"], dtype=np.float64)
            if r["bias"] is not None:          # the offset head of the deformable block: bias, no batch norm (:327-339)
                assert r["epilogue"] is None
                ref = ref + _np(r["bias"]).astype(np.float64)
            ref = _epi(ref, r["epilogue"])
            n_kp += 1
        elif r["op"] == "kpconv_deform":
            ref = ok.kpconv_deform_ops(_np(r["q"])[rows], _np(r["s"]), _np(r["idx"])[rows], _np(r["f"]), _np(r["Kp"]),
                                       _np(r["off"])[rows], None if r["mod"] is None else _np(r["mod"])[rows],
                                       _np(r["W"]), r["extent"], r["infl"], r["mode"], dtype=np.float64)
            ref = _epi(ref, r["epilogue"])
            n_kp += 1
        elif r["op"] == "unary":
            ref = _np(r["x"])[rows].astype(np.float64) @ _np(r["w"]).astype(np.float64)
            res = None if r["residual"] is None else _np(r["residual"])[rows].astype(np.float64)
            ref = _epi(ref, r["epilogue"], res)
        elif r["op"] == "unary_pair":
            (s1, t1), (s2, t2) = r["a1"], r["a2"]
            y = (_np(r["x1"])[rows].astype(np.float64) @ _np(r["w1"]).astype(np.float64)) * _np(s1).astype(np.float64) \
                + _np(t1).astype(np.float64)
            y = y + (_np(r["x2"])[rows].astype(np.float64) @ _np(r["w2"]).astype(np.float64)) \
                * _np(s2).astype(np.float64) + _np(t2).astype(np.float64)
            ref = np.where(y > 0, y, r["alpha"] * y) if r["alpha"] is not None else y
        elif r["op"] == "max_pool":
            x = _np(r["x"])
            ref = ok.ind_max_pool(x, _np(r["inds"])[rows])
            assert np.array_equal(out[rows], ref), "ind_max_pool rows differ (exact op)"
            report.append((r["op"], out.shape, 0.0))
            continue
        else:
            raise AssertionError(r["op"])
        err = float(np.abs(out[rows].astype(np.float64) - ref).max()) / denom
        report.append((r["op"], out.shape, err))
        assert err < rtol, "%s %s: sampled-row error %.3g" % (r["op"], out.shape, err)
    if min_kpconv is not None:
        assert n_kp >= min_kpconv, "only %d KPConv launches recorded" % n_kp
    return report
