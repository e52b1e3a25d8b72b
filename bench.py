#!/usr/bin/env python
"""Benchmark of the D3Feat hot path (BASELINE.json): points/sec through the KPFCNN encoder.

    python bench.py --gpus N --steps K --warmup W            # this repository's CUDA path
    python bench.py --impl reference --gpus N ...            # the reference's CPU path (oracle/_ref + restatement)

A step = one pass of the hot path (4 grid subsamplings + 13 radius searches + 10 KPConv + 23 unary convs +
BN/LeakyReLU/pools) over one batch of `--fragments` stacked synthetic 3DMatch-shaped fragments of `--points`
points (BASELINE configs[1], stacked like configs[3]: 8 fragments per GPU). value = level-0 points / second,
whole job. For N > 1 every rank runs its own fragments (weak scaling) and the step ends with the NCCL
all-gather of the per-fragment descriptors.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

LIMITS = [40, 40, 40, 40, 40]      # "max 40 neighbors" (north_star); calibrated caps are 35-40 on real fragments


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--fragments", type=int, default=8, help="fragments stacked per GPU per step")
    ap.add_argument("--points", type=int, default=30000, help="level-0 points per fragment")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-pipeline", action="store_true", help="one batch at a time on one stream")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def kpconv_algorithmic_bytes(Nq, H, K, Cin, Cout):
    """bytes = Nq*H*(4 + 12 + 4*Cin) + 4*Nq*Cout + 4*K*Cin*Cout + 12*Nq."""
    return Nq * H * (4 + 12 + 4 * Cin) + 4 * Nq * Cout + 4 * K * Cin * Cout + 12 * Nq


# ----------------------------------------------------------------------------------------------------
#  reference arm / CPU baseline: the reference's own C++ cores (oracle/_ref) for the pyramid and the numpy
#  restatement of the TF graph for the encoder, on the host cores
# ----------------------------------------------------------------------------------------------------

def cpu_one_fragment(cfg, params, pts, limits, use_ref):
    from oracle import native as on
    from oracle import kpconv_np as ok
    nb = on.ref_batch_neighbors if use_ref else on.port_batch_neighbors
    sb = on.ref_batch_subsampling if use_ref else on.port_batch_subsampling
    lens = np.array([pts.shape[0]], np.int32)
    inputs = ok.descriptor_input_pyramid(cfg, pts, lens, limits, nb, sb)
    inputs["features"] = np.ones((pts.shape[0], 1), np.float32)
    F = ok.EncoderOracle(cfg, params, np.float32).encoder(inputs)
    return F[-1]


def cpu_reference_run(cfg, params, n_points, steps, warmup):
    """Each step: T = min(nproc, 8) fragments in parallel threads (the reference's tf.data map runs
    input_threads = 8 pyramids concurrently, training_3DMatch.py:35; ctypes and BLAS release the GIL)."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import native as on
    from d3feat_b200 import synth
    use_ref = on.have_ref()
    if not use_ref:
        on.port()
    nproc = os.cpu_count() or 1
    T = max(1, min(nproc, 8))
    frags = [synth.room_fragment(100 + i, n_points) for i in range(T)]
    times = []
    with ThreadPoolExecutor(T) as ex:
        for it in range(warmup + steps):
            t0 = time.perf_counter()
            list(ex.map(lambda p: cpu_one_fragment(cfg, params, p, LIMITS, use_ref), frags))
            dt = time.perf_counter() - t0
            if it >= warmup:
                times.append(dt)
    ms = 1000.0 * float(np.mean(times))
    value = T * n_points / (ms / 1000.0)
    kind = "reference" if use_ref else "port"
    sample = ("%d fragments x %d pts per step in %d threads; pyramid = %s, encoder = numpy fp32 restatement "
              "of the TF1 graph (TensorFlow not installable)") % (
        T, n_points, T, "reference C++ cores (oracle/_ref)" if use_ref else "C restatement (oracle/liboracle.so)")
    return value, ms, dict(kind=kind, cores=nproc, threads=T, sample=sample)


# ----------------------------------------------------------------------------------------------------
#  clocks
# ----------------------------------------------------------------------------------------------------

class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", os.environ.get("D3F_BENCH_LMS", "20")], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except OSError:
            pass

    def stop(self):
        out = dict(sm_mhz=None, sm_max_mhz=None, reasons=[])
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        rows = [r.split(",") for r in open(self.f.name).read().strip().splitlines() if r.count(",") >= 8]
        os.unlink(self.f.name)
        if not rows:
            return out
        sm = [float(r[1]) for r in rows if r[1].strip().replace(".", "").isdigit()]
        mx = [float(r[2]) for r in rows if r[2].strip().replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = set()
        for r in rows:
            for i, n in enumerate(names):
                if r[5 + i].strip().lower().startswith("active"):
                    reasons.add(n)
        out.update(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None,
                   reasons=sorted(reasons), samples=len(rows))
        return out


# ----------------------------------------------------------------------------------------------------
def main():
    args = parse()
    from d3feat_b200 import synth
    cfg = synth.Config(architecture=synth.ARCH_ENCODER)
    params = synth.make_params(cfg, seed=0)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    workload = "%d stacked 3DMatch-shaped synthetic fragments x %d pts per GPU, full 5-level KPFCNN encoder" % (
        args.fragments, args.points)
    base = dict(metric="points/sec through KPFCNN encoder", unit="points/s", n_gpus=args.gpus, steps=args.steps,
                warmup=args.warmup, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32",
                data="synthetic",
                config=dict(workload=workload, fragments_per_gpu=args.fragments, points_per_fragment=args.points,
                            levels=5, K=15, neighbor_cols=LIMITS, first_subsampling_dl=0.03,
                            parallelism="fragments sharded, dp%d" % args.gpus))

    if args.impl == "reference":
        if rank != 0:
            return
        # bounded sample: shrink the fragment when many steps are requested so the arm ends within minutes
        n_pts = args.points if (args.steps + args.warmup) <= 8 else max(4000, int(args.points * 8 / (args.steps + args.warmup)))
        value, ms, info = cpu_reference_run(cfg, params, n_pts, args.steps, args.warmup)
        line = dict(base, impl="reference", value=value, ms_per_step=ms,
                    cpu_baseline=dict(value=value, unit="points/s", cores=info["cores"], kind=info["kind"],
                                      sample=info["sample"]),
                    e2e=dict(value=value, unit="points/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                    gpu_launches=0)
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist
    from d3feat_b200 import _lib
    from d3feat_b200 import convolution_ops as co
    from d3feat_b200.encoder import KPFCNN
    from d3feat_b200.distributed import all_gather_descriptors_padded

    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.lib()

    # ---- workload: this rank's fragments (seeded by global fragment id) --------------------------------
    frag_ids = [rank * args.fragments + i for i in range(args.fragments)]
    clouds = [synth.room_fragment(f, args.points) for f in frag_ids]
    P = np.concatenate(clouds, 0)
    L = np.array([c.shape[0] for c in clouds], np.int32)
    n_points = int(P.shape[0])
    P_pin = torch.from_numpy(P).pin_memory()
    L_pin = torch.from_numpy(L).pin_memory()
    P_dev, L_dev = P_pin.to(dev), L_pin.to(dev)
    enc = KPFCNN(cfg, params, LIMITS, device=dev)
    bbox = np.concatenate([P.min(0), P.max(0)]).astype(np.float32)

    gather_cap = max(64, n_points // 128)    # rows reserved per rank for the coarsest-level descriptors (~1.5x actual)

    def step_resident():
        out = enc(P_dev, L_dev, bbox=bbox, decoder=False)
        desc = out["F"][-1]
        if world > 1:      # the one exchange step: NCCL all-gather of the per-fragment descriptors (sync-free)
            desc, _ = all_gather_descriptors_padded(desc, out["inputs"]["lengths"][-1], gather_cap)
        return desc

    def step_e2e():
        # the call a user makes: host buffers in, descriptors out (host) -- H2D and D2H inside the timed region
        p = P_pin.to(dev, non_blocking=True)
        l = L_pin.to(dev, non_blocking=True)
        out = enc(p, l, decoder=False)          # bbox computed on the device (one small D2H read)
        desc = out["F"][-1]
        if world > 1:      # the gathered matrix stays in HBM (that is where a matcher consumes it) ...
            gathered[0], _ = all_gather_descriptors_padded(desc, out["inputs"]["lengths"][-1], gather_cap)
        return desc.cpu()  # ... the host gets this rank's own descriptors

    gathered = [None]      # the most recent all-gathered descriptor matrix (kept alive until the next step replaces it)
    flush_buf = torch.empty((256 << 20,), dtype=torch.uint8, device=dev)     # > 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        n0 = _lib.launch_count()
        t0 = time.perf_counter()
        for a, b in evs:
            flush_buf.fill_(1)            # L2 flush between timed iterations (outside the event bracket)
            a.record()
            fn()
            b.record()
        barrier()
        wall = (time.perf_counter() - t0) * 1000.0 / steps
        launches = (_lib.launch_count() - n0) // max(steps, 1)
        ms = float(np.mean([a.elapsed_time(b) for a, b in evs]))
        if world > 1:
            tms = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(tms, op=dist.ReduceOp.MAX)
            ms = float(tms.item())
        return ms, wall, launches

    def timed_pipelined(steps, warmup, e2e):
        """K steps of the two-stream pipeline: encoder(i) on one stream while the pyramid of batch i+1 is built on
        the other (encoder.BatchPipeline). The timed region holds exactly K encoders and K pyramids (the first
        encoder consumes the primed pyramid, the last step builds one more), the L2 flush of every step, and for
        e2e the H2D copy of each batch's points and the D2H copy of each batch's descriptors."""
        from d3feat_b200.encoder import BatchPipeline

        def post(inputs, desc):
            if world > 1:      # the one exchange step: NCCL all-gather of the per-fragment descriptors (sync-free);
                # the gathered matrix stays in HBM, the step returns this rank's own descriptors
                gathered[0], _ = all_gather_descriptors_padded(desc, inputs["lengths"][-1], gather_cap)
            return desc

        pipe = BatchPipeline(enc, decoder=False, post=post)
        src_p, src_l, src_bbox = (P_pin, L_pin, None) if e2e else (P_dev, L_dev, bbox)
        host_out = None

        def one(k_flush):
            nonlocal host_out
            res = pipe.step(src_p, src_l, src_bbox, pre=(lambda: flush_buf.fill_(1)) if k_flush else None)
            if e2e:
                with torch.cuda.stream(pipe.s_enc):
                    if host_out is None:
                        host_out = torch.empty(res.shape, dtype=res.dtype, pin_memory=True)
                    host_out.copy_(res, non_blocking=True)
            return res

        pipe.prime(src_p, src_l, src_bbox)
        for _ in range(warmup):
            one(False)
        # untimed settling: a fresh process can see one-off stalls (allocator growth, the previous process's context
        # still being torn down); keep warming up until five consecutive steps run within 1.5x of the fastest seen
        # (with several ranks every step holds a collective, so the count must be the same everywhere: fixed)
        best, calm = float("inf"), 0
        for it in range(40):
            t_s = time.perf_counter()
            one(False)
            dt = time.perf_counter() - t_s
            best = min(best, dt)
            calm = calm + 1 if dt < 1.5 * best else 0
            if (world == 1 and calm >= 5) or (world > 1 and it >= 9):
                break
        pipe.drain()
        barrier()
        n0 = _lib.launch_count()
        t0 = time.perf_counter()
        marks = []
        for _ in range(steps):
            one(True)
            marks.append(time.perf_counter())
        pipe.drain()
        barrier()
        ms = (time.perf_counter() - t0) * 1000.0 / steps      # synchronised on both sides: device-bound wall time
        step_marks.append([round((b - a) * 1000.0, 2) for a, b in zip([t0] + marks[:-1], marks)])
        launches = (_lib.launch_count() - n0) // max(steps, 1)
        if world > 1:
            tms = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(tms, op=dist.ReduceOp.MAX)
            ms = float(tms.item())
        return ms, launches

    step_marks = []        # host-side time between consecutive pipeline steps (diagnostic: shows one-off stalls)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    seq_ms, _, _ = timed(step_resident, args.steps, max(args.warmup, 3))      # un-pipelined latency of one batch
    if args.no_pipeline:
        ms, wall_ms, launches = timed(step_resident, args.steps, 1)
        clocks = sampler.stop() if sampler else None
        ms_e2e, wall_e2e, _ = timed(step_e2e, args.steps, 1)
    else:
        ms, launches = timed_pipelined(args.steps, max(args.warmup, 3), False)
        wall_ms = ms
        clocks = sampler.stop() if sampler else None
        ms_e2e, _ = timed_pipelined(args.steps, max(args.warmup, 3), True)
    total_points = n_points * world
    value = total_points / (ms / 1000.0)
    e2e_value = total_points / (ms_e2e / 1000.0)
    d2h = int(step_e2e().numel() * 4)      # bytes of the host tensor the e2e step returns (per rank)

    # ---- roofline of the dominant kernel: the largest KPConv (level 0, 32 -> 32 on all level-0 points) --
    roof = None
    if rank == 0:
        out = enc(P_dev, L_dev, bbox=bbox, decoder=False)
        inp = out["inputs"]
        q = inp["points"][0]
        idx = inp["neighbors"][0]
        Nq, H = idx.shape
        feat = torch.randn((Nq, 32), device=dev)
        Kp = enc.store.get("layer_0/resnetb_1/conv2/kernel_points")
        W = enc.store.get("layer_0/resnetb_1/conv2/weights")
        extent = cfg.KP_extent * (cfg.first_subsampling_dl * cfg.density_parameter) / cfg.density_parameter
        for _ in range(3):
            co.KPConv_ops(q, q, idx, feat, Kp, W, extent, "linear", "sum")
        reps = 10
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
        for a, b in evs:
            flush_buf.fill_(1)
            a.record()
            co.KPConv_ops(q, q, idx, feat, Kp, W, extent, "linear", "sum")
            b.record()
        torch.cuda.synchronize()
        kms = float(np.mean([a.elapsed_time(b) for a, b in evs]))
        abytes = kpconv_algorithmic_bytes(Nq, H, 15, 32, 32)
        peak, peak_src = peaks()
        ach = abytes / (kms * 1e-3) / 1e9
        # DRAM bytes of the op from the committed ncu capture (profiles/r1_v4_ncu_metrics.txt), per 21760-query chunk:
        # stage-1 kernel 8.07 MB read + 0.33 MB written (the gathered rows are L2 hits), contraction 42.04 MB read
        # (the chunk's wf -- ncu flushes L2 between kernels; back to back it is an L2 hit as well)
        n_chunks = -(-int(Nq) // 21760)
        traffic = int((8.07e6 + 0.33e6 + 42.04e6) * n_chunks) if (args.fragments, args.points) == (8, 30000) else None
        roof = dict(bound="hbm", kernel="kpconv level-0 32->32 (mma.sync stage-1 gather/correlation + tcgen05 contraction)", achieved=ach,
                    peak=peak, unit="GB/s", frac=ach / peak, traffic=traffic, peak_source=peak_src,
                    algorithmic_bytes_per_launch=abytes, ms_per_launch=kms, Nq=int(Nq), H=int(H))

    cpu = None
    if rank == 0 and args.gpus == 1 and not args.no_cpu_baseline:
        v, cms, info = cpu_reference_run(cfg, params, args.points, 1, 0)
        cpu = dict(value=v, unit="points/s", cores=info["cores"], kind=info["kind"], sample=info["sample"],
                   ms_per_step=cms)

    if rank == 0:
        line = dict(base, value=value, ms_per_step=ms, wall_ms_per_step=wall_ms,
                    e2e=dict(value=e2e_value, unit="points/s", h2d_bytes_per_step=int(P.nbytes + L.nbytes) * world,
                             d2h_bytes_per_step=d2h, ms_per_step=ms_e2e),
                    gpu_launches=int(launches), clocks=clocks, roofline=roof, cpu_baseline=cpu)
        line["config"]["l2"] = "256 MiB L2 flush at the start of every timed step"
        line["config"]["e2e_output"] = ("every rank returns its own fragments' descriptors to its host; with N > 1 the "
                                        "all-gathered matrix stays in HBM")
        line["config"]["pipeline"] = ("one batch at a time" if args.no_pipeline else
                                      "two streams: pyramid(i+1) || encoder(i) (encoder.BatchPipeline)")
        line["single_batch_latency_ms"] = seq_ms
        if step_marks:
            line["pipeline_host_step_ms"] = dict(resident=step_marks[0], e2e=step_marks[-1])
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
