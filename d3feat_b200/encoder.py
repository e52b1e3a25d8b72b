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

    def build_inputs_static(self, buffers):
        """Sync-free pyramid over the batch already sitting in buffers.points0 / lengths0 / n0 (see
        pyramid.descriptor_input, static form). Every tensor is a whole capacity-sized buffer."""
        inputs = pyramid.descriptor_input(self.config, buffers.points0, buffers.lengths0, self.limits, buffers=buffers,
                                          static=True)
        inputs["features"] = buffers.features0
        return inputs

    def build_inputs(self, stacked_points, stacked_lengths, features=None, bbox=None, buffers=None):
        pts = stacked_points
        if not torch.is_tensor(pts):
            pts = torch.as_tensor(np.ascontiguousarray(pts, np.float32))
        pts = pts.to(self.device, non_blocking=True)
        lens = stacked_lengths
        if not torch.is_tensor(lens):
            lens = torch.as_tensor(np.ascontiguousarray(lens, np.int32))
        lens = lens.to(self.device, non_blocking=True)
        inputs = pyramid.descriptor_input(self.config, pts, lens, self.limits, bbox=bbox, buffers=buffers)
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

    def describe(self, inputs, F, with_scores=False):
        """Decoder -> descriptors [N,32]; with_scores=True -> (descriptors, detection scores [N,1]) -- the two arrays
        tester.generate_descriptor dumps per fragment (utils/tester.py)."""
        with use_params(self.store):
            return nb.assemble_FCNN_decoder(inputs, self.config, F, 1.0, with_scores=with_scores)

    def __call__(self, stacked_points, stacked_lengths, features=None, bbox=None, decoder=None):
        inputs = self.build_inputs(stacked_points, stacked_lengths, features, bbox)
        F = self.encode(inputs)
        use_dec = self.has_decoder if decoder is None else decoder
        desc, scores = self.describe(inputs, F, with_scores=True) if use_dec else (None, None)
        return dict(inputs=inputs, F=F, descriptors=desc, scores=scores)


class BatchPipeline:
    """Throughput mode: the input pyramid of batch i+1 is built on a second CUDA stream while the encoder of batch i
    runs -- the overlap the reference gets from tf.data prefetch (its CPU pyramid runs ahead of the GPU model,
    datasets/common.py:744-763). Usage:

        pipe = BatchPipeline(enc)
        pipe.prime(points0, lengths0, bbox0)            # pyramid of the first batch
        for i in range(K):
            out = pipe.step(points_next, lengths_next, bbox_next)   # encoder(i) || pyramid(i+1); returns F of batch i
        pipe.drain()
    """

    def __init__(self, enc, decoder=False, post=None):
        self.enc = enc
        self.decoder = decoder
        self.post = post                      # optional callable(inputs, F) run on the encoder stream (e.g. all-gather)
        dev = enc.device
        self.s_pyr = torch.cuda.Stream(device=dev)
        self.s_enc = torch.cuda.Stream(device=dev)
        self.ready = torch.cuda.Event()
        self.pending = None
        self.keep = []                        # keeps the previous batch's tensors alive until its kernels are done
        # Ring of pre-allocated pyramid slots: slot (i mod DEPTH) is rewritten by pyramid(i + DEPTH) only after the
        # host has waited for encoder(i) (self.done), so at most DEPTH - 1 encoders are ever queued behind the host
        # and steady-state batches allocate nothing for the pyramid.
        self.slots = [None] * self.DEPTH
        self.done = [None] * self.DEPTH
        self.n_built = 0

    DEPTH = 3

    def _slot(self, n_points, n_clouds):
        k = self.n_built % self.DEPTH
        self.n_built += 1
        if self.done[k] is not None:
            self.done[k].synchronize()        # the encoder that read this slot DEPTH batches ago has finished
        buf = self.slots[k]
        if buf is None or not buf.fits(n_points, n_clouds, self.enc.limits):
            buf = pyramid.PyramidBuffers(self.enc.config, self.enc.limits, int(n_points * 1.05) + 64, n_clouds,
                                         self.enc.device)
            self.slots[k] = buf
        return k, buf

    def _build(self, points, lengths, bbox, inputs_ready):
        k, buf = self._slot(int(points.shape[0]), int(lengths.shape[0]))
        self.s_pyr.wait_event(inputs_ready)   # caller-produced CUDA inputs are complete before the pyramid reads them
        with torch.cuda.stream(self.s_pyr):
            inputs = self.enc.build_inputs(points, lengths, bbox=bbox, buffers=buf)
            self.ready.record(self.s_pyr)
        for t in (points, lengths):           # caller tensors are read on the pyramid stream
            if torch.is_tensor(t) and t.is_cuda:
                t.record_stream(self.s_pyr)
        inputs["_slot"] = k
        for v in inputs.values():             # the pyramid's tensors are consumed on the encoder stream
            for t in (v if isinstance(v, (list, tuple)) else [v]):
                if torch.is_tensor(t) and t.is_cuda:
                    t.record_stream(self.s_enc)
        return inputs

    def _mark_inputs(self):
        """Event on the caller's current stream: everything the caller enqueued so far (the next batch's CUDA inputs)."""
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.enc.device))
        return ev

    def prime(self, points, lengths, bbox=None):
        self.pending = self._build(points, lengths, bbox, self._mark_inputs())

    def step(self, next_points=None, next_lengths=None, next_bbox=None, pre=None):
        """Enqueue encoder(current batch) on the encoder stream, then build the pyramid of the next batch on the
        pyramid stream (the host blocks only in that stream's size read-backs). Returns the current batch's result.

        Stream contract: the result is produced on the private encoder stream; before returning, the CALLER's
        current stream is made to wait for it (an event, no host sync) and the result's memory is tied to that
        stream, so `res.cpu()` or a kernel launched on the caller's stream reads finished data."""
        inputs = self.pending
        cur = torch.cuda.current_stream(self.enc.device)
        # taken BEFORE `cur` is made to wait for this batch's encoder: the next pyramid then depends on the caller's
        # inputs only, not on encoder(i) -- the pyramid(i+1) || encoder(i) overlap is preserved
        inputs_ready = self._mark_inputs() if next_points is not None else None
        with torch.cuda.stream(self.s_enc):
            self.s_enc.wait_event(self.ready)
            if pre is not None:
                pre()
            F = self.enc.encode(inputs)
            res = self.enc.describe(inputs, F) if self.decoder else F[-1]
            if self.post is not None:
                res = self.post(inputs, res)
            ev = torch.cuda.Event()
            ev.record(self.s_enc)
            self.done[inputs["_slot"]] = ev
        cur.wait_event(ev)
        for t in (res if isinstance(res, (list, tuple)) else [res]):
            if torch.is_tensor(t) and t.is_cuda:
                t.record_stream(cur)
        self.result_event = ev
        self.keep = [inputs, F]
        self.pending = (self._build(next_points, next_lengths, next_bbox, inputs_ready)
                        if next_points is not None else None)
        return res

    def drain(self):
        self.s_pyr.synchronize()
        self.s_enc.synchronize()
        self.keep = []


class GraphPipeline:
    """Throughput / latency mode without the host in the loop: the whole step is two CUDA graph launches.

    The pyramid is built in its static form (capacity-sized launches, level sizes stay in device memory, no
    device->host read), the encoder's kernels take their row counts from the same device counters, so the launch
    sequence of a step depends only on the shape BUCKET (number of clouds, per-level capacities, scene bounds), not on
    the batch. Per ring slot the pyramid and the encoder are captured once and then replayed:

        pipe = GraphPipeline.for_batch(enc, points0, lengths0)     # one exact (synchronising) pass sizes the bucket
        pipe.prime(points, lengths)                                # H2D / D2D copy into the slot + pyramid graph
        for ...:
            res, counts = pipe.step(next_points, next_lengths)     # encoder graph (i) || pyramid graph (i+1)
        pipe.drain(); pipe.check()                                 # status bits: capacity overflow / points out of bounds

    `res` is the slot's static output buffer [capacity of the last level, C] and `counts` the device int32 level sizes
    (rows beyond counts[-1] are undefined); both stay valid until the slot is reused DEPTH steps later. Mirrors the
    overlap the reference gets from tf.data prefetch (datasets/common.py:744-763)."""

    DEPTH = 4

    def __init__(self, enc, capacities, n_clouds, bbox, decoder=False, post=None, encoder_streams=2):
        self.enc, self.decoder, self.post = enc, decoder, post
        dev = enc.device
        self.caps = [int(c) for c in capacities]
        self.n_clouds = int(n_clouds)
        self.bbox = np.ascontiguousarray(bbox, np.float32)
        self.s_pyr = torch.cuda.Stream(device=dev)
        # Encoders of consecutive batches alternate between `encoder_streams` streams: the deep pyramid levels (a few
        # thousand rows) cannot fill 148 SMs on their own, so the tail of encoder(i) runs under the level-0 kernels of
        # encoder(i + 1). Results still come back in batch order (each step waits for its own batch's event).
        self.s_encs = [torch.cuda.Stream(device=dev) for _ in range(max(1, int(encoder_streams)))]
        self.s_enc = self.s_encs[0]
        self.n_stepped = 0
        self.DEPTH = len(self.s_encs) + 2     # encoders in flight + the pyramid being built + one slot of slack
        self.slots = [pyramid.PyramidBuffers(enc.config, enc.limits, self.caps, self.n_clouds, dev, bbox=self.bbox)
                      for _ in range(self.DEPTH)]
        self.g_pyr = [None] * self.DEPTH
        self.g_enc = [None] * self.DEPTH
        self.out = [None] * self.DEPTH          # (inputs, F, res) captured per slot
        self.ready = [torch.cuda.Event() for _ in range(self.DEPTH)]
        self.done = [None] * self.DEPTH
        self.kernels_per_step = 0
        self.n_loaded = 0
        self.pending = None

    @classmethod
    def for_batch(cls, enc, points, lengths, slack=1.125, margin=0.05, **kw):
        """Bucket from a representative batch: one exact pass gives the level sizes (capacities = sizes x slack) and
        the scene bounds (its bbox inflated by `margin` of the extent on every side)."""
        inputs = enc.build_inputs(points, lengths)
        sizes = [int(p.shape[0]) for p in inputs["points"]]
        pts = inputs["points"][0]
        bb = pyramid.ops.host_bbox(pts)
        ext = np.maximum(bb[3:] - bb[:3], 1e-3)
        bb = np.concatenate([bb[:3] - margin * ext, bb[3:] + margin * ext]).astype(np.float32)
        return cls(enc, pyramid.bucket_capacities(sizes, slack), int(lengths.shape[0]), bb, **kw)

    # ---- one slot -----------------------------------------------------------------------------------------
    def _run_pyramid(self, k):
        return self.enc.build_inputs_static(self.slots[k])

    def _run_encoder(self, inputs):
        F = self.enc.encode(inputs)
        res = self.enc.describe(inputs, F) if self.decoder else F[-1]
        return F, res

    def _capture(self, k):
        """Eager warm-up of both halves on slot k (lazy one-time work: weight packing, BN folding, kernel attributes,
        the library's auxiliary stream), then the two captures."""
        from . import _lib
        for _ in range(2):
            with torch.cuda.stream(self.s_pyr):
                inputs = self._run_pyramid(k)
            self.s_enc.wait_stream(self.s_pyr)
            with torch.cuda.stream(self.s_enc):
                self._run_encoder(inputs)
            self.s_pyr.wait_stream(self.s_enc)
        torch.cuda.synchronize(self.enc.device)
        n0 = _lib.launch_count()
        gp = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gp, stream=self.s_pyr):
            inputs = self._run_pyramid(k)
        ge = torch.cuda.CUDAGraph()
        with torch.cuda.graph(ge, stream=self.s_enc):
            F, res = self._run_encoder(inputs)
        self.kernels_per_step = _lib.launch_count() - n0
        self.g_pyr[k], self.g_enc[k], self.out[k] = gp, ge, (inputs, F, res)

    def _load(self, points, lengths, inputs_ready):
        k = self.n_loaded % self.DEPTH
        self.n_loaded += 1
        if self.done[k] is not None:
            self.done[k].synchronize()        # bounds the host's run-ahead to DEPTH steps (normally long complete)
        buf = self.slots[k]
        n0 = int(points.shape[0])
        if n0 > buf.caps[0] or int(lengths.shape[0]) != self.n_clouds:
            raise ValueError("GraphPipeline: batch (%d points, %d clouds) does not fit the bucket (%d, %d)" % (
                n0, int(lengths.shape[0]), buf.caps[0], self.n_clouds))
        if self.g_pyr[k] is None:             # first use of the slot: fill it, then capture its two graphs
            buf.points0[:n0].copy_(torch.as_tensor(points), non_blocking=True)
            buf.lengths0.copy_(torch.as_tensor(lengths), non_blocking=True)
            buf.n0.fill_(n0)
            torch.cuda.synchronize(self.enc.device)
            self._capture(k)
        self.s_pyr.wait_event(inputs_ready)
        with torch.cuda.stream(self.s_pyr):
            buf.points0[:n0].copy_(torch.as_tensor(points), non_blocking=True)
            buf.lengths0.copy_(torch.as_tensor(lengths), non_blocking=True)
            buf.n0.fill_(n0)
            self.g_pyr[k].replay()
            self.ready[k].record(self.s_pyr)
        for t in (points, lengths):
            if torch.is_tensor(t) and t.is_cuda:
                t.record_stream(self.s_pyr)
        return k

    def _mark_inputs(self):
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.enc.device))
        return ev

    def prime(self, points, lengths):
        self.pending = self._load(points, lengths, self._mark_inputs())

    def step(self, next_points=None, next_lengths=None, pre=None):
        """Replay encoder(current batch), then load + replay the pyramid of the next batch on the other stream.
        Returns (result buffer, device level counts) of the current batch, ordered on the caller's stream."""
        k = self.pending
        cur = torch.cuda.current_stream(self.enc.device)
        inputs_ready = self._mark_inputs() if next_points is not None else None
        inputs, F, res = self.out[k]
        s_enc = self.s_encs[self.n_stepped % len(self.s_encs)]
        self.n_stepped += 1
        self.s_enc = s_enc                    # the stream this step's result is produced on
        with torch.cuda.stream(s_enc):
            s_enc.wait_event(self.ready[k])
            if pre is not None:
                pre()
            self.g_enc[k].replay()
            if self.post is not None:
                res = self.post(inputs, res)
            ev = torch.cuda.Event()
            ev.record(s_enc)
            self.done[k] = ev
        cur.wait_event(ev)
        self.pending = self._load(next_points, next_lengths, inputs_ready) if next_points is not None else None
        return res, self.slots[k].counts

    def drain(self):
        self.s_pyr.synchronize()
        for s in self.s_encs:
            s.synchronize()

    def check(self):
        """Raise if any replayed batch overflowed the bucket (synchronises)."""
        self.drain()
        for k, buf in enumerate(self.slots):
            st = int(buf.status.item())
            if st:
                raise RuntimeError("GraphPipeline: slot %d status %d (%s)" % (
                    k, st, "points outside the scene bounds" if st & 1 else "a level exceeded its capacity"))
