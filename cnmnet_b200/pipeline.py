"""End-to-end inference wrapper mirroring the reference's eval loops (eval.py:440-455 for 3 views,
eval.py:635-668 / 885-937 for 5 / 7 views):

    (ref_img [B,3,H,W], src_imgs [B,V,3,H,W], ref_cam [B,2,4,4], src_cams [B,V,2,4,4])
        -> depth [B,H,W], normal [B,3,H,W], prob [B,1,H,W]

All V source views are stacked view-major on the batch axis of one DepthNet plan (the cost volume
kernel shares each reference tile across them), group means + the refine input come from
cnm_refine_aggregate, then DepthRefineNet, depth = 1/idepth and depth->normal.  Inputs are copied
into static buffers and the whole step is replayed as one CUDA graph.
"""
import torch

from . import _lib
from .depthnet.depthNet_model import DepthRefineNet, depthNet


class CNMPipeline:
    def __init__(self, depth_network, depth_refine_network, k_size=9, use_graph=True):
        if not isinstance(depth_network, depthNet) or not isinstance(depth_refine_network, DepthRefineNet):
            raise TypeError("CNMPipeline expects cnmnet_b200 depthNet and DepthRefineNet modules")
        if depth_network.precision != depth_refine_network.precision:
            raise ValueError("both networks must use the same precision")
        self.depth_network, self.refine_network = depth_network, depth_refine_network
        self.k_size, self.use_graph = int(k_size), use_graph
        self._state = {}

    # -- static state for one shape -----------------------------------------------------------
    def _prepare(self, B, V, H, W, device, lane=0):
        key = (B, V, H, W, str(device), lane)
        st = self._state.get(key)
        if st is not None:
            # weights may have changed since the plan was packed: repack, and drop a graph captured with the old ones
            # (head biases and the f16f8 accumulator scales are launch arguments, frozen at capture)
            dplan = self.depth_network.plan(V * B, B, H, W, lane=lane)
            rplan = self.refine_network.plan(B, H, W, V // 2, False, lane=lane)
            epochs = (dplan.weights_epoch, rplan.weights_epoch)
            if st["epochs"] != epochs:
                st["graph"], st["epochs"] = None, epochs
            return st
        if V < 2 or V % 2:
            raise ValueError("the refine network takes two view groups: V must be even and >= 2 (got %d)" % V)
        f32 = dict(dtype=torch.float32, device=device)
        st = {
            "ref": torch.empty((B, 3, H, W), **f32), "src": torch.empty((V, B, 3, H, W), **f32),
            "ref_cam": torch.empty((B, 2, 4, 4), **f32), "src_cam": torch.empty((V, B, 2, 4, 4), **f32),
            "Kinv": torch.empty((B, 3, 3), **f32), "graph": None,
            # the three outputs of a step live in ONE buffer [depth B*HW | normal 3*B*HW | prob B*HW]: a multi-GPU run
            # gathers it with a single collective (cnmnet_b200.dist.OutputGatherer)
            "packed": torch.empty((5 * B * H * W,), **f32),
        }
        n = B * H * W
        st["depth"] = st["packed"][:n].view(B, H, W)
        st["normal"] = st["packed"][n:4 * n].view(B, 3, H, W)
        st["prob"] = st["packed"][4 * n:].view(B, 1, H, W)
        dplan = self.depth_network.plan(V * B, B, H, W, lane=lane)
        rplan = self.refine_network.plan(B, H, W, V // 2, False, lane=lane)
        dplan.bind(st["ref"], st["src"].view(V * B, 3, H, W), st["ref_cam"], st["src_cam"].view(V * B, 2, 4, 4))
        rplan.bind(dplan.disp1.view(V, B, H, W), dplan.iconv1)
        st["dplan"], st["rplan"] = dplan, rplan
        st["epochs"] = (dplan.weights_epoch, rplan.weights_epoch)
        rplan.bind_head_output("prob", st["prob"])     # the prob head writes straight into the packed block
        st["idepth"] = rplan.out["depth"]
        self._state[key] = st
        return st

    def _launch(self, st, stream):
        lib = _lib.load()
        B, _, H, W = st["ref"].shape
        st["dplan"].run(stream)
        st["rplan"].run(stream)
        _lib.check(lib.cnm_reciprocal(st["idepth"].data_ptr(), st["depth"].data_ptr(), B * H * W, stream), "reciprocal")
        _lib.check(lib.cnm_intrinsics_inverse(st["ref_cam"].data_ptr(), st["Kinv"].data_ptr(), B, stream), "Kinv")
        _lib.check(lib.cnm_depth2normal(st["depth"].data_ptr(), st["Kinv"].data_ptr(), st["normal"].data_ptr(), None, B,
                                        H, W, self.k_size, stream), "depth2normal")

    def step(self, st):
        """Runs one step on the static buffers (graph replay when enabled)."""
        if not self.use_graph:
            self._launch(st, _lib.stream_ptr())
            return
        if st["graph"] is None:
            # warm-up outside capture (lazy module / attribute initialisation), then capture
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                self._launch(st, side.cuda_stream)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            before = _lib.launch_count()
            with torch.cuda.graph(g):
                self._launch(st, _lib.stream_ptr())
            st["graph"] = g
            st["launches_per_step"] = _lib.launch_count() - before
        st["graph"].replay()

    def launches_per_step(self, st):
        if st.get("launches_per_step") is None:
            before = _lib.launch_count()
            self._launch(st, _lib.stream_ptr())
            st["launches_per_step"] = _lib.launch_count() - before
        return st["launches_per_step"]

    def load_inputs(self, st, ref_img, src_imgs, ref_cam, src_cams, non_blocking=True):
        st["ref"].copy_(ref_img, non_blocking=non_blocking)
        st["src"].copy_(src_imgs.transpose(0, 1), non_blocking=non_blocking)
        st["ref_cam"].copy_(ref_cam, non_blocking=non_blocking)
        st["src_cam"].copy_(src_cams.transpose(0, 1), non_blocking=non_blocking)

    def static_state(self, B, V, H, W, device=None):
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        return self._prepare(B, V, H, W, dev)

    def infer_stream(self, batches, after_step=None, lanes=2):
        """Streams host batches through the pipeline with transfers overlapped with compute.

        `batches` yields (ref_img [B,3,H,W], src_imgs [B,V,3,H,W], ref_cam [B,2,4,4], src_cams [B,V,2,4,4]) CPU tensors
        (pinned memory gives asynchronous copies); yields (depth, normal, prob) as pinned CPU tensors, one tuple per
        batch, in order.  A yielded tuple is valid until `lanes` more results have been requested (its pinned buffers are then
        overwritten asynchronously by a later batch; with lanes=1 at the very next one): clone what must be kept.

        Batch i runs on lane i % lanes: every lane owns a complete buffer set + CUDA graph and its own compute
        stream, so consecutive batches overlap on the GPU -- the second lane fills what one step leaves idle (the
        partial last wave of every convolution, the launch gaps between the 75 kernels of a step, the layers with
        fewer tiles than SMs).  A copy stream uploads batch i+lanes and another downloads batch i-1 meanwhile; CUDA
        events order the streams, the host only blocks when a result is handed out.  `after_step(st)`, if given, is
        enqueued on the lane's compute stream after every graph replay (bench.py gathers outputs across ranks
        there).  lanes=1 reproduces the strictly sequential schedule."""
        if lanes < 1:
            raise ValueError("lanes must be >= 1")
        caller = torch.cuda.current_stream()
        states = stage_in = stage_out = host_out = comp = up = down = None
        ev = {}
        pending = []
        nslots = 2 * lanes

        def new_event(name, i):
            ev[(name, i)] = torch.cuda.Event()
            return ev[(name, i)]

        try:
            for i, (ref_img, src_imgs, ref_cam, src_cams) in enumerate(batches):
                if states is None:
                    B, V = src_imgs.shape[:2]
                    H, W = ref_img.shape[2:]
                    dev = next(self.depth_network.parameters()).device
                    states = [self._prepare(B, V, H, W, dev, lane) for lane in range(lanes)]
                    st0 = states[0]
                    key = ("stream_bufs", lanes)
                    if key not in st0:
                        # staging keeps the caller's [B,V,...] layout so the uploads are plain contiguous DMA transfers; the
                        # view-major transpose happens in the device-side copy into the static buffers.  Allocated once per
                        # shape: pinning host memory costs milliseconds.  Separate copy streams: an upload never queues
                        # behind a download.
                        f32 = dict(dtype=torch.float32, device=dev)
                        st0[key] = (
                            [[torch.empty((B, 3, H, W), **f32), torch.empty((B, V, 3, H, W), **f32),
                              torch.empty((B, 2, 4, 4), **f32), torch.empty((B, V, 2, 4, 4), **f32)] for _ in range(nslots)],
                            [[torch.empty_like(st0[k]) for k in ("depth", "normal", "prob")] for _ in range(nslots)],
                            [[torch.empty(st0[k].shape, dtype=torch.float32).pin_memory() for k in ("depth", "normal", "prob")]
                             for _ in range(nslots)],
                            [torch.cuda.Stream() for _ in range(lanes)], torch.cuda.Stream(), torch.cuda.Stream())
                    stage_in, stage_out, host_out, comp, up, down = st0[key]
                    for st in states:
                        if self.use_graph and st["graph"] is None:
                            self.step(st)  # capture outside the pipelined region
                    # work the caller enqueued earlier (and results of an earlier call still in flight) comes first
                    for s_ in comp + [up, down]:
                        s_.wait_stream(caller)
                if tuple(ref_img.shape) != tuple(stage_in[0][0].shape) or tuple(src_imgs.shape) != tuple(stage_in[0][1].shape):
                    raise ValueError("infer_stream: every batch of one stream must have the same shape (got %s / %s after %s / %s)"
                                     % (tuple(ref_img.shape), tuple(src_imgs.shape), tuple(stage_in[0][0].shape),
                                        tuple(stage_in[0][1].shape)))
                lane, b = i % lanes, i % nslots
                st, cs = states[lane], comp[lane]
                with torch.cuda.stream(up):  # upload batch i (after compute consumed this staging slot nslots batches ago)
                    if ("in_free", i - nslots) in ev:
                        up.wait_event(ev.pop(("in_free", i - nslots)))
                    for dst, src in zip(stage_in[b], (ref_img, src_imgs, ref_cam, src_cams)):
                        dst.copy_(src, non_blocking=True)
                    new_event("in", i).record(up)
                with torch.cuda.stream(cs):
                    cs.wait_event(ev.pop(("in", i)))
                    st["ref"].copy_(stage_in[b][0], non_blocking=True)
                    st["src"].copy_(stage_in[b][1].transpose(0, 1), non_blocking=True)
                    st["ref_cam"].copy_(stage_in[b][2], non_blocking=True)
                    st["src_cam"].copy_(stage_in[b][3].transpose(0, 1), non_blocking=True)
                    new_event("in_free", i).record(cs)
                    self.step(st)
                    if after_step is not None:
                        after_step(st)
                    if ("out_free", i - nslots) in ev:
                        cs.wait_event(ev.pop(("out_free", i - nslots)))
                    for dst, k in zip(stage_out[b], ("depth", "normal", "prob")):
                        dst.copy_(st[k], non_blocking=True)
                    new_event("out", i).record(cs)
                with torch.cuda.stream(down):  # download batch i while later batches compute
                    down.wait_event(ev.pop(("out", i)))
                    for dst, src in zip(host_out[b], stage_out[b]):
                        dst.copy_(src, non_blocking=True)
                    new_event("out_free", i).record(down)
                    done = torch.cuda.Event()
                    done.record(down)
                pending.append((done, host_out[b]))
                if len(pending) > lanes:  # hand out the oldest batch; `lanes` batches stay in flight
                    d, outs = pending.pop(0)
                    d.synchronize()
                    yield tuple(outs)
            for d, outs in pending:
                d.synchronize()
                yield tuple(outs)
        finally:
            # also when the consumer stops early (generator closed): later work on the caller's stream must not overtake
            # what is still running on the lane / copy streams
            if comp is not None:
                for s_ in comp + [up, down]:
                    caller.wait_stream(s_)

    def __call__(self, ref_img, src_imgs, ref_cam, src_cams):
        if ref_img.dim() != 4 or src_imgs.dim() != 5 or src_imgs.shape[0] != ref_img.shape[0]:
            raise ValueError("expected ref_img [B,3,H,W] and src_imgs [B,V,3,H,W]")
        B, V = src_imgs.shape[:2]
        H, W = ref_img.shape[2:]
        dev = next(self.depth_network.parameters()).device
        st = self._prepare(B, V, H, W, dev)
        self.load_inputs(st, ref_img, src_imgs, ref_cam, src_cams)
        self.step(st)
        return st["depth"].clone(), st["normal"].clone(), st["prob"].clone()
