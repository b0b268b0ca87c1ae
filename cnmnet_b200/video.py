"""Video-stream front end: turns one posed RGB sequence into the reference's frame tuples and streams them through
CNMPipeline.infer_stream.

Reference behaviour mirrored here (eval.py): every evaluated frame `index` is a reference view whose sources are the
frames at fixed offsets -- (+10, -10) for 3 views (eval.py:408-415), (+10, -10, +5, -5) for 5 views (eval.py:581-592),
(+10, -10, +5, -5, +20, -20) for 7 views (eval.py:817-830) -- and only every 3rd (3 / 5 views) or 9th (7 views) frame
is evaluated.  The source order matters: even positions form view group 1, odd positions group 2 of the refine input
(eval.py:659-663, 920-929), which is exactly the [B, V, ...] order CNMPipeline takes.

Divergence, on purpose: the reference's 5- and 7-view loops start at index 10 although they read `index - 20`
(eval.py:581, 817), so their first frames silently wrap to the END of the sequence through Python's negative
indexing.  Here a frame is used as a reference view only if all its sources exist.
"""
import torch

VIEW_OFFSETS = {3: (10, -10), 5: (10, -10, 5, -5), 7: (10, -10, 5, -5, 20, -20)}
EVERY = {3: 3, 5: 3, 7: 9}


def reference_tuples(num_frames, views=3, every=None, offsets=None):
    """-> list of (ref index, (source indices)) in evaluation order."""
    offs = tuple(VIEW_OFFSETS[views] if offsets is None else offsets)
    if not offs or len(offs) % 2:
        raise ValueError("the refine network takes two view groups: an even number of source offsets is needed")
    every = EVERY.get(len(offs) + 1, 1) if every is None else int(every)
    lo, hi = max(0, -min(offs)), num_frames - max(0, max(offs))
    return [(i, tuple(i + o for o in offs)) for i in range(lo, hi) if i % every == 0]


class VideoDepthStream:
    """Streams a posed sequence through a CNMPipeline.

        stream = VideoDepthStream(pipe, batch=5, views=3)
        for index, depth, normal, prob in stream.run(frames, cams):   # frames [T,3,H,W], cams [T,2,4,4] (CPU)
            ...

    Frames are gathered into pinned host batches of `batch` tuples (the last batch is padded by repeating its final
    tuple and the padding is dropped from the output), uploads / compute / downloads overlap as in infer_stream.
    Results are CPU tensors, one (index, depth [H,W], normal [3,H,W], prob [1,H,W]) per evaluated frame, in order."""

    def __init__(self, pipeline, batch=5, views=3, every=None, offsets=None):
        self.pipe, self.batch, self.views, self.every, self.offsets = pipeline, int(batch), views, every, offsets

    def batches(self, frames, cams):
        """-> (list of index lists, generator of pinned (ref, srcs, ref_cam, src_cams) batches)."""
        if frames.dim() != 4 or frames.shape[1] != 3 or cams.shape != (frames.shape[0], 2, 4, 4):
            raise ValueError("expected frames [T,3,H,W] and cams [T,2,4,4]")
        tuples = reference_tuples(frames.shape[0], self.views, self.every, self.offsets)
        B = self.batch
        groups = [tuples[i:i + B] for i in range(0, len(tuples), B)]
        index_lists = [[t[0] for t in g] for g in groups]

        def gen():
            T, _, H, W = frames.shape
            V = len(tuples[0][1]) if tuples else 0
            # rotating pinned staging sets: infer_stream (two lanes) hands a result out -- and asks for the next batch --
            # only after the batch three positions earlier has been consumed; one more set for margin
            pin = [[torch.empty(s, dtype=torch.float32).pin_memory() if torch.cuda.is_available() else
                    torch.empty(s, dtype=torch.float32)
                    for s in ((B, 3, H, W), (B, V, 3, H, W), (B, 2, 4, 4), (B, V, 2, 4, 4))] for _ in range(4)]
            for n, g in enumerate(groups):
                g = list(g) + [g[-1]] * (B - len(g))
                ref_idx = torch.tensor([t[0] for t in g])
                src_idx = torch.tensor([t[1] for t in g])
                bufs = pin[n % 4]
                torch.index_select(frames, 0, ref_idx, out=bufs[0])
                bufs[1].copy_(frames[src_idx.reshape(-1)].view(B, V, 3, H, W))
                torch.index_select(cams, 0, ref_idx, out=bufs[2])
                bufs[3].copy_(cams[src_idx.reshape(-1)].view(B, V, 2, 4, 4))
                yield tuple(bufs)

        return index_lists, gen()

    def run(self, frames, cams):
        index_lists, gen = self.batches(frames.float(), cams.float())
        for idx, (depth, normal, prob) in zip(index_lists, self.pipe.infer_stream(gen)):
            for j, i in enumerate(idx):   # padding rows (j >= len(idx)) are dropped
                yield i, depth[j].clone(), normal[j].clone(), prob[j].clone()
