"""world_size-2 gloo test of the data-parallel plumbing (sharding + output gather + metric reduce)."""
import os
import socket
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, total, q):
    sys.path.insert(0, ROOT)
    from cnmnet_b200.dist import OutputGatherer, gather_outputs, reduce_metrics, shard_range, unpack_outputs
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    a, b = shard_range(total, rank, world)
    full = torch.arange(total * 6, dtype=torch.float32).reshape(total, 2, 3)
    got = gather_outputs(full[a:b].clone(), total)
    m = reduce_metrics(torch.tensor([float(b - a), 1.0]))
    ok = torch.equal(got, full) and m.tolist() == [float(total), float(world)]
    # one packed all-gather per step (depth | normal | prob of B frames per rank), two steps through the two slots
    B, H, W = 2, 3, 4
    g = OutputGatherer(5 * B * H * W, "cpu")
    for step in range(3):
        packed = torch.arange(5 * B * H * W, dtype=torch.float32) + 1000.0 * rank + 7.0 * step
        depth, normal, prob = unpack_outputs(g.result(g.submit(packed)), B, H, W)
        ok = ok and depth.shape == (world * B, H, W) and normal.shape == (world * B, 3, H, W) and prob.shape == (world * B, 1, H, W)
        for r in range(world):
            want = torch.arange(5 * B * H * W, dtype=torch.float32) + 1000.0 * r + 7.0 * step
            ok = ok and torch.equal(depth[r * B:(r + 1) * B].reshape(-1), want[:B * H * W])
            ok = ok and torch.equal(normal[r * B:(r + 1) * B].reshape(-1), want[B * H * W:4 * B * H * W])
            ok = ok and torch.equal(prob[r * B:(r + 1) * B].reshape(-1), want[4 * B * H * W:])
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _run(total):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, total, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok in res), res


def test_gather_even():
    _run(10)


def test_gather_ragged():
    _run(5)
