"""N>1 host path on CPU: world_size-2 gloo process group exercising the sample sharding and the single
all-gather of final latents (bagel_b200/dist.py). The GPU build uses the same code over NCCL."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from bagel_b200 import dist as bdist


def test_shard_range_partitions_everything():
    for n in (0, 1, 7, 8, 16, 33):
        for world in (1, 2, 3, 8):
            spans = [bdist.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # uniform latents: rank r owns samples r*2, r*2+1 of a global batch of 4
        prompts = list(range(4))
        mine = bdist.shard_list(prompts)
        local = torch.stack([torch.full((5, 64), float(s)) for s in mine], 0)
        full = bdist.gather_latents(local)
        ok1 = full.shape == (4, 5, 64) and all(bool((full[i] == float(i)).all()) for i in range(4))
        # ragged latents
        counts = [[3, 1], [2, 4]]
        loc = [torch.full((t, 64), float(10 * rank + j)) for j, t in enumerate(counts[rank])]
        rag = bdist.gather_ragged_latents(loc, counts)
        want = [(3, 0.0), (1, 1.0), (2, 10.0), (4, 11.0)]
        ok2 = len(rag) == 4 and all(r.shape[0] == t and bool((r == v).all()) for r, (t, v) in zip(rag, want))
        q.put((rank, ok1, ok2))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_gather_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=90) for _ in procs]
    for p in procs:
        p.join(30)
    assert sorted(r[0] for r in res) == [0, 1]
    assert all(r[1] and r[2] for r in res), res
