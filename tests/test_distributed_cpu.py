"""Multi-GPU host logic on CPU: world_size 2 over gloo.  Proofs shard by index with no exchange; one gather at the end."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
from bulletproofs_r1cs_gadgets_b200.parallel import shard_range, gather_records  # noqa: E402


def test_shard_range_partitions():
    for total in (0, 1, 5, 8, 65536, 65537):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(r, world, total) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    assert [shard_range(r, 8, 65536) for r in (0, 7)] == [(0, 8192), (57344, 65536)]


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, total, reclen, out_q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    a, b = shard_range(rank, world, total)
    # fake "proof records": record i is filled with a function of its global index
    local = torch.stack([torch.full((reclen,), (7 * i + 3) % 251, dtype=torch.uint8) for i in range(a, b)]) if b > a else torch.zeros((0, reclen), dtype=torch.uint8)
    allr = gather_records(local, total)
    ok = allr.shape == (total, reclen) and all(int(allr[i, 0]) == (7 * i + 3) % 251 and int(allr[i, -1]) == (7 * i + 3) % 251 for i in range(total))
    out_q.put((rank, bool(ok)))
    dist.destroy_process_group()


@pytest.mark.parametrize("total", [8, 7])
def test_gather_records_gloo_world2(total):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, total, 1472, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]
