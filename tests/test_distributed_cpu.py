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
from bulletproofs_r1cs_gadgets_b200.parallel import shard_range, gather_records, prove_batch_sharded, pack_records, unpack_records, record_len  # noqa: E402


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


class _FakeCircuit:
    """stands in for api.Circuit on the CPU: `prove_batch` is a pure function of the inputs (no GPU here); the sharded entry
    point must hand every statement to exactly one rank and return all records in global order, V and status included"""
    m, proof_len = 3, 96

    def prove_batch(self, gens, label, v, v_blinding, entropy, aux=None, pub=None):
        import numpy as np
        B = entropy.shape[0]
        V = (v.astype(np.uint16) + v_blinding).astype(np.uint8)
        proofs = np.repeat(entropy, 3, axis=1)
        status = (entropy[:, 0] % 5 == 0).astype(np.int32) * 3
        return V, proofs, status


def _inputs(first, count):
    import numpy as np
    idx = np.arange(first, first + count, dtype=np.int64)
    v = np.zeros((count, 3, 32), np.uint8); v[:] = (idx % 251)[:, None, None]
    vb = np.zeros((count, 3, 32), np.uint8); vb[:] = ((3 * idx) % 7)[:, None, None]
    ent = np.zeros((count, 32), np.uint8); ent[:] = ((11 * idx + 1) % 253)[:, None]
    return dict(v=v, v_blinding=vb, entropy=ent)


def _sharded_worker(rank, world, port, total, out_q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    V, P, S = prove_batch_sharded(_FakeCircuit(), None, b"x", total, _inputs)
    ref = _FakeCircuit().prove_batch(None, b"x", **_inputs(0, total))
    ok = V.numpy().tobytes() == ref[0].tobytes() and P.numpy().tobytes() == ref[1].tobytes() and S.numpy().tolist() == ref[2].tolist()
    out_q.put((rank, bool(ok)))
    dist.destroy_process_group()


@pytest.mark.parametrize("total", [9, 1])
def test_prove_batch_sharded_gloo_world2(total):
    """total = 1: the second rank owns nothing and still takes part in the gather"""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_sharded_worker, args=(r, 2, port, total, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]


def test_record_round_trip():
    V = torch.arange(2 * 3 * 32, dtype=torch.int64).remainder(256).to(torch.uint8).view(2, 3, 32)
    P = torch.arange(2 * 96, dtype=torch.int64).remainder(251).to(torch.uint8).view(2, 96)
    S = torch.tensor([0, 3], dtype=torch.int32)
    rec = pack_records(V, P, S)
    assert rec.shape == (2, record_len(3, 96))
    V2, P2, S2 = unpack_records(rec, 3, 96)
    assert torch.equal(V, V2) and torch.equal(P, P2) and torch.equal(S, S2)
