"""Multi-GPU plumbing of the hot path: proofs of a batch are independent, so rank r proves the index range
shard_range(r, world, total) with no exchange during proving; the only collective is the gather of the
fixed-size proof records {commitments || proof || status} at the end (NCCL over NVLink on GPUs,
gloo in the CPU tests).  `prove_batch_sharded` is the product-level entry; bench.py uses the same shard / pack / gather
functions around the streaming device-pointer calls."""
import numpy as np
import torch
import torch.distributed as dist


def shard_range(rank, world, total):
    """contiguous, balanced partition of [0, total): the first total % world ranks get one extra proof"""
    base, extra = divmod(total, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def gather_records(local, total, group=None):
    """local: uint8 tensor [count_r, record_len] of this rank's proofs (count_r from shard_range).
    Returns the [total, record_len] tensor of all proofs in global index order on every rank."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return local
    rank = dist.get_rank(group)
    base, extra = divmod(total, world)
    width = base + (1 if extra else 0)
    padded = torch.zeros((width, local.shape[1]), dtype=local.dtype, device=local.device)
    padded[: local.shape[0]] = local
    out = torch.empty((world * width, local.shape[1]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, padded, group=group)
    parts = []
    for r in range(world):
        a, b = shard_range(r, world, total)
        parts.append(out[r * width: r * width + (b - a)])
    del rank
    return torch.cat(parts, dim=0)


def record_len(m, proof_len):
    """bytes of one proof record: m compressed commitments, the proof, the int32 status (little endian)"""
    return 32 * m + proof_len + 4


def pack_records(V, proofs, status):
    """V [count, m, 32] uint8, proofs [count, proof_len] uint8, status [count] int32 (torch tensors, same device) -> [count, record_len]"""
    count = proofs.shape[0]
    if count == 0:
        return torch.zeros((0, V.shape[1] * 32 + proofs.shape[1] + 4), dtype=torch.uint8, device=proofs.device)
    st = status.to(torch.int32).contiguous().view(torch.uint8).view(count, 4)
    return torch.cat([V.reshape(count, -1), proofs, st], dim=1)


def unpack_records(records, m, proof_len):
    count = records.shape[0]
    if count == 0:
        return (torch.zeros((0, m, 32), dtype=torch.uint8, device=records.device), torch.zeros((0, proof_len), dtype=torch.uint8, device=records.device),
                torch.zeros((0,), dtype=torch.int32, device=records.device))
    V = records[:, : 32 * m].reshape(count, m, 32)
    proofs = records[:, 32 * m: 32 * m + proof_len]
    status = records[:, 32 * m + proof_len:].contiguous().view(torch.int32).view(count)
    return V, proofs, status


def prove_batch_sharded(circuit, gens, label, total, inputs_for_range, device=None, group=None):
    """Proves statements 0 .. total-1 over every rank of the process group: rank r builds the inputs of its own index range
    (`inputs_for_range(first, count)` -> dict(v, v_blinding, entropy[, aux][, pub]) of uint8 arrays, as workloads.*.inputs),
    proves them with `circuit.prove_batch`, and the ranks exchange the records once.  Returns (V, proofs, status) of ALL `total`
    statements in global order, as torch tensors on `device` (default: cuda when the backend is nccl, else cpu), on every rank.
    Mirrors what a batch of the reference's prove() calls would produce one after the other (src/gadget_vsmt_2.rs:289-350)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if (dist.is_initialized() and dist.get_backend(group) == "nccl") else torch.device("cpu")
    a, b = shard_range(rank, world, total)
    m, plen = circuit.m, circuit.proof_len
    if b > a:
        inp = inputs_for_range(a, b - a)
        V, proofs, status = circuit.prove_batch(gens, label, inp["v"], inp["v_blinding"], inp["entropy"], aux=inp.get("aux"), pub=inp.get("pub"))
    else:
        V, proofs, status = np.zeros((0, m, 32), np.uint8), np.zeros((0, plen), np.uint8), np.zeros((0,), np.int32)
    local = pack_records(torch.from_numpy(np.ascontiguousarray(V)), torch.from_numpy(np.ascontiguousarray(proofs)),
                         torch.from_numpy(np.ascontiguousarray(status))).to(device)
    return unpack_records(gather_records(local, total, group=group), m, plen)
