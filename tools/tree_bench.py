"""Device-side sparse Merkle tree: batched update / get / witness-row throughput at depth 32 with the full
Poseidon 4+140+4 inverse permutation, next to the oracle's C hash on one host core (the reference does depth hashes per
update, one key at a time: src/gadget_vsmt_2.rs:63-98).  Prints one JSON line per measurement."""
import json, os, random, sys, time
HERE = os.path.dirname(os.path.abspath(__file__)); sys.path.insert(0, os.path.join(HERE, ".."))
import numpy as np, torch
from bulletproofs_r1cs_gadgets_b200 import api, trees
from oracle import c_oracle as CO, tree_pyref as TP
api.load()
depth = 32
K = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 16
pp = api.PoseidonParams()
rnd = random.Random(7)
def rand_bytes(n, seed):
    r = np.random.default_rng(seed).integers(0, 256, size=(n, 32), dtype=np.uint8); r[:, 31] &= 0x0f; return r
def timed(f):
    torch.cuda.synchronize(); t = time.perf_counter(); out = f(); torch.cuda.synchronize(); return out, time.perf_counter() - t
# hash kernel alone
for n in (1 << 12, 1 << 16, 1 << 18):
    xl, xr = rand_bytes(n, 1), rand_bytes(n, 2)
    trees.poseidon_hash_2_batch(pp, xl[:64], xr[:64])
    api.profile_enable(True)
    out, dt = timed(lambda: trees.poseidon_hash_2_batch(pp, xl, xr))
    api.profile_enable(False); prof = api.profile_report()
    kms = prof["KPoseidonHash2Batch"][1]
    print(json.dumps({"op": "poseidon_hash_2_batch", "count": n, "kernel_ms": round(kms, 3), "hashes_per_s_kernel": round(n / kms * 1e3), "hashes_per_s_host_buffers": round(n / dt)}))
# CPU: the oracle's C hash, one core
CO.build(); CO.poseidon_set_params(open(os.path.join(HERE, "..", "bulletproofs_r1cs_gadgets_b200", "data", "poseidon_constants.bin"), "rb").read())
h2 = TP.c_oracle_hash2(CO, 1)
t = time.perf_counter(); x = 1
for i in range(200): x = h2(x, i)
cpu_hash_s = (time.perf_counter() - t) / 200
assert int.from_bytes(trees.poseidon_hash_2_batch(pp, [5], [6])[0].tobytes(), "little") == h2(5, 6)
print(json.dumps({"op": "cpu_oracle_hash2", "cores": 1, "hashes_per_s": round(1 / cpu_hash_s, 1), "updates_per_s_depth32": round(1 / (cpu_hash_s * depth), 2)}))
# batched updates into an empty and into a populated tree
tree = trees.DeviceVsmt2(pp, depth)
for rep in range(2):
    keys = [rnd.randrange(2 ** depth) for _ in range(K)]; vals = rand_bytes(K, 10 + rep)
    n0 = tree.num_nodes
    api.profile_enable(True)
    root, dt = timed(lambda: tree.update_batch(keys, vals))
    api.profile_enable(False); prof = api.profile_report()
    hk = prof["KTreeHashLevel"]
    print(json.dumps({"op": "update_batch", "tree_nodes_before": n0, "count": K, "s": round(dt, 4), "updates_per_s": round(K / dt), "hashes": int(hk[2]), "hash_kernel_ms": round(hk[1], 2),
                      "hash_launches": hk[0], "hashes_per_s_kernel": round(hk[2] / hk[1] * 1e3), "vs_cpu_core": round(K / dt * cpu_hash_s * depth, 1), "tree_nodes_after": tree.num_nodes}))
keys_q = keys
(lv, pf), dt = timed(lambda: tree.get_batch(keys_q))
assert lv.tobytes() == vals.tobytes() or len(set(keys_q)) < K
print(json.dumps({"op": "get_batch", "count": K, "s": round(dt, 4), "gets_per_s": round(K / dt)}))
d_idx = torch.tensor(keys_q, dtype=torch.int64, device="cuda")
d_v = torch.zeros((K, 2 * depth + 5, 32), dtype=torch.uint8, device="cuda"); d_pub = torch.zeros((K, 1, 32), dtype=torch.uint8, device="cuda")
tree.witness_rows_device(d_idx, d_v, d_pub)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record(); tree.witness_rows_device(d_idx, d_v, d_pub, torch.cuda.current_stream().cuda_stream); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print(json.dumps({"op": "witness_rows_device", "count": K, "ms": round(ms, 3), "proof_inputs_per_s": round(K / ms * 1e3), "GB_per_s_written": round(K * (2 * depth + 6) * 32 / ms / 1e6, 1)}))
# one sampled path checked by the oracle
path = [int.from_bytes(pf[0, j].tobytes(), "little") for j in range(depth)]
chk = TP.VanillaSparseMerkleTree.__new__(TP.VanillaSparseMerkleTree); chk.depth, chk.hash2, chk.root = depth, h2, tree.root
assert chk.verify_proof(keys_q[0], int.from_bytes(lv[0].tobytes(), "little"), path)
print(json.dumps({"check": "device path verifies under the oracle's verify_proof", "ok": True}))
