"""Throughput of the 4-ary tree membership circuit (reference src/gadget_vsmt_4.rs, Poseidon 4:1): levels = 16 covers the same
2^32 leaves as BASELINE's depth-32 binary tree.  Device-resident inputs, CUDA events; proofs checked by the combined verifier."""
import ctypes as C, json, os, sys
HERE = os.path.dirname(os.path.abspath(__file__)); sys.path.insert(0, os.path.join(HERE, ".."))
import numpy as np, torch
from bulletproofs_r1cs_gadgets_b200 import api, workloads
lib = api.load()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
levels = int(sys.argv[2]) if len(sys.argv) > 2 else 16
gens = api.Gens(16384 if levels <= 16 else 32768)
wl = workloads.Vsmt4(gens, levels=levels); circ = wl.circuit
inp = wl.inputs(0, B, with_root=B <= 256)
d = {k: torch.from_numpy(inp[k]).cuda() for k in ("v", "v_blinding", "entropy", "aux", "pub")}
dV = torch.empty((B, circ.m, 32), dtype=torch.uint8, device="cuda"); dP = torch.empty((B, circ.proof_len), dtype=torch.uint8, device="cuda"); dS = torch.empty(B, dtype=torch.int32, device="cuda")
p = lambda t: C.c_void_p(t.data_ptr())
def run():
    rc = lib.bp_prove_batch_device(gens._h, circ._h, C.c_uint32(B), api._buf(wl.label), C.c_size_t(len(wl.label)), p(d["v"]), p(d["v_blinding"]), p(d["entropy"]), p(d["aux"]), p(d["pub"]),
                                   None, None, None, p(dV), p(dP), p(dS), C.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0, rc
for _ in range(2): run()
torch.cuda.synchronize()
api.profile_enable(True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); run(); run(); e1.record(); torch.cuda.synchronize()
api.profile_enable(False)
ms = e0.elapsed_time(e1) / 2
assert not dS.cpu().numpy().any()
prof = api.profile_report()
out = {"metric": "R1CS proofs/sec (Poseidon 4:1 VSMT-4, %d levels)" % levels, "value": B / ms * 1e3, "batch": B, "ms_per_step": ms, "n": circ.n, "q": circ.q, "m": circ.m,
       "kernel_ms_per_step": {k: round(v[1] / 2, 2) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][1])[:8]}}
if B <= 256:
    st, comb = circ.verify_batch_combined(gens, wl.label, dV.cpu().numpy(), dP.cpu().numpy(), inp["entropy"], pub=inp["pub"])
    out["combined_verification"] = comb; out["structural_status_clean"] = not st.any()
print(json.dumps(out))
