"""Throughput of batched verification (Verifier::verify) for VSMT-2 depth-32 proofs: prove a batch once, then time bp_verify_batch_device."""
import ctypes as C, json, os, sys
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, ".."))
import numpy as np, torch
from bulletproofs_r1cs_gadgets_b200 import api, workloads
lib = api.load()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
gens = api.Gens(32768)
wl = workloads.Vsmt2(gens, depth=32)
inp = wl.inputs(0, B, with_root=False)
roots = wl.inputs(0, 2)["pub"]
V, P, st = wl.circuit.prove_batch(gens, wl.label, inp["v"], inp["v_blinding"], inp["entropy"])
assert not st.any()
pub = np.zeros((B, 1, 32), np.uint8); pub[:2] = roots   # only the first two roots are real: the others must be rejected
d = {k: torch.from_numpy(a).cuda() for k, a in dict(V=V, P=P, ent=inp["entropy"], pub=pub).items()}
dS = torch.zeros(B, dtype=torch.int32, device="cuda")
p = lambda t: C.c_void_p(t.data_ptr())
def run():
    rc = lib.bp_verify_batch_device(gens._h, wl.circuit._h, C.c_uint32(B), api._buf(wl.label), C.c_size_t(len(wl.label)), p(d["V"]), p(d["P"]), p(d["ent"]), p(d["pub"]), p(dS),
                                    C.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0
for _ in range(2): run()
torch.cuda.synchronize()
api.profile_enable(True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); run(); e1.record(); torch.cuda.synchronize()
api.profile_enable(False)
ms = e0.elapsed_time(e1)
s = dS.cpu().numpy()
prof = api.profile_report()
result = {"metric": "R1CS verifications/sec (Poseidon VSMT-2 depth-32)", "value": B / ms * 1e3, "batch": B, "ms": ms, "accepted_first_two": s[:2].tolist(),
                  "rejected_rest": bool((s[2:] == 3).all()), "kernel_ms": {k: round(v[1], 2) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][1])[:8]}}
print(json.dumps(result))

# cross-proof batched verification of the same batch with the true roots (all valid), then with one wrong root
pub_ok = wl.inputs(0, B)["pub"]
d_ok = torch.from_numpy(pub_ok).cuda(); d_bad = d_ok.clone(); d_bad[B // 2, 0, 0] ^= 1
dC = torch.zeros(1, dtype=torch.int32, device="cuda")
def runc(dpub):
    rc = lib.bp_verify_batch_combined_device(gens._h, wl.circuit._h, C.c_uint32(B), api._buf(wl.label), C.c_size_t(len(wl.label)), p(d["V"]), p(d["P"]), p(d["ent"]), p(dpub), p(dS), p(dC),
                                             C.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0, rc
for _ in range(2): runc(d_ok)
torch.cuda.synchronize()
api.profile_enable(True)
e0.record(); runc(d_ok); e1.record(); torch.cuda.synchronize()
api.profile_enable(False)
ms = e0.elapsed_time(e1); ok = int(dC.item()); st_ok = not dS.cpu().numpy().any()
prof = api.profile_report()
runc(d_bad); torch.cuda.synchronize(); bad = int(dC.item())
print(json.dumps({"metric": "R1CS verifications/sec, cross-proof combined check (Poseidon VSMT-2 depth-32)", "value": B / ms * 1e3, "batch": B, "ms": ms,
                  "combined_all_valid": ok, "structural_status_clean": st_ok, "combined_with_one_wrong_root": bad,
                  "kernel_ms": {k: round(v[1], 2) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][1])[:8]}}))
