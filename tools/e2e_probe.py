"""Times the host-buffer call (bp_prove_batch) against the device-buffer call on the same inputs, several repetitions each."""
import ctypes as C, os, sys, time
HERE = os.path.dirname(os.path.abspath(__file__)); sys.path.insert(0, os.path.join(HERE, ".."))
import numpy as np, torch
from bulletproofs_r1cs_gadgets_b200 import api, workloads
lib = api.load()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
gens = api.Gens(32768); wl = workloads.Vsmt2(gens, depth=32); circ = wl.circuit
inp = wl.inputs(0, B, with_root=False)
d = {k: torch.from_numpy(inp[k]).cuda() for k in ("v", "v_blinding", "entropy")}
dV = torch.empty((B, circ.m, 32), dtype=torch.uint8, device="cuda"); dP = torch.empty((B, circ.proof_len), dtype=torch.uint8, device="cuda"); dS = torch.empty(B, dtype=torch.int32, device="cuda")
p = lambda t: C.c_void_p(t.data_ptr())
def dev():
    assert lib.bp_prove_batch_device(gens._h, circ._h, C.c_uint32(B), api._buf(wl.label), C.c_size_t(len(wl.label)), p(d["v"]), p(d["v_blinding"]), p(d["entropy"]), None, None, None, None, None, p(dV), p(dP), p(dS), C.c_void_p(torch.cuda.current_stream().cuda_stream)) == 0
dev(); torch.cuda.synchronize()
for name, fn in (("device", dev), ("host", lambda: circ.prove_batch(gens, wl.label, inp["v"], inp["v_blinding"], inp["entropy"])), ("device", dev), ("host", lambda: circ.prove_batch(gens, wl.label, inp["v"], inp["v_blinding"], inp["entropy"]))):
    torch.cuda.synchronize(); t0 = time.time(); fn(); torch.cuda.synchronize(); print(name, round((time.time() - t0) * 1e3, 1), "ms", flush=True)
