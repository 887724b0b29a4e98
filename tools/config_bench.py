"""Throughput of BASELINE.json's smaller configurations on one B200 (device-resident inputs, CUDA events):
config 1 gadget_bound_check (64-bit), config 2 gadget_poseidon 2:1 (1024 proofs, cube and inverse S-box),
config 4 gadget_mimc (8192 proofs = one GPU's share of 32768 over 4 GPUs).  Bit-exactness of these shapes against the C oracle
is what tests/test_gpu.py checks; here every proof of the batch is re-checked by the combined verifier."""
import ctypes as C, json, os, sys
HERE = os.path.dirname(os.path.abspath(__file__)); sys.path.insert(0, os.path.join(HERE, ".."))
import numpy as np, torch
from bulletproofs_r1cs_gadgets_b200 import api, workloads
lib = api.load()
gens = api.Gens(2048)
p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None

def bench(name, wl, B, gen_inputs):
    circ = wl.circuit
    inp = gen_inputs(B)
    d = {k: torch.from_numpy(a).cuda() for k, a in inp.items()}
    dV = torch.empty((B, circ.m, 32), dtype=torch.uint8, device="cuda"); dP = torch.empty((B, circ.proof_len), dtype=torch.uint8, device="cuda"); dS = torch.empty(B, dtype=torch.int32, device="cuda")
    def run():
        rc = lib.bp_prove_batch_device(gens._h, circ._h, C.c_uint32(B), api._buf(wl.label), C.c_size_t(len(wl.label)), p(d["v"]), p(d["v_blinding"]), p(d["entropy"]), p(d.get("aux")), p(d.get("pub")),
                                       None, None, None, p(dV), p(dP), p(dS), C.c_void_p(torch.cuda.current_stream().cuda_stream))
        assert rc == 0, rc
    for _ in range(3): run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 3
    e0.record()
    for _ in range(reps): run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    assert not dS.cpu().numpy().any()
    st, comb = circ.verify_batch_combined(gens, wl.label, dV.cpu().numpy(), dP.cpu().numpy(), inp["entropy"], pub=inp.get("pub"))
    print(json.dumps({"config": name, "batch": B, "n": circ.n, "q": circ.q, "m": circ.m, "ms_per_step": round(ms, 3), "proofs_per_s": round(B / ms * 1e3, 1),
                      "combined_verification": comb, "structural_status_clean": bool(not st.any())}), flush=True)

wl = workloads.BoundCheck(gens)
bench("1: gadget_bound_check 64-bit (n=128)", wl, 1024, lambda B: wl.inputs(0, B))
for sbox, nm in ((api.SBOX_CUBE, "cube"), (api.SBOX_INVERSE, "inverse")):
    wl = workloads.PoseidonHash2(gens, sbox)
    bench("2: gadget_poseidon 2:1 %s S-box" % nm, wl, 1024, lambda B, wl=wl: wl.inputs(0, B))
wl = workloads.Mimc(gens)
bench("4: gadget_mimc 322 rounds, one GPU's share of 32768", wl, 8192, lambda B: wl.inputs(0, B))
