#!/usr/bin/env python
"""Benchmark of the hot path: batched Bulletproofs R1CS proving, Poseidon VSMT-2 depth-32 membership proofs.

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torchrun, one rank per GPU)
    python bench.py --impl reference ...                     (the CPU oracle port on all host cores)

A step = one pass of the prover over one batch of `--batch` synthetic proofs per GPU (BASELINE.json
config 5 shape: depth 32, inverse S-box, n = 18176 multipliers, N = 32768, m = 69 commitments).
Batches are streamed with two in flight (bp_prove_stream_*): a step enqueues the first phase of the next batch and
the MSM / inner-product phase of the current one, so every timed step holds one whole batch of work.
`value` = proofs/s with inputs resident in HBM; `e2e` = the same through the host-buffer C-ABI calls
(bp_prove_stream_begin_host / _finish_host: H2D of inputs and D2H of commitments + proofs inside the timed region);
`single_call_value_per_gpu` = one plain bp_prove_batch_device call on rank 0 (nothing overlapped across batches).
After the timed region every proof of the last batch goes through the cross-proof combined verifier (`verified`), and the other
BASELINE.json configurations are measured (`configs`: Poseidon 2:1 x 1024 cube / inverse, MiMC 8192 per GPU -- 32768 over 4 GPUs
under torchrun --, the MSM sweep 2^10..2^22, per-proof and combined verification).  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

METRIC = "R1CS proofs/sec (Poseidon VSMT-2 depth-32)"
UNIT = "proofs/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=int(os.environ.get("BP_BENCH_BATCH", "8192")), help="proofs per GPU per step")
    ap.add_argument("--depth", type=int, default=32)
    ap.add_argument("--cpu-sample", type=int, default=0, help="proofs in the cpu_baseline sample (0 = one per host core)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the other BASELINE configs / verifier measurements after the headline")
    ap.add_argument("--no-verify", action="store_true", help="skip the combined verification of the whole timed batch")
    return ap.parse_args()


class ClockSampler:
    """samples nvidia-smi during the timed region (B200_PROFILING.md clocks line)"""

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 7:
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def oracle_circuit(depth):
    """constraint system of the VSMT-2 circuit recorded by the independent Python oracle (root value irrelevant to the prover)"""
    from oracle import bp_pyref as R, gadgets_pyref as G, c_oracle as CO
    pp = G.PoseidonParams()
    vf = R.Verifier(R.Transcript(b"VSMT"))
    vs = [vf.commit(bytes(32)) for _ in range(1 + 2 * depth + 4)]
    G.vanilla_merkle_tree_verif_gadget(vf, depth, 0, vs[0], vs[1:1 + depth], vs[1 + depth:1 + 2 * depth], vs[1 + 2 * depth:], pp)
    blob = open(os.path.join(HERE, "bulletproofs_r1cs_gadgets_b200", "data", "poseidon_constants.bin"), "rb").read()
    CO.poseidon_set_params(blob)
    return CO.Circuit.from_cs(vf, len(vs))


def cpu_prove(circ, depth, inputs, count, nthreads):
    """times the C oracle (native witness + prove) on `count` proofs over `nthreads` host threads"""
    from oracle import c_oracle as CO
    N = 1
    while N < circ.n:
        N *= 2
    CO.lib().bpo_ensure_gens(N)
    t0 = time.time()
    status, V, proofs = CO.prove_batch(circ, count, inputs["v"][:count], inputs["v_blinding"][:count], inputs["entropy"][:count], b"VSMT", N, nthreads,
                                       witness_kind=1, depth=depth)
    dt = time.time() - t0
    assert not status.any()
    return dt, V, proofs


def run_reference(args):
    """--impl reference: the reference itself is Rust with un-vendored crates and cannot be built in this image
    (no cargo/rustc), so this arm times the oracle's C port of the same algorithm on all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from bulletproofs_r1cs_gadgets_b200 import workloads  # input generator only (pure hashlib); no GPU library is loaded
    cores = os.cpu_count() or 1
    sample = args.cpu_sample or cores
    circ = oracle_circuit(args.depth)
    wl = workloads.Vsmt2.__new__(workloads.Vsmt2)
    wl.depth, wl.cfg = args.depth, b"vsmt2/%d" % args.depth

    class _M:
        m = 1 + 2 * args.depth + 4
    wl.circuit = _M()
    inputs = wl.inputs(0, sample, with_root=False)
    times = []
    for i in range(args.warmup + args.steps):
        dt, _, _ = cpu_prove(circ, args.depth, inputs, sample, cores)
        if i >= args.warmup:
            times.append(dt)
    total = sum(times)
    value = sample * len(times) / total
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1000 * total / len(times), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64 (F_l, GF(2^255-19))",
            "data": "synthetic", "impl": "reference",
            "config": {"workload": "gadget_vsmt_2 depth-%d membership proofs, inverse S-box (n=%d, N=%d, m=%d, q=%d)"
                                   % (args.depth, circ.n, 1 << (circ.n - 1).bit_length(), circ.m, circ.q),  # the product arm's workload string
                       "proofs_per_step": sample, "note": "C port of the reference algorithm (oracle/bp_oracle.c); curve25519-dalek itself is not buildable here"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": "%d proofs per step, one thread per proof" % sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def _ptr(t):
    import ctypes as C
    return C.c_void_p(t.data_ptr()) if t is not None else None


def timed_device_batches(lib, api, gens, wl, inp_np, reps, dev, warm=3):
    """one plain bp_prove_batch_device call per rep on device-resident inputs: (ms per call, V, proofs) -- the small configurations"""
    import ctypes as C
    import torch
    circ = wl.circuit
    B = inp_np["entropy"].shape[0]
    d = {k: torch.from_numpy(a).to(dev) for k, a in inp_np.items()}
    dV = torch.empty((B, circ.m, 32), dtype=torch.uint8, device=dev)
    dP = torch.empty((B, circ.proof_len), dtype=torch.uint8, device=dev)
    dS = torch.empty(B, dtype=torch.int32, device=dev)

    def run():
        rc = lib.bp_prove_batch_device(gens._h, circ._h, C.c_uint32(B), api._buf(wl.label), C.c_size_t(len(wl.label)), _ptr(d["v"]), _ptr(d["v_blinding"]),
                                       _ptr(d["entropy"]), _ptr(d.get("aux")), _ptr(d.get("pub")), None, None, None, _ptr(dV), _ptr(dP), _ptr(dS),
                                       C.c_void_p(torch.cuda.current_stream().cuda_stream))
        assert rc == 0, rc
    for _ in range(warm):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        run()
    e1.record()
    torch.cuda.synchronize()
    assert not dS.cpu().numpy().any()
    return e0.elapsed_time(e1) / reps, dV, dP


def streamed_device_batches(lib, api, gens, wl, inp_np, steps, dev, warm=1):
    """the headline's pipeline for another circuit: two batches in flight (bp_prove_stream_begin / _finish), device-resident inputs;
    (ms per batch, V, proofs) -- both slots prove the same statements and must agree"""
    import ctypes as C
    import torch
    circ = wl.circuit
    B = inp_np["entropy"].shape[0]
    d = {k: torch.from_numpy(a).to(dev) for k, a in inp_np.items()}
    outs = [(torch.empty((B, circ.m, 32), dtype=torch.uint8, device=dev), torch.empty((B, circ.proof_len), dtype=torch.uint8, device=dev),
             torch.zeros(B, dtype=torch.int32, device=dev)) for _ in range(2)]
    lbuf = api._buf(wl.label)

    def begin(slot):
        o = outs[slot]
        rc = lib.bp_prove_stream_begin(gens._h, circ._h, C.c_int32(slot), C.c_uint32(B), lbuf, C.c_size_t(len(wl.label)), _ptr(d["v"]), _ptr(d["v_blinding"]),
                                       _ptr(d["entropy"]), _ptr(d.get("aux")), _ptr(d.get("pub")), None, None, None, _ptr(o[0]), _ptr(o[1]), _ptr(o[2]),
                                       C.c_void_p(torch.cuda.current_stream().cuda_stream))
        assert rc == 0, rc

    def finish(slot):
        rc = lib.bp_prove_stream_finish(gens._h, circ._h, C.c_int32(slot), C.c_void_p(torch.cuda.current_stream().cuda_stream))
        assert rc == 0, rc
    k = 0
    begin(0)
    for _ in range(warm):
        begin((k + 1) % 2); finish(k % 2); k += 1
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        begin((k + 1) % 2); finish(k % 2); k += 1
    e1.record()
    torch.cuda.synchronize()
    finish(k % 2)
    torch.cuda.synchronize()
    for o in outs:
        assert not o[2].cpu().numpy().any()
    assert torch.equal(outs[0][1], outs[1][1]), "the two stream slots disagree"
    return e0.elapsed_time(e1) / steps, outs[0][0], outs[0][1]


def combined_verify_device(lib, api, gens, circ, label, dV, dP, dE, dPub, dev, piece=2048):
    """cross-proof combined verification of device-resident proofs in pieces that fit one device chunk: (all combined verdicts 0, structural statuses all 0)"""
    import ctypes as C
    import torch
    B = dP.shape[0]
    dS = torch.zeros(B, dtype=torch.int32, device=dev)
    dC = torch.zeros(1, dtype=torch.int32, device=dev)
    ok = True
    for a in range(0, B, piece):
        b = min(B, a + piece)
        rc = lib.bp_verify_batch_combined_device(gens._h, circ._h, C.c_uint32(b - a), api._buf(label), C.c_size_t(len(label)), _ptr(dV[a:b]), _ptr(dP[a:b]), _ptr(dE[a:b]),
                                                 _ptr(dPub[a:b]) if dPub is not None else None, _ptr(dS[a:b]), _ptr(dC), C.c_void_p(torch.cuda.current_stream().cuda_stream))
        assert rc == 0, rc
        torch.cuda.synchronize()
        ok = ok and int(dC.item()) == 0
    return ok, not bool(dS.cpu().numpy().any())


def extra_configs(args, lib, api, workloads, parallel, gens32k, wl5, dev, world, rank, dist, hbm_peak):
    """BASELINE.json configs 2, 3, 4 and the verifier, measured after the headline (device-resident inputs, CUDA events).
    Config 4 is sharded like the headline (8192 MiMC proofs per GPU: 32768 over 4 GPUs under torchrun --gpus 4); the rest runs on rank 0."""
    import ctypes as C
    import numpy as np
    import torch
    out = {}
    g2k = api.Gens(2048)
    # config 4: gadget_mimc, 8192 proofs per GPU (reference src/gadget_mimc.rs:92-175)
    wl = workloads.Mimc(g2k)
    a, b = parallel.shard_range(rank, world, world * 8192)
    inp = wl.inputs(a, b - a)
    if world > 1:
        dist.barrier()
    ms, dV, dP = timed_device_batches(lib, api, g2k, wl, inp, 3, dev)
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    okc, oks = combined_verify_device(lib, api, g2k, wl.circuit, wl.label, dV, dP, torch.from_numpy(inp["entropy"]).to(dev), torch.from_numpy(inp["pub"]).to(dev), dev, piece=8192)
    out["mimc_8192_per_gpu"] = {"proofs_per_s": world * 8192 / ms * 1e3, "ms_per_step": ms, "proofs_per_step": world * 8192, "n": wl.circuit.n, "gpus": world,
                                "verified_combined_rank0": bool(okc and oks)}
    if rank != 0:
        return out
    # config 2: gadget_poseidon 2:1 preimage, 1024 proofs, both S-boxes (reference src/gadget_poseidon.rs:691-785)
    for sbox, nm in ((api.SBOX_CUBE, "cube"), (api.SBOX_INVERSE, "inverse")):
        wl = workloads.PoseidonHash2(g2k, sbox)
        inp = wl.inputs(0, 1024)
        ms, dV, dP = timed_device_batches(lib, api, g2k, wl, inp, 3, dev)
        okc, oks = combined_verify_device(lib, api, g2k, wl.circuit, wl.label, dV, dP, torch.from_numpy(inp["entropy"]).to(dev), torch.from_numpy(inp["pub"]).to(dev), dev, piece=1024)
        out["poseidon_1024_" + nm] = {"proofs_per_s": 1024 / ms * 1e3, "ms_per_step": ms, "n": wl.circuit.n, "verified_combined": bool(okc and oks)}
    del g2k
    # verifier on depth-32 membership proofs: per-proof (Verifier::verify, reference src/gadget_vsmt_2.rs:395) and cross-proof combined
    Bv = 1024
    inp = wl5.inputs(0, Bv, with_root=False)
    pub = wl5.roots_batch(inp["v"])
    d = {k: torch.from_numpy(v).to(dev) for k, v in inp.items() if k in ("v", "v_blinding", "entropy")}
    dV = torch.empty((Bv, wl5.circuit.m, 32), dtype=torch.uint8, device=dev)
    dP = torch.empty((Bv, wl5.circuit.proof_len), dtype=torch.uint8, device=dev)
    dS = torch.empty(Bv, dtype=torch.int32, device=dev)
    dC = torch.zeros(1, dtype=torch.int32, device=dev)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    lbl, ll = api._buf(wl5.label), C.c_size_t(len(wl5.label))
    assert lib.bp_prove_batch_device(gens32k._h, wl5.circuit._h, C.c_uint32(Bv), lbl, ll, _ptr(d["v"]), _ptr(d["v_blinding"]), _ptr(d["entropy"]), None, None,
                                     None, None, None, _ptr(dV), _ptr(dP), _ptr(dS), st) == 0
    dpub = torch.from_numpy(pub).to(dev)
    dbad = dpub.clone()
    dbad[Bv // 2, 0, 0] ^= 1

    def per_proof(p_):
        assert lib.bp_verify_batch_device(gens32k._h, wl5.circuit._h, C.c_uint32(Bv), lbl, ll, _ptr(dV), _ptr(dP), _ptr(d["entropy"]), _ptr(p_), _ptr(dS), st) == 0

    def comb(p_):
        assert lib.bp_verify_batch_combined_device(gens32k._h, wl5.circuit._h, C.c_uint32(Bv), lbl, ll, _ptr(dV), _ptr(dP), _ptr(d["entropy"]), _ptr(p_), _ptr(dS), _ptr(dC), st) == 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for name, fn in (("verify_per_proof", per_proof), ("verify_combined", comb)):
        fn(dpub)
        torch.cuda.synchronize()
        e0.record()
        fn(dpub)
        fn(dpub)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 2
        good = (not dS.cpu().numpy().any()) and (name == "verify_per_proof" or int(dC.item()) == 0)
        fn(dbad)
        torch.cuda.synchronize()
        sb = dS.cpu().numpy()
        rejected = (sb[Bv // 2] == 3 and not np.delete(sb, Bv // 2).any()) if name == "verify_per_proof" else int(dC.item()) == 3
        out[name] = {"verifications_per_s": Bv / ms * 1e3, "ms_per_step": ms, "batch": Bv, "depth": 32, "all_valid_accepted": bool(good), "one_wrong_root_rejected": bool(rejected)}
    del d, dV, dP, dpub, dbad
    # the reference's OWN test configuration (src/gadget_vsmt_2.rs:23,262-399): TreeDepth = 253, n = 143704, N = 262144 generators
    # (shift table only), m = 511; byte parity with the C oracle is tests/test_gpu.py::test_vsmt2_depth253_reference_configuration
    g18 = api.Gens(1 << 18)
    wl253 = workloads.Vsmt2(g18, depth=253)
    # No 8-bit direct tables at this capacity: four rounds over the shift table, then the level-4 generators from 5-bit fold tables
    # (41 GB, built on the first call; csrc/engine.cu ensure_fold_table).  512 proofs per batch in device chunks of 128, two batches in
    # flight like the headline: the MSM phase of a batch covers the first phase of the next (253 sequential Poseidon hashes per proof
    # in the witness program and 287k sequential RNG draws: ~2.7 s of latency whatever the batch size).  The depth-32 workspace
    # (120 GB with both stream slots) is released first.
    wl5.circuit.release_workspace()
    Bd, ch = 512, 128
    inp = wl253.inputs(0, Bd, with_root=False)
    os.environ["BP_B200_CHUNK"] = str(ch)
    try:
        ms, dV, dP = streamed_device_batches(lib, api, g18, wl253, {k: inp[k] for k in ("v", "v_blinding", "entropy")}, 2, dev, warm=1)
        ms1, _, dP1 = timed_device_batches(lib, api, g18, wl253, {k: inp[k][:ch] for k in ("v", "v_blinding", "entropy")}, 1, dev, warm=0)
        assert torch.equal(dP1, dP[:ch]), "streamed and plain calls disagree"
        pub = torch.from_numpy(wl253.roots_batch(inp["v"])).to(dev)
        okc, oks = combined_verify_device(lib, api, g18, wl253.circuit, wl253.label, dV, dP, torch.from_numpy(inp["entropy"]).to(dev), pub, dev, piece=ch)
    finally:
        del os.environ["BP_B200_CHUNK"]
    out["vsmt2_depth253_reference_config"] = {"proofs_per_s": Bd / ms * 1e3, "ms_per_step": ms, "batch": Bd, "device_chunk": ch, "n": wl253.circuit.n, "N": 1 << 18,
                                               "m": wl253.circuit.m, "proof_bytes": wl253.circuit.proof_len, "verified_combined": bool(okc and oks),
                                               "single_call_of_%d" % ch: {"proofs_per_s": ch / ms1 * 1e3, "ms": ms1,
                                                                     "note": "one plain call: bound by its exposed first phase, not by the MSMs"}}
    del dP1
    del g18, wl253, dV, dP, pub
    # config 3: ristretto MSM microbenchmark, one instance over the first 2^k generators of chain G, uniform scalars
    import hashlib
    msm = {}
    gbig = api.Gens(1 << 22)
    for lg in (10, 12, 14, 16, 18, 20, 22):
        n = 1 << lg
        gm = gens32k if n <= 32768 else gbig  # up to 2^15 rows: the capacity-32768 generators with their direct tables (what a proof uses)
        raw = np.frombuffer(hashlib.shake_256(b"msm-bench/%d" % n).digest(32 * n), dtype=np.uint8).reshape(n, 32).copy()
        raw[:, 31] &= 0x0f  # < 2^252 < l: canonical without a big-integer reduction
        d_in = torch.from_numpy(raw).to(dev)
        d_out = torch.zeros(32, dtype=torch.uint8, device=dev)
        for _ in range(2):
            assert lib.bp_msm_gens_device(gm._h, n, _ptr(d_in), _ptr(d_out), st) == 0
        torch.cuda.synchronize()
        reps = 5 if lg <= 18 else 2
        e0.record()
        for _ in range(reps):
            lib.bp_msm_gens_device(gm._h, n, _ptr(d_in), _ptr(d_out), st)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        gbs = 64.0 * n / (ms / 1e3) / 1e9
        msm["2^%d" % lg] = {"ms": round(ms, 4), "Mterms_per_s": round(n / ms / 1e3, 2), "GB_per_s": round(gbs, 3), "frac_of_hbm": round(gbs / hbm_peak, 6)}
    out["msm_sweep"] = {"unit": "64 algorithmic bytes per term (32 B scalar + 32 B compressed point)", "hbm_peak_GB_per_s": hbm_peak, "sizes": msm}
    del gbig
    return out


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    import numpy as np
    import torch
    import ctypes as C
    from bulletproofs_r1cs_gadgets_b200 import api, workloads, parallel

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    lib = api.load()

    B = args.batch
    gens = api.Gens(32768 if args.depth == 32 else 1 << 20)
    wl = workloads.Vsmt2(gens, depth=args.depth)
    circ = wl.circuit
    if gens.capacity < wl.gens_capacity:
        gens = api.Gens(wl.gens_capacity)
    # rank r proves the statements parallel.shard_range(r, world, world * B): independent proofs, no exchange during proving
    first, last = parallel.shard_range(rank, world, world * B)
    assert last - first == B
    inp = wl.inputs(first, B, with_root=False)
    pin = {k: torch.from_numpy(v).pin_memory() for k, v in inp.items() if k in ("v", "v_blinding", "entropy")}
    d = {k: t.to(dev) for k, t in pin.items()}
    m, plen = circ.m, circ.proof_len
    # two sets of output buffers: two batches are in flight (bp_prove_stream_*, include/bp_b200.h)
    outs = [(torch.empty((B, m, 32), dtype=torch.uint8, device=dev), torch.empty((B, plen), dtype=torch.uint8, device=dev),
             torch.empty((B,), dtype=torch.int32, device=dev)) for _ in range(2)]
    d_V, d_P, d_S = outs[0]
    gathered = [None]
    label = b"VSMT"
    lbuf = api._buf(label)
    ptr = _ptr

    def begin(slot):
        stream = torch.cuda.current_stream().cuda_stream
        o = outs[slot]
        rc = lib.bp_prove_stream_begin(gens._h, circ._h, C.c_int32(slot), C.c_uint32(B), lbuf, C.c_size_t(len(label)), ptr(d["v"]), ptr(d["v_blinding"]),
                                       ptr(d["entropy"]), None, None, None, None, None, ptr(o[0]), ptr(o[1]), ptr(o[2]), C.c_void_p(stream))
        if rc != 0:
            raise api.R1CSError(rc, "bp_prove_stream_begin")

    def finish(slot):
        stream = torch.cuda.current_stream().cuda_stream
        rc = lib.bp_prove_stream_finish(gens._h, circ._h, C.c_int32(slot), C.c_void_p(stream))
        if rc != 0:
            raise api.R1CSError(rc, "bp_prove_stream_finish")
        if world > 1:  # the one collective of the path: every rank receives all records {commitments || proof || status}
            gathered[0] = parallel.gather_records(parallel.pack_records(*outs[slot]), world * B)

    # One step = one whole batch: the first phase (commitments, witness program, blinding draws) of the NEXT batch is enqueued,
    # then the rest (every MSM, the inner-product argument) of the CURRENT one.  K steps hold K first phases and K second phases.
    seq = [0]
    begin(0)

    def step_device():
        k = seq[0]
        begin((k + 1) % 2)
        finish(k % 2)
        seq[0] = k + 1

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step_device()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    # Timed region: CUDA events around the DOMINANT kernel's launches only (mode 2) -- events around all ~700 launches of a step
    # cost 2.6 % of it (measured: 1298 vs 1264 proofs/s).  The per-kernel breakdown comes from one extra, untimed step below.
    api.profile_enable(0 if os.environ.get("BP_BENCH_NO_PROFILE") else 2)
    l0 = api.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step_device()
    e1.record()
    barrier()
    api.profile_enable(False)
    clocks = sampler.stop()
    launches = api.launch_count() - l0
    prof_timed = api.profile_report()
    sorted_items = prof_timed.pop("@sorted_items", (0, 0.0, 0.0))[2]
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    # one more step with events around EVERY launch: kernel_ms_per_step / kernel_time_shares (not part of `value`)
    api.profile_enable(1)
    step_device()
    barrier()
    api.profile_enable(False)
    prof = api.profile_report()
    prof.pop("@sorted_items", None)
    prof_steps = 1
    finish(seq[0] % 2)  # drain the batch whose first phase the last step enqueued
    barrier()
    for o in outs:
        assert not o[2].cpu().numpy().any(), "prover status %s" % o[2][:8]
    assert torch.equal(outs[0][1], outs[1][1]), "the two stream slots disagree"
    if world > 1:  # the gathered records hold this rank's proofs at its own index range
        gV, gP, gS = parallel.unpack_records(gathered[0], m, plen)
        assert torch.equal(gP[first:last], outs[seq[0] % 2][1]) and torch.equal(gV[first:last], outs[seq[0] % 2][0]) and not gS.cpu().numpy().any()
        del gV, gP, gS
        gathered[0] = None
    value = world * B * args.steps / (ms / 1000.0)
    # the same batch as ONE plain call (no overlap between batches), for comparison
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    d_P1 = torch.empty_like(d_P)
    g0.record()
    rc = lib.bp_prove_batch_device(gens._h, circ._h, C.c_uint32(B), lbuf, C.c_size_t(len(label)), ptr(d["v"]), ptr(d["v_blinding"]), ptr(d["entropy"]),
                                   None, None, None, None, None, ptr(outs[1][0]), ptr(d_P1), ptr(outs[1][2]), C.c_void_p(torch.cuda.current_stream().cuda_stream))
    g1.record()
    barrier()
    assert rc == 0 and torch.equal(d_P1, d_P), "streamed and plain calls disagree"
    single_call_value = B / (g0.elapsed_time(g1) / 1000.0)
    del d_P1

    # ---- every proof of the timed batch through the cross-proof combined verifier (roots hashed on the device, level by level) ----
    verified = None
    if not args.no_verify:
        tv = time.time()
        d_pub = torch.from_numpy(wl.roots_batch(inp["v"])).to(dev)
        okc, oks = combined_verify_device(lib, api, gens, circ, label, d_V, d_P, d["entropy"], d_pub, dev)
        d_bad = d_pub.clone()
        d_bad[B // 3, 0, 0] ^= 1
        piece = (B // 3) // 2048 * 2048
        badc, _ = combined_verify_device(lib, api, gens, circ, label, d_V[piece:piece + 2048], d_P[piece:piece + 2048], d["entropy"][piece:piece + 2048],
                                         d_bad[piece:piece + 2048], dev)
        verified = {"proofs": B, "combined_ok": bool(okc), "structural_status_clean": bool(oks), "wrong_root_rejected": bool(not badc),
                    "api": "bp_verify_batch_combined_device in pieces of 2048 proofs", "seconds": round(time.time() - tv, 2)}
        if world > 1:
            t = torch.tensor([1 if (okc and oks and not badc) else 0], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            verified["all_ranks_ok"] = bool(int(t.item()))
        del d_pub, d_bad

    # ---- e2e: the host-buffer C-ABI call (H2D inputs + D2H outputs inside the timed region), over the same number of steps ----
    e2e = None
    if not args.no_e2e:
        # pinned host buffers in, pinned host buffers out; every step copies its inputs H2D and its results D2H and waits for them
        hin = [pin[k].numpy() for k in ("v", "v_blinding", "entropy")]
        hout = [(torch.empty((B, m, 32), dtype=torch.uint8).pin_memory().numpy(), torch.empty((B, plen), dtype=torch.uint8).pin_memory().numpy(),
                 torch.empty((B,), dtype=torch.int32).pin_memory().numpy()) for _ in range(2)]
        ps = api.ProveStream(circ, gens, label, stream=torch.cuda.current_stream().cuda_stream)
        e2e_steps = args.steps
        ps.begin(0, *hin)
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        for k in range(e2e_steps):
            ps.begin((k + 1) % 2, *hin)
            V_h, P_h, S_h = ps.finish(k % 2, out=hout[k % 2])
        f1.record()
        barrier()
        ems = f0.elapsed_time(f1)
        ps.finish(e2e_steps % 2, out=hout[e2e_steps % 2])
        if world > 1:
            t = torch.tensor([ems], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ems = float(t.item())
        assert not S_h.any()
        assert P_h.tobytes() == d_P.cpu().numpy().tobytes(), "host-buffer and device-buffer paths disagree"
        e2e = {"value": world * B * e2e_steps / (ems / 1000.0), "unit": UNIT, "steps": e2e_steps,
               "h2d_bytes_per_step": int(sum(a.nbytes for a in hin)), "d2h_bytes_per_step": int(V_h.nbytes + P_h.nbytes + S_h.nbytes),
               "api": "bp_prove_stream_begin_host / _finish_host (pinned host buffers; H2D of the inputs and D2H of V, proofs, status every step)"}
        del hout

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(HERE, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    hbm_peak, peak_src = (peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json)") if "hbm_gbs" in peaks else (6650.0, "fallback (B200_PROFILING.md)")

    # ---- the other BASELINE configs and the verifier (all ranks take part in the sharded MiMC run) ----
    configs = None
    if not args.no_extras:
        # the depth-32 workspace stays until the depth-253 configuration needs the room (extra_configs releases it there)
        configs = extra_configs(args, lib, api, workloads, parallel, gens, wl, dev, world, rank, dist, hbm_peak)
        if world > 1:
            dist.barrier()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (algorithmic bytes: DESIGN.md section 4) ----
    n, N, k = circ.n, 1 << (circ.n - 1).bit_length(), (circ.n - 1).bit_length()
    J = min(int(os.environ.get("BP_B200_UNFOLD", "4")), k)  # unfold_rounds() in csrc/engine.cu
    # algorithmic bytes: 64 B per multiscalar term (32 B scalar + 32 B compressed point), 96 B per folded point
    terms_sorted = 2 * (2 * n + 1) + J * 2 * (N + 1)                              # A_I, S + the unfolded rounds, per proof (sorted-bucket MSM)
    terms_table = n + 1                                                           # A_O (0/1 scalars, direct tables)
    terms_bucket = sum(2 * (2 * (N >> (j + 1)) + 1) for j in range(J, k))         # L_j, R_j of the folded rounds (per-proof points)
    fold_outputs = sum(2 * (N >> (j + 1)) for j in range(J, k) if (N >> (j + 1)) > 1)
    alg_bytes = {"KBucketAccumulate": 64.0 * terms_sorted * B * args.steps, "KMsmTable": 64.0 * terms_table * B * args.steps,
                 "KMsmAccumulate": 64.0 * terms_bucket * B * args.steps,
                 "KFoldTable": 64.0 * 2 * N * B * args.steps, "KFoldGens": 96.0 * fold_outputs * B * args.steps}
    wait_ms = prof.pop("@wait_first_phase", (0, 0.0, 0.0))[1]
    total_kernel_ms = sum(v[1] for v in prof.values()) or 1.0
    shares = {kname: round(v[1] / total_kernel_ms, 4) for kname, v in sorted(prof.items(), key=lambda kv: -kv[1][1])[:8]}
    # DRAM traffic per launch of the same launch geometry from the committed ncu --set full capture (profiles/), if there is one
    traffic = {}
    for fn in ("r02_ncu_traffic.json", "r01_ncu_traffic.json"):
        try:
            traffic = json.load(open(os.path.join(HERE, "profiles", fn)))
            break
        except (OSError, ValueError):
            pass

    def roof(kname):
        # the dominant kernel is timed inside the timed region (args.steps steps); the others in the extra profiled step
        src, nsteps = (prof_timed, args.steps) if kname in prof_timed else (prof, prof_steps)
        if kname not in src or kname not in alg_bytes:
            return None
        launches_k, ms_k, _ = src[kname]
        ach = alg_bytes[kname] * nsteps / args.steps / (ms_k / 1000.0) / 1e9
        tr = traffic.get(kname, {})
        return {"kernel": kname, "bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak,
                "traffic": tr.get("dram_bytes_per_launch"), "traffic_note": tr.get("note"),
                "peak_source": peak_src, "launches": launches_k, "avg_launch_ms": ms_k / launches_k, "share_of_step": shares.get(kname)}
    dominant = max(((kn, v) for kn, v in prof.items() if kn in alg_bytes), key=lambda kv: kv[1][1])[0] if prof else None
    roofline = roof(dominant) or roof("KBucketAccumulate")
    if roofline is not None:
        roofline["timed_with"] = "CUDA events around this kernel's launches inside the timed region" if dominant in prof_timed else "CUDA events in the extra profiled step"
    # What binds these kernels is the multiplier pipe.  Numerator: the EXACT number of point additions the kernel did (length of the
    # sorted item lists, counted on the device) x 504 IMAD.WIDE per addition (7 field multiplications x 72, csrc/fe25519.h);
    # denominator: the instruction peak measured on this pool (tools/intpipe_bench.cu, profiles/r01_intpipe_peak.jsonl).
    roofline_int = None
    if "KBucketAccumulate" in prof_timed and sorted_items:
        imad_peak = 8.171e12
        try:
            for line in open(os.path.join(HERE, "profiles", "r01_intpipe_peak.jsonl")):
                r = json.loads(line)
                if "carry chains" in r.get("op", ""):
                    imad_peak = r["Tops_per_s"] * 1e12
        except (OSError, ValueError):
            pass
        t_s = prof_timed["KBucketAccumulate"][1] / 1000.0
        ach = sorted_items * 504.0 / t_s
        roofline_int = {"kernel": "KBucketAccumulate", "bound": "multiplier pipe (IMAD.WIDE.U32)", "achieved": ach / 1e12, "peak": imad_peak / 1e12, "unit": "T IMAD.WIDE/s",
                        "frac": ach / imad_peak, "additions": int(sorted_items), "G_additions_per_s": sorted_items / t_s / 1e9,
                        "peak_source": "measured instruction peak, mad.lo.cc/madc.hi.cc carry chains (profiles/r01_intpipe_peak.jsonl)",
                        "note": "additions counted exactly on the device (one per non-zero digit); 7 field multiplications x 72 IMAD.WIDE each; "
                                "the pipe also issues the kernel's IMAD.MOV / IMAD.X, which this numerator leaves out"}

    # ---- cpu_baseline: the oracle port on a bounded sample of the same workload ----
    cpu = None
    if not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        sample = args.cpu_sample or min(cores, B)
        oc = oracle_circuit(args.depth)
        dt, V_o, P_o = cpu_prove(oc, args.depth, inp, sample, cores)
        same = P_o.tobytes() == d_P[:sample].cpu().numpy().tobytes() and V_o.tobytes() == d_V[:sample].cpu().numpy().tobytes()
        cpu = {"value": sample / dt, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": "first %d proofs of the step, one thread per proof (oracle/bp_oracle.c)" % sample, "bit_exact_vs_gpu": bool(same)}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64 (F_l Montgomery 4x64, GF(2^255-19) 8x32-bit saturated)",
            "data": "synthetic",
            "config": {"workload": "gadget_vsmt_2 depth-%d membership proofs, inverse S-box (n=%d, N=%d, m=%d, q=%d)" % (args.depth, circ.n, N, circ.m, circ.q),
                       "proofs_per_gpu_per_step": B, "global_proofs_per_step": world * B,
                       "note": "BASELINE config 5 is 65536 proofs over 8 GPUs = 8192 per GPU; a step is one GPU's share",
                       "parallelism": "proofs sharded over %d GPU(s) (parallel.shard_range), one NCCL all-gather of the records {V || proof || status} per step" % world,
                       "pipeline": "two batches in flight per GPU: a step enqueues the first phase (commitments, witness program, blinding draws) of batch k+1, "
                                   "then the MSM / inner-product phase of batch k (bp_prove_stream_begin / _finish); single_call_value_per_gpu = one plain bp_prove_batch_device call on rank 0",
                       "l2": "per-step working set (GBs of scalars, digits, buckets, folded generators) exceeds the 126 MB L2"},
            "single_call_value_per_gpu": single_call_value, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "roofline_msm": roof("KBucketAccumulate"), "roofline_int": roofline_int,
            "kernel_time_shares": shares, "kernel_ms_per_step": {kn: round(v[1] / prof_steps, 3) for kn, v in sorted(prof.items(), key=lambda kv: -kv[1][1])},
            "kernel_ms_note": "from one extra step with CUDA events around every launch (that step runs ~2.6 %% slower than the timed ones); stream wait for the first phase in that step: %.3f ms" % wait_ms,
            "verified": verified, "configs": configs, "cpu_baseline": cpu}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
