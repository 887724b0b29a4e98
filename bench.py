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
Prints ONE JSON line on rank 0.
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
            "config": {"workload": "gadget_vsmt_2 depth-%d membership proofs, inverse S-box (n=%d, m=%d)" % (args.depth, circ.n, circ.m),
                       "proofs_per_step": sample, "note": "C port of the reference algorithm (oracle/bp_oracle.c); curve25519-dalek itself is not buildable here"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": "%d proofs per step, one thread per proof" % sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    import numpy as np
    import torch
    import ctypes as C
    from bulletproofs_r1cs_gadgets_b200 import api, workloads

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
    # rank r proves proofs [r*B, (r+1)*B): independent statements, no exchange during proving
    inp = wl.inputs(rank * B, B, with_root=False)
    pin = {k: torch.from_numpy(v).pin_memory() for k, v in inp.items() if k in ("v", "v_blinding", "entropy")}
    d = {k: t.to(dev) for k, t in pin.items()}
    m, plen = circ.m, circ.proof_len
    # two sets of output buffers: two batches are in flight (bp_prove_stream_*, include/bp_b200.h)
    outs = [(torch.empty((B, m, 32), dtype=torch.uint8, device=dev), torch.empty((B, plen), dtype=torch.uint8, device=dev),
             torch.empty((B,), dtype=torch.int32, device=dev)) for _ in range(2)]
    d_V, d_P, d_S = outs[0]
    gathered = torch.empty((world * B, plen), dtype=torch.uint8, device=dev) if world > 1 else None
    label = b"VSMT"
    lbuf = api._buf(label)

    def ptr(t):
        return C.c_void_p(t.data_ptr())

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
        if world > 1:  # the one collective of the path: gather the fixed-size proof records
            dist.all_gather_into_tensor(gathered, outs[slot][1])

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
    api.profile_enable(True)
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
    prof = api.profile_report()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    finish(seq[0] % 2)  # drain the batch whose first phase the last step enqueued
    barrier()
    for o in outs:
        assert not o[2].cpu().numpy().any(), "prover status %s" % o[2][:8]
    assert torch.equal(outs[0][1], outs[1][1]), "the two stream slots disagree"
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

    # ---- e2e: the host-buffer C-ABI call (H2D inputs + D2H outputs inside the timed region) ----
    e2e = None
    if not args.no_e2e:
        # pinned host buffers in, pinned host buffers out; every step copies its inputs H2D and its results D2H and waits for them
        hin = [pin[k].numpy() for k in ("v", "v_blinding", "entropy")]
        hout = [(torch.empty((B, m, 32), dtype=torch.uint8).pin_memory().numpy(), torch.empty((B, plen), dtype=torch.uint8).pin_memory().numpy(),
                 torch.empty((B,), dtype=torch.int32).pin_memory().numpy()) for _ in range(2)]
        ps = api.ProveStream(circ, gens, label, stream=torch.cuda.current_stream().cuda_stream)
        e2e_steps = 2  # the device path above already warmed every kernel, table and workspace
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

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel and of the MSM kernel (algorithmic bytes: DESIGN.md section 4) ----
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(HERE, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    hbm_peak, peak_src = (peaks["hbm_gbs"], "measured") if "hbm_gbs" in peaks else (6650.0, "fallback")
    n, N, k = circ.n, 1 << (circ.n - 1).bit_length(), (circ.n - 1).bit_length()
    J = min(int(os.environ.get("BP_B200_UNFOLD", "4")), k)  # unfold_rounds() in csrc/engine.cu
    SB_WINDOWS = 17                                          # 15-bit windows (csrc/kernels.h)
    # algorithmic bytes: 64 B per multiscalar term (32 B scalar + 32 B compressed point), 96 B per folded point
    terms_sorted = 2 * (2 * n + 1) + J * 2 * (N + 1)                              # A_I, S + the unfolded rounds, per proof (sorted-bucket MSM)
    terms_table = n + 1                                                           # A_O (0/1 scalars, direct tables)
    terms_bucket = sum(2 * (2 * (N >> (j + 1)) + 1) for j in range(J, k))         # L_j, R_j of the folded rounds (per-proof points)
    fold_outputs = sum(2 * (N >> (j + 1)) for j in range(J, k) if (N >> (j + 1)) > 1)
    alg_bytes = {"KBucketAccumulate": 64.0 * terms_sorted * B * args.steps, "KMsmTable": 64.0 * terms_table * B * args.steps,
                 "KMsmAccumulate": 64.0 * terms_bucket * B * args.steps,
                 "KFoldTable": 64.0 * 2 * N * B * args.steps, "KFoldGens": 96.0 * fold_outputs * B * args.steps}
    total_kernel_ms = sum(v[1] for v in prof.values()) or 1.0
    shares = {kname: round(v[1] / total_kernel_ms, 4) for kname, v in sorted(prof.items(), key=lambda kv: -kv[1][1])[:8]}
    # DRAM traffic per launch of the same launch geometry from the committed ncu --set full capture (profiles/), if there is one
    traffic = {}
    try:
        traffic = json.load(open(os.path.join(HERE, "profiles", "r01_ncu_traffic.json")))
    except (OSError, ValueError):
        pass

    def roof(kname):
        if kname not in prof or kname not in alg_bytes:
            return None
        launches_k, ms_k, _ = prof[kname]
        ach = alg_bytes[kname] / (ms_k / 1000.0) / 1e9
        tr = traffic.get(kname, {})
        return {"kernel": kname, "bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak,
                "traffic": tr.get("dram_bytes_per_launch"), "traffic_note": tr.get("note"),
                "peak_source": peak_src, "launches": launches_k, "avg_launch_ms": ms_k / launches_k, "share_of_step": shares.get(kname)}
    dominant = max(((kn, v) for kn, v in prof.items() if kn in alg_bytes), key=lambda kv: kv[1][1])[0] if prof else None
    roofline = roof(dominant) or roof("KBucketAccumulate")
    # the roof that actually binds these kernels is the integer pipe: mixed point additions per second against the
    # measured peak of the same addition in isolation (profiles/r01_field_microbench_3way.jsonl, ge_madd, 32 warps/SM)
    roofline_int = None
    if "KBucketAccumulate" in prof:
        # rows that actually reach the kernel: the three equal left wires / two equal right wires of every inverse S-box share one
        # row of A_I (188 S-boxes per level), and the N - n padding rows of L_0 are one row
        rows_sorted = terms_sorted - 3 * 188 * args.depth - (N - n - 1)
        adds = float(rows_sorted) * SB_WINDOWS * B * args.steps  # one addition per non-zero digit (upper bound: zero scalars add nothing)
        ach = adds / (prof["KBucketAccumulate"][1] / 1000.0) / 1e9
        roofline_int = {"kernel": "KBucketAccumulate", "bound": "integer pipe (IMAD.WIDE)", "achieved": ach, "peak": 12.57, "unit": "G mixed additions/s",
                        "frac": ach / 12.57, "peak_source": "measured: ge_madd microbenchmark on this pool (tools/fe_bench2)",
                        "note": "additions counted as rows x 17 windows; rows with a zero scalar (a third of a_R, the padded half of round 0) contribute none, so this is an upper bound"}

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
                       "note": "BASELINE config 5 is 65536 proofs over 8 GPUs = 8192 per GPU; a step is one GPU's share", "parallelism": "proofs sharded over %d GPU(s), NCCL all-gather of proof bytes" % world,
                       "pipeline": "two batches in flight per GPU: a step enqueues the first phase (commitments, witness program, blinding draws) of batch k+1, "
                                   "then the MSM / inner-product phase of batch k (bp_prove_stream_begin / _finish); single_call_value_per_gpu = one plain bp_prove_batch_device call on rank 0",
                       "l2": "per-step working set (GBs of scalars, digits, buckets, folded generators) exceeds the 126 MB L2"},
            "single_call_value_per_gpu": single_call_value, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "roofline_msm": roof("KBucketAccumulate"), "roofline_int": roofline_int,
            "kernel_time_shares": shares, "kernel_ms_per_step": {kn: round(v[1] / args.steps, 3) for kn, v in sorted(prof.items(), key=lambda kv: -kv[1][1])[:12]},
            "cpu_baseline": cpu}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
