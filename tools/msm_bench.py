"""BASELINE.json config 3: ristretto255 MSM microbenchmark, 2^10..2^22 generators of chain G, one MSM instance per launch.
Writes gpurun_out/msm_bench.jsonl.  GB/s = 64 algorithmic bytes per term / CUDA-event time; peak from MEASURED_PEAKS.json."""
import ctypes as C, json, os, sys, hashlib
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, ".."))
import numpy as np, torch
from bulletproofs_r1cs_gadgets_b200 import api
L = api.L
lib = api.load()
peak = json.load(open(os.path.join(HERE, "..", "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(HERE, "..", "MEASURED_PEAKS.json")) else 6650.0
max_log = int(sys.argv[1]) if len(sys.argv) > 1 else 22
out = open(os.path.join(HERE, "..", "gpurun_out", "msm_bench.jsonl"), "w")

def scalars(n, kind):
    raw = np.frombuffer(hashlib.shake_256(b"msm-bench/%d" % n).digest(32 * n), dtype=np.uint8).reshape(n, 32).copy()
    raw[:, 31] &= 0x0f  # < 2^252 < l: uniform enough for a throughput test, canonical without a big-int reduction
    if kind == "bits":
        raw[:, 1:] = 0; raw[:, 0] &= 1
    elif kind == "u64":
        raw[:, 8:] = 0
    elif kind == "half_zero":
        raw[::2] = 0
    return raw

for cap_log in ([15, max_log] if max_log > 15 else [max_log]):
    cap = 1 << cap_log
    gens = api.Gens(cap)
    for lg in range(10, cap_log + 1):
        if cap_log > 15 and lg <= 15:
            continue
        n = 1 << lg
        for kind in (["uniform", "bits", "u64", "half_zero"] if lg in (10, 15, 20) else ["uniform"]):
            d_in = torch.from_numpy(scalars(n, kind)).cuda()
            d_out = torch.zeros(32, dtype=torch.uint8, device="cuda")
            st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
            for _ in range(3):
                assert lib.bp_msm_gens_device(gens._h, n, C.c_void_p(d_in.data_ptr()), C.c_void_p(d_out.data_ptr()), st) == 0
            torch.cuda.synchronize()
            reps = 5 if lg <= 18 else 2
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                lib.bp_msm_gens_device(gens._h, n, C.c_void_p(d_in.data_ptr()), C.c_void_p(d_out.data_ptr()), st)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            gbs = 64.0 * n / (ms / 1e3) / 1e9
            row = {"log2_n": lg, "scalars": kind, "path": ("sorted buckets on the 15-bit shift table, split into sub-instances" if n >= 32768 else "direct 8-bit tables" if cap_log <= 15 else "bucket method"), "ms": round(ms, 4),
                   "Mterms_per_s": round(n / ms / 1e3, 2), "GB_per_s": round(gbs, 3), "hbm_peak_GB_per_s": peak, "frac_of_hbm": round(gbs / peak, 6),
                   "result": d_out.cpu().numpy().tobytes().hex()[:16]}
            print(json.dumps(row)); out.write(json.dumps(row) + "\n"); out.flush()
    del gens
