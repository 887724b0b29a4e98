"""Per-kernel time of the MSM microbenchmark entry (bp_msm_gens_device) at a few sizes.  usage: python tools/msm_profile.py (GPU box)"""
import hashlib
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
from bulletproofs_r1cs_gadgets_b200 import api  # noqa: E402

lib = api.load()
dev = torch.device("cuda", 0)
sizes = [int(x) for x in sys.argv[1:]] or [10, 14, 16, 20, 22]
g32 = api.Gens(32768)
gbig = api.Gens(1 << max(max(sizes), 16))
st = torch.cuda.current_stream().cuda_stream
for lg in sizes:
    n = 1 << lg
    gm = g32 if n <= 32768 else gbig
    raw = np.frombuffer(hashlib.shake_256(b"msm-bench/%d" % n).digest(32 * n), dtype=np.uint8).reshape(n, 32).copy()
    raw[:, 31] &= 0x0f
    d_in = torch.from_numpy(raw).to(dev)
    d_out = torch.zeros(32, dtype=torch.uint8, device=dev)
    for _ in range(2):
        assert lib.bp_msm_gens_device(gm._h, n, d_in.data_ptr(), d_out.data_ptr(), st) == 0
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    e0.record()
    for _ in range(reps):
        lib.bp_msm_gens_device(gm._h, n, d_in.data_ptr(), d_out.data_ptr(), st)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    api.profile_enable(1)
    lib.bp_msm_gens_device(gm._h, n, d_in.data_ptr(), d_out.data_ptr(), st)
    torch.cuda.synchronize()
    api.profile_enable(0)
    rep = api.profile_report()
    print(json.dumps({"log2_n": lg, "ms": round(ms, 4), "Mterms_per_s": round(n / ms / 1e3, 1),
                      "kernels": {k: [v[0], round(v[1], 4)] for k, v in sorted(rep.items(), key=lambda kv: -kv[1][1])}}))
