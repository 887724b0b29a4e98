"""Per-kernel time of the reference's own configuration (TreeDepth = 253, N = 262144): one batch with CUDA events around every launch.
usage: python tools/depth253_profile.py [batch] [chunk]   (GPU box)"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from bulletproofs_r1cs_gadgets_b200 import api, workloads  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
ch = int(sys.argv[2]) if len(sys.argv) > 2 else B
os.environ["BP_B200_CHUNK"] = str(ch)
g = api.Gens(1 << 18)
wl = workloads.Vsmt2(g, depth=253)
inp = wl.inputs(0, B, with_root=False)
for rep in range(2):
    api.profile_enable(1 if rep else 0)
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    V, P, st = wl.circuit.prove_batch(g, wl.label, inp["v"], inp["v_blinding"], inp["entropy"])
    t1.record(); torch.cuda.synchronize()
    assert not st.any()
    print("call %d: %.1f ms for %d proofs" % (rep, t0.elapsed_time(t1), B), file=sys.stderr)
api.profile_enable(0)
rep = api.profile_report()
rows = sorted(((v[1], k, v[0]) for k, v in rep.items()), reverse=True)
print(json.dumps({"batch": B, "chunk": ch, "unfold": os.environ.get("BP_B200_UNFOLD", "4"), "fold_bits": os.environ.get("BP_B200_FOLD_BITS", "auto"),
                  "kernel_ms": {k: round(ms, 2) for ms, k, _ in rows[:24]}, "free_GB": round(torch.cuda.mem_get_info()[0] / 1e9, 1)}))
