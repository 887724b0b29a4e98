"""debug aid: bp_msm_gens_device at n rows under different sort / indexing switches must agree"""
import os, sys, subprocess, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import numpy as np, torch, hashlib
    from bulletproofs_r1cs_gadgets_b200 import api
    lib = api.load()
    out = {}
    for n in (40000, 65536, 1 << 20):
        g = api.Gens(max(n, 1 << 16))
        raw = np.frombuffer(hashlib.shake_256(b"dbg/%d" % n).digest(32 * n), dtype=np.uint8).reshape(n, 32).copy()
        raw[:, 31] &= 0x0f
        d_in = torch.from_numpy(raw).cuda(); d_out = torch.zeros(32, dtype=torch.uint8, device="cuda")
        rc = lib.bp_msm_gens_device(g._h, n, d_in.data_ptr(), d_out.data_ptr(), torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        out[n] = (rc, bytes(d_out.cpu().numpy()).hex())
        del g
    print(json.dumps(out))
else:
    for env in ({}, {"BP_B200_NO_REL": "1"}, {"BP_B200_SORT": "staged"}, {"BP_B200_SORT": "staged", "BP_B200_NO_REL": "1"}):
        e = dict(os.environ); e.update(env)
        r = subprocess.run([sys.executable, __file__, "child"], env=e, capture_output=True, text=True)
        print(env, r.stdout.strip()[-400:], r.stderr.strip()[-300:])
