import sys, time, random, os, ctypes as C
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch
from oracle import bp_pyref as R, gadgets_pyref as G, c_oracle as CO
from bulletproofs_r1cs_gadgets_b200 import api
lib = api.load()
L = R.L
rnd = random.Random(5); rs = lambda: rnd.randrange(L)
gens = api.Gens(512)
for n, kind in [(1,'one'),(3,'rand'),(300,'rand'),(300,'bits'),(7,'big')]:
    if kind=='rand': sc=[rnd.randrange(L) for _ in range(n)]
    elif kind=='bits': sc=[rnd.randrange(2) for _ in range(n)]
    elif kind=='one': sc=[1]
    else: sc=[L-1, L-2, 128, 129, 2**252, 255, 256][:n]
    arr = torch.from_numpy(api.scalars_to_array(sc)).cuda()
    out = torch.zeros(32, dtype=torch.uint8, device='cuda')
    rc = lib.bp_msm_gens_device(gens._h, n, C.c_void_p(arr.data_ptr()), C.c_void_p(out.data_ptr()), None)
    torch.cuda.synchronize()
    Gs = R.BulletproofGens(300).G(n)
    exp = R.ristretto_encode(R.msm(sc, Gs))
    print(n, kind, rc, out.cpu().numpy().tobytes()==exp, flush=True)
consts = [rs() for _ in range(5)]
xl, xr = rs(), rs(); img = G.mimc(xl, xr, consts)
bl = [rs(), rs()]
ent = bytes(range(32))
p = api.Prover(gens, b"MiMC")
V0, v0 = p.commit(xl, bl[0]); V1, v1 = p.commit(xr, bl[1])
p.mimc_gadget(v0, v1, consts, img)
proof = p.prove(ent)
op = R.Prover(R.PedersenGens(), R.Transcript(b"MiMC")); oV0, ov0 = op.commit(xl, bl[0]); oV1, ov1 = op.commit(xr, bl[1])
G.mimc_gadget(op, ov0, ov1, 5, consts, img)
oproof = R.proof_to_bytes(op.prove(R.BulletproofGens(128), ent))
print("V", V0==oV0, V1==oV1)
print([proof[i:i+32]==oproof[i:i+32] for i in range(0,len(proof),32)])
