import sys, time, random, os, ctypes as C, struct
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
from oracle import bp_pyref as R, gadgets_pyref as G, c_oracle as CO
from bulletproofs_r1cs_gadgets_b200 import api
lib = api.load(sys.argv[1] if len(sys.argv) > 1 else None)
L = R.L
os.environ["BP_B200_DEBUG_DUMP"] = "/tmp/bp_dump.bin"
rnd = random.Random(5); rs = lambda: rnd.randrange(L)
gens = api.Gens(128)
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
tr = {}
oproof = R.proof_to_bytes(op.prove(R.BulletproofGens(128), ent, trace=tr))
d = {}
f = open("/tmp/bp_dump.bin", "rb").read(); off = 0
while off < len(f):
    name = f[off:off+24].rstrip(b"\0").decode(); cnt = struct.unpack("<Q", f[off+24:off+32])[0]; off += 32
    if name == "raw_states":
        raw = f[off:]; break
    d[name] = [int.from_bytes(f[off+32*i:off+32*i+32], "little") for i in range(cnt)]; off += 32*cnt
n = 10
ot = R.Transcript(b"MiMC"); ot.append_message(b"dom-sep", b"r1cs v1"); ot.append_point(b"V", oV0); ot.append_point(b"V", oV1); ot.append_u64(b"m", 2)
def show(tag, raw208, st):
    print(tag, "state", raw208[:200] == bytes(st.state), "pos", raw208[200], st.pos, "pos_begin", raw208[201], st.pos_begin, "flags", raw208[202], st.cur_flags)
show("ts", raw[:208], ot.strobe)
show("rng", raw[208:416], ot.build_rng(bl, ent).strobe)
rr = raw[208:416]; orr = bytes(ot.build_rng(bl, ent).strobe.state)
print("rng first32 == entropy", rr[:32] == ent, "ndiff", sum(x != y for x, y in zip(rr[:200], orr)), rr[:40].hex())
print("Vdev", raw[416:448] == oV0, raw[448:480] == oV1)
print("v", d["v"] == [xl, xr], "vbl", d["vbl"] == bl)
print("sL", d["rand1"][3:3+n] == tr["sL"], "sR", d["rand1"][3+n:3+2*n] == tr["sR"])
print("wit", d["wit"][:n] == tr["aL"], d["wit"][n:2*n] == tr["aR"], d["wit"][2*n:] == tr["aO"])
print("y", d["chal"][0] == tr["y"], "z", d["chal"][1] == tr["z"], "u", d["chal"][3] == tr["u"], "x", d["chal"][4] == tr["x"], "w", d["chal"][5] == tr["w"])
print("wL", d["w_all"][:n] == tr["wL"], "wR", d["w_all"][n:2*n] == tr["wR"], "wO", d["w_all"][2*n:3*n] == tr["wO"], "wV", d["w_all"][3*n:3*n+2] == tr["wV"])
print("t", d["t"] == [tr["t"][j] for j in range(1, 7)])
print([proof[i:i+32]==oproof[i:i+32] for i in range(0,len(proof),32)])
