import sys, os, random, hashlib, ctypes as C
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import numpy as np
from oracle import bp_pyref as R
from bulletproofs_r1cs_gadgets_b200 import api
lib = api.load(sys.argv[1] if len(sys.argv) > 1 else None)
def st(which, data, outlen):
    out = (C.c_uint8 * outlen)()
    rc = lib.bp_selftest_device(which, api._buf(data), C.c_size_t(len(data)), out, C.c_size_t(outlen))
    assert rc == 0, rc
    return bytes(out)
rnd = random.Random(3)
x = bytes(rnd.randrange(256) for _ in range(200))
print("keccak", st(7, x, 200) == bytes(R.keccak_f(bytearray(x))))
lab = b"test protocol"
print("merlin", st(0, bytes([len(lab)]) + lab + b"some data", 32).hex() == "d5a21972d0d5fe320c0d263fac7fffb8145aa640af6e9bca177c03c7efcf0615")
a = rnd.randrange(R.L); b = rnd.randrange(R.L)
print("inv", int.from_bytes(st(1, a.to_bytes(32,'little'), 32),'little') == pow(a, R.L-2, R.L))
w = bytes(rnd.randrange(256) for _ in range(64))
print("wide", int.from_bytes(st(2, w, 32),'little') == int.from_bytes(w,'little') % R.L)
print("mul", int.from_bytes(st(4, a.to_bytes(32,'little')+b.to_bytes(32,'little'), 32),'little') == a*b % R.L)
r = st(3, R.BASEPOINT_COMPRESSED, 33); print("ristretto", r[32] == 1 and r[:32] == R.BASEPOINT_COMPRESSED)
print("uniform", st(5, w, 32) == R.ristretto_encode(R.from_uniform_bytes(w)))
t = R.Transcript(b"rngtest"); rng = t.build_rng([a], w[:32])
exp = b"".join(rng.random_scalar().to_bytes(32,'little') for _ in range(4))
print("rng", st(6, a.to_bytes(32,'little') + w[:32], 128) == exp)

# replica of the prover's transcript start (crosses the STROBE rate boundary several times)
V0, V1 = R.ristretto_encode(R.pt_mul(5, R.BASEPOINT)), R.ristretto_encode(R.pt_mul(7, R.BASEPOINT))
ot = R.Transcript(b"MiMC"); ot.append_message(b"dom-sep", b"r1cs v1"); ot.append_point(b"V", V0); ot.append_point(b"V", V1); ot.append_u64(b"m", 2)
out = st(8, V0 + V1 + a.to_bytes(32,'little') + b.to_bytes(32,'little') + w[:32], 480)
print("tsstart ts", out[:200] == bytes(ot.strobe.state), out[200:203] == bytes([ot.strobe.pos, ot.strobe.pos_begin, ot.strobe.cur_flags]))
rg = ot.build_rng([a, b], w[:32])
print("tsstart rng", out[208:408] == bytes(rg.strobe.state), out[408:411] == bytes([rg.strobe.pos, rg.strobe.pos_begin, rg.strobe.cur_flags]))
print("tsstart draws", out[416:480] == b"".join(rg.random_scalar().to_bytes(32,'little') for _ in range(2)))

out = (C.c_uint8 * 416)()
inp = V0 + V1 + a.to_bytes(32,'little') + b.to_bytes(32,'little') + w[:32]
rc = lib.bp_selftest_tsstart(2, api._buf(inp), api._buf(b"MiMC"), C.c_size_t(4), out); out = bytes(out)
print("real KTsStart ts", rc, out[:200] == bytes(ot.strobe.state))
rg2 = ot.build_rng([a, b], w[:32])
print("real KTsStart rng", out[208:408] == bytes(rg2.strobe.state), "first32==entropy", out[208:240] == w[:32])
