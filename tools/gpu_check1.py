"""First GPU bring-up check: product library vs the oracles (dev tool; tests/ holds the real suite)."""
import sys, time, random, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
from oracle import bp_pyref as R, gadgets_pyref as G, c_oracle as CO
from bulletproofs_r1cs_gadgets_b200 import api
L = R.L
rnd = random.Random(5); rs = lambda: rnd.randrange(L)
t0 = time.time(); gens = api.Gens(2048); print("gens(2048) %.2fs" % (time.time() - t0), flush=True)
B_, Bb = gens.pedersen()
print("pedersen ok", B_ == R.BASEPOINT_COMPRESSED, Bb.hex() == "8c9240b456a9e6dc65c377a1048d745f94a08cdb7f44cbcd7b46f34048871134")
g = gens.export(0, 2048)
og = np.zeros((2048, 32), np.uint8); CO.lib().bpo_gens_compressed(0, 2048, og.ctypes.data_as(CO.u8p))
print("G chain ok", g.tobytes() == og.tobytes())
v, r = rs(), rs()
print("commit ok", gens.commit(v, r) == R.ristretto_encode(R.PedersenGens().commit(v, r)))
consts = [rs() for _ in range(5)]
xl, xr = rs(), rs(); img = G.mimc(xl, xr, consts)
bl = [rs(), rs()]
ent = bytes(range(32))
p = api.Prover(gens, b"MiMC")
V0, v0 = p.commit(xl, bl[0]); V1, v1 = p.commit(xr, bl[1])
p.mimc_gadget(v0, v1, consts, img)
t0 = time.time(); proof = p.prove(ent); print("mimc5 prove %.3fs" % (time.time() - t0))
op = R.Prover(R.PedersenGens(), R.Transcript(b"MiMC")); oV0, ov0 = op.commit(xl, bl[0]); oV1, ov1 = op.commit(xr, bl[1])
G.mimc_gadget(op, ov0, ov1, 5, consts, img)
oproof = R.proof_to_bytes(op.prove(R.BulletproofGens(128), ent))
print("mimc5 proof matches python oracle:", proof == oproof)
# poseidon 2:1 full rounds, batch via witness program, compare to C oracle with native witness
pp = api.PoseidonParams()
blob = open(os.path.join(os.path.dirname(api.__file__), "data", "poseidon_constants.bin"), "rb").read()
CO.poseidon_set_params(blob)
for sbox, name in ((api.SBOX_CUBE, "cube"), (api.SBOX_INVERSE, "inverse")):
    vf = api.Verifier(gens, b"Poseidon_hash_2")
    # structure only: commitments are placeholders on the recording side
    xs = [vf.commit(bytes(32)) for _ in range(2)]
    st = [vf.commit(bytes(32)) for _ in range(4)]
    # expected hash differs per proof -> for the batch test use the same inputs in every proof
    a, b = rs(), rs()
    h = pp.hash_2(a, b, sbox)
    vf.poseidon_hash_2_gadget(pp, xs[0], xs[1], st, sbox, h)
    circ = vf.compile()
    print(name, "n", circ.n, "q", circ.q, "m", circ.m, "aux", circ.num_aux, "tape", circ.has_witness_program)
    Bn = 4
    vals = api.scalars_to_array([a, b, 0, 101, 0, 0] * Bn).reshape(Bn, 6, 32)
    bls = api.scalars_to_array(sum(([rs(), rs(), 0, 0, 0, 0] for _ in range(Bn)), [])).reshape(Bn, 6, 32)
    ents = np.frombuffer(bytes(rnd.randrange(256) for _ in range(32 * Bn)), np.uint8).reshape(Bn, 32)
    t0 = time.time(); V, proofs, status = circ.prove_batch(gens, b"Poseidon_hash_2", vals, bls, ents); dt = time.time() - t0
    # oracle circuit from python recorder
    ov = R.Verifier(R.Transcript(b"x")); ovs = [ov.commit(bytes(32)) for _ in range(6)]
    opp = G.PoseidonParams()
    G.poseidon_hash_2_gadget(ov, ovs[0], ovs[1], ovs[2:], opp, sbox, h)
    oc = CO.Circuit.from_cs(ov, 6)
    aL, aR, aO, hh = CO.poseidon_hash2_witness(a.to_bytes(32, "little"), b.to_bytes(32, "little"), sbox, oc.n)
    ok = True
    for i in range(Bn):
        rc, oV, opf = CO.prove(oc, aL, aR, aO, vals[i], bls[i], b"Poseidon_hash_2", ents[i].tobytes(), 2048)
        ok &= (rc == 0 and oV.tobytes() == V[i].tobytes() and opf == proofs[i].tobytes())
    print(name, "batch of %d: %.3fs status %s matches C oracle: %s" % (Bn, dt, status.tolist(), ok), flush=True)
print("launches", api.launch_count())
