"""Generates tests/golden/proofs.json with the independent big-int oracle (oracle/bp_pyref.py + gadgets_pyref.py).

The reference (Rust, un-vendored crates) cannot run in this image and holds no golden proofs of its own, so these
vectors pin the restated protocol: both oracles and the CUDA product must reproduce them byte for byte.
Run from the repo root:  python tests/golden/make_golden.py
"""
import json, os, random, sys
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import bp_pyref as R, gadgets_pyref as G

L = R.L
rnd = random.Random(20261017)
rs = lambda: rnd.randrange(L)
out = {"cases": []}


def run(name, label, values, blindings, cap, build, params, verifies=True):
    pc, bp = R.PedersenGens(), R.BulletproofGens(cap)
    p = R.Prover(pc, R.Transcript(label))
    Vs, vs = [], []
    for v, b in zip(values, blindings):
        V, var = p.commit(v, b); Vs.append(V); vs.append(var)
    build(p, vs, True)
    entropy = bytes(rnd.randrange(256) for _ in range(32))
    proof = R.proof_to_bytes(p.prove(bp, entropy))
    vf = R.Verifier(R.Transcript(label)); vv = [vf.commit(V) for V in Vs]; build(vf, vv, False)
    try:
        ok = vf.verify(R.proof_from_bytes(proof), pc, bp, bytes(32))
    except R.VerificationError:
        ok = False
    assert ok == verifies
    out["cases"].append({"name": name, "label": label.decode(), "gens_capacity": cap, "params": params,
                         "values": [hex(v) for v in values], "blindings": [hex(b) for b in blindings], "entropy": entropy.hex(),
                         "verifies": verifies, "n": p.num_multipliers(), "q": p.num_constraints(), "commitments": [V.hex() for V in Vs], "proof": proof.hex()})
    print(name, "n", p.num_multipliers(), "q", p.num_constraints(), len(proof))


# MiMC, 5 rounds
consts = [rs() for _ in range(5)]
xl, xr = rs(), rs()
img = G.mimc(xl, xr, consts)
run("mimc5", b"MiMC", [xl, xr], [rs(), rs()], 16, lambda cs, v, pr: G.mimc_gadget(cs, v[0], v[1], 5, consts, img),
    {"constants": [hex(c) for c in consts], "image": hex(img)})
# bound check, 8 bits (n = 16 = N: no padding)
v, mn, mx = 77, 10, 200
run("bound8", b"BoundsTest", [v, v - mn, mx - v], [rs(), rs(), rs()], 16,
    lambda cs, vs, pr: G.bound_check_gadget(cs, (vs[0], v if pr else None), (vs[1], v - mn if pr else None), (vs[2], mx - v if pr else None), mx, mn, 8),
    {"v": v, "min": mn, "max": mx, "bit_size": 8})
# Poseidon 2:1 with reduced rounds (2 + 3 + 2), both S-boxes
pp = G.PoseidonParams(6, 2, 2, 3)
for sbox, nm in ((G.CUBE, "cube"), (G.INVERSE, "inverse")):
    a, b = rs(), rs()
    h = G.poseidon_hash_2(a, b, pp, sbox)
    run("poseidon_2_3_2_" + nm, b"Poseidon_hash_2_" + nm.encode(), [a, b, 0, 101, 0, 0], [rs(), rs(), 0, 0, 0, 0], 128,
        lambda cs, vs, pr: G.poseidon_hash_2_gadget(cs, vs[0], vs[1], vs[2:6], pp, sbox, h),
        {"full_b": 2, "full_e": 2, "partial": 3, "sbox": sbox, "hash": hex(h)})
# inverse S-box with a zero input in the first round, lane 1: xl = -round_key[1]  (Scalar::invert(0) == 0 path)
a, b = (-pp.round_keys[1]) % L, rs()
h = G.poseidon_hash_2(a, b, pp, G.INVERSE)
run("poseidon_2_3_2_inverse_zero_input", b"Poseidon_hash_2_inverse", [a, b, 0, 101, 0, 0], [rs(), rs(), 0, 0, 0, 0], 128,
    lambda cs, vs, pr: G.poseidon_hash_2_gadget(cs, vs[0], vs[1], vs[2:6], pp, G.INVERSE, h),
    {"full_b": 2, "full_e": 2, "partial": 3, "sbox": 1, "hash": hex(h)}, verifies=False)  # the circuit forces S-box inputs to be non-zero
# VSMT-2, depth 2, reduced rounds
depth = 2
leaf, bits, sibs = rs(), [1, 0], [rs(), rs()]
root = G.vsmt_root_from_path(leaf, bits, sibs, pp)
vals = [leaf] + bits + sibs + [0, 101, 0, 0]
bl = [rs() for _ in range(1 + 2 * depth)] + [0, 0, 0, 0]
run("vsmt2_depth2_2_3_2", b"VSMT", vals, bl, 256,
    lambda cs, vs, pr: G.vanilla_merkle_tree_verif_gadget(cs, depth, root, vs[0], vs[1:1 + depth], vs[1 + depth:1 + 2 * depth], vs[1 + 2 * depth:], pp),
    {"full_b": 2, "full_e": 2, "partial": 3, "depth": depth, "root": hex(root)})
json.dump(out, open(os.path.join(HERE, "proofs.json"), "w"), indent=1)
