"""Shared drivers for the parity tests: the same flows run against the emulation build (CPU, -m "not gpu") and the
product library (GPU, -m gpu).  The oracle is only ever the checker."""
import json
import os
import random

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
from oracle import bp_pyref as R, gadgets_pyref as G, c_oracle as CO  # noqa: E402

L = R.L
POSEIDON_BLOB = open(os.path.join(ROOT, "bulletproofs_r1cs_gadgets_b200", "data", "poseidon_constants.bin"), "rb").read()


def golden_cases():
    return json.load(open(os.path.join(ROOT, "tests", "golden", "proofs.json")))["cases"]


def golden_builder(case):
    """returns build(cs, vars, is_prover) for the Python-oracle gadget layer"""
    p = case["params"]
    name = case["name"]
    if name.startswith("mimc"):
        consts = [int(c, 16) for c in p["constants"]]
        return lambda cs, v, pr: G.mimc_gadget(cs, v[0], v[1], len(consts), consts, int(p["image"], 16))
    if name.startswith("bound"):
        v, mn, mx, bs = p["v"], p["min"], p["max"], p["bit_size"]
        return lambda cs, vs, pr: G.bound_check_gadget(cs, (vs[0], v if pr else None), (vs[1], v - mn if pr else None),
                                                       (vs[2], mx - v if pr else None), mx, mn, bs)
    pp = G.PoseidonParams(6, p["full_b"], p["full_e"], p["partial"])
    if name.startswith("poseidon"):
        return lambda cs, vs, pr: G.poseidon_hash_2_gadget(cs, vs[0], vs[1], vs[2:6], pp, p["sbox"], int(p["hash"], 16))
    if name.startswith("vsmt2"):
        d = p["depth"]
        return lambda cs, vs, pr: G.vanilla_merkle_tree_verif_gadget(cs, d, int(p["root"], 16), vs[0], vs[1:1 + d], vs[1 + d:1 + 2 * d], vs[1 + 2 * d:], pp)
    raise KeyError(name)


def product_builder(api, case):
    """same circuits through the product's own gadget entry points (C++ host layer behind the C-ABI)"""
    p = case["params"]
    name = case["name"]
    if name.startswith("mimc"):
        consts = [int(c, 16) for c in p["constants"]]
        return lambda cs, v, pr: cs.mimc_gadget(v[0], v[1], consts, int(p["image"], 16))
    if name.startswith("bound"):
        v, mn, mx, bs = p["v"], p["min"], p["max"], p["bit_size"]
        return lambda cs, vs, pr: cs.bound_check_gadget(vs[0], vs[1], vs[2], mx, mn, bs, values=(v, v - mn, mx - v) if pr else None)
    pp = api.PoseidonParams(6, p["full_b"], p["full_e"], p["partial"])
    if name.startswith("poseidon"):
        return lambda cs, vs, pr: cs.poseidon_hash_2_gadget(pp, vs[0], vs[1], vs[2:6], p["sbox"], int(p["hash"], 16))
    if name.startswith("vsmt2"):
        d = p["depth"]
        return lambda cs, vs, pr: cs.vsmt2_verif_gadget(pp, d, int(p["root"], 16), vs[0], vs[1:1 + d], vs[1 + d:1 + 2 * d], vs[1 + 2 * d:])
    raise KeyError(name)


def oracle_prove_case(case):
    """Python oracle prover on a golden case -> (commitments, proof bytes, prover)"""
    build = golden_builder(case)
    p = R.Prover(R.PedersenGens(), R.Transcript(case["label"].encode()))
    Vs, vs = [], []
    for v, b in zip(case["values"], case["blindings"]):
        V, var = p.commit(int(v, 16), int(b, 16)); Vs.append(V); vs.append(var)
    build(p, vs, True)
    return Vs, p


def c_oracle_prove_case(case):
    Vs, p = oracle_prove_case(case)
    circ = CO.Circuit.from_cs(p, len(Vs))
    vals = CO.scalars_to_array([int(v, 16) for v in case["values"]], L)
    bls = CO.scalars_to_array([int(b, 16) for b in case["blindings"]], L)
    rc, V, proof = CO.prove(circ, CO.scalars_to_array(p.aL, L), CO.scalars_to_array(p.aR, L), CO.scalars_to_array(p.aO, L), vals, bls,
                            case["label"].encode(), bytes.fromhex(case["entropy"]), case["gens_capacity"])
    return rc, V, proof, circ


def product_tier1_prove(api, gens, case):
    build = product_builder(api, case)
    p = api.Prover(gens, case["label"].encode())
    Vs, vs = [], []
    for v, b in zip(case["values"], case["blindings"]):
        V, var = p.commit(int(v, 16), int(b, 16)); Vs.append(V); vs.append(var)
    build(p, vs, True)
    return Vs, p.prove(bytes.fromhex(case["entropy"])), p


def product_tier1_verify(api, gens, case, Vs, proof, entropy=bytes(32)):
    build = product_builder(api, case)
    v = api.Verifier(gens, case["label"].encode())
    vs = [v.commit(V) for V in Vs]
    build(v, vs, False)
    try:
        v.verify(proof, entropy)
        return 0
    except api.R1CSError as e:
        return e.code


def oracle_generic_prover(api_cs_cls, *a):
    raise NotImplementedError


def rand_scalars(seed, n):
    rnd = random.Random(seed)
    return [rnd.randrange(L) for _ in range(n)]


def selftest(api, which, data, outlen):
    import ctypes as C
    out = (C.c_uint8 * outlen)()
    rc = api.load().bp_selftest_device(which, api._buf(data), C.c_size_t(len(data)), out, C.c_size_t(outlen))
    assert rc == 0, rc
    return bytes(out)


def run_primitive_selftests(api):
    """device primitives vs the big-int oracle, through bp_selftest_device"""
    import hashlib
    rnd = random.Random(3)
    x = bytes(rnd.randrange(256) for _ in range(200))
    assert selftest(api, 7, x, 200) == bytes(R.keccak_f(bytearray(x)))
    lab = b"test protocol"
    assert selftest(api, 0, bytes([len(lab)]) + lab + b"some data", 32).hex() == "d5a21972d0d5fe320c0d263fac7fffb8145aa640af6e9bca177c03c7efcf0615"
    for _ in range(4):
        a, b = rnd.randrange(L), rnd.randrange(L)
        assert int.from_bytes(selftest(api, 1, a.to_bytes(32, "little"), 32), "little") == pow(a, L - 2, L)
        assert int.from_bytes(selftest(api, 4, a.to_bytes(32, "little") + b.to_bytes(32, "little"), 32), "little") == a * b % L
        w = bytes(rnd.randrange(256) for _ in range(64))
        assert int.from_bytes(selftest(api, 2, w, 32), "little") == int.from_bytes(w, "little") % L
        assert selftest(api, 5, w, 32) == R.ristretto_encode(R.from_uniform_bytes(w))
    assert int.from_bytes(selftest(api, 1, bytes(32), 32), "little") == 0  # invert(0) == 0
    for edge in (L - 1, 1, 2 ** 252, 2 ** 256 - 1):
        assert int.from_bytes(selftest(api, 1, (edge % 2 ** 256).to_bytes(32, "little"), 32), "little") == pow(edge % L, L - 2, L)
    r = selftest(api, 3, R.BASEPOINT_COMPRESSED, 33)
    assert r[32] == 1 and r[:32] == R.BASEPOINT_COMPRESSED
    # RFC 9496 one-way map vector
    h = hashlib.sha512(b"Ristretto is traditionally a short shot of espresso coffee").digest()
    assert selftest(api, 5, h, 32).hex() == "3066f82a1a747d45120d1740f14358531a8f04bbffe6a819f86dfe50f44a0a46"
    # non-canonical / invalid encodings are rejected (RFC 9496 section A.3 style): s >= p, negative s, non-square
    for bad in (bytes([0xed] + [0xff] * 30 + [0x7f]), bytes([1] + [0] * 31), bytes([0xff] * 32)):
        assert selftest(api, 3, bad, 33)[32] == 0
    a = rnd.randrange(L); w = bytes(rnd.randrange(256) for _ in range(32))
    t = R.Transcript(b"rngtest"); rng = t.build_rng([a], w)
    exp = b"".join(rng.random_scalar().to_bytes(32, "little") for _ in range(4))
    assert selftest(api, 6, a.to_bytes(32, "little") + w, 128) == exp


def check_golden_tier1(api, gens):
    for case in golden_cases():
        Vs, proof, _ = product_tier1_prove(api, gens, case)
        assert [V.hex() for V in Vs] == case["commitments"], case["name"]
        assert proof.hex() == case["proof"], case["name"]
        want = 0 if case["verifies"] else 3
        assert product_tier1_verify(api, gens, case, Vs, proof) == want, case["name"]


def tamper_cases(proof):
    """(description, bytes) variants of a valid proof that must all be rejected"""
    out = []
    n = len(proof)
    for off in (3, 40, 70, 200, 330, 360, 392, 424, 450, 482, n - 50, n - 20):
        b = bytearray(proof); b[off] ^= 1
        out.append(("flip@%d" % off, bytes(b)))
    b = bytearray(proof); b[0:32] = bytes(32)
    out.append(("A_I1 identity", bytes(b)))
    b = bytearray(proof); b[352:384] = (L + 5).to_bytes(32, "little")
    out.append(("t_x non-canonical", bytes(b)))
    out.append(("truncated", proof[:-32]))
    out.append(("extended", proof + bytes(64)))
    return out
