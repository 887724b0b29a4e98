"""Kernel-body tests on the CPU: the CUDA sources compiled with g++ -DBP_HOST_EMUL (tests/emul).  Every kernel functor,
the engine orchestration, the recorder and the C-ABI run here exactly as on the device, one "thread" at a time, and
are compared with the oracle.  This is test infrastructure -- the product library has no CPU path."""
import numpy as np
import pytest

import helpers as H
from helpers import R, G, CO, L


@pytest.fixture(scope="module")
def api(emul_so):
    from bulletproofs_r1cs_gadgets_b200 import api as a
    a._lib = None
    a.load(emul_so)
    yield a
    a._lib = None


@pytest.fixture(scope="module")
def gens(api):
    return api.Gens(256)


def test_primitives(api):
    H.run_primitive_selftests(api)


def test_generators(api, gens, oracle_lib):
    B, Bb = gens.pedersen()
    assert B == R.BASEPOINT_COMPRESSED and Bb == R.ristretto_encode(R.PedersenGens().B_blinding)
    og = np.zeros((256, 32), np.uint8); oracle_lib.lib().bpo_gens_compressed(0, 256, og.ctypes.data_as(CO.u8p))
    oh = np.zeros((256, 32), np.uint8); oracle_lib.lib().bpo_gens_compressed(1, 256, oh.ctypes.data_as(CO.u8p))
    assert gens.export(0, 256).tobytes() == og.tobytes() and gens.export(1, 256).tobytes() == oh.tobytes()
    for v, r in [(0, 0), (1, 0), (0, 1), (L - 1, L - 1)] + [tuple(H.rand_scalars(4, 2))]:
        assert gens.commit(v, r) == R.ristretto_encode(R.PedersenGens().commit(v, r))


def _msm_host_pointers(api, gens, arr):
    """emulation build: "device" pointers are host pointers"""
    out = np.zeros(32, np.uint8)
    rc = api.load().bp_msm_gens_device(gens._h, arr.shape[0], arr.ctypes.data_as(api.u8p), out.ctypes.data_as(api.u8p), None)
    return rc, out.tobytes()


def test_msm_entry(api, gens, call=_msm_host_pointers):
    rnd_sets = {"one": [1], "rand3": H.rand_scalars(1, 3), "rand200": H.rand_scalars(2, 200), "bits": [x & 1 for x in H.rand_scalars(3, 200)],
                "edge": [L - 1, L - 2, 127, 128, 129, 255, 256, 2 ** 252, 0], "zeros": [0, 0, 0]}
    for name, sc in rnd_sets.items():
        rc, out = call(api, gens, api.scalars_to_array(sc))
        assert rc == 0
        assert out == R.ristretto_encode(R.msm(sc, R.BulletproofGens(256).G(len(sc)))), name


def test_golden_proofs_tier1(api, gens):
    H.check_golden_tier1(api, gens)


def test_python_gadget_code_drives_product_cs(api, gens):
    """the oracle's generic gadget functions run unchanged against the product's ConstraintSystem mirror (tier-1 C-ABI):
    same recorded circuit, same proof bytes"""
    case = next(c for c in H.golden_cases() if c["name"] == "poseidon_2_3_2_inverse")
    build = H.golden_builder(case)

    class Adapter:  # maps the oracle's (kind, index) variables / LC objects onto the product API
        def __init__(self, cs): self.cs = cs
        def _lc(self, lc): return api.LinearCombination([(api.Variable(k, i), c) for (k, i), c in R.LC.of(lc).terms])
        def _v(self, v): return (v.kind, v.index)
        def multiply(self, l, r): return tuple(self._v(x) for x in self.cs.multiply(self._lc(l), self._lc(r)))
        def constrain(self, lc): self.cs.constrain(self._lc(lc))
        def evaluate_lc(self, lc): return self.cs.evaluate_lc(self._lc(lc))
        def allocate_single(self, a):
            v, o = self.cs.allocate_single(a)
            return self._v(v), (self._v(o) if o is not None else None)
        def allocate_multiplier(self, a): return tuple(self._v(x) for x in self.cs.allocate_multiplier(a))
    p = api.Prover(gens, case["label"].encode())
    vs = []
    for v, b in zip(case["values"], case["blindings"]):
        V, var = p.commit(int(v, 16), int(b, 16)); vs.append((var.kind, var.index))
    build(Adapter(p), vs, True)
    assert (p.num_multipliers(), p.num_constraints()) == (case["n"], case["q"])
    assert p.prove(bytes.fromhex(case["entropy"])).hex() == case["proof"]


def test_verifier_rejects_tampering(api, gens):
    case = next(c for c in H.golden_cases() if c["name"] == "vsmt2_depth2_2_3_2")
    Vs, proof, _ = H.product_tier1_prove(api, gens, case)
    assert H.product_tier1_verify(api, gens, case, Vs, proof) == 0
    assert H.product_tier1_verify(api, gens, case, Vs, proof, entropy=bytes(range(32))) == 0  # verifier entropy is free
    for what, bad in H.tamper_cases(proof):
        rc = H.product_tier1_verify(api, gens, case, Vs, bad)
        assert rc in (2, 3), (what, rc)
    Vbad = list(Vs); Vbad[0], Vbad[1] = Vbad[1], Vbad[0]
    assert H.product_tier1_verify(api, gens, case, Vbad, proof) == 3
    case2 = dict(case, label="VSMT-other")
    assert H.product_tier1_verify(api, gens, case2, Vs, proof) == 3


def test_batch_witness_program_and_public_inputs(api, gens):
    """tier 2: circuit compiled once, witness program on the device, per-proof public inputs; bytes equal the oracle's"""
    pp = api.PoseidonParams(6, 2, 2, 3)
    opp = G.PoseidonParams(6, 2, 2, 3)
    for sbox in (api.SBOX_CUBE, api.SBOX_INVERSE):
        rec = api.Verifier(gens, b"P"); xs = [rec.commit(bytes(32)) for _ in range(2)]; st = [rec.commit(bytes(32)) for _ in range(4)]
        hv = rec.public_input()
        rec.poseidon_hash_2_gadget(pp, xs[0], xs[1], st, sbox, hv)
        circ = rec.compile()
        assert circ.has_witness_program and circ.num_public == 1 and circ.num_aux == 0
        Bn = 3
        sc = H.rand_scalars(10 + sbox, 4 * Bn)
        ins = [(sc[4 * i], sc[4 * i + 1]) for i in range(Bn)]
        ins[1] = ((-opp.round_keys[1]) % L, ins[1][1])  # proof 1 hits Scalar::invert(0) in round 0 (invalid statement for the inverse S-box)
        hs = [G.poseidon_hash_2(a, b, opp, sbox) for a, b in ins]
        vals = api.scalars_to_array(sum(([a, b, 0, 101, 0, 0] for a, b in ins), [])).reshape(Bn, 6, 32)
        bls = api.scalars_to_array(sum(([sc[4 * i + 2], sc[4 * i + 3], 0, 0, 0, 0] for i in range(Bn)), [])).reshape(Bn, 6, 32)
        ents = np.arange(32 * Bn, dtype=np.uint8).reshape(Bn, 32)
        pub = api.scalars_to_array(hs).reshape(Bn, 1, 32)
        V, proofs, status = circ.prove_batch(gens, b"P", vals, bls, ents, pub=pub)
        assert not status.any()
        for i in range(Bn):
            op = R.Prover(R.PedersenGens(), R.Transcript(b"P"))
            ops = [op.commit(int.from_bytes(vals[i][j].tobytes(), "little"), int.from_bytes(bls[i][j].tobytes(), "little")) for j in range(6)]
            G.poseidon_hash_2_gadget(op, ops[0][1], ops[1][1], [o[1] for o in ops[2:]], opp, sbox, hs[i])
            assert b"".join(o[0] for o in ops) == V[i].tobytes()
            assert R.proof_to_bytes(op.prove(R.BulletproofGens(256), ents[i].tobytes())) == proofs[i].tobytes()
        ok = circ.verify_batch(gens, b"P", V, proofs, ents, pub=pub)
        assert ok.tolist() == ([0, 3, 0] if sbox == api.SBOX_INVERSE else [0, 0, 0])
        bad = pub.copy(); bad[2, 0, 0] ^= 1
        assert circ.verify_batch(gens, b"P", V, proofs, ents, pub=bad)[2] == 3


def test_batch_aux_inputs_bound_check(api, gens):
    """allocate_multiplier assignments (range-proof bits) travel as auxiliary inputs of the witness program"""
    from bulletproofs_r1cs_gadgets_b200 import workloads
    wl = workloads.BoundCheck(gens, vmin=10, vmax=200, bit_size=8)
    assert (wl.circuit.n, wl.circuit.q, wl.circuit.m, wl.circuit.num_aux) == (16, 37, 3, 32)
    inp = wl.inputs(0, 3)
    V, proofs, status = wl.circuit.prove_batch(gens, wl.label, inp["v"], inp["v_blinding"], inp["entropy"], aux=inp["aux"])
    assert not status.any()
    for i in range(3):
        v = int.from_bytes(inp["v"][i][0].tobytes(), "little")
        op = R.Prover(R.PedersenGens(), R.Transcript(b"BoundsTest"))
        ops = [op.commit(int.from_bytes(inp["v"][i][j].tobytes(), "little"), int.from_bytes(inp["v_blinding"][i][j].tobytes(), "little")) for j in range(3)]
        G.bound_check_gadget(op, (ops[0][1], v), (ops[1][1], v - 10), (ops[2][1], 200 - v), 200, 10, 8)
        assert R.proof_to_bytes(op.prove(R.BulletproofGens(16), inp["entropy"][i].tobytes())) == proofs[i].tobytes()
    assert not wl.circuit.verify_batch(gens, wl.label, V, proofs, inp["entropy"]).any()
    # a value outside the range cannot be proven: a = v - min wraps, its bits no longer recompose
    bad = {k: a.copy() for k, a in inp.items()}
    bad["v"][0, 0, 0] ^= 0x80
    V2, proofs2, st2 = wl.circuit.prove_batch(gens, wl.label, bad["v"], bad["v_blinding"], bad["entropy"], aux=bad["aux"])
    assert wl.circuit.verify_batch(gens, wl.label, V2, proofs2, bad["entropy"]).tolist()[0] == 3


def test_vsmt4_membership(api, gens, levels=2, params=(6, 2, 2, 3), count=3, c_oracle_prover=False):
    """4-ary sparse Merkle membership (reference src/gadget_vsmt_4.rs:199-312, Poseidon 4:1 of src/gadget_poseidon.rs:488-551):
    native root equal to the oracle's, commitments and proof bytes equal to the oracle's prover running the oracle's restatement
    of the gadget, verifier accepts, wrong root and wrong index digit rejected"""
    from bulletproofs_r1cs_gadgets_b200 import workloads
    pp, opp = api.PoseidonParams(*params), G.PoseidonParams(*params)
    wl = workloads.Vsmt4(gens, levels=levels, params=pp)
    assert wl.circuit.num_aux == 4 * levels and wl.circuit.num_public == 1 and wl.circuit.m == 2 + 3 * levels + 2
    inp = wl.inputs(0, count)
    V, P, st = wl.circuit.prove_batch(gens, wl.label, inp["v"], inp["v_blinding"], inp["entropy"], aux=inp["aux"], pub=inp["pub"])
    assert not st.any()
    cap = 1
    while cap < wl.circuit.n:
        cap *= 2
    for i in range(count):
        leaf, digits, sibs = wl.witness_values(i)
        root = G.vsmt4_root_from_path(leaf, digits, sibs, opp)
        assert root == int.from_bytes(inp["pub"][i, 0].tobytes(), "little")
        op = R.Prover(R.PedersenGens(), R.Transcript(b"VSMT"))
        ops = [op.commit(int.from_bytes(inp["v"][i][j].tobytes(), "little"), int.from_bytes(inp["v_blinding"][i][j].tobytes(), "little")) for j in range(wl.circuit.m)]
        vs = [o[1] for o in ops]
        G.vanilla_merkle_tree_4_verif_gadget(op, levels, root, vs[0], vs[1], digits, vs[2:2 + 3 * levels], vs[2 + 3 * levels:], opp)
        assert (len(op.aL), op.num_constraints()) == (wl.circuit.n, wl.circuit.q)
        assert b"".join(o[0] for o in ops) == V[i].tobytes()
        if c_oracle_prover:
            # large circuits: the Python oracle records the circuit and the witness, the C oracle (fast) produces the proof
            oc = CO.Circuit.from_cs(op, wl.circuit.m)
            rc, oV, oP = CO.prove(oc, api.scalars_to_array(op.aL), api.scalars_to_array(op.aR), api.scalars_to_array(op.aO), inp["v"][i], inp["v_blinding"][i],
                                  b"VSMT", inp["entropy"][i].tobytes(), cap)
            assert rc == 0 and oV.tobytes() == V[i].tobytes() and oP == P[i].tobytes()
        else:
            assert R.proof_to_bytes(op.prove(R.BulletproofGens(cap), inp["entropy"][i].tobytes())) == P[i].tobytes()
    assert not wl.circuit.verify_batch(gens, wl.label, V, P, inp["entropy"], pub=inp["pub"]).any()
    bad = inp["pub"].copy(); bad[1, 0, 0] ^= 1
    assert wl.circuit.verify_batch(gens, wl.label, V, P, inp["entropy"], pub=bad).tolist() == [0, 3] + [0] * (count - 2)
    # a prover lying about one index digit (its bits still sum to the committed index only if unchanged): rejected
    aux2 = inp["aux"].copy(); aux2[0, 0, 0] ^= 1; aux2[0, 1, 0] ^= 1
    V2, P2, st2 = wl.circuit.prove_batch(gens, wl.label, inp["v"], inp["v_blinding"], inp["entropy"], aux=aux2, pub=inp["pub"])
    assert wl.circuit.verify_batch(gens, wl.label, V2, P2, inp["entropy"], pub=inp["pub"])[0] == 3
    # tier 1: the same gadget through the one-at-a-time Prover / Verifier mirror
    leaf, digits, sibs = wl.witness_values(0)
    root = int.from_bytes(inp["pub"][0, 0].tobytes(), "little")
    pr = api.Prover(gens, b"VSMT")
    pv = [pr.commit(int.from_bytes(inp["v"][0][j].tobytes(), "little"), int.from_bytes(inp["v_blinding"][0][j].tobytes(), "little"))[1] for j in range(wl.circuit.m)]
    pr.vsmt4_verif_gadget(pp, levels, root, pv[0], pv[1], digits, pv[2:2 + 3 * levels], pv[2 + 3 * levels:])
    assert pr.prove(inp["entropy"][0].tobytes()) == P[0].tobytes()


def test_sparse_merkle_tree_and_membership_from_a_real_tree(api, gens):
    """the reference's two VSMT tests (src/gadget_vsmt_2.rs:223-259 data structure, :262-399 membership proof) at a small depth:
    update/get/verify_proof of the HashMap tree, then a membership proof whose witness comes from the tree instead of a synthetic path"""
    from bulletproofs_r1cs_gadgets_b200 import trees
    pp, opp = api.PoseidonParams(6, 2, 2, 3), G.PoseidonParams(6, 2, 2, 3)
    depth = 3  # 3 x (4 + 81) multipliers fit the 256 generators of the test fixture
    tree = trees.VanillaSparseMerkleTree(pp, depth)
    # empty-subtree recurrence equals the oracle's native hash
    e = 0
    for i in range(depth):
        e = G.poseidon_hash_2(e, e, opp, G.INVERSE)
        assert tree.empty_tree_hashes[i + 1] == e
    vals = {k: H.rand_scalars(40 + k, 1)[0] for k in (1, 2, 7, 4)}
    for k, v in vals.items():
        tree.update(k, v)
    for k, v in vals.items():
        proof = []
        assert tree.get(k, proof) == v and len(proof) == depth
        assert tree.verify_proof(k, v, proof) and tree.verify_proof(k, v, proof, tree.root)
        assert not tree.verify_proof(k, (v + 1) % L, proof)
    assert tree.get(5) == 0  # untouched leaf
    # membership proof for key 7 (reference test flow: commit leaf, index bits LSB first, siblings leaf level first, statics)
    k = 7; proof = []; leaf = tree.get(k, proof); proof.reverse()
    bits = [(k >> i) & 1 for i in range(depth)]
    pr = api.Prover(gens, b"VSMT")
    bl = H.rand_scalars(77, 1 + 2 * depth)
    com = [pr.commit(x, b) for x, b in zip([leaf] + bits + proof, bl)]
    Vs, cvars = [c[0] for c in com], [c[1] for c in com]
    statics = pr.allocate_statics(4)
    pr.vsmt2_verif_gadget(pp, depth, tree.root, cvars[0], cvars[1:1 + depth], cvars[1 + depth:], statics)
    pf = pr.prove(bytes(range(32)))
    for root, ok in ((tree.root, True), ((tree.root + 1) % L, False)):
        vf = api.Verifier(gens, b"VSMT")
        vv = [vf.commit(V) for V in Vs]
        vst = vf.allocate_statics(4)
        vf.vsmt2_verif_gadget(pp, depth, root, vv[0], vv[1:1 + depth], vv[1 + depth:], vst)
        if ok:
            assert vf.verify(pf, bytes(32))
        else:
            with pytest.raises(api.R1CSError) as e:
                vf.verify(pf, bytes(32))
            assert e.value.code == 3


def test_device_tree_depth5(api, gens): test_device_tree(api, gens, depth=5, nkeys=11, prove=False)
def test_device_tree_depth63(api, gens): test_device_tree(api, gens, depth=63, nkeys=9, prove=False, seed=77)  # widest index of the 64-bit entry points
def test_device_tree_depth253(api, gens): test_device_tree(api, gens, depth=253, nkeys=9, prove=False, seed=253)  # the reference's TreeDepth: 256-bit keys


def test_device_tree(api, gens, oracle_lib=None, depth=3, params=(6, 2, 2, 3), nkeys=6, prove=True, seed=900):
    """device-side batched sparse Merkle tree (bp_vsmt2_*) against the oracle's restatement of the reference's
    VanillaSparseMerkleTree (oracle/tree_pyref.py, reference src/gadget_vsmt_2.rs:27-166) applied one key at a time:
    empty-subtree hashes, roots after batched updates (with a repeated key and a later overwrite), get (leaf + siblings root -> leaf),
    an untouched key, verify_proof on the device's paths, the circuit's witness rows, and a membership proof driven from the tree"""
    import random
    from bulletproofs_r1cs_gadgets_b200 import trees, workloads
    from oracle import tree_pyref as TP
    pp = api.PoseidonParams(*params)
    if oracle_lib is not None and params == (6, 4, 4, 140):  # full-size permutation: the C oracle's hash (the Python one takes ~0.1 s)
        oracle_lib.poseidon_set_params(H.POSEIDON_BLOB)
        h2 = TP.c_oracle_hash2(oracle_lib, 1)
    else:
        opp = G.PoseidonParams(*params)
        h2 = lambda a, b: G.poseidon_hash_2(a, b, opp, G.INVERSE)
    ref = TP.VanillaSparseMerkleTree(h2, depth)
    dev = trees.DeviceVsmt2(pp, depth)
    assert dev.empty_tree_hashes == ref.empty_tree_hashes and dev.root == ref.root and dev.num_nodes == 0
    rnd = random.Random(seed)
    keys = [0, 2 ** depth - 1, 1] + [rnd.randrange(2 ** depth) for _ in range(nkeys - 3)]
    vals = H.rand_scalars(seed, len(keys))
    keys.append(keys[4]); vals.append(vals[0] + 5)  # a repeated key inside one batch: the last value stays
    first = len(keys) // 2
    for a, b in ((0, first), (first, len(keys))):
        for k, v in zip(keys[a:b], vals[a:b]):
            ref.update(k, v)
        assert dev.update_batch(keys[a:b], vals[a:b]) == ref.root == dev.root
    over = {keys[1]: 7, keys[5]: 0}  # overwrite existing leaves (one with the empty value)
    for k, v in over.items():
        ref.update(k, v)
    assert dev.update_batch(list(over), list(over.values())) == ref.root
    assert dev.update_batch([], []) == ref.root
    n_before = dev.num_nodes
    assert dev.update_batch(list(over), list(over.values())) == ref.root and dev.num_nodes == n_before  # idempotent, no new nodes
    untouched = next(k for k in range(2 ** depth) if k not in keys)
    q = sorted(set(keys)) + [untouched]
    leaves, proofs = dev.get_batch(q)
    for i, k in enumerate(q):
        path = []
        assert int.from_bytes(leaves[i].tobytes(), "little") == ref.get(k, path)
        got = [int.from_bytes(proofs[i, j].tobytes(), "little") for j in range(depth)]
        assert got == path
        assert ref.verify_proof(k, ref.get(k), got)
    assert int.from_bytes(leaves[-1].tobytes(), "little") == 0
    plist = []
    assert dev.get(q[2], plist) == ref.get(q[2]) and len(plist) == depth
    # witness rows = the reference prover's commit order: leaf, bits LSB first, siblings leaf level first, statics
    v, pub = dev.witness_rows(q)
    for i, k in enumerate(q):
        path = []; leaf = ref.get(k, path); path.reverse()
        want = [leaf] + [(k >> j) & 1 for j in range(depth)] + path + [0, 101, 0, 0]
        assert v[i].tobytes() == api.scalars_to_array(want).tobytes()
        assert pub[i, 0].tobytes() == api.scalar_bytes(ref.root)
    with pytest.raises(api.R1CSError):
        dev.update_batch([2 ** depth], [1])
    # batched hash entry
    xs, ys = H.rand_scalars(seed + 1, 5), H.rand_scalars(seed + 2, 5)
    hb = trees.poseidon_hash_2_batch(pp, [0] + xs, [0] + ys)
    assert [int.from_bytes(hb[i].tobytes(), "little") for i in range(6)] == [h2(a, b) for a, b in zip([0] + xs, [0] + ys)]
    if not prove:
        return
    # membership proofs straight from the device tree: rows -> prove_batch -> verify_batch (root as the public input)
    wl = workloads.Vsmt2(gens, depth=depth, params=pp)
    B = 3
    vb = api.scalars_to_array(H.rand_scalars(seed + 3, B * wl.circuit.m)).reshape(B, wl.circuit.m, 32)
    vb[:, -4:] = 0  # statics are committed with blinding 0 (reference src/gadget_poseidon.rs:556-571)
    ent = np.frombuffer(bytes(range(B * 32)), dtype=np.uint8).reshape(B, 32)
    V, P, st = wl.circuit.prove_batch(gens, wl.label, v[:B], vb, ent, pub=pub[:B])
    assert not st.any()
    assert not wl.circuit.verify_batch(gens, wl.label, V, P, ent, pub=pub[:B]).any()
    bad = pub[:B].copy(); bad[1, 0, 0] ^= 1
    assert wl.circuit.verify_batch(gens, wl.label, V, P, ent, pub=bad).tolist() == [0, 3, 0]


def test_explicit_witness_equals_witness_program(api, gens, oracle_lib):
    from bulletproofs_r1cs_gadgets_b200 import workloads
    wl = workloads.Mimc(gens, rounds=6)
    inp = wl.inputs(5, 2)
    V1, P1, s1 = wl.circuit.prove_batch(gens, wl.label, inp["v"], inp["v_blinding"], inp["entropy"], pub=inp["pub"])
    cb = CO.scalars_to_array(wl.constants, L).tobytes()
    wit = [oracle_lib.mimc_witness(inp["v"][i][0].tobytes(), inp["v"][i][1].tobytes(), cb, wl.circuit.n) for i in range(2)]
    aL, aR, aO = (np.stack([w[j] for w in wit]) for j in range(3))
    V2, P2, s2 = wl.circuit.prove_batch(gens, wl.label, inp["v"], inp["v_blinding"], inp["entropy"], witness=(aL, aR, aO))
    assert P1.tobytes() == P2.tobytes() and V1.tobytes() == V2.tobytes() and not s1.any() and not s2.any()
    assert not wl.circuit.verify_batch(gens, wl.label, V1, P1, inp["entropy"], pub=inp["pub"]).any()


def test_streamed_batches_equal_plain_batches(api, gens):
    """bp_prove_stream_*: two batches in flight (begin 0, begin 1, finish 0, begin 0 again, finish 1, finish 0) give the bytes of
    three plain bp_prove_batch calls; a slot cannot be begun twice or finished when idle; chunked batches stream too"""
    import os
    from bulletproofs_r1cs_gadgets_b200 import workloads
    wl = workloads.Mimc(gens, rounds=6)
    batches = [wl.inputs(10 * i, 3 + i) for i in range(3)]
    plain = [wl.circuit.prove_batch(gens, wl.label, b["v"], b["v_blinding"], b["entropy"], pub=b["pub"]) for b in batches]
    for chunk in (None, "2"):
        if chunk:
            os.environ["BP_B200_CHUNK"] = chunk
        try:
            st = api.ProveStream(wl.circuit, gens, wl.label)
            args = lambda b: (b["v"], b["v_blinding"], b["entropy"], None, b["pub"])
            st.begin(0, *args(batches[0]))
            st.begin(1, *args(batches[1]))
            with pytest.raises(api.R1CSError):
                st.begin(1, *args(batches[2]))
            got0 = st.finish(0)
            st.begin(0, *args(batches[2]))
            got1 = st.finish(1)
            got2 = st.finish(0)
            with pytest.raises(api.R1CSError):
                st.finish(0)
        finally:
            os.environ.pop("BP_B200_CHUNK", None)
        for got, want in zip((got0, got1, got2), plain):
            assert got[0].tobytes() == want[0].tobytes() and got[1].tobytes() == want[1].tobytes() and not got[2].any()
    # a plain call while a streamed batch is in flight uses the other slot
    st = api.ProveStream(wl.circuit, gens, wl.label)
    st.begin(0, batches[0]["v"], batches[0]["v_blinding"], batches[0]["entropy"], None, batches[0]["pub"])
    again = wl.circuit.prove_batch(gens, wl.label, batches[1]["v"], batches[1]["v_blinding"], batches[1]["entropy"], pub=batches[1]["pub"])
    assert again[1].tobytes() == plain[1][1].tobytes()
    assert st.finish(0)[1].tobytes() == plain[0][1].tobytes()


def test_skewed_digit_distributions(api, gens, oracle_lib, n=100, cap=256):
    """explicit witnesses whose scalars all share their digits (every row of A_I lands in one or two buckets of the sorted-bucket
    MSM: the single-coarse-group path of the two-pass sort at n >= 4096 on the GPU), negative digits only, and zeros; proof bytes
    against the C oracle's prover on the same (unsatisfied -- the prover does not care) circuit"""
    kind, idx = [1, 0, 2, 3], [0, 0, n - 1, n // 2]
    coeff = api.scalars_to_array([1, L - 1, 7, 9])
    cons_ptr = [0, 2, 4]
    circ = api.Circuit.from_arrays(n, 1, cons_ptr, kind, idx, coeff)
    oc = CO.Circuit(n, 1, cons_ptr, kind, idx, coeff)
    rows = [([5] * n, [7] * n, [35] * n), ([2 ** 120 + 3] * n, [L - 1] * n, [0] * n), ([0] * n, [1] * n, [2 ** 252] * n)]
    B = len(rows)
    wit = [np.stack([api.scalars_to_array(r[j]) for r in rows]) for j in range(3)]
    v = api.scalars_to_array(H.rand_scalars(31, B)).reshape(B, 1, 32)
    vb = api.scalars_to_array(H.rand_scalars(32, B)).reshape(B, 1, 32)
    ent = np.frombuffer(bytes(range(100, 100 + 32 * B)), dtype=np.uint8).reshape(B, 32)
    V, P, st = circ.prove_batch(gens, b"skew", v, vb, ent, witness=tuple(wit))
    assert not st.any()
    for i in range(B):
        rc, oV, oP = CO.prove(oc, wit[0][i], wit[1][i], wit[2][i], v[i], vb[i], b"skew", ent[i].tobytes(), cap)
        assert rc == 0 and oV.tobytes() == V[i].tobytes() and oP == P[i].tobytes(), i


def test_slots_with_very_many_terms(api, gens, oracle_lib, n=20, cap=256):
    """variables that appear in hundreds of constraints (the constant slot of a Poseidon circuit holds one term per round key): their
    flattened weights are summed in parts of 256 terms (KFlattenParts / KFlattenSum); proof bytes against the C oracle's prover, and
    the verifiers (which flatten the same way) accept"""
    q = 1300
    rng = np.random.RandomState(5)
    kind, idx, co, cons_ptr = [], [], [], [0]
    for k in range(q):  # constraint k:  c1 * V_0 + c2 * (constant) + c3 * aL[k % n]  (+ c4 * aO[3] in every other one)
        kind += [0, 4, 1]; idx += [0, 0, k % n]; co += [int(x) for x in rng.randint(1, 2 ** 31, 3)]
        if k % 2:
            kind.append(3); idx.append(3); co.append(L - 1 - k)
        cons_ptr.append(len(kind))
    coeff = api.scalars_to_array(co)
    circ = api.Circuit.from_arrays(n, 1, cons_ptr, kind, idx, coeff)
    oc = CO.Circuit(n, 1, cons_ptr, kind, idx, coeff)
    B = 3
    wit = [np.stack([api.scalars_to_array(H.rand_scalars(40 + 3 * i + j, n)) for i in range(B)]) for j in range(3)]
    v = api.scalars_to_array(H.rand_scalars(31, B)).reshape(B, 1, 32)
    vb = api.scalars_to_array(H.rand_scalars(32, B)).reshape(B, 1, 32)
    ent = np.frombuffer(bytes(range(50, 50 + 32 * B)), dtype=np.uint8).reshape(B, 32)
    api.profile_enable(1)
    V, P, st = circ.prove_batch(gens, b"long", v, vb, ent, witness=tuple(wit))
    api.profile_enable(0)
    assert "KFlattenParts" in api.profile_report()
    assert not st.any()
    for i in range(B):
        rc, oV, oP = CO.prove(oc, wit[0][i], wit[1][i], wit[2][i], v[i], vb[i], b"long", ent[i].tobytes(), cap)
        assert rc == 0 and oV.tobytes() == V[i].tobytes() and oP == P[i].tobytes(), i
    # (the circuit is not satisfied by random wires: the verifiers must reject, identically, after flattening the same weights)
    assert circ.verify_batch(gens, b"long", V, P, ent).tolist() == [3] * B
    assert [CO.verify(oc, V[i], P[i].tobytes(), b"long", ent[i].tobytes(), cap) for i in range(B)] == [3] * B


def test_chunking_is_invisible(api, gens, monkeypatch):
    from bulletproofs_r1cs_gadgets_b200 import workloads
    wl = workloads.Mimc(gens, rounds=3)
    inp = wl.inputs(0, 5)
    V1, P1, _ = wl.circuit.prove_batch(gens, wl.label, inp["v"], inp["v_blinding"], inp["entropy"], pub=inp["pub"])
    monkeypatch.setenv("BP_B200_CHUNK", "2")  # ragged: 2 + 2 + 1
    V2, P2, _ = wl.circuit.prove_batch(gens, wl.label, inp["v"], inp["v_blinding"], inp["entropy"], pub=inp["pub"])
    assert P1.tobytes() == P2.tobytes() and V1.tobytes() == V2.tobytes()
    st = wl.circuit.verify_batch(gens, wl.label, V2, P2, inp["entropy"], pub=inp["pub"])
    assert not st.any()


def test_msm_path_choice_is_invisible(api, gens, monkeypatch):
    """the direct 8-bit tables (small instances) and the sorted-bucket path on the 15-bit shift table (>= 8192 rows per
    instance by default) must give the same proof bytes and the same verdicts: force each path on small circuits.  The
    sorted path also collapses the padding rows of L_0 into one row and, for inverse-S-box Poseidon circuits proven from the
    witness program, the equal-scalar rows of A_I into generator sums -- both are covered here."""
    from bulletproofs_r1cs_gadgets_b200 import workloads
    for wl in (workloads.Mimc(gens, rounds=5), workloads.PoseidonHash2(gens, api.SBOX_INVERSE, params=api.PoseidonParams(6, 2, 2, 3))):
        inp = wl.inputs(0, 3)
        outs = []
        for min_rows in ("1000000000", "0"):
            monkeypatch.setenv("BP_B200_SORTED_MIN_ROWS", min_rows)
            V, P, st = wl.circuit.prove_batch(gens, wl.label, inp["v"], inp["v_blinding"], inp["entropy"], pub=inp["pub"])
            assert not st.any()
            assert not wl.circuit.verify_batch(gens, wl.label, V, P, inp["entropy"], pub=inp["pub"]).any()
            bad = inp["pub"].copy(); bad[1, 0, 0] ^= 1
            assert wl.circuit.verify_batch(gens, wl.label, V, P, inp["entropy"], pub=bad).tolist() == [0, 3, 0]
            outs.append((V.tobytes(), P.tobytes()))
        assert outs[0] == outs[1], wl.name
        monkeypatch.setenv("BP_B200_NO_MERGE", "1"); monkeypatch.setenv("BP_B200_NO_PADSUM", "1")
        V, P, st = wl.circuit.prove_batch(gens, wl.label, inp["v"], inp["v_blinding"], inp["entropy"], pub=inp["pub"])
        assert (V.tobytes(), P.tobytes()) == outs[0]
        monkeypatch.delenv("BP_B200_NO_MERGE"); monkeypatch.delenv("BP_B200_NO_PADSUM")


def test_shift_table_only_generators(api, gens, monkeypatch):
    """Generators above ~65k capacity (the reference's own configuration, BulletproofGens::new(819200, 1) at
    src/gadget_vsmt_2.rs:290, needs 262144) carry the 15-bit shift table but no direct 8-bit tables: commitments through the sorted
    path, every inner-product round on folded generators.  Forced here at small capacity; the proofs must be the bytes the
    fully tabulated generators produce."""
    from bulletproofs_r1cs_gadgets_b200 import workloads
    monkeypatch.setenv("BP_B200_NO_DIRECT_TABLE", "1")
    lean = api.Gens(256)
    monkeypatch.delenv("BP_B200_NO_DIRECT_TABLE")
    for wl in (workloads.Mimc(gens, rounds=5), workloads.PoseidonHash2(gens, api.SBOX_INVERSE, params=api.PoseidonParams(6, 2, 2, 3)),
               workloads.Vsmt2(gens, depth=2, params=api.PoseidonParams(6, 2, 2, 3))):
        inp = wl.inputs(0, 3)
        want = wl.circuit.prove_batch(gens, wl.label, inp["v"], inp["v_blinding"], inp["entropy"], pub=inp["pub"])
        for min_rows in ("0", "1000000000"):  # sorted path for the commitments / bucket method on the generators
            monkeypatch.setenv("BP_B200_SORTED_MIN_ROWS", min_rows)
            V, P, st = wl.circuit.prove_batch(lean, wl.label, inp["v"], inp["v_blinding"], inp["entropy"], pub=inp["pub"])
            assert not st.any() and V.tobytes() == want[0].tobytes() and P.tobytes() == want[1].tobytes(), (wl.name, min_rows)
            assert not wl.circuit.verify_batch(lean, wl.label, V, P, inp["entropy"], pub=inp["pub"]).any()
            stc, comb = wl.circuit.verify_batch_combined(lean, wl.label, V, P, inp["entropy"], pub=inp["pub"])
            assert comb == 0 and not stc.any()
            bad = inp["pub"].copy(); bad[2, 0, 1] ^= 2
            assert wl.circuit.verify_batch(lean, wl.label, V, P, inp["entropy"], pub=bad).tolist() == [0, 0, 3]
        monkeypatch.delenv("BP_B200_SORTED_MIN_ROWS")


def test_fold_tables_of_large_capacities(api, gens, monkeypatch):
    """Without the 8-bit direct tables the level-4 generators are materialised from FOLD tables with narrower windows (4..7 bits,
    chosen by a memory budget; built on first use for the circuit's N) after four rounds over the shift table -- the path of the
    reference's own configuration (N = 262144, src/gadget_vsmt_2.rs:23,290).  Every window width must give the bytes the fully
    tabulated generators produce; a zero budget falls back to folding from round 0."""
    from bulletproofs_r1cs_gadgets_b200 import workloads
    monkeypatch.setenv("BP_B200_NO_DIRECT_TABLE", "1")
    lean = api.Gens(256)
    monkeypatch.delenv("BP_B200_NO_DIRECT_TABLE")
    monkeypatch.setenv("BP_B200_SORTED_MIN_ROWS", "0")
    wls = (workloads.Mimc(gens, rounds=5), workloads.Vsmt2(gens, depth=2, params=api.PoseidonParams(6, 2, 2, 3)))
    for wl in wls:
        inp = wl.inputs(0, 3)
        want = wl.circuit.prove_batch(gens, wl.label, inp["v"], inp["v_blinding"], inp["entropy"], pub=inp["pub"])
        for bits in ("4", "5", "6", "7", None):
            if bits is None:  # budget 0 on generators that have no fold table yet
                monkeypatch.delenv("BP_B200_FOLD_BITS")
                monkeypatch.setenv("BP_B200_FOLD_TABLE_GB", "0")
                monkeypatch.setenv("BP_B200_NO_DIRECT_TABLE", "1")
                target = api.Gens(256)
                monkeypatch.delenv("BP_B200_NO_DIRECT_TABLE")
            else:
                monkeypatch.setenv("BP_B200_FOLD_BITS", bits)
                target = lean
            api.profile_enable(1)
            V, P, st = wl.circuit.prove_batch(target, wl.label, inp["v"], inp["v_blinding"], inp["entropy"], pub=inp["pub"])
            api.profile_enable(0)
            ran = api.profile_report()
            if wl.circuit.n > 16:  # (the 5-round MiMC circuit has four rounds in all: nothing left to materialise)
                assert ("KFoldTable" in ran) == (bits is not None), (bits, sorted(ran))
            assert not st.any() and V.tobytes() == want[0].tobytes() and P.tobytes() == want[1].tobytes(), (wl.name, bits)
        monkeypatch.delenv("BP_B200_FOLD_TABLE_GB")
    # a smaller circuit on the same generators reads the table that was built for the larger one (table capacity > its N)
    small = workloads.Mimc(gens, rounds=20)
    assert small.circuit.n == 40 and wls[1].circuit.n > 64
    inp = small.inputs(0, 2)
    want = small.circuit.prove_batch(gens, small.label, inp["v"], inp["v_blinding"], inp["entropy"], pub=inp["pub"])
    monkeypatch.setenv("BP_B200_FOLD_BITS", "5")
    wls[1].circuit.prove_batch(lean, wls[1].label, *[wls[1].inputs(0, 1)[k] for k in ("v", "v_blinding", "entropy")], pub=wls[1].inputs(0, 1)["pub"])
    api.profile_enable(1)
    V, P, st = small.circuit.prove_batch(lean, small.label, inp["v"], inp["v_blinding"], inp["entropy"], pub=inp["pub"])
    api.profile_enable(0)
    ran = api.profile_report()
    assert "KFoldTable" in ran and "KTableBuild" not in ran, sorted(ran)
    assert not st.any() and V.tobytes() == want[0].tobytes() and P.tobytes() == want[1].tobytes()
    monkeypatch.delenv("BP_B200_FOLD_BITS")
    # the workspace can be dropped between batches and comes back on the next call
    wl = wls[0]
    wl.circuit.release_workspace()
    inp = wl.inputs(0, 3)
    V, P, st = wl.circuit.prove_batch(lean, wl.label, inp["v"], inp["v_blinding"], inp["entropy"], pub=inp["pub"])
    assert not st.any() and P.tobytes() == want_first(wl, gens, inp)


def want_first(wl, gens, inp):
    return wl.circuit.prove_batch(gens, wl.label, inp["v"], inp["v_blinding"], inp["entropy"], pub=inp["pub"])[1].tobytes()


def test_combined_verification(api, gens):
    """cross-proof batched verification: one combined verdict for the batch, per-proof status for structural failures only"""
    from bulletproofs_r1cs_gadgets_b200 import workloads
    wl = workloads.Mimc(gens, rounds=5)
    Bn = 37  # ragged against the 32-proof chunks of the scalar combination
    inp = wl.inputs(0, Bn)
    V, P, st = wl.circuit.prove_batch(gens, wl.label, inp["v"], inp["v_blinding"], inp["entropy"], pub=inp["pub"])
    assert not st.any()
    st, comb = wl.circuit.verify_batch_combined(gens, wl.label, V, P, inp["entropy"], pub=inp["pub"])
    assert not st.any() and comb == 0
    # a wrong public input (image of the hash) in one proof: structurally fine, combined check fails
    bad = inp["pub"].copy(); bad[33, 0, 0] ^= 1
    st, comb = wl.circuit.verify_batch_combined(gens, wl.label, V, P, inp["entropy"], pub=bad)
    assert not st.any() and comb == 3
    # a tampered scalar of the proof (t_x): same
    P2 = P.copy(); P2[5, 352] ^= 1
    st, comb = wl.circuit.verify_batch_combined(gens, wl.label, V, P2, inp["entropy"], pub=inp["pub"])
    assert comb == 3
    # structural failures are reported per proof and left out of the combination: the rest still passes
    P3 = P.copy(); P3[7, 352:384] = 0xff           # non-canonical t_x -> FormatError
    V3 = V.copy(); V3[9, 0, :] = 0xff              # undecodable commitment -> VerificationError
    st, comb = wl.circuit.verify_batch_combined(gens, wl.label, V3, P3, inp["entropy"], pub=inp["pub"])
    assert st[7] == 2 and st[9] == 3 and not np.delete(st, [7, 9]).any() and comb == 0
    # the combined verdict agrees with per-proof verification
    assert not wl.circuit.verify_batch(gens, wl.label, V, P, inp["entropy"], pub=inp["pub"]).any()
    # single proof
    st, comb = wl.circuit.verify_batch_combined(gens, wl.label, V[:1], P[:1], inp["entropy"][:1], pub=inp["pub"][:1])
    assert comb == 0 and not st.any()
    # Forgery regression: one valid proof submitted twice with the inner-product scalar a moved by +d and -d,
    # the SAME verifier entropy in both slots.  Weights drawn from each proof's own transcript RNG were equal for the two slots
    # (a and b are never absorbed by the transcript) and the two errors cancelled; weights derived from the digest of the whole
    # batch differ per slot and depend on a, so the combination must fail -- as each proof does on its own.
    off = wl.circuit.proof_len - 64  # the proof ends with a, b
    a = int.from_bytes(P[0, off:off + 32].tobytes(), "little")
    d = 0x1234567
    Pf = np.stack([P[0], P[0]]).copy()
    Pf[0, off:off + 32] = np.frombuffer(((a + d) % L).to_bytes(32, "little"), np.uint8)
    Pf[1, off:off + 32] = np.frombuffer(((a - d) % L).to_bytes(32, "little"), np.uint8)
    Vf, pubf, entf = np.stack([V[0], V[0]]), np.stack([inp["pub"][0], inp["pub"][0]]), np.stack([inp["entropy"][0], inp["entropy"][0]])
    assert wl.circuit.verify_batch(gens, wl.label, Vf, Pf, entf, pub=pubf).tolist() == [3, 3]
    st, comb = wl.circuit.verify_batch_combined(gens, wl.label, Vf, Pf, entf, pub=pubf)
    assert comb == 3
    # the same two slots holding the untouched proof pass; weights differ between slots although entropy and proof are equal
    st, comb = wl.circuit.verify_batch_combined(gens, wl.label, Vf, np.stack([P[0], P[0]]), entf, pub=pubf)
    assert comb == 0 and not st.any()


def test_static_commitments_are_checked_by_the_batch_verifiers(api, gens):
    """the Poseidon statics (0, 101, 0, 0 with blinding 0) are commitments the reference's verifier computes
    itself (allocate_statics_for_verifier, src/gadget_poseidon.rs:580-608).  A prover who chooses other values for those slots
    gets a proof that is VALID for its own V -- the batch verifiers must reject it because V differs from the fixed bytes."""
    from bulletproofs_r1cs_gadgets_b200 import workloads
    pp = api.PoseidonParams(6, 2, 2, 3)
    wl = workloads.PoseidonHash2(gens, sbox=api.SBOX_INVERSE, params=pp)
    inp = wl.inputs(0, 3)
    v, pub = inp["v"].copy(), inp["pub"].copy()
    v[1, 3] = api.scalars_to_array([77])[0]          # proof 1: padding lane 77 instead of 101 ...
    xl, xr = (int.from_bytes(v[1, j].tobytes(), "little") for j in (0, 1))
    image = G.poseidon_permutation([0, xl, xr, 77, 0, 0], G.PoseidonParams(6, 2, 2, 3), G.INVERSE)[1]
    pub[1, 0] = api.scalars_to_array([image])[0]     # ... and the image that goes with it: a true statement about a DIFFERENT hash
    V, P, st = wl.circuit.prove_batch(gens, wl.label, v, inp["v_blinding"], inp["entropy"], pub=pub)
    assert not st.any()
    assert V[0, 3].tobytes() == gens.commit(101, 0) and V[1, 3].tobytes() == gens.commit(77, 0)
    # the proof itself is sound for its own commitments: a verifier told nothing about fixed slots accepts it
    loose = api.Verifier(gens, wl.label)
    lv = [loose.commit(bytes(32)) for _ in range(6)]
    loose.poseidon_hash_2_gadget(pp, lv[0], lv[1], lv[2:6], api.SBOX_INVERSE, loose.public_input())
    assert loose.compile().verify_batch(gens, wl.label, V, P, inp["entropy"], pub=pub).tolist() == [0, 0, 0]
    # the circuit built with allocate_statics knows what slots 2..5 must hold
    assert wl.circuit.verify_batch(gens, wl.label, V, P, inp["entropy"], pub=pub).tolist() == [0, 3, 0]
    st, comb = wl.circuit.verify_batch_combined(gens, wl.label, V, P, inp["entropy"], pub=pub)
    assert st.tolist() == [0, 3, 0] and comb == 0   # reported per proof, left out of the combination


def test_wire_format(api, gens):
    """tagged wire form: one-phase proofs drop the three identity commitments; round trip; bad tags rejected"""
    case = H.golden_cases()[0]
    proof = bytes.fromhex(case["proof"])
    assert proof[96:192] == bytes(96)
    wire = api.proof_to_wire(proof)
    assert len(wire) == 1 + len(proof) - 96 and wire[0] == 0 and wire[1:97] == proof[:96] and wire[97:] == proof[192:]
    assert api.proof_from_wire(wire) == proof
    two = bytearray(proof); two[100] = 1            # pretend a second-phase commitment is present
    w2 = api.proof_to_wire(bytes(two))
    assert w2[0] == 1 and len(w2) == 1 + len(proof) and api.proof_from_wire(w2) == bytes(two)
    for bad in (b"", bytes([2]) + wire[1:], wire[:-1], wire + b"\x00" * 32):
        with pytest.raises(api.R1CSError) as e:
            api.proof_from_wire(bad)
        assert e.value.code == 2


def test_error_codes(api, gens):
    small = api.Gens(8)
    case = H.golden_cases()[0]  # mimc5: n = 10 -> N = 16 > 8
    build = H.product_builder(api, case)
    p = api.Prover(small, b"MiMC")
    vs = [p.commit(int(v, 16), int(b, 16))[1] for v, b in zip(case["values"], case["blindings"])]
    build(p, vs, True)
    with pytest.raises(api.R1CSError) as e:
        p.prove(bytes(32))
    assert e.value.code == 1  # InvalidGeneratorsLength
    p2 = api.Prover(gens, b"x")
    with pytest.raises(api.R1CSError) as e:
        p2.allocate_multiplier(None)
    assert e.value.code == 4  # MissingAssignment
    v = api.Verifier(gens, b"x")
    assert v.evaluate_lc(api.LinearCombination.of(5)) is None  # Option::None on the verifier
    assert v.allocate_multiplier(None)[2].kind == api.VAR_MULT_OUT
    with pytest.raises(api.R1CSError) as e:
        v.constrain(api.LinearCombination([(api.Variable(api.VAR_MULT_LEFT, 99), 1)]))
    assert e.value.code == 6
    # empty batch is a no-op; circuits without multipliers prove and verify (N = 1, no inner-product rounds)
    pr = api.Prover(gens, b"lin"); V, a = pr.commit(7, 3); pr.constrain(a - 7)
    proof = pr.prove(bytes(32))
    assert len(proof) == 32 * 16
    vf = api.Verifier(gens, b"lin"); b = vf.commit(V); vf.constrain(b - 7)
    assert vf.verify(proof, bytes(32))
    vf = api.Verifier(gens, b"lin"); b = vf.commit(V); vf.constrain(b - 8)
    with pytest.raises(api.R1CSError):
        vf.verify(proof, bytes(32))
    # wrongly shaped batch arrays are refused by the host layer (the C-ABI takes bare pointers): proof bytes -> FormatError,
    # everything else -> InvalidArgument; a witness program that reads public inputs refuses to run without them
    from bulletproofs_r1cs_gadgets_b200 import workloads
    wl = workloads.Mimc(gens, rounds=3)
    inp = wl.inputs(0, 2)
    V2, P2, st = wl.circuit.prove_batch(gens, wl.label, inp["v"], inp["v_blinding"], inp["entropy"], pub=inp["pub"])
    assert not st.any()
    for kw, code in ((dict(V=V2[:, :1]), 6), (dict(proofs=P2[:, :-32]), 2), (dict(pub=inp["pub"][:1]), 6), (dict(V=V2[:1]), 6)):
        args = dict(V=V2, proofs=P2, pub=inp["pub"]); args.update(kw)
        for fn in (wl.circuit.verify_batch, wl.circuit.verify_batch_combined):
            with pytest.raises(api.R1CSError) as e:
                fn(gens, wl.label, args["V"], args["proofs"], inp["entropy"], pub=args["pub"])
            assert e.value.code == code
    with pytest.raises(api.R1CSError) as e:
        wl.circuit.prove_batch(gens, wl.label, inp["v"][:, :1], inp["v_blinding"], inp["entropy"], pub=inp["pub"])
    assert e.value.code == 6


def test_single_multiplier_and_allocate_single(api, gens):
    """factors-style circuit p*q = r (reference src/factors.rs:12-21) and the fork's allocate_single pairing"""
    pq = (17, 19)
    pr = api.Prover(gens, b"Factors")
    cp = pr.commit(pq[0], 11); cq = pr.commit(pq[1], 12); cr = pr.commit(pq[0] * pq[1], 13)
    _, _, o = pr.multiply(cp[1] + 0, cq[1] + 0)
    pr.constrain(o - cr[1])
    l, none = pr.allocate_single(5)
    r, out = pr.allocate_single(9)
    assert none is None and out.kind == api.VAR_MULT_OUT and pr.evaluate_lc(api.LinearCombination.of(out)) == 45
    pr.constrain(out - 45)
    proof = pr.prove(bytes(32))
    op = R.Prover(R.PedersenGens(), R.Transcript(b"Factors"))
    a = op.commit(17, 11)[1]; b = op.commit(19, 12)[1]; c = op.commit(17 * 19, 13)[1]
    _, _, oo = op.multiply(R.LC.of(a) + 0, R.LC.of(b) + 0)
    op.constrain(R.LC.of(oo) - R.LC.of(c))
    op.allocate_single(5); _, o2 = op.allocate_single(9)
    op.constrain(R.LC.of(o2) - 45)
    assert R.proof_to_bytes(op.prove(R.BulletproofGens(256), bytes(32))) == proof
