"""Parity tests proper: the CUDA library on a real B200, through the C-ABI, against the oracle.
Run with  python -m pytest tests -m gpu.  The small-case bodies are shared with tests/test_host_emul.py."""
import ctypes as C
import os

import numpy as np
import pytest

import helpers as H
import test_host_emul as E
import test_libsodium_pin as S
from helpers import R, G, CO, L

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def api(product_so):
    from bulletproofs_r1cs_gadgets_b200 import api as a
    a._lib = None
    a.load(product_so)
    return a


@pytest.fixture(scope="module")
def gens(api):
    return api.Gens(256)


@pytest.fixture(scope="module")
def gens_big(api):
    return api.Gens(32768)


def test_primitives(api): E.test_primitives(api)
def test_generators(api, gens, oracle_lib): E.test_generators(api, gens, oracle_lib)
def _msm_device_pointers(api, gens, arr):
    import torch
    d_in = torch.from_numpy(arr).cuda()
    d_out = torch.zeros(32, dtype=torch.uint8, device="cuda")
    rc = api.load().bp_msm_gens_device(gens._h, arr.shape[0], C.c_void_p(d_in.data_ptr()), C.c_void_p(d_out.data_ptr()), None)
    torch.cuda.synchronize()
    return rc, d_out.cpu().numpy().tobytes()


def test_msm_entry(api, gens): E.test_msm_entry(api, gens, call=_msm_device_pointers)


# third-party pin (libsodium, tests/test_libsodium_pin.py): device scalar field, Pedersen commitments, and the MSM kernels of
# both paths (direct tables at 200 rows; sort / KBucketAccumulate / KBucketReduce at 40000 rows, ragged last sub-instance)
def test_libsodium_scalar_field(api): S.check_device_scalar_field(api, 1000)
def test_libsodium_commit_and_one_way_map(api, gens): S.check_device_from_uniform_and_commit(api, gens, 16)
def test_libsodium_msm_table_path(api, gens): S.check_device_msm(api, gens, 200, _msm_device_pointers)
def test_libsodium_msm_sorted_path(api): S.check_device_msm(api, api.Gens(1 << 16), 40000, _msm_device_pointers)
def test_golden_proofs_tier1(api, gens): E.test_golden_proofs_tier1(api, gens)
def test_python_gadget_code_drives_product_cs(api, gens): E.test_python_gadget_code_drives_product_cs(api, gens)
def test_verifier_rejects_tampering(api, gens): E.test_verifier_rejects_tampering(api, gens)
def test_batch_witness_program_and_public_inputs(api, gens): E.test_batch_witness_program_and_public_inputs(api, gens)
def test_batch_aux_inputs_bound_check(api, gens): E.test_batch_aux_inputs_bound_check(api, gens)
def test_explicit_witness_equals_witness_program(api, gens, oracle_lib): E.test_explicit_witness_equals_witness_program(api, gens, oracle_lib)
def test_slots_with_very_many_terms(api, gens, oracle_lib): E.test_slots_with_very_many_terms(api, gens, oracle_lib)
def test_chunking_is_invisible(api, gens, monkeypatch): E.test_chunking_is_invisible(api, gens, monkeypatch)
def test_msm_path_choice_is_invisible(api, gens, monkeypatch): E.test_msm_path_choice_is_invisible(api, gens, monkeypatch)
def test_combined_verification(api, gens): E.test_combined_verification(api, gens)
def test_shift_table_only_generators(api, gens, monkeypatch): E.test_shift_table_only_generators(api, gens, monkeypatch)
def test_fold_tables_of_large_capacities(api, gens, monkeypatch): E.test_fold_tables_of_large_capacities(api, gens, monkeypatch)
def test_wire_format(api, gens): E.test_wire_format(api, gens)
def test_static_commitments_are_checked_by_the_batch_verifiers(api, gens): E.test_static_commitments_are_checked_by_the_batch_verifiers(api, gens)
def test_vsmt4_membership_small(api, gens): E.test_vsmt4_membership(api, gens)
def test_vsmt4_membership_reference_parameters(api, gens_big, oracle_lib): E.test_vsmt4_membership(api, gens_big, levels=3, params=(6, 4, 4, 140), count=2, c_oracle_prover=True)
def test_sparse_merkle_tree_and_membership_from_a_real_tree(api, gens): E.test_sparse_merkle_tree_and_membership_from_a_real_tree(api, gens)
def test_device_tree_small(api, gens): E.test_device_tree(api, gens)
def test_device_tree_depth5(api, gens): E.test_device_tree_depth5(api, gens)
def test_device_tree_depth63(api, gens): E.test_device_tree_depth63(api, gens)
def test_device_tree_depth253(api, gens): E.test_device_tree_depth253(api, gens)
def test_device_tree_depth253_reference_parameters(api, gens, oracle_lib):
    """the reference's own tree (TreeDepth = 253, src/gadget_vsmt_2.rs:23, Poseidon 4+140+4 inverse) on the device against the oracle's
    one-key-at-a-time tree hashing with the C oracle: roots, leaves, sibling paths, witness rows"""
    E.test_device_tree(api, gens, oracle_lib=oracle_lib, depth=253, params=(6, 4, 4, 140), nkeys=8, prove=False, seed=2530)
def test_device_tree_reference_parameters(api, gens_big, oracle_lib):
    """depth 32, Poseidon 4+140+4 inverse: device tree vs the oracle's one-key-at-a-time tree hashing with the C oracle; membership proofs from the tree"""
    E.test_device_tree(api, gens_big, oracle_lib=oracle_lib, depth=32, params=(6, 4, 4, 140), nkeys=24, prove=True, seed=901)
def test_streamed_batches_equal_plain_batches(api, gens): E.test_streamed_batches_equal_plain_batches(api, gens)
def test_skewed_digit_distributions_small(api, gens, oracle_lib): E.test_skewed_digit_distributions(api, gens, oracle_lib)
def test_skewed_digit_distributions_sorted_path(api, gens_big, oracle_lib): E.test_skewed_digit_distributions(api, gens_big, oracle_lib, n=6000, cap=8192)
def test_error_codes(api, gens): E.test_error_codes(api, gens)
def test_single_multiplier_and_allocate_single(api, gens): E.test_single_multiplier_and_allocate_single(api, gens)


@pytest.mark.parametrize("path", ["tables", "buckets"])
def test_msm_entry_larger_sizes(api, gens_big, oracle_lib, path, monkeypatch):
    """MSM microbenchmark entry (BASELINE config 3) against the C oracle's Pippenger, device buffers; both the fixed-base
    table path and the bucket method (the path capacities above ~49k generators use), including their split reductions"""
    import torch
    if path == "buckets":
        monkeypatch.setenv("BP_B200_MSM_BUCKET", "1")
    og = np.zeros((8192, 32), np.uint8)
    oracle_lib.lib().bpo_gens_compressed(0, 8192, og.ctypes.data_as(CO.u8p))
    for n, seed in ((1, 1), (33, 2), (1024, 3), (4096, 4), (8192, 5)):
        sc = H.rand_scalars(seed, n)
        if n == 1024:
            sc = [s if i % 2 else 0 for i, s in enumerate(sc)]  # 50 % zeros (inverse-S-box a_R shape)
        arr = api.scalars_to_array(sc)
        d_in = torch.from_numpy(arr).cuda()
        d_out = torch.zeros(32, dtype=torch.uint8, device="cuda")
        assert api.load().bp_msm_gens_device(gens_big._h, n, C.c_void_p(d_in.data_ptr()), C.c_void_p(d_out.data_ptr()), None) == 0
        torch.cuda.synchronize()
        out = np.zeros(32, np.uint8)
        assert oracle_lib.lib().bpo_msm(n, arr.ctypes.data_as(CO.u8p), og.ctypes.data_as(CO.u8p), out.ctypes.data_as(CO.u8p)) == 0
        assert d_out.cpu().numpy().tobytes() == out.tobytes(), n


def test_msm_entry_split_sorted_path(api, oracle_lib):
    """n >= 32768 goes through the split sorted-bucket path (sub-instances of consecutive rows through the batched sort /
    accumulate / reduce kernels, partial results summed): bit-exact against the C oracle's Pippenger at 32768 and 40000
    generators (ragged last sub-instance), and linear in the scalars at 2^17 (sum of two MSMs = MSM of the sums)"""
    import torch
    g = api.Gens(1 << 17)
    og = np.zeros((40000, 32), np.uint8)
    oracle_lib.lib().bpo_ensure_gens(40000)
    oracle_lib.lib().bpo_gens_compressed(0, 40000, og.ctypes.data_as(CO.u8p))

    def msm(sc):
        arr = api.scalars_to_array(sc)
        d_in = torch.from_numpy(arr).cuda()
        d_out = torch.zeros(32, dtype=torch.uint8, device="cuda")
        assert api.load().bp_msm_gens_device(g._h, len(sc), C.c_void_p(d_in.data_ptr()), C.c_void_p(d_out.data_ptr()), None) == 0
        torch.cuda.synchronize()
        return arr, d_out.cpu().numpy().tobytes()
    for n, seed in ((32768, 11), (40000, 12)):
        sc = H.rand_scalars(seed, n)
        sc[5] = 0; sc[6] = 1; sc[7] = L - 1; sc[n - 1] = 2 ** 252
        arr, got = msm(sc)
        out = np.zeros(32, np.uint8)
        assert oracle_lib.lib().bpo_msm(n, arr.ctypes.data_as(CO.u8p), og.ctypes.data_as(CO.u8p), out.ctypes.data_as(CO.u8p)) == 0
        assert got == out.tobytes(), n
    n = 1 << 17
    a, b = H.rand_scalars(21, n), H.rand_scalars(22, n)
    _, Pa = msm(a); _, Pb = msm(b); _, Pab = msm([(x + y) % L for x, y in zip(a, b)])
    assert R.ristretto_encode(R.pt_add(R.ristretto_decode(Pa), R.ristretto_decode(Pb))) == Pab


def _oracle_circuit(build, m, label):
    v = R.Verifier(R.Transcript(label)); vs = [v.commit(bytes(32)) for _ in range(m)]; build(v, vs)
    return CO.Circuit.from_cs(v, m)


def _check_batch_against_c_oracle(api, gens, wl, ocirc, witness_fn, B, cap, first=0):
    inp = wl.inputs(first, B)
    V, P, st = wl.circuit.prove_batch(gens, wl.label, inp["v"], inp["v_blinding"], inp["entropy"], aux=inp.get("aux"), pub=inp.get("pub"))
    assert not st.any()
    assert (ocirc.n, ocirc.q, ocirc.m) == (wl.circuit.n, wl.circuit.q, wl.circuit.m)
    for i in range(B):
        aL, aR, aO = witness_fn(inp, i)
        rc, oV, oP = CO.prove(ocirc, aL, aR, aO, inp["v"][i], inp["v_blinding"][i], wl.label, inp["entropy"][i].tobytes(), cap)
        assert rc == 0 and oV.tobytes() == V[i].tobytes(), "commitments differ from the oracle (proof %d)" % i
        assert oP == P[i].tobytes(), "proof bytes differ from the oracle (proof %d)" % i
    ok = wl.circuit.verify_batch(gens, wl.label, V, P, inp["entropy"], pub=inp.get("pub"))
    assert not ok.any()
    return inp, V, P


@pytest.mark.parametrize("sbox", [0, 1])
def test_poseidon_2to1_reference_parameters(api, gens_big, oracle_lib, sbox):
    """BASELINE config 2 shape: width 6, 4+140+4 rounds, both S-boxes (reference src/gadget_poseidon.rs:691-785, :889-894)"""
    from bulletproofs_r1cs_gadgets_b200 import workloads
    oracle_lib.poseidon_set_params(H.POSEIDON_BLOB)
    wl = workloads.PoseidonHash2(gens_big, sbox)
    assert (wl.circuit.n, wl.circuit.q) == ((376, 753), (564, 1317))[sbox]
    pp = G.PoseidonParams()
    # the constant in the last constraint does not enter the proof: record the oracle circuit with hash = 0
    oc = _oracle_circuit(lambda cs, v: G.poseidon_hash_2_gadget(cs, v[0], v[1], v[2:], pp, sbox, 0), 6, wl.label)

    def wit(inp, i):
        aL, aR, aO, h = oracle_lib.poseidon_hash2_witness(inp["v"][i][0].tobytes(), inp["v"][i][1].tobytes(), sbox, oc.n)
        assert h == inp["pub"][i][0].tobytes()
        return aL, aR, aO
    _check_batch_against_c_oracle(api, gens_big, wl, oc, wit, 4, 1024)


def test_mimc_322(api, gens_big, oracle_lib):
    """BASELINE config 4 shape (reference src/gadget_mimc.rs:92-175)"""
    from bulletproofs_r1cs_gadgets_b200 import workloads
    wl = workloads.Mimc(gens_big)
    assert (wl.circuit.n, wl.circuit.q, wl.circuit.m) == (644, 1289, 2)
    oc = _oracle_circuit(lambda cs, v: G.mimc_gadget(cs, v[0], v[1], 322, wl.constants, 0), 2, wl.label)
    cb = CO.scalars_to_array(wl.constants, L).tobytes()

    def wit(inp, i):
        aL, aR, aO, img = oracle_lib.mimc_witness(inp["v"][i][0].tobytes(), inp["v"][i][1].tobytes(), cb, oc.n)
        assert img == inp["pub"][i][0].tobytes()
        return aL, aR, aO
    _check_batch_against_c_oracle(api, gens_big, wl, oc, wit, 3, 1024)


def test_bound_check_64bit(api, gens, oracle_lib):
    """BASELINE config 1 (reference src/gadget_bound_check.rs:49-116): n = 128 = N, no padding"""
    from bulletproofs_r1cs_gadgets_b200 import workloads
    wl = workloads.BoundCheck(gens)
    assert (wl.circuit.n, wl.circuit.q, wl.circuit.m) == (128, 261, 3)
    oc = _oracle_circuit(lambda cs, v: G.bound_check_gadget(cs, (v[0], None), (v[1], None), (v[2], None), 2 ** 64 - 1, 0, 64), 3, wl.label)

    def wit(inp, i):
        vals = [int.from_bytes(inp["v"][i][j].tobytes(), "little") for j in range(3)]
        p = R.Prover(R.PedersenGens(), R.Transcript(b"x")); vs = [p.commit(v, 0)[1] for v in vals]
        G.bound_check_gadget(p, (vs[0], vals[0]), (vs[1], vals[1]), (vs[2], vals[2]), 2 ** 64 - 1, 0, 64)
        return tuple(CO.scalars_to_array(a, L) for a in (p.aL, p.aR, p.aO))
    _check_batch_against_c_oracle(api, gens, wl, oc, wit, 3, 128)


@pytest.mark.parametrize("depth,B", [(3, 3), (32, 2)])
def test_vsmt2_membership(api, gens_big, oracle_lib, depth, B):
    """BASELINE config 5 (headline): Poseidon VSMT-2 membership, inverse S-box; depth 32 is the full-size case"""
    from bulletproofs_r1cs_gadgets_b200 import workloads
    oracle_lib.poseidon_set_params(H.POSEIDON_BLOB)
    wl = workloads.Vsmt2(gens_big, depth=depth)
    assert (wl.circuit.n, wl.circuit.q, wl.circuit.m) == (568 * depth, 1324 * depth + 1, 2 * depth + 5)
    pp = G.PoseidonParams()
    oc = _oracle_circuit(lambda cs, v: G.vanilla_merkle_tree_verif_gadget(cs, depth, 0, v[0], v[1:1 + depth], v[1 + depth:1 + 2 * depth], v[1 + 2 * depth:], pp),
                         2 * depth + 5, wl.label)

    def wit(inp, i):
        leaf, bits, sibs = wl.witness_values(i)
        aL, aR, aO, root = oracle_lib.vsmt2_witness(depth, api.scalar_bytes(leaf), bits, api.scalars_to_array(sibs), oc.n)
        assert root == inp["pub"][i][0].tobytes()
        return aL, aR, aO
    N = 1
    while N < oc.n:
        N *= 2
    inp, V, P = _check_batch_against_c_oracle(api, gens_big, wl, oc, wit, B, N)
    # wrong root, wrong sibling commitment, tampered proof bytes are rejected, proof by proof
    bad = inp["pub"].copy(); bad[0, 0, 5] ^= 4
    assert wl.circuit.verify_batch(gens_big, wl.label, V, P, inp["entropy"], pub=bad).tolist() == [3] + [0] * (B - 1)
    V2 = V.copy(); V2[B - 1, 1 + depth] = V2[B - 1, 2 + depth]
    assert wl.circuit.verify_batch(gens_big, wl.label, V2, P, inp["entropy"], pub=inp["pub"]).tolist() == [0] * (B - 1) + [3]
    P2 = P.copy(); P2[0, 500] ^= 1
    assert wl.circuit.verify_batch(gens_big, wl.label, V, P2, inp["entropy"], pub=inp["pub"])[0] == 3


def test_vsmt2_depth32_batch_properties(api, gens_big):
    """full-size batch: prove -> verify round trip, determinism, host-buffer and device-buffer entry points agree,
    results independent of how the batch is chunked"""
    import torch
    from bulletproofs_r1cs_gadgets_b200 import workloads
    wl = workloads.Vsmt2(gens_big, depth=32)
    B = 24
    inp = wl.inputs(1000, B, with_root=False)
    roots = wl.inputs(1000, 3)["pub"]  # native roots for the first three proofs only (host Poseidon is slow)
    V, P, st = wl.circuit.prove_batch(gens_big, wl.label, inp["v"], inp["v_blinding"], inp["entropy"])
    assert not st.any()
    V_again, P_again, _ = wl.circuit.prove_batch(gens_big, wl.label, inp["v"], inp["v_blinding"], inp["entropy"])
    assert P.tobytes() == P_again.tobytes() and V.tobytes() == V_again.tobytes()
    os.environ["BP_B200_CHUNK"] = "7"
    try:
        V_c, P_c, _ = wl.circuit.prove_batch(gens_big, wl.label, inp["v"], inp["v_blinding"], inp["entropy"])
    finally:
        del os.environ["BP_B200_CHUNK"]
    assert P.tobytes() == P_c.tobytes() and V.tobytes() == V_c.tobytes()
    d = {k: torch.from_numpy(inp[k]).cuda() for k in ("v", "v_blinding", "entropy")}
    dV = torch.empty((B, wl.circuit.m, 32), dtype=torch.uint8, device="cuda")
    dP = torch.empty((B, wl.circuit.proof_len), dtype=torch.uint8, device="cuda")
    dS = torch.empty((B,), dtype=torch.int32, device="cuda")
    p = lambda t: C.c_void_p(t.data_ptr())
    rc = api.load().bp_prove_batch_device(gens_big._h, wl.circuit._h, C.c_uint32(B), api._buf(wl.label), C.c_size_t(len(wl.label)), p(d["v"]), p(d["v_blinding"]),
                                          p(d["entropy"]), None, None, None, None, None, p(dV), p(dP), p(dS), C.c_void_p(torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    assert rc == 0 and dP.cpu().numpy().tobytes() == P.tobytes() and not dS.cpu().numpy().any()
    ok = wl.circuit.verify_batch(gens_big, wl.label, V[:3], P[:3], inp["entropy"][:3], pub=roots)
    assert not ok.any()
    assert len({P[i].tobytes() for i in range(B)}) == B  # distinct statements, distinct proofs
    assert P.shape[1] == 1472 and all(P[i, 96:192].tobytes() == bytes(96) for i in range(B))  # single-phase: A_I2, A_O2, S2 are the identity


def test_vsmt2_depth32_chunk_border_vs_c_oracle(api, gens_big, oracle_lib, monkeypatch):
    """depth-32 parity on a batch of 64 proofs that crosses device-chunk borders (chunks of 24, 24, 16): every
    proof and every commitment byte-equal to the C oracle's (native witness + prove, all host cores), on the streamed entry points
    the bench uses as well as on the plain call (reference flow src/gadget_vsmt_2.rs:262-399)"""
    import torch
    from bulletproofs_r1cs_gadgets_b200 import workloads
    oracle_lib.poseidon_set_params(H.POSEIDON_BLOB)
    wl = workloads.Vsmt2(gens_big, depth=32)
    B = 64
    inp = wl.inputs(5000, B, with_root=False)
    pp = G.PoseidonParams()
    oc = _oracle_circuit(lambda cs, v: G.vanilla_merkle_tree_verif_gadget(cs, 32, 0, v[0], v[1:33], v[33:65], v[65:], pp), 69, wl.label)
    st_o, V_o, P_o = CO.prove_batch(oc, B, inp["v"], inp["v_blinding"], inp["entropy"], wl.label, 32768, os.cpu_count() or 1, witness_kind=1, depth=32)
    assert not st_o.any()
    monkeypatch.setenv("BP_B200_CHUNK", "24")
    V, P, st = wl.circuit.prove_batch(gens_big, wl.label, inp["v"], inp["v_blinding"], inp["entropy"])
    assert not st.any()
    assert V.tobytes() == V_o.tobytes(), "commitments differ from the C oracle"
    bad = [i for i in range(B) if P[i].tobytes() != P_o[i].tobytes()]
    assert not bad, "proofs differ from the C oracle: %s" % bad
    ps = api.ProveStream(wl.circuit, gens_big, wl.label)
    ps.begin(0, inp["v"][:40], inp["v_blinding"][:40], inp["entropy"][:40])
    ps.begin(1, inp["v"][40:], inp["v_blinding"][40:], inp["entropy"][40:])
    Va, Pa, sa = ps.finish(0)
    Vb, Pb, sb = ps.finish(1)
    assert not sa.any() and not sb.any()
    assert Pa.tobytes() + Pb.tobytes() == P_o.tobytes() and Va.tobytes() + Vb.tobytes() == V_o.tobytes()
    # the whole batch through both verifiers with device-hashed roots; one wrong root is found by each
    monkeypatch.delenv("BP_B200_CHUNK")  # one combination per call: the 64 proofs must share a device chunk
    pub = wl.roots_batch(inp["v"])
    assert pub[:2].tobytes() == wl.inputs(5000, 2)["pub"].tobytes()  # device-batched roots = the host's level-by-level hashing
    assert not wl.circuit.verify_batch(gens_big, wl.label, V, P, inp["entropy"], pub=pub).any()
    stc, comb = wl.circuit.verify_batch_combined(gens_big, wl.label, V, P, inp["entropy"], pub=pub)
    assert comb == 0 and not stc.any()
    badpub = pub.copy(); badpub[41, 0, 3] ^= 0x10
    assert wl.circuit.verify_batch(gens_big, wl.label, V, P, inp["entropy"], pub=badpub).tolist() == [0] * 41 + [3] + [0] * 22
    stc, comb = wl.circuit.verify_batch_combined(gens_big, wl.label, V, P, inp["entropy"], pub=badpub)
    assert comb == 3 and not stc.any()


def test_vsmt2_depth253_reference_configuration(api, oracle_lib):
    """The reference's own test configuration (src/gadget_vsmt_2.rs:23,262-399): TreeDepth = 253, inverse S-box, 4+140+4 rounds,
    n = 143704 multipliers -> N = 262144 of 819200 generators (shift table + 6-bit fold tables, no direct tables), m = 511 commitments,
    label b"VSMT".
    Two proofs, byte-equal to the C oracle's; both verifiers accept them and reject a wrong root."""
    from bulletproofs_r1cs_gadgets_b200 import workloads
    oracle_lib.poseidon_set_params(H.POSEIDON_BLOB)
    depth, B, cap = 253, 2, 1 << 18
    g = api.Gens(819200)  # BulletproofGens::new(819200, 1), src/gadget_vsmt_2.rs:290: capacity well above the N = 2^18 the proof uses (the C oracle builds the first 2^18)
    wl = workloads.Vsmt2(g, depth=depth)
    assert (wl.circuit.n, wl.circuit.q, wl.circuit.m, wl.circuit.proof_len) == (143704, 334973, 511, 1664)
    inp = wl.inputs(7000, B, with_root=False)
    pp = G.PoseidonParams()
    oc = _oracle_circuit(lambda cs, v: G.vanilla_merkle_tree_verif_gadget(cs, depth, 0, v[0], v[1:1 + depth], v[1 + depth:1 + 2 * depth], v[1 + 2 * depth:], pp),
                         2 * depth + 5, wl.label)
    assert (oc.n, oc.q, oc.m) == (wl.circuit.n, wl.circuit.q, wl.circuit.m)
    st_o, V_o, P_o = CO.prove_batch(oc, B, inp["v"], inp["v_blinding"], inp["entropy"], wl.label, cap, B, witness_kind=1, depth=depth)
    assert not st_o.any()
    V, P, st = wl.circuit.prove_batch(g, wl.label, inp["v"], inp["v_blinding"], inp["entropy"])
    assert not st.any()
    assert V.tobytes() == V_o.tobytes(), "commitments differ from the C oracle"
    assert P.tobytes() == P_o.tobytes(), "proof bytes differ from the C oracle"
    pub = wl.roots_batch(inp["v"])
    assert not wl.circuit.verify_batch(g, wl.label, V, P, inp["entropy"], pub=pub).any()
    stc, comb = wl.circuit.verify_batch_combined(g, wl.label, V, P, inp["entropy"], pub=pub)
    assert comb == 0 and not stc.any()
    bad = pub.copy(); bad[1, 0, 0] ^= 1
    assert wl.circuit.verify_batch(g, wl.label, V, P, inp["entropy"], pub=bad).tolist() == [0, 3]
    assert wl.circuit.verify_batch_combined(g, wl.label, V, P, inp["entropy"], pub=bad)[1] == 3


def test_msm_2_pow_20_linearity_and_splitting(api):
    """MSM parity at the benchmarked size (BASELINE config 3, 2^20 generators): linear in the scalars, and equal to the sum of
    the results over two halves of the rows (the second half shifted in by zero scalars) -- size-independent properties"""
    import torch
    n = 1 << 20
    g = api.Gens(n)
    import hashlib

    def scal(tag):
        raw = np.frombuffer(hashlib.shake_256(b"msm-2^20/" + tag).digest(32 * n), dtype=np.uint8).reshape(n, 32).copy()
        raw[:, 31] &= 0x0f
        return raw

    def msm(arr):
        d_in = torch.from_numpy(np.ascontiguousarray(arr)).cuda()
        d_out = torch.zeros(32, dtype=torch.uint8, device="cuda")
        assert api.load().bp_msm_gens_device(g._h, arr.shape[0], C.c_void_p(d_in.data_ptr()), C.c_void_p(d_out.data_ptr()), None) == 0
        torch.cuda.synchronize()
        return d_out.cpu().numpy().tobytes()
    a, b = scal(b"a"), scal(b"b")
    ia = [int.from_bytes(a[i].tobytes(), "little") for i in range(n)]
    ib = [int.from_bytes(b[i].tobytes(), "little") for i in range(n)]
    s = api.scalars_to_array([(x + y) % L for x, y in zip(ia, ib)])
    Pa, Pb, Ps = msm(a), msm(b), msm(s)
    add = lambda x, y: R.ristretto_encode(R.pt_add(R.ristretto_decode(x), R.ristretto_decode(y)))
    assert add(Pa, Pb) == Ps
    lo, hi = a.copy(), a.copy()
    lo[n // 2:] = 0; hi[:n // 2] = 0
    assert add(msm(lo), msm(hi)) == Pa
    assert msm(a[:n // 2]) == msm(lo)  # trailing zero rows change nothing


def test_device_tree_large_batch_properties(api, oracle_lib):
    """2^14 random updates at depth 32 (full Poseidon): the root does not depend on how the updates are batched, get returns what
    update stored, paths fetched from the device verify against the device root under the ORACLE's verify_proof, and the witness
    rows written into device buffers equal the host-buffer ones"""
    import random
    import torch
    from bulletproofs_r1cs_gadgets_b200 import trees
    from oracle import tree_pyref as TP
    depth, K = 32, 1 << 14
    pp = api.PoseidonParams()
    rnd = random.Random(4242)
    keys = [rnd.randrange(2 ** depth) for _ in range(K)]
    keys[100] = keys[7]  # one repeated key
    vals = api.scalars_to_array(H.rand_scalars(4243, K))
    one = trees.DeviceVsmt2(pp, depth)
    r_one = one.update_batch(keys, vals)
    parts = trees.DeviceVsmt2(pp, depth)
    for a in range(0, K, 5000):
        r_parts = parts.update_batch(keys[a:a + 5000], vals[a:a + 5000])
    assert r_one == r_parts == one.root == parts.root
    assert one.num_nodes == parts.num_nodes and K * (depth - 14) < one.num_nodes <= K * (depth + 1)
    leaves, proofs = one.get_batch(keys)
    last = {k: i for i, k in enumerate(keys)}
    assert all(leaves[i].tobytes() == vals[last[k]].tobytes() for i, k in enumerate(keys))
    oracle_lib.poseidon_set_params(H.POSEIDON_BLOB)
    checker = TP.VanillaSparseMerkleTree.__new__(TP.VanillaSparseMerkleTree)
    checker.depth, checker.hash2, checker.root = depth, TP.c_oracle_hash2(oracle_lib, 1), r_one
    for i in (0, 7, 100, K - 1):
        path = [int.from_bytes(proofs[i, j].tobytes(), "little") for j in range(depth)]
        assert checker.verify_proof(keys[i], int.from_bytes(leaves[i].tobytes(), "little"), path)
        assert not checker.verify_proof(keys[i] ^ 1, int.from_bytes(leaves[i].tobytes(), "little"), path)
    v, pub = one.witness_rows(keys[:64])
    d_idx = torch.tensor(keys[:64], dtype=torch.int64, device="cuda")
    d_v = torch.zeros((64, 2 * depth + 5, 32), dtype=torch.uint8, device="cuda")
    d_pub = torch.zeros((64, 1, 32), dtype=torch.uint8, device="cuda")
    one.witness_rows_device(d_idx, d_v, d_pub, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert d_v.cpu().numpy().tobytes() == v.tobytes() and d_pub.cpu().numpy().tobytes() == pub.tobytes()
