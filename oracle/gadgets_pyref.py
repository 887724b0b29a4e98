"""TEST INFRASTRUCTURE ONLY -- Python restatement of the reference's gadget layer.

Each function cites the reference file:line it follows (paths relative to
/root/reference/src).  Generic over the Prover/Verifier of oracle/bp_pyref.py, exactly
as the reference gadgets are generic over `CS: ConstraintSystem`.
"""
import os
from .bp_pyref import L, LC, Var, VAR_ONE, inv_mod_l, COMMITTED

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..",
                     "bulletproofs_r1cs_gadgets_b200", "data", "poseidon_constants.bin")


class PoseidonParams:
    """gadget_poseidon.rs:27-94; constants loaded with the LE semantics of scalar_utils.rs:232-237."""

    def __init__(self, width=6, full_rounds_beginning=4, full_rounds_end=4, partial_rounds=140):
        blob = open(_DATA, "rb").read()
        vals = [int.from_bytes(blob[i: i + 32], "little") for i in range(0, len(blob), 32)]
        assert width == 6
        self.width = width
        self.full_rounds_beginning = full_rounds_beginning
        self.full_rounds_end = full_rounds_end
        self.partial_rounds = partial_rounds
        total = full_rounds_beginning + partial_rounds + full_rounds_end
        assert total * width <= 960
        self.MDS = [[vals[i * 6 + j] for j in range(6)] for i in range(6)]
        self.round_keys = vals[36: 36 + total * width]


CUBE, INVERSE = 0, 1


def apply_sbox(sbox, x):  # gadget_poseidon.rs:120-125 (Scalar::invert(0) == 0)
    return x * x % L * x % L if sbox == CUBE else inv_mod_l(x)


def poseidon_permutation(inp, params, sbox):  # gadget_poseidon.rs:189-280
    w = params.width
    st = [v % L for v in inp]
    off = 0
    fb, pr, fe = params.full_rounds_beginning, params.partial_rounds, params.full_rounds_end
    for rnd in range(fb + pr + fe):
        full = rnd < fb or rnd >= fb + pr
        for i in range(w):
            st[i] = (st[i] + params.round_keys[off]) % L
            off += 1
            if full:
                st[i] = apply_sbox(sbox, st[i])
        if not full:
            st[w - 1] = apply_sbox(sbox, st[w - 1])
        st = [sum(st[j] * params.MDS[i][j] for j in range(w)) % L for i in range(w)]
    return st


PADDING_CONST = 101
ZERO_CONST = 0


def poseidon_hash_2(xl, xr, params, sbox):  # gadget_poseidon.rs:428-443
    return poseidon_permutation([ZERO_CONST, xl, xr, PADDING_CONST, ZERO_CONST, ZERO_CONST], params, sbox)[1]


def poseidon_hash_4(inputs, params, sbox):  # gadget_poseidon.rs:488-503
    return poseidon_permutation([ZERO_CONST] + list(inputs) + [PADDING_CONST], params, sbox)[1]


def constrain_lc_with_scalar(cs, lc, scalar):  # r1cs_utils.rs:51-53
    cs.constrain(LC.of(lc) - LC.of(scalar % L))


def is_nonzero_gadget(cs, x_var, x_inv_var):  # gadget_zero_nonzero.rs:46-66
    x_lc = LC.of(x_var)
    y_lc = LC.of(1)
    one_minus_y = LC.of(VAR_ONE) - y_lc
    _, _, o1 = cs.multiply(x_lc, one_minus_y)
    cs.constrain(LC.of(o1))
    _, _, o2 = cs.multiply(x_lc, LC([(x_inv_var, 1)]))
    cs.constrain(LC.of(o2) - y_lc)


def synthesize_sbox(cs, sbox, input_lc, round_key):
    inp = LC.of(input_lc) + LC.of(round_key)
    if sbox == CUBE:  # gadget_poseidon.rs:141-150
        i, _, sqr = cs.multiply(inp, inp)
        _, _, cube = cs.multiply(LC.of(sqr), LC.of(i))
        return cube
    # gadget_poseidon.rs:153-185
    val_l = cs.evaluate_lc(inp)
    val_r = None if val_l is None else inv_mod_l(val_l)
    var_l, _ = cs.allocate_single(val_l)
    var_r, var_o = cs.allocate_single(val_r)
    is_nonzero_gadget(cs, var_l, var_r)
    constrain_lc_with_scalar(cs, LC.of(var_o), 1)
    return var_r


def poseidon_permutation_constraints(cs, inputs, params, sbox):  # gadget_poseidon.rs:282-399
    w = params.width
    st = [LC.of(x) for x in inputs]
    off = 0
    fb, pr, fe = params.full_rounds_beginning, params.partial_rounds, params.full_rounds_end
    for rnd in range(fb + pr + fe):
        full = rnd < fb or rnd >= fb + pr
        outs = []
        for i in range(w):
            rk = params.round_keys[off]
            off += 1
            if full or i == w - 1:
                outs.append(LC.of(synthesize_sbox(cs, sbox, st[i], rk)))
            else:
                outs.append(st[i] + LC.of(rk))
        nxt = [LC() for _ in range(w)]
        for j in range(w):
            for i in range(w):
                nxt[i] = nxt[i] + outs[j] * params.MDS[i][j]
        st = [lc.simplify() for lc in nxt] if not full else nxt
    return st


def poseidon_permutation_gadget(cs, input_vars, params, sbox, output):  # gadget_poseidon.rs:402-420
    out = poseidon_permutation_constraints(cs, [LC.of(v) for v in input_vars], params, sbox)
    for i in range(params.width):
        constrain_lc_with_scalar(cs, out[i], output[i])


def poseidon_hash_2_constraints(cs, xl, xr, statics, params, sbox):  # gadget_poseidon.rs:445-468
    inputs = [statics[0], xl, xr] + list(statics[1:])
    return poseidon_permutation_constraints(cs, inputs, params, sbox)[1]


def poseidon_hash_2_gadget(cs, xl_var, xr_var, static_vars, params, sbox, output):  # :470-486
    h = poseidon_hash_2_constraints(cs, LC.of(xl_var), LC.of(xr_var), [LC.of(s) for s in static_vars], params, sbox)
    constrain_lc_with_scalar(cs, h, output)


def poseidon_hash_4_gadget(cs, in_vars, static_vars, params, sbox, output):  # :505-551
    inputs = [LC.of(static_vars[0])] + [LC.of(v) for v in in_vars] + [LC.of(s) for s in static_vars[1:]]
    h = poseidon_permutation_constraints(cs, inputs, params, sbox)[1]
    constrain_lc_with_scalar(cs, h, output)


def allocate_statics_for_prover(prover, num_statics):  # gadget_poseidon.rs:554-578
    out = [prover.commit(ZERO_CONST, 0)[1], prover.commit(PADDING_CONST, 0)[1]]
    for _ in range(2, num_statics):
        out.append(prover.commit(ZERO_CONST, 0)[1])
    return out


def allocate_statics_for_verifier(verifier, num_statics, pc_gens):  # gadget_poseidon.rs:581-608
    from .bp_pyref import ristretto_encode
    pad = ristretto_encode(pc_gens.commit(PADDING_CONST, 0))
    zero = ristretto_encode(pc_gens.commit(ZERO_CONST, 0))
    out = [verifier.commit(zero), verifier.commit(pad)]
    for _ in range(2, num_statics):
        out.append(verifier.commit(zero))
    return out


def vanilla_merkle_tree_verif_gadget(cs, depth, root, leaf_var, bit_vars, proof_vars, static_vars, params):
    """gadget_vsmt_2.rs:171-209."""
    statics = [LC.of(s) for s in static_vars]
    prev = LC()
    for i in range(depth):
        leaf_lc = LC.of(leaf_var) if i == 0 else prev
        one_minus = LC.of(VAR_ONE) - LC.of(bit_vars[i])
        _, _, l1 = cs.multiply(one_minus, leaf_lc)
        _, _, l2 = cs.multiply(LC.of(bit_vars[i]), LC.of(proof_vars[i]))
        left = LC.of(l1) + LC.of(l2)
        _, _, r1 = cs.multiply(LC.of(bit_vars[i]), leaf_lc)
        _, _, r2 = cs.multiply(one_minus, LC.of(proof_vars[i]))
        right = LC.of(r1) + LC.of(r2)
        prev = poseidon_hash_2_constraints(cs, left, right, statics, params, INVERSE)
    constrain_lc_with_scalar(cs, prev, root)


def poseidon_hash_4_constraints(cs, inputs4, statics, params, sbox):  # gadget_poseidon.rs:505-530
    inputs = [statics[0]] + list(inputs4) + list(statics[1:])
    return poseidon_permutation_constraints(cs, inputs, params, sbox)[1]


def vanilla_merkle_tree_4_verif_gadget(cs, levels, root, leaf_var, leaf_index_var, index_digits, proof_vars, static_vars, params):
    """gadget_vsmt_4.rs:199-312 for `levels` levels (the reference hard-codes 4 * LeafIndexBytes); index_digits = base-4 digits of
    the leaf index, least significant first (None on the verifier); proof_vars = 3 * levels siblings, popped from the END."""
    statics = [LC.of(s) for s in static_vars]
    prev = LC.of(leaf_var)
    nodes = list(proof_vars)
    index_terms = [(leaf_index_var, L - 1)]
    exp_4 = 1
    for lvl in range(levels):
        bits = []
        for which in range(2):
            if index_digits is None:
                b, b_1, o = cs.allocate_multiplier(None)
            else:
                bit = (index_digits[lvl] >> which) & 1
                b, b_1, o = cs.allocate_multiplier((bit, 1 - bit))
            cs.constrain(LC.of(o))
            cs.constrain(LC.of(b) + (LC.of(b_1) - LC.of(1)))
            bits.append((b, b_1))
        (b0, b0_1), (b1, b1_1) = bits
        index_terms.append((b1, 2 * exp_4 % L))
        index_terms.append((b0, exp_4))
        N3 = LC.of(nodes.pop()); N2 = LC.of(nodes.pop()); N1 = LC.of(nodes.pop())
        mul = lambda a, b: LC.of(cs.multiply(LC.of(a), LC.of(b))[2])
        b0_1_b1_1 = mul(b0_1, b1_1); b0_1_b1 = mul(b0_1, b1); b0_b1_1 = mul(b0, b1_1); b0_b1 = mul(b0, b1)
        c0_1 = mul(b0_1_b1_1, prev); c0_2 = mul(b0, N1); c0_3 = mul(b0_1_b1, N1)
        c0 = c0_1 + c0_2 + c0_3
        c1_1 = mul(b0_1_b1_1, N1); c1_2 = mul(b0_b1_1, prev); c1_3 = mul(b0_1_b1, N2); c1_4 = mul(b0_b1, N2)
        c1 = c1_1 + c1_2 + c1_3 + c1_4
        c2_1 = mul(b1_1, N2); c2_2 = mul(b0_1_b1, prev); c2_3 = mul(b0_b1, N3)
        c2 = c2_1 + c2_2 + c2_3
        c3_1 = mul(b1_1, N3); c3_2 = mul(b0_1_b1, N3); c3_3 = mul(b0_b1, prev)
        c3 = c3_1 + c3_2 + c3_3
        prev = poseidon_hash_4_constraints(cs, [c0, c1, c2, c3], statics, params, INVERSE)
        exp_4 = exp_4 * 4 % L
    cs.constrain(LC(index_terms))
    constrain_lc_with_scalar(cs, prev, root)


def vsmt4_root_from_path(leaf, digits, siblings3, params):
    """native root of a synthetic 4-ary path: at each level the running node sits at position `digit` among its three siblings
    (N1, N2, N3), arrangements of gadget_vsmt_4.rs:176-181"""
    cur = leaf
    for d, (n1, n2, n3) in zip(digits, siblings3):
        children = [n1, n2, n3]
        children.insert(d, cur)
        cur = poseidon_hash_4(children, params, INVERSE)
    return cur


def vsmt_root_from_path(leaf, bits, siblings, params):
    """Native root for a synthetic path, orientation of gadget_vsmt_2.rs:134-147 / :192-200."""
    cur = leaf
    for b, s in zip(bits, siblings):
        cur = poseidon_hash_2(s, cur, params, INVERSE) if b else poseidon_hash_2(cur, s, params, INVERSE)
    return cur


def mimc(xl, xr, constants):  # gadget_mimc.rs:19-39
    for c in constants:
        t = (xl + c) % L
        xl, xr = (t * t % L * t + xr) % L, xl
    return xl


def mimc_gadget(cs, left_var, right_var, rounds, constants, image):  # gadget_mimc.rs:41-79
    lv, rv = LC.of(left_var), LC.of(right_var)
    for j in range(rounds):
        lpc = lv + LC([(VAR_ONE, constants[j] % L)])
        l, _, l_sqr = cs.multiply(lpc, lpc)
        _, _, l_cube = cs.multiply(LC.of(l_sqr), LC.of(l))
        lv, rv = LC.of(l_cube) + rv, lv
    constrain_lc_with_scalar(cs, lv, image)


def positive_no_gadget(cs, v_var, v_assignment, bit_size):  # r1cs_utils.rs:20-48
    terms = [(v_var, L - 1)]
    exp_2 = 1
    for i in range(bit_size):
        if v_assignment is None:
            a, b, o = cs.allocate_multiplier(None)
        else:
            bit = (v_assignment >> i) & 1
            a, b, o = cs.allocate_multiplier((1 - bit, bit))
        cs.constrain(LC.of(o))
        cs.constrain(LC.of(a) + (LC.of(b) - LC.of(1)))
        terms.append((b, exp_2))
        exp_2 = exp_2 * 2 % L
    cs.constrain(LC(terms))


def bound_check_gadget(cs, v, a, b, vmax, vmin, bit_size):
    """gadget_bound_check.rs:18-45; v, a, b are (variable, assignment-or-None)."""
    cs.constrain(LC.of(v[0]) - LC.of(vmin) - LC.of(a[0]))
    cs.constrain(LC.of(vmax) - LC.of(v[0]) - LC.of(b[0]))
    constrain_lc_with_scalar(cs, LC.of(a[0]) + LC.of(b[0]), vmax - vmin)
    positive_no_gadget(cs, a[0], a[1], bit_size)
    positive_no_gadget(cs, b[0], b[1], bit_size)
