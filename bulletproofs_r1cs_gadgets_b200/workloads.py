"""Named workloads of BASELINE.json: circuit construction through the gadget API and synthetic inputs.

Synthetic inputs follow SURVEY.md section 8(d): every scalar is
SHAKE256("bp-b200/" || config || u64le(proof_index) || u64le(item_index)) -> 64 bytes -> reduced mod l,
so the oracle, the CPU baseline and the GPU all see the same bytes.  The driver code below is the
batched analogue of the reference's test drivers (reference src/gadget_vsmt_2.rs:262-399,
src/gadget_poseidon.rs:691-785, src/gadget_mimc.rs:92-175, src/gadget_bound_check.rs:49-116).
"""
import hashlib

import numpy as np

from . import api

L = api.L


def synth_scalar(config, proof, item):
    h = hashlib.shake_256(b"bp-b200/" + config + int(proof).to_bytes(8, "little") + int(item).to_bytes(8, "little")).digest(64)
    return int.from_bytes(h, "little") % L


def synth_bytes32(config, proof, item):
    return hashlib.shake_256(b"bp-b200/" + config + int(proof).to_bytes(8, "little") + int(item).to_bytes(8, "little")).digest(32)


class Workload:
    """circuit + label + per-proof input generator"""

    def __init__(self, name, label, circuit, gens_capacity):
        self.name, self.label, self.circuit, self.gens_capacity = name, label, circuit, gens_capacity

    def inputs(self, first, count):
        """returns dict(v, v_blinding, entropy[, aux][, pub]) of uint8 arrays for proofs first..first+count-1"""
        raise NotImplementedError


def _next_pow2(n):
    N = 1
    while N < n:
        N *= 2
    return N


class Vsmt2(Workload):
    """BASELINE config 5: depth-`depth` membership proof in a binary sparse Merkle tree, Poseidon 2:1, inverse S-box.
    Commit order (reference src/gadget_vsmt_2.rs:296-330): leaf, index bits LSB first, siblings leaf level first, 4 statics."""

    def __init__(self, gens_for_recording, depth=32, params=None, public_root=True):
        self.depth = depth
        self.params = params or api.PoseidonParams()
        rec = api.Verifier(gens_for_recording, b"VSMT")
        zero = bytes(32)
        leaf = rec.commit(zero)
        bits = [rec.commit(zero) for _ in range(depth)]
        nodes = [rec.commit(zero) for _ in range(depth)]
        statics = rec.allocate_statics(4)  # the verifier computes these four commitments itself (reference src/gadget_poseidon.rs:580-608)
        root = rec.public_input()
        rec.vsmt2_verif_gadget(self.params, depth, root, leaf, bits, nodes, statics)
        circ = rec.compile()
        super().__init__("vsmt2_depth%d" % depth, b"VSMT", circ, _next_pow2(circ.n))
        self.cfg = b"vsmt2/%d" % depth

    def witness_values(self, p):
        d = self.depth
        leaf = synth_scalar(self.cfg, p, 0)
        bits = [synth_bytes32(self.cfg, p, 1 + i)[0] & 1 for i in range(d)]
        sibs = [synth_scalar(self.cfg, p, 1 + d + i) for i in range(d)]
        return leaf, bits, sibs

    def root(self, leaf, bits, sibs):
        cur = leaf  # orientation of reference src/gadget_vsmt_2.rs:192-200
        for b, s in zip(bits, sibs):
            cur = self.params.hash_2(s, cur, api.SBOX_INVERSE) if b else self.params.hash_2(cur, s, api.SBOX_INVERSE)
        return cur

    def roots_batch(self, v):
        """Merkle roots of a batch of committed-value rows v [count][m][32] (leaf, bits, siblings, statics), every level hashed
        as one device batch: the public inputs the verifier needs when `inputs(..., with_root=False)` skipped them"""
        d = self.depth
        cur = np.ascontiguousarray(v[:, 0])
        for i in range(d):
            bit = (v[:, 1 + i, 0] & 1).astype(bool)[:, None]
            sib = v[:, 1 + d + i]
            cur = self.params.hash_2_batch(np.where(bit, sib, cur), np.where(bit, cur, sib), api.SBOX_INVERSE)
        return cur.reshape(-1, 1, 32)

    def inputs(self, first, count, with_root=True):
        d, m = self.depth, self.circuit.m
        v = np.zeros((count, m, 32), dtype=np.uint8)
        vb = np.zeros((count, m, 32), dtype=np.uint8)
        ent = np.zeros((count, 32), dtype=np.uint8)
        pub = np.zeros((count, 1, 32), dtype=np.uint8)
        for i in range(count):
            p = first + i
            leaf, bits, sibs = self.witness_values(p)
            vals = [leaf] + bits + sibs + [0, 101, 0, 0]
            bl = [synth_scalar(self.cfg, p, 1000 + j) for j in range(1 + 2 * d)] + [0, 0, 0, 0]
            v[i] = api.scalars_to_array(vals)
            vb[i] = api.scalars_to_array(bl)
            ent[i] = np.frombuffer(synth_bytes32(self.cfg, p, 2000), dtype=np.uint8)
            if with_root:
                pub[i, 0] = np.frombuffer(api.scalar_bytes(self.root(leaf, bits, sibs)), dtype=np.uint8)
        return dict(v=v, v_blinding=vb, entropy=ent, pub=pub)


class Vsmt4(Workload):
    """Membership proof in a 4-ary sparse Merkle tree, Poseidon 4:1, inverse S-box (reference src/gadget_vsmt_4.rs:199-312,
    test at :362-483).  Commit order of the reference test: leaf value, leaf index, 3 * levels siblings, 2 statics.  The bits of the
    index digits are allocate_multiplier inputs, i.e. per-proof auxiliary inputs of the batched circuit: b0, 1-b0, b1, 1-b1 per level."""

    def __init__(self, gens_for_recording, levels=16, params=None):
        self.levels = levels
        self.params = params or api.PoseidonParams()
        rec = api.Verifier(gens_for_recording, b"VSMT")
        zero = bytes(32)
        leaf = rec.commit(zero)
        index = rec.commit(zero)
        nodes = [rec.commit(zero) for _ in range(3 * levels)]
        statics = rec.allocate_statics(2)
        root = rec.public_input()
        rec.vsmt4_verif_gadget(self.params, levels, root, leaf, index, None, nodes, statics)
        circ = rec.compile()
        super().__init__("vsmt4_levels%d" % levels, b"VSMT", circ, _next_pow2(circ.n))
        self.cfg = b"vsmt4/%d" % levels

    def witness_values(self, p):
        lv = self.levels
        leaf = synth_scalar(self.cfg, p, 0)
        digits = [synth_bytes32(self.cfg, p, 1 + i)[0] & 3 for i in range(lv)]
        # sibling triples in the order the gadget consumes them: level 0 (leaf level) first
        sibs = [tuple(synth_scalar(self.cfg, p, 1 + lv + 3 * i + j) for j in range(3)) for i in range(lv)]
        return leaf, digits, sibs

    def root(self, leaf, digits, sibs):
        cur = leaf  # arrangements of reference src/gadget_vsmt_4.rs:176-181
        for d, (n1, n2, n3) in zip(digits, sibs):
            children = [n1, n2, n3]
            children.insert(d, cur)
            cur = self.params.hash_4(children, api.SBOX_INVERSE)
        return cur

    def inputs(self, first, count, with_root=True):
        lv, m = self.levels, self.circuit.m
        v = np.zeros((count, m, 32), dtype=np.uint8)
        vb = np.zeros((count, m, 32), dtype=np.uint8)
        ent = np.zeros((count, 32), dtype=np.uint8)
        pub = np.zeros((count, 1, 32), dtype=np.uint8)
        aux = np.zeros((count, 4 * lv, 32), dtype=np.uint8)
        for i in range(count):
            p = first + i
            leaf, digits, sibs = self.witness_values(p)
            index = sum(d << (2 * j) for j, d in enumerate(digits))
            # the gadget pops (N3, N2, N1) of the level it is processing from the END of the vector
            nodes = [x for triple in reversed(sibs) for x in triple]
            vals = [leaf, index] + nodes + [0, 101]
            bl = [synth_scalar(self.cfg, p, 1000 + j) for j in range(2 + 3 * lv)] + [0, 0]
            v[i] = api.scalars_to_array(vals)
            vb[i] = api.scalars_to_array(bl)
            ent[i] = np.frombuffer(synth_bytes32(self.cfg, p, 2000), dtype=np.uint8)
            a = []
            for d in digits:
                b0, b1 = d & 1, (d >> 1) & 1
                a += [b0, 1 - b0, b1, 1 - b1]
            aux[i] = api.scalars_to_array(a)
            if with_root:
                pub[i, 0] = np.frombuffer(api.scalar_bytes(self.root(leaf, digits, sibs)), dtype=np.uint8)
        return dict(v=v, v_blinding=vb, entropy=ent, pub=pub, aux=aux)


class PoseidonHash2(Workload):
    """BASELINE config 2: Poseidon 2:1 preimage proofs (reference src/gadget_poseidon.rs:691-785)."""

    def __init__(self, gens_for_recording, sbox, params=None):
        self.sbox = sbox
        self.params = params or api.PoseidonParams()
        label = b"Poseidon_hash_2_cube" if sbox == api.SBOX_CUBE else b"Poseidon_hash_2_inverse"
        rec = api.Verifier(gens_for_recording, label)
        zero = bytes(32)
        xs = [rec.commit(zero) for _ in range(2)]
        statics = rec.allocate_statics(4)
        h = rec.public_input()
        rec.poseidon_hash_2_gadget(self.params, xs[0], xs[1], statics, sbox, h)
        circ = rec.compile()
        super().__init__("poseidon2_%s" % ("cube" if sbox == api.SBOX_CUBE else "inverse"), label, circ, _next_pow2(circ.n))
        self.cfg = b"poseidon2/%d" % sbox

    def inputs(self, first, count):
        v = np.zeros((count, 6, 32), dtype=np.uint8)
        vb = np.zeros((count, 6, 32), dtype=np.uint8)
        ent = np.zeros((count, 32), dtype=np.uint8)
        pub = np.zeros((count, 1, 32), dtype=np.uint8)
        for i in range(count):
            p = first + i
            xl, xr = synth_scalar(self.cfg, p, 0), synth_scalar(self.cfg, p, 1)
            v[i] = api.scalars_to_array([xl, xr, 0, 101, 0, 0])
            vb[i] = api.scalars_to_array([synth_scalar(self.cfg, p, 10), synth_scalar(self.cfg, p, 11), 0, 0, 0, 0])
            ent[i] = np.frombuffer(synth_bytes32(self.cfg, p, 2000), dtype=np.uint8)
            pub[i, 0] = np.frombuffer(api.scalar_bytes(self.params.hash_2(xl, xr, self.sbox)), dtype=np.uint8)
        return dict(v=v, v_blinding=vb, entropy=ent, pub=pub)


class Mimc(Workload):
    """BASELINE config 4: MiMC-322 preimage proofs (reference src/gadget_mimc.rs:92-175)."""

    def __init__(self, gens_for_recording, rounds=322):
        self.rounds = rounds
        self.cfg = b"mimc/%d" % rounds
        self.constants = [synth_scalar(self.cfg, 0xFFFFFFFF, j) for j in range(rounds)]  # fixed for the batch (gadget_mimc.rs:96)
        rec = api.Verifier(gens_for_recording, b"MiMC")
        zero = bytes(32)
        a, b = rec.commit(zero), rec.commit(zero)
        img = rec.public_input()
        rec.mimc_gadget(a, b, self.constants, img)
        circ = rec.compile()
        super().__init__("mimc_%d" % rounds, b"MiMC", circ, _next_pow2(circ.n))

    def image(self, xl, xr):
        for c in self.constants:  # gadget_mimc.rs:19-39
            t = (xl + c) % L
            xl, xr = (t * t % L * t + xr) % L, xl
        return xl

    def inputs(self, first, count):
        v = np.zeros((count, 2, 32), dtype=np.uint8)
        vb = np.zeros((count, 2, 32), dtype=np.uint8)
        ent = np.zeros((count, 32), dtype=np.uint8)
        pub = np.zeros((count, 1, 32), dtype=np.uint8)
        for i in range(count):
            p = first + i
            xl, xr = synth_scalar(self.cfg, p, 0), synth_scalar(self.cfg, p, 1)
            v[i] = api.scalars_to_array([xl, xr])
            vb[i] = api.scalars_to_array([synth_scalar(self.cfg, p, 10), synth_scalar(self.cfg, p, 11)])
            ent[i] = np.frombuffer(synth_bytes32(self.cfg, p, 2000), dtype=np.uint8)
            pub[i, 0] = np.frombuffer(api.scalar_bytes(self.image(xl, xr)), dtype=np.uint8)
        return dict(v=v, v_blinding=vb, entropy=ent, pub=pub)


class BoundCheck(Workload):
    """BASELINE config 1: min <= v <= max with two `bit_size`-bit range constraints (reference src/gadget_bound_check.rs:49-116).
    The bit assignments of allocate_multiplier (src/r1cs_utils.rs:29-32) are the auxiliary inputs of the witness program."""

    def __init__(self, gens_for_recording, vmin=0, vmax=2 ** 64 - 1, bit_size=64):
        self.vmin, self.vmax, self.bit_size = vmin, vmax, bit_size
        rec = api.Verifier(gens_for_recording, b"BoundsTest")
        zero = bytes(32)
        v, a, b = rec.commit(zero), rec.commit(zero), rec.commit(zero)
        rec.bound_check_gadget(v, a, b, vmax, vmin, bit_size)
        circ = rec.compile()
        super().__init__("bound_check_%d" % bit_size, b"BoundsTest", circ, _next_pow2(circ.n))
        self.cfg = b"bound/%d" % bit_size

    def inputs(self, first, count):
        n_aux = self.circuit.num_aux
        v = np.zeros((count, 3, 32), dtype=np.uint8)
        vb = np.zeros((count, 3, 32), dtype=np.uint8)
        ent = np.zeros((count, 32), dtype=np.uint8)
        aux = np.zeros((count, n_aux, 32), dtype=np.uint8)
        for i in range(count):
            p = first + i
            val = self.vmin + synth_scalar(self.cfg, p, 0) % (self.vmax - self.vmin)
            a, b = val - self.vmin, self.vmax - val
            v[i] = api.scalars_to_array([val, a, b])
            vb[i] = api.scalars_to_array([synth_scalar(self.cfg, p, 10 + j) for j in range(3)])
            ent[i] = np.frombuffer(synth_bytes32(self.cfg, p, 2000), dtype=np.uint8)
            k = 0
            for q in (a, b):
                for bit_i in range(self.bit_size):
                    bit = (q >> bit_i) & 1
                    aux[i, k, 0] = 1 - bit
                    aux[i, k + 1, 0] = bit
                    k += 2
        return dict(v=v, v_blinding=vb, entropy=ent, aux=aux)


def make(name, gens, **kw):
    if name == "vsmt2":
        return Vsmt2(gens, **kw)
    if name == "poseidon2_cube":
        return PoseidonHash2(gens, api.SBOX_CUBE, **kw)
    if name == "poseidon2_inverse":
        return PoseidonHash2(gens, api.SBOX_INVERSE, **kw)
    if name == "mimc":
        return Mimc(gens, **kw)
    if name == "bound_check":
        return BoundCheck(gens, **kw)
    raise ValueError(name)
