"""Python host layer over the C-ABI (include/bp_b200.h).

Mirrors the surface the reference's gadgets are written against -- `Prover`, `Verifier`,
`LinearCombination`, `Variable`, `PedersenGens`/`BulletproofGens` (one `Gens` object here) -- plus the
batched `Circuit` API that is the reason this library exists.  Everything numeric happens in
`libbp_b200.so` (CUDA, sm_100a); this module only marshals bytes.  If the shared library is missing
the import of any symbol raises: there is no Python or CPU fallback.

Reference call sites mirrored (paths under /root/reference/src): gadget_mimc.rs:99-169 (prove/verify
flow), gadget_poseidon.rs:554-608 (statics), gadget_vsmt_2.rs:171-209,289-395, gadget_bound_check.rs:49-116.
"""
import ctypes as C
import os

import numpy as np

L = 2 ** 252 + 27742317777372353535851937790883648493

BP_OK = 0
ERRORS = {
    1: "InvalidGeneratorsLength", 2: "FormatError", 3: "VerificationError", 4: "MissingAssignment", 5: "GadgetError",
    6: "InvalidArgument", 7: "NoDevice", 8: "CudaError", 9: "OutOfMemory",
}
SBOX_CUBE, SBOX_INVERSE = 0, 1
VAR_COMMITTED, VAR_MULT_LEFT, VAR_MULT_RIGHT, VAR_MULT_OUT, VAR_ONE, VAR_PUBLIC = 0, 1, 2, 3, 4, 5

_PKG = os.path.dirname(os.path.abspath(__file__))
DEFAULT_SO = os.environ.get("BP_B200_LIB") or os.path.join(_PKG, "libbp_b200.so")  # BP_B200_LIB: developer override for kernel-variant builds


class R1CSError(Exception):
    def __init__(self, code, where=""):
        self.code = code
        super().__init__("%s (%d) %s" % (ERRORS.get(code, "Unknown"), code, where))


class bp_var(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("index", C.c_uint32)]


class bp_term(C.Structure):
    _fields_ = [("var", bp_var), ("coeff", C.c_uint8 * 32)]


u8p = C.POINTER(C.c_uint8)
_lib = None


_HEADER = os.path.join(_PKG, "..", "include", "bp_b200.h")
_SCALAR_TYPES = {"int32_t": C.c_int32, "uint32_t": C.c_uint32, "int64_t": C.c_int64, "uint64_t": C.c_uint64, "size_t": C.c_size_t, "int": C.c_int}


def header_prototypes(header=_HEADER):
    """[(name, return type, [parameter type, ...])] of every function include/bp_b200.h declares, as C type strings"""
    import re
    src = re.sub(r"/\*.*?\*/", "", open(header).read(), flags=re.S)
    out = []
    for ret, name, args in re.findall(r"\n\s*([a-z0-9_]+(?:\s*\*)?)\s+(bp_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", src):
        params = []
        for prm in args.split(","):
            prm = " ".join(prm.split())
            if prm in ("void", ""):
                continue
            if "*" in prm or "[" in prm:
                params.append("pointer")
            else:
                params.append(prm.rsplit(" ", 1)[0].replace("const ", ""))
        out.append((name, " ".join(ret.split()), params))
    return out


def _declare_prototypes(lib):
    """restype / argtypes of every entry point, taken from the header itself: ctypes then converts and range-checks every scalar
    argument (a bare Python int would otherwise travel as a C int) and refuses a call with the wrong number of arguments.
    Pointers are declared as void* (arrays, ctypes pointers, integer addresses and None all convert)."""
    if not os.path.exists(_HEADER):
        raise RuntimeError("bp_b200: %s not found (the Python layer takes the C prototypes from it)" % _HEADER)
    for name, ret, params in header_prototypes():
        fn = getattr(lib, name, None)
        if fn is None:
            raise RuntimeError("bp_b200: %s is declared in include/bp_b200.h but missing from the library" % name)
        fn.restype = None if ret == "void" else _SCALAR_TYPES[ret]
        fn.argtypes = [C.c_void_p if p == "pointer" else (bp_var if p == "bp_var" else _SCALAR_TYPES[p]) for p in params]


def load(path=None):
    """Load the shared library (once).  `path` is for the test-only emulation build."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    path = path or DEFAULT_SO
    if not os.path.exists(path):
        raise RuntimeError("bp_b200: %s not found -- build it with `python -c 'import __graft_entry__ as g; g.build()'`; "
                           "there is no CPU fallback" % path)
    lib = C.CDLL(path)
    _declare_prototypes(lib)
    _lib = lib
    return lib


def _check(rc, where=""):
    if rc != BP_OK:
        raise R1CSError(rc, where)


def scalar_bytes(x):
    if isinstance(x, (bytes, bytearray)):
        assert len(x) == 32
        return bytes(x)
    return (int(x) % L).to_bytes(32, "little")


def _buf(b):
    return (C.c_uint8 * len(b)).from_buffer_copy(bytes(b))


def _np_u8(a):
    return np.ascontiguousarray(a, dtype=np.uint8)


def _want(a, shape, what, code=6):
    """the C-ABI takes plain pointers and sizes: a wrongly shaped array would be an out-of-bounds access inside the library, so
    shapes are checked here (FormatError for proof bytes, InvalidArgument otherwise)"""
    if a is None:
        raise R1CSError(4, "%s is required" % what)
    if tuple(a.shape) != tuple(shape):
        raise R1CSError(code, "%s has shape %s, expected %s" % (what, tuple(a.shape), tuple(shape)))
    return a


def scalars_to_array(vals):
    """iterable of ints -> uint8 [len][32]"""
    vals = list(vals)
    if not vals:
        return np.zeros((0, 32), dtype=np.uint8)
    return np.frombuffer(b"".join(scalar_bytes(v) for v in vals), dtype=np.uint8).reshape(-1, 32).copy()


class Variable:
    __slots__ = ("kind", "index")

    def __init__(self, kind, index=0):
        self.kind, self.index = kind, index

    @staticmethod
    def One():
        return Variable(VAR_ONE, 0)

    def __eq__(self, o):
        return isinstance(o, Variable) and (self.kind, self.index) == (o.kind, o.index)

    def __hash__(self):
        return hash((self.kind, self.index))

    def __repr__(self):
        return "Variable(%d,%d)" % (self.kind, self.index)

    def _c(self):
        return bp_var(self.kind, self.index)

    # arithmetic promotes to LinearCombination, as in the reference
    def __add__(self, o):
        return LinearCombination.of(self) + o

    def __sub__(self, o):
        return LinearCombination.of(self) - o

    def __rsub__(self, o):
        return LinearCombination.of(o) - self

    def __radd__(self, o):
        return LinearCombination.of(o) + self

    def __mul__(self, s):
        return LinearCombination.of(self) * s


class LinearCombination:
    """list of (Variable, int coefficient); duplicates allowed (reference LinearCombination)."""

    def __init__(self, terms=None):
        self.terms = list(terms) if terms else []

    @staticmethod
    def of(x):
        if isinstance(x, LinearCombination):
            return x
        if isinstance(x, Variable):
            return LinearCombination([(x, 1)])
        return LinearCombination([(Variable.One(), int(x) % L)])

    def __add__(self, o):
        return LinearCombination(self.terms + LinearCombination.of(o).terms)

    __radd__ = __add__

    def __sub__(self, o):
        return LinearCombination(self.terms + [(v, (-c) % L) for v, c in LinearCombination.of(o).terms])

    def __rsub__(self, o):
        return LinearCombination.of(o) - self

    def __neg__(self):
        return LinearCombination([(v, (-c) % L) for v, c in self.terms])

    def __mul__(self, s):
        return LinearCombination([(v, c * int(s) % L) for v, c in self.terms])

    def get_terms(self):
        return list(self.terms)

    def _c(self):
        n = len(self.terms)
        arr = (bp_term * max(n, 1))()
        for i, (v, c) in enumerate(self.terms):
            arr[i].var.kind, arr[i].var.index = v.kind, v.index
            C.memmove(arr[i].coeff, scalar_bytes(c), 32)
        return arr, n


class PoseidonParams:
    """PoseidonParams::new(width, full_b, full_e, partial) over the reference's constants (gadget_poseidon.rs:27-94)."""

    def __init__(self, width=6, full_rounds_beginning=4, full_rounds_end=4, partial_rounds=140, constants=None):
        if constants is None:
            constants = open(os.path.join(_PKG, "data", "poseidon_constants.bin"), "rb").read()
        self.width, self.full_rounds_beginning, self.full_rounds_end, self.partial_rounds = width, full_rounds_beginning, full_rounds_end, partial_rounds
        self._h = C.c_void_p()
        _check(load().bp_poseidon_params_new(_buf(constants), C.c_size_t(len(constants) // 32), width, full_rounds_beginning, full_rounds_end,
                                             partial_rounds, C.byref(self._h)), "poseidon_params_new")

    def __del__(self):
        if _lib is not None and getattr(self, "_h", None):
            _lib.bp_poseidon_params_free(self._h)
            self._h = None

    def hash_2(self, xl, xr, sbox):
        out = (C.c_uint8 * 32)()
        _check(load().bp_poseidon_hash_2(self._h, _buf(scalar_bytes(xl)), _buf(scalar_bytes(xr)), sbox, out))
        return int.from_bytes(bytes(out), "little")

    def hash_2_batch(self, xl, xr, sbox):
        """Poseidon_hash_2 of `count` independent pairs on the device; xl, xr: uint8 [count][32] -> uint8 [count][32]"""
        xl, xr = _np_u8(xl), _np_u8(xr)
        _want(xl, (xl.shape[0], 32), "xl"); _want(xr, xl.shape, "xr")
        out = np.zeros_like(xl)
        if xl.shape[0]:
            _check(load().bp_poseidon_hash_2_batch(self._h, C.c_int32(sbox), C.c_uint32(xl.shape[0]), xl.ctypes.data_as(u8p), xr.ctypes.data_as(u8p),
                                                   out.ctypes.data_as(u8p)), "poseidon_hash_2_batch")
        return out

    def hash_4(self, xs, sbox):
        """Poseidon_hash_4 (reference src/gadget_poseidon.rs:488-503)"""
        out = (C.c_uint8 * 32)()
        _check(load().bp_poseidon_hash_4(self._h, _buf(b"".join(scalar_bytes(x) for x in xs)), sbox, out))
        return int.from_bytes(bytes(out), "little")


class Gens:
    """PedersenGens::default() + BulletproofGens::new(capacity, 1), resident on the device."""

    def __init__(self, capacity):
        self._h = C.c_void_p()
        _check(load().bp_gens_new(C.c_uint32(capacity), C.byref(self._h)), "gens_new")
        self.capacity = capacity

    def __del__(self):
        if _lib is not None and getattr(self, "_h", None):
            _lib.bp_gens_free(self._h)
            self._h = None

    def pedersen(self):
        B, Bb = (C.c_uint8 * 32)(), (C.c_uint8 * 32)()
        _check(load().bp_gens_pedersen(self._h, B, Bb))
        return bytes(B), bytes(Bb)

    def export(self, which, count):
        out = np.zeros((count, 32), dtype=np.uint8)
        _check(load().bp_gens_export(self._h, which, C.c_uint32(count), out.ctypes.data_as(u8p)))
        return out

    def commit(self, v, r):
        """pc_gens.commit(v, r).compress()"""
        out = (C.c_uint8 * 32)()
        _check(load().bp_pc_commit(self._h, 1, _buf(scalar_bytes(v)), _buf(scalar_bytes(r)), out))
        return bytes(out)


def _vars3(arr):
    return tuple(Variable(arr[i].kind, arr[i].index) for i in range(3))


class ConstraintSystem:
    def __init__(self, gens, label, prover):
        self.gens = gens
        self._h = C.c_void_p()
        fn = load().bp_prover_new if prover else load().bp_verifier_new
        _check(fn(gens._h, _buf(label) if label else None, C.c_size_t(len(label)), C.byref(self._h)))
        self.is_prover = prover

    def __del__(self):
        if _lib is not None and getattr(self, "_h", None):
            _lib.bp_cs_free(self._h)
            self._h = None

    def multiply(self, left, right):
        l, nl = LinearCombination.of(left)._c()
        r, nr = LinearCombination.of(right)._c()
        out = (bp_var * 3)()
        _check(load().bp_cs_multiply(self._h, l, C.c_size_t(nl), r, C.c_size_t(nr), out), "multiply")
        return _vars3(out)

    def allocate_multiplier(self, assignment):
        out = (bp_var * 3)()
        if assignment is None:
            rc = load().bp_cs_allocate_multiplier(self._h, None, None, out)
        else:
            rc = load().bp_cs_allocate_multiplier(self._h, _buf(scalar_bytes(assignment[0])), _buf(scalar_bytes(assignment[1])), out)
        _check(rc, "allocate_multiplier")
        return _vars3(out)

    def allocate_single(self, assignment):
        var, ov, has = bp_var(), bp_var(), C.c_int32(0)
        val = None if assignment is None else _buf(scalar_bytes(assignment))
        _check(load().bp_cs_allocate_single(self._h, val, C.byref(var), C.byref(ov), C.byref(has)), "allocate_single")
        return Variable(var.kind, var.index), (Variable(ov.kind, ov.index) if has.value else None)

    def evaluate_lc(self, lc):
        arr, n = LinearCombination.of(lc)._c()
        out = (C.c_uint8 * 32)()
        rc = load().bp_cs_evaluate_lc(self._h, arr, C.c_size_t(n), out)
        if rc == 4:
            return None
        _check(rc, "evaluate_lc")
        return int.from_bytes(bytes(out), "little")

    def public_input(self, value=None):
        """declares the next per-proof public input of a batched circuit (BP_VAR_PUBLIC)"""
        var = bp_var()
        _check(load().bp_cs_public_input(self._h, None if value is None else _buf(scalar_bytes(value)), C.byref(var)), "public_input")
        return Variable(var.kind, var.index)

    def constrain(self, lc):
        arr, n = LinearCombination.of(lc)._c()
        _check(load().bp_cs_constrain(self._h, arr, C.c_size_t(n)), "constrain")

    def num_constraints(self):
        return int(load().bp_cs_num_constraints(self._h))

    def num_multipliers(self):
        return int(load().bp_cs_num_multipliers(self._h))

    def num_commitments(self):
        return int(load().bp_cs_num_commitments(self._h))

    # ---- the reference's gadgets, executed by the library's host layer against this constraint system
    def allocate_statics(self, num_statics):
        out = (bp_var * num_statics)()
        _check(load().bp_gadget_allocate_statics(self._h, C.c_uint32(num_statics), out), "allocate_statics")
        return [Variable(o.kind, o.index) for o in out]

    def poseidon_hash_2_gadget(self, params, xl, xr, statics, sbox, expected):
        st = (bp_var * len(statics))(*[s._c() for s in statics])
        if isinstance(expected, Variable):
            _check(load().bp_gadget_poseidon_hash_2_public(self._h, params._h, xl._c(), xr._c(), st, C.c_uint32(len(statics)), sbox, expected._c()),
                   "poseidon_hash_2_gadget")
            return
        _check(load().bp_gadget_poseidon_hash_2(self._h, params._h, xl._c(), xr._c(), st, C.c_uint32(len(statics)), sbox,
                                                _buf(scalar_bytes(expected))), "poseidon_hash_2_gadget")

    def vsmt2_verif_gadget(self, params, depth, root, leaf, bits, nodes, statics):
        b = (bp_var * depth)(*[x._c() for x in bits])
        nd = (bp_var * depth)(*[x._c() for x in nodes])
        st = (bp_var * len(statics))(*[s._c() for s in statics])
        if isinstance(root, Variable):
            _check(load().bp_gadget_vsmt2_verif_public(self._h, params._h, C.c_uint32(depth), root._c(), leaf._c(), b, nd, st, C.c_uint32(len(statics))),
                   "vsmt2_verif_gadget")
            return
        _check(load().bp_gadget_vsmt2_verif(self._h, params._h, C.c_uint32(depth), _buf(scalar_bytes(root)), leaf._c(), b, nd, st,
                                            C.c_uint32(len(statics))), "vsmt2_verif_gadget")

    def poseidon_hash_4_gadget(self, params, xs, statics, sbox, expected):
        st = (bp_var * len(statics))(*[s._c() for s in statics])
        xv = (bp_var * 4)(*[x._c() for x in xs])
        if isinstance(expected, Variable):
            _check(load().bp_gadget_poseidon_hash_4_public(self._h, params._h, xv, st, C.c_uint32(len(statics)), sbox, expected._c()), "poseidon_hash_4_gadget")
            return
        _check(load().bp_gadget_poseidon_hash_4(self._h, params._h, xv, st, C.c_uint32(len(statics)), sbox, _buf(scalar_bytes(expected))), "poseidon_hash_4_gadget")

    def vsmt4_verif_gadget(self, params, levels, root, leaf, leaf_index, index_digits, nodes, statics):
        """vanilla_merkle_merkle_tree_4_verif_gadget (reference src/gadget_vsmt_4.rs:199-312); index_digits: base-4 digits, LSB first, or None"""
        nd = (bp_var * (3 * levels))(*[x._c() for x in nodes])
        st = (bp_var * len(statics))(*[s._c() for s in statics])
        dg = _buf(bytes(index_digits)) if index_digits is not None else None
        if isinstance(root, Variable):
            _check(load().bp_gadget_vsmt4_verif_public(self._h, params._h, C.c_uint32(levels), root._c(), leaf._c(), leaf_index._c(), dg, nd, st,
                                                       C.c_uint32(len(statics))), "vsmt4_verif_gadget")
            return
        _check(load().bp_gadget_vsmt4_verif(self._h, params._h, C.c_uint32(levels), _buf(scalar_bytes(root)), leaf._c(), leaf_index._c(), dg, nd, st,
                                            C.c_uint32(len(statics))), "vsmt4_verif_gadget")

    def mimc_gadget(self, left, right, constants, image):
        cb = b"".join(scalar_bytes(c) for c in constants)
        if isinstance(image, Variable):
            _check(load().bp_gadget_mimc_public(self._h, left._c(), right._c(), C.c_uint32(len(constants)), _buf(cb), image._c()), "mimc_gadget")
            return
        _check(load().bp_gadget_mimc(self._h, left._c(), right._c(), C.c_uint32(len(constants)), _buf(cb), _buf(scalar_bytes(image))), "mimc_gadget")

    def bound_check_gadget(self, v, a, b, vmax, vmin, bit_size, values=None):
        has = values is not None
        vv, av, bv = values if has else (0, 0, 0)
        _check(load().bp_gadget_bound_check(self._h, v._c(), a._c(), b._c(), int(has), C.c_uint64(vv), C.c_uint64(av), C.c_uint64(bv),
                                            C.c_uint64(vmax), C.c_uint64(vmin), C.c_uint32(bit_size)), "bound_check_gadget")

    def compile(self):
        return Circuit._from_cs(self)


class Prover(ConstraintSystem):
    """Transcript::new(label); Prover::new(&pc_gens, &mut transcript)"""

    def __init__(self, gens, label):
        super().__init__(gens, label, True)

    def commit(self, v, v_blinding):
        V, var = (C.c_uint8 * 32)(), bp_var()
        _check(load().bp_prover_commit(self._h, _buf(scalar_bytes(v)), _buf(scalar_bytes(v_blinding)), V, C.byref(var)), "commit")
        return bytes(V), Variable(var.kind, var.index)

    def prove(self, entropy):
        """prover.prove(&bp_gens); `entropy` = the 32 bytes the reference takes from thread_rng()"""
        plen = load().bp_cs_proof_len(self._h)
        out = (C.c_uint8 * plen)()
        n = C.c_size_t(plen)
        _check(load().bp_prover_prove(self._h, _buf(entropy), out, C.byref(n)), "prove")
        return bytes(out[: n.value])


class Verifier(ConstraintSystem):
    """Transcript::new(label); Verifier::new(&mut transcript)"""

    def __init__(self, gens, label):
        super().__init__(gens, label, False)

    def commit(self, V):
        var = bp_var()
        _check(load().bp_verifier_commit(self._h, _buf(V), C.byref(var)), "commit")
        return Variable(var.kind, var.index)

    def verify(self, proof, entropy):
        """verifier.verify(&proof, &pc_gens, &bp_gens) -> raises R1CSError on failure"""
        _check(load().bp_verifier_verify(self._h, _buf(proof), C.c_size_t(len(proof)), _buf(entropy)), "verify")
        return True


class Circuit:
    """A compiled constraint system shared by every proof of a batch."""

    def __init__(self):
        self._h = C.c_void_p()

    @staticmethod
    def _from_cs(cs):
        c = Circuit()
        _check(load().bp_circuit_compile(cs._h, C.byref(c._h)), "circuit_compile")
        c._sizes()
        return c

    @staticmethod
    def from_arrays(n, m, cons_ptr, kind, idx, coeff):
        c = Circuit()
        cons_ptr = np.ascontiguousarray(cons_ptr, dtype=np.uint32)
        kind, idx, coeff = _np_u8(kind), np.ascontiguousarray(idx, dtype=np.uint32), _np_u8(coeff)
        _check(load().bp_circuit_from_arrays(C.c_uint32(n), C.c_uint32(m), C.c_uint32(len(cons_ptr) - 1), cons_ptr.ctypes.data_as(C.POINTER(C.c_uint32)),
                                             kind.ctypes.data_as(u8p), idx.ctypes.data_as(C.POINTER(C.c_uint32)), coeff.ctypes.data_as(u8p),
                                             C.byref(c._h)), "circuit_from_arrays")
        c._sizes()
        return c

    def _sizes(self):
        l = load()
        self.n = l.bp_circuit_num_multipliers(self._h)
        self.q = l.bp_circuit_num_constraints(self._h)
        self.m = l.bp_circuit_num_commitments(self._h)
        self.num_aux = l.bp_circuit_num_aux(self._h)
        self.num_public = l.bp_circuit_num_public(self._h)
        self.has_witness_program = bool(l.bp_circuit_has_witness_program(self._h))
        self.proof_len = l.bp_circuit_proof_len(self._h)

    def __del__(self):
        if _lib is not None and getattr(self, "_h", None):
            _lib.bp_circuit_free(self._h)
            self._h = None

    def release_workspace(self):
        """frees the device workspace kept between batch calls (the next call allocates it again)"""
        _check(load().bp_circuit_release_workspace(self._h), "circuit_release_workspace")

    def prove_batch(self, gens, label, v, v_blinding, entropy, aux=None, pub=None, witness=None):
        """v, v_blinding: uint8 [B][m][32]; entropy [B][32]; aux [B][num_aux][32]; pub [B][num_public][32]; witness = (aL,aR,aO) each [B][n][32] or None.
        Returns (V [B][m][32], proofs [B][proof_len], status [B]).  Host buffers in, host buffers out."""
        v, v_blinding, entropy = _np_u8(v), _np_u8(v_blinding), _np_u8(entropy)
        B = entropy.shape[0]
        V = np.zeros((B, self.m, 32), dtype=np.uint8)
        proofs = np.zeros((B, self.proof_len), dtype=np.uint8)
        status = np.zeros(B, dtype=np.int32)
        p = lambda a: a.ctypes.data_as(u8p) if a is not None else None
        aux = _np_u8(aux) if aux is not None else None
        pub = _np_u8(pub) if pub is not None else None
        wl = [_np_u8(w) for w in witness] if witness is not None else [None, None, None]
        self._check_prover_shapes(B, v, v_blinding, entropy, aux, pub, wl)
        _check(load().bp_prove_batch(gens._h, self._h, C.c_uint32(B), _buf(label) if label else None, C.c_size_t(len(label)), p(v), p(v_blinding),
                                     p(entropy), p(aux), p(pub), p(wl[0]), p(wl[1]), p(wl[2]), p(V), p(proofs), status.ctypes.data_as(C.POINTER(C.c_int32))),
               "prove_batch")
        return V, proofs, status

    def _check_prover_shapes(self, B, v, v_blinding, entropy, aux, pub, wl):
        _want(entropy, (B, 32), "entropy"); _want(v, (B, self.m, 32), "v"); _want(v_blinding, (B, self.m, 32), "v_blinding")
        if wl[0] is not None:
            for w, name in zip(wl, ("aL", "aR", "aO")):
                _want(w, (B, self.n, 32), name)
        elif self.num_aux:
            _want(aux, (B, self.num_aux, 32), "aux")
        if pub is not None:
            _want(pub, (B, self.num_public, 32), "pub")

    def _check_verifier_shapes(self, B, V, proofs, entropy, pub):
        _want(entropy, (B, 32), "entropy"); _want(V, (B, self.m, 32), "V"); _want(proofs, (B, self.proof_len), "proofs", code=2)
        if self.num_public:
            _want(pub, (B, self.num_public, 32), "pub")

    def verify_batch(self, gens, label, V, proofs, entropy, pub=None):
        V, proofs, entropy = _np_u8(V), _np_u8(proofs), _np_u8(entropy)
        pub = _np_u8(pub) if pub is not None else None
        B = entropy.shape[0]
        self._check_verifier_shapes(B, V, proofs, entropy, pub)
        status = np.zeros(B, dtype=np.int32)
        _check(load().bp_verify_batch(gens._h, self._h, C.c_uint32(B), _buf(label) if label else None, C.c_size_t(len(label)), V.ctypes.data_as(u8p),
                                      proofs.ctypes.data_as(u8p), entropy.ctypes.data_as(u8p), pub.ctypes.data_as(u8p) if pub is not None else None,
                                      status.ctypes.data_as(C.POINTER(C.c_int32))), "verify_batch")
        return status


class ProveStream:
    """Pipeline of batches over one circuit (bp_prove_stream_* of include/bp_b200.h): `begin(slot, ...)` launches the
    latency-bound first phase of a batch, `finish(slot)` the rest.  Enqueue begin(batch k+1) before finish(batch k) and the first
    phase of the next batch runs beside the MSM phase of the current one.  Host buffers (numpy, or pinned torch tensors' numpy
    views); the device-pointer form is bp_prove_stream_begin / _finish."""

    def __init__(self, circuit, gens, label, stream=None):
        self.circuit, self.gens, self.label, self.stream = circuit, gens, bytes(label), stream
        self._shape = [None, None]

    def begin(self, slot, v, v_blinding, entropy, aux=None, pub=None):
        v, v_blinding, entropy = _np_u8(v), _np_u8(v_blinding), _np_u8(entropy)
        aux = _np_u8(aux) if aux is not None else None
        pub = _np_u8(pub) if pub is not None else None
        B = entropy.shape[0]
        self.circuit._check_prover_shapes(B, v, v_blinding, entropy, aux, pub, [None, None, None])
        p = lambda a: a.ctypes.data_as(u8p) if a is not None else None
        _check(load().bp_prove_stream_begin_host(self.gens._h, self.circuit._h, C.c_int32(slot), C.c_uint32(B), _buf(self.label) if self.label else None,
                                                 C.c_size_t(len(self.label)), p(v), p(v_blinding), p(entropy), p(aux), p(pub),
                                                 C.c_void_p(self.stream) if self.stream else None), "prove_stream_begin_host")
        self._shape[slot] = (B, (v, v_blinding, entropy, aux, pub))  # keeps the host buffers alive until finish

    def finish(self, slot, out=None):
        """returns (V, proofs, status) of the slot's batch; `out` = preallocated (V, proofs, status) arrays"""
        if self._shape[slot] is None:
            raise R1CSError(6, "prove_stream_finish_host: slot not begun")
        B = self._shape[slot][0]
        c = self.circuit
        V, proofs, status = out if out is not None else (np.zeros((B, c.m, 32), dtype=np.uint8), np.zeros((B, c.proof_len), dtype=np.uint8),
                                                         np.zeros(B, dtype=np.int32))
        _want(V, (B, c.m, 32), "out V"); _want(proofs, (B, c.proof_len), "out proofs"); _want(status, (B,), "out status")
        if V.dtype != np.uint8 or proofs.dtype != np.uint8 or status.dtype != np.int32 or not (V.flags.c_contiguous and proofs.flags.c_contiguous):
            raise R1CSError(6, "out buffers must be contiguous uint8 / int32 arrays")
        _check(load().bp_prove_stream_finish_host(self.gens._h, c._h, C.c_int32(slot), V.ctypes.data_as(u8p), proofs.ctypes.data_as(u8p),
                                                  status.ctypes.data_as(C.POINTER(C.c_int32)), C.c_void_p(self.stream) if self.stream else None),
               "prove_stream_finish_host")
        self._shape[slot] = None
        return V, proofs, status


def _verify_batch_combined(self, gens, label, V, proofs, entropy, pub=None):
    """cross-proof batched verification: returns (status [B] of the structural checks, combined verdict 0 / 3)"""
    V, proofs, entropy = _np_u8(V), _np_u8(proofs), _np_u8(entropy)
    pub = _np_u8(pub) if pub is not None else None
    B = entropy.shape[0]
    self._check_verifier_shapes(B, V, proofs, entropy, pub)
    status = np.zeros(B, dtype=np.int32)
    combined = C.c_int32(0)
    f = load().bp_verify_batch_combined
    f.restype = C.c_int32
    _check(f(gens._h, self._h, C.c_uint32(B), _buf(label) if label else None, C.c_size_t(len(label)), V.ctypes.data_as(u8p),
             proofs.ctypes.data_as(u8p), entropy.ctypes.data_as(u8p), pub.ctypes.data_as(u8p) if pub is not None else None,
             status.ctypes.data_as(C.POINTER(C.c_int32)), C.byref(combined)), "verify_batch_combined")
    return status, int(combined.value)


Circuit.verify_batch_combined = _verify_batch_combined


def proof_to_wire(proof):
    """untagged field tuple -> tagged wire form (App. A.7)"""
    proof = bytes(proof)
    out = (C.c_uint8 * (len(proof) + 1))()
    f = load().bp_proof_to_wire
    f.restype = C.c_int64
    n = f(_buf(proof), C.c_size_t(len(proof)), out, C.c_size_t(len(proof) + 1))
    if n < 0:
        raise R1CSError(-n, "proof_to_wire")
    return bytes(out[:n])


def proof_from_wire(wire):
    wire = bytes(wire)
    out = (C.c_uint8 * (len(wire) + 96))()
    f = load().bp_proof_from_wire
    f.restype = C.c_int64
    n = f(_buf(wire), C.c_size_t(len(wire)), out, C.c_size_t(len(wire) + 96))
    if n < 0:
        raise R1CSError(-n, "proof_from_wire")
    return bytes(out[:n])


def launch_count():
    return int(load().bp_launch_count())


def profile_enable(on=True):
    """True / 1: CUDA events around every kernel launch (costs ~2.5 % of a depth-32 step); 2: around the dominant kernel
    (KBucketAccumulate) only; False / 0: off"""
    load().bp_profile_enable(int(on))


def profile_report():
    """{kernel: (launches, total_ms, threads)} since the last report; call after a device synchronise.
    The pseudo-kernel "@sorted_items" carries, in the `threads` field, the exact number of point additions KBucketAccumulate did."""
    buf = C.create_string_buffer(1 << 16)
    n = load().bp_profile_report(buf, C.c_size_t(len(buf)))
    out = {}
    for line in buf.raw[:n].decode().splitlines():
        name, launches, ms, threads = line.split()
        out[name] = (int(launches), float(ms), float(threads))
    return out
