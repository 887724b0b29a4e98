"""TEST INFRASTRUCTURE ONLY -- ctypes loader for the C oracle (oracle/bp_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module.  It also turns a constraint system recorded by oracle/bp_pyref.py
into the flat CSR arrays both the C oracle and the product's batched entry points take.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libbp_oracle.so")
_lib = None

u8p = C.POINTER(C.c_uint8)
u32p = C.POINTER(C.c_uint32)


def build(force=False):
    if force or not os.path.exists(_SO) or any(
            os.path.getmtime(os.path.join(_HERE, f)) > os.path.getmtime(_SO)
            for f in ("bp_oracle.c", "ed25519_ref.h", "merlin_ref.h")):
        subprocess.check_call(["make", "-C", _HERE, "-B", "_build/libbp_oracle.so"], stdout=subprocess.DEVNULL)
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
        _lib.bpo_proof_len.restype = C.c_size_t
        _lib.bpo_proof_len.argtypes = [C.c_uint32]
        _lib.bpo_init()
    return _lib


def _b(x):
    """bytes-like / numpy -> (keepalive, u8 pointer)"""
    if isinstance(x, np.ndarray):
        a = np.ascontiguousarray(x, dtype=np.uint8)
    else:
        a = np.frombuffer(bytes(x), dtype=np.uint8)
    return a, a.ctypes.data_as(u8p)


def sc_bytes(x, L):
    return (int(x) % L).to_bytes(32, "little")


class Circuit:
    """Flat CSR view of a recorded constraint system (kinds as in bp_pyref: 0 committed,1 L,2 R,3 O,4 one)."""

    def __init__(self, n, m, cons_ptr, kind, idx, coeff):
        self.n, self.m = int(n), int(m)
        self.cons_ptr = np.ascontiguousarray(cons_ptr, dtype=np.uint32)
        self.kind = np.ascontiguousarray(kind, dtype=np.uint8)
        self.idx = np.ascontiguousarray(idx, dtype=np.uint32)
        self.coeff = np.ascontiguousarray(coeff, dtype=np.uint8).reshape(-1, 32)
        self.q = len(self.cons_ptr) - 1

    @staticmethod
    def from_cs(cs, m):
        """cs: bp_pyref Prover or Verifier after the gadget ran."""
        from .bp_pyref import L
        ptr, kind, idx, coeff = [0], [], [], bytearray()
        for lc in cs.constraints:
            for (k, i), c in lc.terms:
                kind.append(k)
                idx.append(i)
                coeff += (c % L).to_bytes(32, "little")
            ptr.append(len(kind))
        return Circuit(cs.num_multipliers(), m, ptr, kind, idx, np.frombuffer(bytes(coeff), dtype=np.uint8))

    def c_args(self):
        return (C.c_uint32(self.n), C.c_uint32(self.m), C.c_uint32(self.q),
                self.cons_ptr.ctypes.data_as(u32p), self.kind.ctypes.data_as(u8p),
                self.idx.ctypes.data_as(u32p), self.coeff.ctypes.data_as(u8p))


def scalars_to_array(vals, L):
    return np.frombuffer(b"".join((int(v) % L).to_bytes(32, "little") for v in vals), dtype=np.uint8).reshape(-1, 32).copy() \
        if len(vals) else np.zeros((0, 32), dtype=np.uint8)


def prove(circ, aL, aR, aO, v, vbl, label, entropy, gens_capacity):
    """aL.. : uint8 arrays [n][32]; v, vbl: [m][32].  Returns (status, V[m][32], proof bytes)."""
    l = lib()
    plen = l.bpo_proof_len(circ.n)
    V = np.zeros((circ.m, 32), dtype=np.uint8)
    proof = np.zeros(plen, dtype=np.uint8)
    keep = [_b(x) for x in (aL, aR, aO, v, vbl, label, entropy)]
    rc = l.bpo_prove(*circ.c_args(), keep[0][1], keep[1][1], keep[2][1], keep[3][1], keep[4][1],
                     keep[5][1], C.c_uint32(len(label)), keep[6][1], C.c_uint32(gens_capacity),
                     V.ctypes.data_as(u8p), proof.ctypes.data_as(u8p))
    return rc, V, proof.tobytes()


def verify(circ, V, proof, label, entropy, gens_capacity):
    l = lib()
    keep = [_b(x) for x in (V, proof, label, entropy)]
    return l.bpo_verify(*circ.c_args(), keep[0][1], keep[1][1], C.c_size_t(len(proof)), keep[2][1],
                        C.c_uint32(len(label)), keep[3][1], C.c_uint32(gens_capacity))


def prove_batch(circ, B, v, vbl, entropy, label, gens_capacity, nthreads, witness_kind=0, depth=0, aL=None, aR=None, aO=None):
    l = lib()
    plen = l.bpo_proof_len(circ.n)
    V = np.zeros((B, circ.m, 32), dtype=np.uint8)
    proofs = np.zeros((B, plen), dtype=np.uint8)
    status = np.zeros(B, dtype=np.int32)
    z = np.zeros(1, dtype=np.uint8)
    keep = [_b(x if x is not None else z) for x in (aL, aR, aO, v, vbl, label, entropy)]
    l.bpo_prove_batch(*circ.c_args(), C.c_uint32(B), C.c_int(witness_kind), C.c_uint32(depth),
                      keep[0][1], keep[1][1], keep[2][1], keep[3][1], keep[4][1], keep[5][1], C.c_uint32(len(label)),
                      keep[6][1], C.c_uint32(gens_capacity), C.c_uint32(nthreads),
                      V.ctypes.data_as(u8p), proofs.ctypes.data_as(u8p), C.c_size_t(plen),
                      status.ctypes.data_as(C.POINTER(C.c_int)))
    return status, V, proofs


def poseidon_set_params(blob, width=6, fb=4, fe=4, pr=140):
    a, p = _b(blob)
    rc = lib().bpo_poseidon_set_params(p, C.c_uint32(len(a) // 32), C.c_uint32(width), C.c_uint32(fb), C.c_uint32(fe), C.c_uint32(pr))
    assert rc == 0
    return rc


def vsmt2_witness(depth, leaf, bits, sibs, n):
    """returns aL,aR,aO [n][32] and root bytes"""
    l = lib()
    aL = np.zeros((n, 32), dtype=np.uint8)
    aR = np.zeros((n, 32), dtype=np.uint8)
    aO = np.zeros((n, 32), dtype=np.uint8)
    root = np.zeros(32, dtype=np.uint8)
    kb = [_b(leaf), _b(bytes(bits)), _b(sibs)]
    l.bpo_vsmt2_witness.restype = C.c_uint32
    got = l.bpo_vsmt2_witness(C.c_uint32(depth), kb[0][1], kb[1][1], kb[2][1], aL.ctypes.data_as(u8p), aR.ctypes.data_as(u8p),
                              aO.ctypes.data_as(u8p), root.ctypes.data_as(u8p))
    assert got == n, (got, n)
    return aL, aR, aO, root.tobytes()


def poseidon_hash2_witness(xl, xr, inverse, n):
    l = lib()
    aL = np.zeros((n, 32), dtype=np.uint8)
    aR = np.zeros((n, 32), dtype=np.uint8)
    aO = np.zeros((n, 32), dtype=np.uint8)
    h = np.zeros(32, dtype=np.uint8)
    kb = [_b(xl), _b(xr)]
    l.bpo_poseidon_hash2_witness.restype = C.c_uint32
    got = l.bpo_poseidon_hash2_witness(kb[0][1], kb[1][1], C.c_int(inverse), aL.ctypes.data_as(u8p), aR.ctypes.data_as(u8p),
                                       aO.ctypes.data_as(u8p), h.ctypes.data_as(u8p))
    assert got == n, (got, n)
    return aL, aR, aO, h.tobytes()


def mimc_witness(xl, xr, constants, n):
    l = lib()
    rounds = len(constants) // 32
    aL = np.zeros((n, 32), dtype=np.uint8)
    aR = np.zeros((n, 32), dtype=np.uint8)
    aO = np.zeros((n, 32), dtype=np.uint8)
    img = np.zeros(32, dtype=np.uint8)
    kb = [_b(xl), _b(xr), _b(constants)]
    l.bpo_mimc_witness.restype = C.c_uint32
    got = l.bpo_mimc_witness(kb[0][1], kb[1][1], C.c_uint32(rounds), kb[2][1], aL.ctypes.data_as(u8p), aR.ctypes.data_as(u8p),
                             aO.ctypes.data_as(u8p), img.ctypes.data_as(u8p))
    assert got == n
    return aL, aR, aO, img.tobytes()


def call_bytes(name, outlen, *ins):
    """generic helper for the KAT entry points taking byte buffers and one output buffer"""
    l = lib()
    out = np.zeros(outlen, dtype=np.uint8)
    keep = [_b(x) for x in ins]
    rc = getattr(l, name)(*[k[1] for k in keep], out.ctypes.data_as(u8p))
    return rc, out.tobytes()
