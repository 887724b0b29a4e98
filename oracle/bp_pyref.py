"""TEST INFRASTRUCTURE ONLY -- big-integer Python restatement of the reference hot path.

This file is the slow, obviously-correct cross-check for the C oracle
(`oracle/bp_oracle.c`) and, through it, for the CUDA product.  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline leg may import it.

PARITY UNPINNED: the reference (lovesh/bulletproofs-r1cs-gadgets) holds no golden
proof bytes, and its arithmetic lives in un-vendored crates (bulletproofs fork
branch "smt", curve25519-dalek 2.x, merlin 2.x -- reference Cargo.toml:8,18,22-26).
What is pinned: RFC 9496 ristretto255 vectors, the Merlin test vector, the
reference's own constants file and gadget shapes (see tests/).

Restated algorithms (reference call sites in brackets):
  * scalar field mod l, GF(2^255-19), ristretto255 (RFC 9496)    [Cargo.toml:8]
  * Merlin / STROBE-128 transcripts and transcript RNG             [Cargo.toml:18]
  * PedersenGens / BulletproofGens generator chains                [gadget_vsmt_2.rs:289-290]
  * R1CS ConstraintSystem, Prover::prove, Verifier::verify, IPA    [gadget_vsmt_2.rs:294-347,356-395]
  * gadgets: see oracle/gadgets_pyref.py
"""
import hashlib

# ----------------------------------------------------------------------------
# fields
# ----------------------------------------------------------------------------
P = 2**255 - 19
L = 2**252 + 27742317777372353535851937790883648493
D = (-121665 * pow(121666, P - 2, P)) % P
SQRT_M1 = pow(2, (P - 1) // 4, P)
ONE_MINUS_D_SQ = (1 - D * D) % P
D_MINUS_ONE_SQ = ((D - 1) * (D - 1)) % P


def _is_neg(x):
    return (x % P) & 1


def _abs(x):
    x %= P
    return P - x if x & 1 else x


def sqrt_ratio_m1(u, v):
    """RFC 9496 section 4.2 SQRT_RATIO_M1."""
    u %= P
    v %= P
    v3 = v * v % P * v % P
    v7 = v3 * v3 % P * v % P
    r = u * v3 % P * pow(u * v7 % P, (P - 5) // 8, P) % P
    check = v * r % P * r % P
    correct = check == u
    flipped = check == (-u) % P
    flipped_i = check == (-u * SQRT_M1) % P
    if flipped or flipped_i:
        r = r * SQRT_M1 % P
    r = _abs(r)
    return (correct or flipped), r


_, SQRT_AD_MINUS_ONE = sqrt_ratio_m1((-1 * D - 1) % P, 1)  # sqrt(a*d - 1), a = -1
_ok, INVSQRT_A_MINUS_D = sqrt_ratio_m1(1, (-1 - D) % P)
assert _ok
# RFC 9496 fixes the sign of SQRT_AD_MINUS_ONE by its decimal constant:
SQRT_AD_MINUS_ONE = 25063068953384623474111414158702152701244531502492656460079210482610430750235
assert SQRT_AD_MINUS_ONE * SQRT_AD_MINUS_ONE % P == (-D - 1) % P


# ----------------------------------------------------------------------------
# points: extended twisted Edwards (X, Y, Z, T), a = -1
# ----------------------------------------------------------------------------
IDENT = (0, 1, 1, 0)


def pt_add(p, q):
    X1, Y1, Z1, T1 = p
    X2, Y2, Z2, T2 = q
    A = (Y1 - X1) * (Y2 - X2) % P
    B = (Y1 + X1) * (Y2 + X2) % P
    C = T1 * 2 * D % P * T2 % P
    Dd = Z1 * 2 * Z2 % P
    E, F, G, H = B - A, Dd - C, Dd + C, B + A
    return (E * F % P, G * H % P, F * G % P, E * H % P)


def pt_dbl(p):
    return pt_add(p, p)


def pt_neg(p):
    X, Y, Z, T = p
    return ((-X) % P, Y, Z, (-T) % P)


def pt_mul(s, p):
    s %= L
    q = IDENT
    while s:
        if s & 1:
            q = pt_add(q, p)
        p = pt_dbl(p)
        s >>= 1
    return q


def pt_eq(p, q):
    # ristretto equality
    X1, Y1, _, _ = p
    X2, Y2, _, _ = q
    return (X1 * Y2 - Y1 * X2) % P == 0 or (Y1 * Y2 - X1 * X2) % P == 0


def ristretto_decode(b):
    s = int.from_bytes(b, "little")
    if s >= P or (s & 1):
        return None
    ss = s * s % P
    u1 = (1 - ss) % P
    u2 = (1 + ss) % P
    u2s = u2 * u2 % P
    v = (-(D * u1 % P * u1) - u2s) % P
    ok, invsqrt = sqrt_ratio_m1(1, v * u2s % P)
    den_x = invsqrt * u2 % P
    den_y = invsqrt * den_x % P * v % P
    x = _abs(2 * s * den_x % P)
    y = u1 * den_y % P
    t = x * y % P
    if (not ok) or _is_neg(t) or y == 0:
        return None
    return (x, y, 1, t)


def ristretto_encode(p):
    X0, Y0, Z0, T0 = p
    u1 = (Z0 + Y0) * (Z0 - Y0) % P
    u2 = X0 * Y0 % P
    _, invsqrt = sqrt_ratio_m1(1, u1 * u2 % P * u2 % P)
    den1 = invsqrt * u1 % P
    den2 = invsqrt * u2 % P
    z_inv = den1 * den2 % P * T0 % P
    ix0 = X0 * SQRT_M1 % P
    iy0 = Y0 * SQRT_M1 % P
    enchanted = den1 * INVSQRT_A_MINUS_D % P
    rotate = _is_neg(T0 * z_inv)
    if rotate:
        x, y, den_inv = iy0, ix0, enchanted
    else:
        x, y, den_inv = X0, Y0, den2
    if _is_neg(x * z_inv):
        y = (-y) % P
    s = _abs(den_inv * ((Z0 - y) % P))
    return s.to_bytes(32, "little")


def elligator(t):
    """RFC 9496 MAP."""
    r = SQRT_M1 * t % P * t % P
    u = (r + 1) * ONE_MINUS_D_SQ % P
    v = (-1 - r * D) % P * ((r + D) % P) % P
    was_square, s = sqrt_ratio_m1(u, v)
    s_prime = (-_abs(s * t)) % P
    if not was_square:
        s = s_prime
        c = r
    else:
        c = P - 1
    N = (c * ((r - 1) % P) % P * D_MINUS_ONE_SQ - v) % P
    w0 = 2 * s * v % P
    w1 = N * SQRT_AD_MINUS_ONE % P
    w2 = (1 - s * s) % P
    w3 = (1 + s * s) % P
    return (w0 * w3 % P, w2 * w1 % P, w1 * w3 % P, w0 * w2 % P)


def from_uniform_bytes(b):
    assert len(b) == 64
    r0 = int.from_bytes(b[:32], "little") & ((1 << 255) - 1)
    r1 = int.from_bytes(b[32:], "little") & ((1 << 255) - 1)
    return pt_add(elligator(r0 % P), elligator(r1 % P))


BASEPOINT_COMPRESSED = bytes.fromhex("e2f2ae0a6abc4e71a884a961c500515f58e30b6aa582dd8db6a65945e08d2d76")
BASEPOINT = ristretto_decode(BASEPOINT_COMPRESSED)


def msm(scalars, points):
    """Pippenger (variable time); canonical encodings make algorithm choice irrelevant."""
    n = len(scalars)
    if n == 0:
        return IDENT
    c = 4 if n < 32 else (6 if n < 512 else (8 if n < 8192 else 10))
    scalars = [s % L for s in scalars]
    res = IDENT
    nwin = (253 + c - 1) // c
    for w in range(nwin - 1, -1, -1):
        for _ in range(c):
            res = pt_dbl(res)
        buckets = [None] * (1 << c)
        sh = w * c
        mask = (1 << c) - 1
        for s, p in zip(scalars, points):
            d = (s >> sh) & mask
            if d:
                buckets[d] = p if buckets[d] is None else pt_add(buckets[d], p)
        run = None
        tot = None
        for d in range(mask, 0, -1):
            if buckets[d] is not None:
                run = buckets[d] if run is None else pt_add(run, buckets[d])
            if run is not None:
                tot = run if tot is None else pt_add(tot, run)
        if tot is not None:
            res = pt_add(res, tot)
    return res


# ----------------------------------------------------------------------------
# Keccak-f[1600], STROBE-128, Merlin
# ----------------------------------------------------------------------------
_RC = [
    0x0000000000000001, 0x0000000000008082, 0x800000000000808A, 0x8000000080008000,
    0x000000000000808B, 0x0000000080000001, 0x8000000080008081, 0x8000000000008009,
    0x000000000000008A, 0x0000000000000088, 0x0000000080008009, 0x000000008000000A,
    0x000000008000808B, 0x800000000000008B, 0x8000000000008089, 0x8000000000008003,
    0x8000000000008002, 0x8000000000000080, 0x000000000000800A, 0x800000008000000A,
    0x8000000080008081, 0x8000000000008080, 0x0000000080000001, 0x8000000080008008,
]
_ROT = [[0, 36, 3, 41, 18], [1, 44, 10, 45, 2], [62, 6, 43, 15, 61], [28, 55, 25, 21, 56], [27, 20, 39, 8, 14]]
_M64 = (1 << 64) - 1


def _rol(x, n):
    n %= 64
    return ((x << n) | (x >> (64 - n))) & _M64 if n else x


def keccak_f(state_bytes):
    A = [[int.from_bytes(state_bytes[8 * (x + 5 * y): 8 * (x + 5 * y) + 8], "little") for y in range(5)] for x in range(5)]
    for rnd in range(24):
        C = [A[x][0] ^ A[x][1] ^ A[x][2] ^ A[x][3] ^ A[x][4] for x in range(5)]
        Dv = [C[(x - 1) % 5] ^ _rol(C[(x + 1) % 5], 1) for x in range(5)]
        A = [[A[x][y] ^ Dv[x] for y in range(5)] for x in range(5)]
        B = [[0] * 5 for _ in range(5)]
        for x in range(5):
            for y in range(5):
                B[y][(2 * x + 3 * y) % 5] = _rol(A[x][y], _ROT[x][y])
        A = [[B[x][y] ^ ((~B[(x + 1) % 5][y]) & B[(x + 2) % 5][y]) for y in range(5)] for x in range(5)]
        A[0][0] ^= _RC[rnd]
    out = bytearray(200)
    for x in range(5):
        for y in range(5):
            out[8 * (x + 5 * y): 8 * (x + 5 * y) + 8] = A[x][y].to_bytes(8, "little")
    return out


STROBE_R = 166
FLAG_I, FLAG_A, FLAG_C, FLAG_T, FLAG_M, FLAG_K = 1, 2, 4, 8, 16, 32


class Strobe128:
    def __init__(self, protocol_label=None):
        if protocol_label is None:
            return
        st = bytearray(200)
        st[0:6] = bytes([1, STROBE_R + 2, 1, 0, 1, 96])
        st[6:18] = b"STROBEv1.0.2"
        self.state = keccak_f(st)
        self.pos = 0
        self.pos_begin = 0
        self.cur_flags = 0
        self.meta_ad(protocol_label, False)

    def clone(self):
        c = Strobe128()
        c.state = bytearray(self.state)
        c.pos, c.pos_begin, c.cur_flags = self.pos, self.pos_begin, self.cur_flags
        return c

    def _run_f(self):
        self.state[self.pos] ^= self.pos_begin
        self.state[self.pos + 1] ^= 0x04
        self.state[STROBE_R + 1] ^= 0x80
        self.state = keccak_f(self.state)
        self.pos = 0
        self.pos_begin = 0

    def _absorb(self, data):
        for b in data:
            self.state[self.pos] ^= b
            self.pos += 1
            if self.pos == STROBE_R:
                self._run_f()

    def _overwrite(self, data):
        for b in data:
            self.state[self.pos] = b
            self.pos += 1
            if self.pos == STROBE_R:
                self._run_f()

    def _squeeze(self, n):
        out = bytearray()
        for _ in range(n):
            out.append(self.state[self.pos])
            self.state[self.pos] = 0
            self.pos += 1
            if self.pos == STROBE_R:
                self._run_f()
        return bytes(out)

    def _begin_op(self, flags, more):
        if more:
            assert self.cur_flags == flags
            return
        assert flags & FLAG_T == 0
        old_begin = self.pos_begin
        self.pos_begin = self.pos + 1
        self.cur_flags = flags
        self._absorb(bytes([old_begin, flags]))
        if (flags & (FLAG_C | FLAG_K)) and self.pos != 0:
            self._run_f()

    def meta_ad(self, data, more):
        self._begin_op(FLAG_M | FLAG_A, more)
        self._absorb(data)

    def ad(self, data, more):
        self._begin_op(FLAG_A, more)
        self._absorb(data)

    def prf(self, n, more=False):
        self._begin_op(FLAG_I | FLAG_A | FLAG_C, more)
        return self._squeeze(n)

    def key(self, data, more=False):
        self._begin_op(FLAG_A | FLAG_C, more)
        self._overwrite(data)


def _u32le(n):
    return n.to_bytes(4, "little")


class Transcript:
    def __init__(self, label):
        self.strobe = Strobe128(b"Merlin v1.0")
        self.append_message(b"dom-sep", label)

    def append_message(self, label, msg):
        self.strobe.meta_ad(label, False)
        self.strobe.meta_ad(_u32le(len(msg)), True)
        self.strobe.ad(msg, False)

    def append_u64(self, label, x):
        self.append_message(label, x.to_bytes(8, "little"))

    def challenge_bytes(self, label, n):
        self.strobe.meta_ad(label, False)
        self.strobe.meta_ad(_u32le(n), True)
        return self.strobe.prf(n, False)

    # bulletproofs TranscriptProtocol
    def append_scalar(self, label, s):
        self.append_message(label, (s % L).to_bytes(32, "little"))

    def append_point(self, label, pbytes):
        self.append_message(label, pbytes)

    def validate_and_append_point(self, label, pbytes):
        if pbytes == bytes(32):
            raise VerificationError("identity point")
        self.append_message(label, pbytes)

    def challenge_scalar(self, label):
        return int.from_bytes(self.challenge_bytes(label, 64), "little") % L

    def build_rng(self, witness_blindings, entropy32):
        st = self.strobe.clone()
        for vb in witness_blindings:
            st.meta_ad(b"v_blinding", False)
            st.meta_ad(_u32le(32), True)
            st.key((vb % L).to_bytes(32, "little"), False)
        st.meta_ad(b"rng", False)
        st.key(entropy32, False)
        return TranscriptRng(st)


class TranscriptRng:
    def __init__(self, strobe):
        self.strobe = strobe

    def fill_bytes(self, n):
        self.strobe.meta_ad(_u32le(n), False)
        return self.strobe.prf(n, False)

    def random_scalar(self):
        return int.from_bytes(self.fill_bytes(64), "little") % L


# ----------------------------------------------------------------------------
# generators
# ----------------------------------------------------------------------------
class PedersenGens:
    def __init__(self):
        self.B = BASEPOINT
        self.B_blinding = from_uniform_bytes(hashlib.sha3_512(BASEPOINT_COMPRESSED).digest())

    def commit(self, v, r):
        return pt_add(pt_mul(v, self.B), pt_mul(r, self.B_blinding))


_GEN_CACHE = {}


def generators_chain(label, n):
    key = bytes(label)
    have = _GEN_CACHE.get(key, [])
    if len(have) < n:
        xof = hashlib.shake_256(b"GeneratorsChain" + key).digest(64 * n)
        for i in range(len(have), n):
            have.append(from_uniform_bytes(xof[64 * i: 64 * i + 64]))
        _GEN_CACHE[key] = have
    return have[:n]


class BulletproofGens:
    def __init__(self, capacity, parties=1):
        self.capacity = capacity

    def G(self, n):
        return generators_chain(b"G" + _u32le(0), n)

    def H(self, n):
        return generators_chain(b"H" + _u32le(0), n)


# ----------------------------------------------------------------------------
# R1CS
# ----------------------------------------------------------------------------
class R1CSError(Exception):
    pass


class VerificationError(R1CSError):
    pass


class MissingAssignment(R1CSError):
    pass


class InvalidGeneratorsLength(R1CSError):
    pass


# Variable = (kind, index); kinds:
COMMITTED, MUL_LEFT, MUL_RIGHT, MUL_OUT, ONE = 0, 1, 2, 3, 4


def Var(kind, idx=0):
    return (kind, idx)


VAR_ONE = (ONE, 0)


class LC:
    """LinearCombination: list of (Variable, coeff); duplicates allowed."""

    def __init__(self, terms=None):
        self.terms = list(terms) if terms else []

    @staticmethod
    def of(x):
        if isinstance(x, LC):
            return x
        if isinstance(x, tuple):
            return LC([(x, 1)])
        return LC([(VAR_ONE, x % L)])

    def __add__(self, o):
        return LC(self.terms + LC.of(o).terms)

    def __sub__(self, o):
        return LC(self.terms + [(v, (-c) % L) for v, c in LC.of(o).terms])

    def __neg__(self):
        return LC([(v, (-c) % L) for v, c in self.terms])

    def __mul__(self, s):
        return LC([(v, c * s % L) for v, c in self.terms])

    def simplify(self):
        d = {}
        for v, c in self.terms:
            d[v] = (d.get(v, 0) + c) % L
        return LC(list(d.items()))


def inv_mod_l(x):
    return pow(x % L, L - 2, L)


class ConstraintSystemBase:
    def __init__(self):
        self.constraints = []
        self.pending = None

    def constrain(self, lc):
        self.constraints.append(LC.of(lc))

    def num_constraints(self):
        return len(self.constraints)


class Prover(ConstraintSystemBase):
    def __init__(self, pc_gens, transcript):
        super().__init__()
        self.pc_gens = pc_gens
        self.t = transcript
        self.t.append_message(b"dom-sep", b"r1cs v1")
        self.v = []
        self.v_blinding = []
        self.aL, self.aR, self.aO = [], [], []

    def commit(self, v, v_blinding):
        i = len(self.v)
        self.v.append(v % L)
        self.v_blinding.append(v_blinding % L)
        V = ristretto_encode(self.pc_gens.commit(v, v_blinding))
        self.t.append_point(b"V", V)
        return V, Var(COMMITTED, i)

    def evaluate_lc(self, lc):
        acc = 0
        for (k, i), c in LC.of(lc).terms:
            val = (self.v[i] if k == COMMITTED else self.aL[i] if k == MUL_LEFT else
                   self.aR[i] if k == MUL_RIGHT else self.aO[i] if k == MUL_OUT else 1)
            acc += val * c
        return acc % L

    def multiply(self, left, right):
        left, right = LC.of(left), LC.of(right)
        l = self.evaluate_lc(left)
        r = self.evaluate_lc(right)
        i = len(self.aL)
        self.aL.append(l)
        self.aR.append(r)
        self.aO.append(l * r % L)
        self.constrain(left - Var(MUL_LEFT, i))
        self.constrain(right - Var(MUL_RIGHT, i))
        return Var(MUL_LEFT, i), Var(MUL_RIGHT, i), Var(MUL_OUT, i)

    def allocate_single(self, assignment):
        if assignment is None:
            raise MissingAssignment()
        if self.pending is None:
            i = len(self.aL)
            self.pending = i
            self.aL.append(assignment % L)
            self.aR.append(0)
            self.aO.append(0)
            return Var(MUL_LEFT, i), None
        i = self.pending
        self.pending = None
        self.aR[i] = assignment % L
        self.aO[i] = self.aL[i] * self.aR[i] % L
        return Var(MUL_RIGHT, i), Var(MUL_OUT, i)

    def allocate_multiplier(self, assignment):
        if assignment is None:
            raise MissingAssignment()
        l, r = assignment
        i = len(self.aL)
        self.aL.append(l % L)
        self.aR.append(r % L)
        self.aO.append(l * r % L)
        return Var(MUL_LEFT, i), Var(MUL_RIGHT, i), Var(MUL_OUT, i)

    def num_multipliers(self):
        return len(self.aL)

    def _flatten(self, z):
        n, m = len(self.aL), len(self.v)
        wL, wR, wO, wV = [0] * n, [0] * n, [0] * n, [0] * m
        ez = z
        for lc in self.constraints:
            for (k, i), c in lc.terms:
                if k == MUL_LEFT:
                    wL[i] = (wL[i] + ez * c) % L
                elif k == MUL_RIGHT:
                    wR[i] = (wR[i] + ez * c) % L
                elif k == MUL_OUT:
                    wO[i] = (wO[i] + ez * c) % L
                elif k == COMMITTED:
                    wV[i] = (wV[i] - ez * c) % L
            ez = ez * z % L
        return wL, wR, wO, wV

    def prove(self, bp_gens, entropy32, trace=None):
        t = self.t
        m = len(self.v)
        t.append_u64(b"m", m)
        rng = t.build_rng(self.v_blinding, entropy32)
        n = len(self.aL)
        if bp_gens.capacity < n:
            raise InvalidGeneratorsLength()
        i_b, o_b, s_b = rng.random_scalar(), rng.random_scalar(), rng.random_scalar()
        sL = [rng.random_scalar() for _ in range(n)]
        sR = [rng.random_scalar() for _ in range(n)]
        Bb = self.pc_gens.B_blinding
        N = 1
        while N < n:
            N *= 2
        if bp_gens.capacity < N:
            raise InvalidGeneratorsLength()
        G, H = bp_gens.G(N), bp_gens.H(N)
        A_I1 = ristretto_encode(msm([i_b] + self.aL + self.aR, [Bb] + G[:n] + H[:n]))
        A_O1 = ristretto_encode(msm([o_b] + self.aO, [Bb] + G[:n]))
        S1 = ristretto_encode(msm([s_b] + sL + sR, [Bb] + G[:n] + H[:n]))
        t.append_point(b"A_I1", A_I1)
        t.append_point(b"A_O1", A_O1)
        t.append_point(b"S1", S1)
        t.append_message(b"dom-sep", b"r1cs-1phase")
        ident = bytes(32)
        t.append_point(b"A_I2", ident)
        t.append_point(b"A_O2", ident)
        t.append_point(b"S2", ident)
        y = t.challenge_scalar(b"y")
        z = t.challenge_scalar(b"z")
        wL, wR, wO, wV = self._flatten(z)
        y_inv = inv_mod_l(y)
        exp_y_inv = [1] * N
        for i in range(1, N):
            exp_y_inv[i] = exp_y_inv[i - 1] * y_inv % L
        l1, l2, l3, r0, r1, r3 = [], [], [], [], [], []
        ey = 1
        for i in range(n):
            l1.append((self.aL[i] + exp_y_inv[i] * wR[i]) % L)
            l2.append(self.aO[i])
            l3.append(sL[i])
            r0.append((wO[i] - ey) % L)
            r1.append((ey * self.aR[i] + wL[i]) % L)
            r3.append(ey * sR[i] % L)
            ey = ey * y % L

        def ip(a, b):
            return sum(x * yv for x, yv in zip(a, b)) % L

        t1 = ip(l1, r0)
        t2 = (ip(l1, r1) + ip(l2, r0)) % L
        t3 = (ip(l2, r1) + ip(l3, r0)) % L
        t4 = (ip(l1, r3) + ip(l3, r1)) % L
        t5 = ip(l2, r3)
        t6 = ip(l3, r3)
        tb = {j: rng.random_scalar() for j in (1, 3, 4, 5, 6)}
        pc = self.pc_gens
        T = {j: ristretto_encode(pc.commit(tv, tb[j])) for j, tv in ((1, t1), (3, t3), (4, t4), (5, t5), (6, t6))}
        for j in (1, 3, 4, 5, 6):
            t.append_point(b"T_%d" % j, T[j])
        u = t.challenge_scalar(b"u")
        x = t.challenge_scalar(b"x")
        tb[2] = ip(wV, self.v_blinding)
        tpoly = {1: t1, 2: t2, 3: t3, 4: t4, 5: t5, 6: t6}
        t_x = sum(tpoly[j] * pow(x, j, L) for j in range(1, 7)) % L
        t_x_blinding = sum(tb[j] * pow(x, j, L) for j in range(1, 7)) % L
        x2, x3 = x * x % L, x * x * x % L
        l_vec = [(l1[i] * x + l2[i] * x2 + l3[i] * x3) % L for i in range(n)] + [0] * (N - n)
        r_vec = [(r0[i] + r1[i] * x + r3[i] * x3) % L for i in range(n)]
        for i in range(n, N):
            r_vec.append((-ey) % L)
            ey = ey * y % L
        e_blinding = x * (i_b + x * (o_b + x * s_b)) % L
        t.append_scalar(b"t_x", t_x)
        t.append_scalar(b"t_x_blinding", t_x_blinding)
        t.append_scalar(b"e_blinding", e_blinding)
        w = t.challenge_scalar(b"w")
        Q = pt_mul(w, pc.B)
        G_factors = [1] * n + [u] * (N - n)
        H_factors = [exp_y_inv[i] * G_factors[i] % L for i in range(N)]
        if trace is not None:
            trace.update(dict(y=y, z=z, u=u, x=x, w=w, sL=sL, sR=sR, aL=list(self.aL), aR=list(self.aR), aO=list(self.aO),
                              l_vec=list(l_vec), r_vec=list(r_vec), t=tpoly, wL=wL, wR=wR, wO=wO, wV=wV))
        Ls, Rs, a, b = ipa_create(t, Q, G_factors, H_factors, list(G), list(H), l_vec, r_vec)
        return dict(A_I1=A_I1, A_O1=A_O1, S1=S1, A_I2=ident, A_O2=ident, S2=ident,
                    T_1=T[1], T_3=T[3], T_4=T[4], T_5=T[5], T_6=T[6],
                    t_x=t_x, t_x_blinding=t_x_blinding, e_blinding=e_blinding, L=Ls, R=Rs, a=a, b=b)


def ipa_create(t, Q, Gf, Hf, G, H, a, b):
    n = len(G)
    t.append_message(b"dom-sep", b"ipp v1")
    t.append_u64(b"n", n)
    Ls, Rs = [], []
    first = True
    while n != 1:
        n //= 2
        aL, aR, bL, bR = a[:n], a[n:], b[:n], b[n:]
        GL, GR, HL, HR = G[:n], G[n:], H[:n], H[n:]
        cL = sum(x * yv for x, yv in zip(aL, bR)) % L
        cR = sum(x * yv for x, yv in zip(aR, bL)) % L
        if first:
            Lp = msm([aL[i] * Gf[n + i] for i in range(n)] + [bR[i] * Hf[i] for i in range(n)] + [cL], GR + HL + [Q])
            Rp = msm([aR[i] * Gf[i] for i in range(n)] + [bL[i] * Hf[n + i] for i in range(n)] + [cR], GL + HR + [Q])
        else:
            Lp = msm(aL + bR + [cL], GR + HL + [Q])
            Rp = msm(aR + bL + [cR], GL + HR + [Q])
        Lc, Rc = ristretto_encode(Lp), ristretto_encode(Rp)
        Ls.append(Lc)
        Rs.append(Rc)
        t.append_point(b"L", Lc)
        t.append_point(b"R", Rc)
        u = t.challenge_scalar(b"u")
        ui = inv_mod_l(u)
        a = [(aL[i] * u + ui * aR[i]) % L for i in range(n)]
        b = [(bL[i] * ui + u * bR[i]) % L for i in range(n)]
        if first:
            G = [msm([ui * Gf[i], u * Gf[n + i]], [GL[i], GR[i]]) for i in range(n)]
            H = [msm([u * Hf[i], ui * Hf[n + i]], [HL[i], HR[i]]) for i in range(n)]
            first = False
        else:
            G = [msm([ui, u], [GL[i], GR[i]]) for i in range(n)]
            H = [msm([u, ui], [HL[i], HR[i]]) for i in range(n)]
    return Ls, Rs, a[0], b[0]


class Verifier(ConstraintSystemBase):
    def __init__(self, transcript):
        super().__init__()
        self.t = transcript
        self.t.append_message(b"dom-sep", b"r1cs v1")
        self.V = []
        self.num_vars = 0

    def commit(self, V):
        i = len(self.V)
        self.V.append(bytes(V))
        self.t.append_point(b"V", bytes(V))
        return Var(COMMITTED, i)

    def evaluate_lc(self, lc):
        return None

    def _alloc(self):
        i = self.num_vars
        self.num_vars += 1
        return i

    def multiply(self, left, right):
        left, right = LC.of(left), LC.of(right)
        i = self._alloc()
        self.constrain(left - Var(MUL_LEFT, i))
        self.constrain(right - Var(MUL_RIGHT, i))
        return Var(MUL_LEFT, i), Var(MUL_RIGHT, i), Var(MUL_OUT, i)

    def allocate_single(self, assignment):
        if self.pending is None:
            i = self._alloc()
            self.pending = i
            return Var(MUL_LEFT, i), None
        i = self.pending
        self.pending = None
        return Var(MUL_RIGHT, i), Var(MUL_OUT, i)

    def allocate_multiplier(self, assignment):
        i = self._alloc()
        return Var(MUL_LEFT, i), Var(MUL_RIGHT, i), Var(MUL_OUT, i)

    def num_multipliers(self):
        return self.num_vars

    def verify(self, proof, pc_gens, bp_gens, entropy32):
        t = self.t
        m = len(self.V)
        t.append_u64(b"m", m)
        t.validate_and_append_point(b"A_I1", proof["A_I1"])
        t.validate_and_append_point(b"A_O1", proof["A_O1"])
        t.validate_and_append_point(b"S1", proof["S1"])
        t.append_message(b"dom-sep", b"r1cs-1phase")
        n = self.num_vars
        N = 1
        while N < n:
            N *= 2
        if bp_gens.capacity < N:
            raise InvalidGeneratorsLength()
        t.append_point(b"A_I2", proof["A_I2"])
        t.append_point(b"A_O2", proof["A_O2"])
        t.append_point(b"S2", proof["S2"])
        y = t.challenge_scalar(b"y")
        z = t.challenge_scalar(b"z")
        for j in (1, 3, 4, 5, 6):
            t.validate_and_append_point(b"T_%d" % j, proof["T_%d" % j])
        u = t.challenge_scalar(b"u")
        x = t.challenge_scalar(b"x")
        t.append_scalar(b"t_x", proof["t_x"])
        t.append_scalar(b"t_x_blinding", proof["t_x_blinding"])
        t.append_scalar(b"e_blinding", proof["e_blinding"])
        w = t.challenge_scalar(b"w")
        # flatten
        wL, wR, wO, wV, wc = [0] * n, [0] * n, [0] * n, [0] * m, 0
        ez = z
        for lc in self.constraints:
            for (k, i), c in lc.terms:
                if k == MUL_LEFT:
                    wL[i] = (wL[i] + ez * c) % L
                elif k == MUL_RIGHT:
                    wR[i] = (wR[i] + ez * c) % L
                elif k == MUL_OUT:
                    wO[i] = (wO[i] + ez * c) % L
                elif k == COMMITTED:
                    wV[i] = (wV[i] - ez * c) % L
                else:
                    wc = (wc - ez * c) % L
            ez = ez * z % L
        # ipp verification scalars
        lg = N.bit_length() - 1
        if len(proof["L"]) != lg or len(proof["R"]) != lg:
            raise VerificationError("ipp length")
        t.append_message(b"dom-sep", b"ipp v1")
        t.append_u64(b"n", N)
        ch = []
        for Lc, Rc in zip(proof["L"], proof["R"]):
            t.validate_and_append_point(b"L", Lc)
            t.validate_and_append_point(b"R", Rc)
            ch.append(t.challenge_scalar(b"u"))
        ch_inv = [inv_mod_l(c) for c in ch]
        allinv = 1
        for c in ch_inv:
            allinv = allinv * c % L
        ch_sq = [c * c % L for c in ch]
        ch_inv_sq = [c * c % L for c in ch_inv]
        s = [allinv]
        for i in range(1, N):
            lg_i = i.bit_length() - 1
            k = 1 << lg_i
            s.append(s[i - k] * ch_sq[(lg - 1) - lg_i] % L)
        a, b = proof["a"], proof["b"]
        y_inv = inv_mod_l(y)
        yinv = [1] * N
        for i in range(1, N):
            yinv[i] = yinv[i - 1] * y_inv % L
        ynwR = [wR[i] * yinv[i] % L for i in range(n)] + [0] * (N - n)
        delta = sum(ynwR[i] * wL[i] for i in range(n)) % L
        uf = [1] * n + [u] * (N - n)
        g_sc = [uf[i] * (x * ynwR[i] - a * s[i]) % L for i in range(N)]
        wLp = wL + [0] * (N - n)
        wOp = wO + [0] * (N - n)
        h_sc = [uf[i] * (yinv[i] * (x * wLp[i] + wOp[i] - b * s[N - 1 - i]) - 1) % L for i in range(N)]
        rng = t.build_rng([], entropy32)
        r = rng.random_scalar()
        xx = x * x % L
        rxx = r * xx % L
        xxx = x * xx % L
        T_sc = [r * x % L, rxx * x % L, rxx * xx % L, rxx * xxx % L, rxx * xx % L * xx % L]
        scalars = [x, xx, xxx, u * x % L, u * xx % L, u * xxx % L] + [wv * rxx % L for wv in wV] + T_sc
        scalars.append((w * (proof["t_x"] - a * b) + r * (xx * (wc + delta) - proof["t_x"])) % L)
        scalars.append((-proof["e_blinding"] - r * proof["t_x_blinding"]) % L)
        scalars += g_sc + h_sc + ch_sq + ch_inv_sq
        pts_c = [proof[k] for k in ("A_I1", "A_O1", "S1", "A_I2", "A_O2", "S2")] + self.V + \
                [proof["T_%d" % j] for j in (1, 3, 4, 5, 6)]
        pts = []
        for c in pts_c:
            p = ristretto_decode(c)
            if p is None:
                raise VerificationError("bad point")
            pts.append(p)
        pts += [pc_gens.B, pc_gens.B_blinding] + bp_gens.G(N) + bp_gens.H(N)
        for c in list(proof["L"]) + list(proof["R"]):
            p = ristretto_decode(c)
            if p is None:
                raise VerificationError("bad point")
            pts.append(p)
        chk = msm(scalars, pts)
        if not pt_eq(chk, IDENT):
            raise VerificationError("mega check failed")
        return True


def proof_to_bytes(proof):
    """Untagged field tuple: 14 elements, then L0,R0,...,a,b."""
    out = b""
    for k in ("A_I1", "A_O1", "S1", "A_I2", "A_O2", "S2", "T_1", "T_3", "T_4", "T_5", "T_6"):
        out += proof[k]
    for k in ("t_x", "t_x_blinding", "e_blinding"):
        out += (proof[k] % L).to_bytes(32, "little")
    for Lc, Rc in zip(proof["L"], proof["R"]):
        out += Lc + Rc
    out += (proof["a"] % L).to_bytes(32, "little") + (proof["b"] % L).to_bytes(32, "little")
    return out


def proof_from_bytes(buf):
    if len(buf) % 32 or len(buf) < 16 * 32:
        raise R1CSError("format")
    el = [buf[i: i + 32] for i in range(0, len(buf), 32)]
    names = ("A_I1", "A_O1", "S1", "A_I2", "A_O2", "S2", "T_1", "T_3", "T_4", "T_5", "T_6")
    proof = {k: el[i] for i, k in enumerate(names)}
    for i, k in enumerate(("t_x", "t_x_blinding", "e_blinding")):
        v = int.from_bytes(el[11 + i], "little")
        if v >= L:
            raise R1CSError("format")
        proof[k] = v
    rest = el[14:]
    k2 = len(rest) - 2
    if k2 % 2:
        raise R1CSError("format")
    proof["L"] = [rest[2 * i] for i in range(k2 // 2)]
    proof["R"] = [rest[2 * i + 1] for i in range(k2 // 2)]
    proof["a"] = int.from_bytes(rest[-2], "little")
    proof["b"] = int.from_bytes(rest[-1], "little")
    return proof
