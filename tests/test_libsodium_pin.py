"""Third-party pin of the arithmetic under the oracles and under the CUDA path.

The reference's curve and scalar arithmetic is curve25519-dalek (un-vendored, reference Cargo.toml:8); no byte produced by it
exists in this image.  What IS here is libsodium (PyNaCl's bundled build): an independent implementation of the same field
F_l (crypto_core_ed25519_scalar_*) and of the same curve (crypto_scalarmult_ed25519_noclamp, crypto_core_ed25519_add) --
without the ristretto255 layer.  A ristretto element is a coset P + E[4] of Edwards points and ristretto_encode is constant on
the coset, so `ristretto-decode -> ed25519-encode -> libsodium -> ed25519-decode -> ristretto-encode` must reproduce the bytes
of our own ristretto scalar multiplication / addition / multiscalar multiplication.  Checked here: the Python oracle, the C
oracle, and (GPU-marked) the device primitives and the product MSM, on random and edge inputs.
"""
import ctypes as C
import random

import numpy as np
import pytest

nb = pytest.importorskip("nacl.bindings")

from helpers import R, CO, L, selftest  # noqa: E402

P = R.P
EDGE_SCALARS = [0, 1, 2, L - 1, L - 2, (L - 1) // 2, (L + 1) // 2, 2 ** 252, 2 ** 252 - 1, 2 ** 128, 2 ** 64 - 1, 2 ** 64, 27742317777372353535851937790883648493]


def sod_bytes(x):
    return (x % L).to_bytes(32, "little")


def ed_encode(p):
    """Edwards point (X:Y:Z:T) -> the 32-byte ed25519 encoding libsodium takes (y, sign of x in bit 255)"""
    X, Y, Z, _ = p
    zi = pow(Z, P - 2, P)
    x, y = X * zi % P, Y * zi % P
    return (y | ((x & 1) << 255)).to_bytes(32, "little")


def ed_decode(b):
    v = int.from_bytes(b, "little")
    sign, y = v >> 255, v & ((1 << 255) - 1)
    u, w = (y * y - 1) % P, (R.D * y * y + 1) % P
    ok, x = R.sqrt_ratio_m1(u, w)
    assert ok, "libsodium returned a point that is not on the curve"
    if (x & 1) != sign:
        x = (-x) % P
    return (x, y, 1, x * y % P)


E4 = [(0, 1, 1, 0), (0, P - 1, 1, 0), (R.SQRT_M1, 0, 1, 0), (P - R.SQRT_M1, 0, 1, 0)]  # the 4-torsion subgroup


def main_subgroup(pt):
    """The representative of the ristretto element pt + E[4] that lies in the prime-order subgroup (libsodium's scalar
    multiplication and addition refuse anything else); the subgroup test is libsodium's own."""
    hits = [q for q in (R.pt_add(pt, t) for t in E4) if nb.crypto_core_ed25519_is_valid_point(ed_encode(q))]
    assert len(hits) == 1, "exactly one of the four coset representatives has prime order"
    return hits[0]


def sodium_point(pt):
    return ed_encode(main_subgroup(pt))


def sodium_mul(s, pt):
    """s * pt through libsodium, as an Edwards point (s != 0 mod l: libsodium refuses to return the identity)"""
    return ed_decode(nb.crypto_scalarmult_ed25519_noclamp(sod_bytes(s), sodium_point(pt)))


def random_points(rnd, count):
    """ristretto points the way the protocol makes them: one-way map of uniform bytes (includes non-prime-order representatives)"""
    return [R.from_uniform_bytes(bytes(rnd.randrange(256) for _ in range(64))) for _ in range(count)]


# ------------------------------------------------------------------------------------------------ F_l
def scalar_cases(rnd, count):
    xs = list(EDGE_SCALARS)
    while len(xs) < count:
        xs.append(rnd.randrange(L))
    return xs


def test_scalar_field_c_oracle_vs_libsodium(oracle_lib):
    """C oracle sc_mul / sc_invert / wide reduction against libsodium on >= 1000 inputs (the Python oracle is Python's own big ints)"""
    lib = oracle_lib.lib()
    rnd = random.Random(20261017)
    xs, ys = scalar_cases(rnd, 1024), scalar_cases(random.Random(7), 1024)
    rnd.shuffle(ys)
    out = np.zeros(32, np.uint8)

    def call(fn, *bufs):
        arrs = [np.frombuffer(b, np.uint8) for b in bufs]
        getattr(lib, fn)(*[a.ctypes.data_as(CO.u8p) for a in arrs], out.ctypes.data_as(CO.u8p))
        return out.tobytes()
    for x, y in zip(xs, ys):
        bx, by = sod_bytes(x), sod_bytes(y)
        assert call("bpo_sc_mul", bx, by) == nb.crypto_core_ed25519_scalar_mul(bx, by) == sod_bytes(x * y)
        if x % L:
            assert call("bpo_sc_invert", bx) == nb.crypto_core_ed25519_scalar_invert(bx) == sod_bytes(pow(x, L - 2, L))
    assert call("bpo_sc_invert", bytes(32)) == bytes(32)  # Scalar::invert(0) = 0; libsodium rejects 0
    wides = [bytes(64), b"\xff" * 64, (L).to_bytes(64, "little"), (L * L).to_bytes(64, "little"), (2 ** 512 - 1).to_bytes(64, "little")]
    wides += [bytes(rnd.randrange(256) for _ in range(64)) for _ in range(1024)]
    for w in wides:
        assert call("bpo_sc_from_wide", w) == nb.crypto_core_ed25519_scalar_reduce(w) == sod_bytes(int.from_bytes(w, "little"))


# ------------------------------------------------------------------------------------------------ the curve
def test_python_oracle_points_vs_libsodium():
    rnd = random.Random(11)
    pts = random_points(rnd, 24) + [R.BASEPOINT]
    scal = [s for s in EDGE_SCALARS if s % L] + [rnd.randrange(1, L) for _ in range(40)]
    n = 0
    for i, s in enumerate(scal):
        pt = pts[i % len(pts)]
        assert R.ristretto_encode(sodium_mul(s, pt)) == R.ristretto_encode(R.pt_mul(s, pt))
        n += 1
    # addition / subtraction of two independent points
    for i in range(24):
        a, b = pts[i], pts[(i + 5) % 24]
        s = ed_decode(nb.crypto_core_ed25519_add(sodium_point(a), sodium_point(b)))
        d = ed_decode(nb.crypto_core_ed25519_sub(sodium_point(a), sodium_point(b)))
        assert R.ristretto_encode(s) == R.ristretto_encode(R.pt_add(a, b))
        assert R.ristretto_encode(d) == R.ristretto_encode(R.pt_add(a, R.pt_neg(b)))
    # the ristretto layer: decode(encode(P)) stays in the coset P + E[4] (equal main-subgroup representatives), the
    # basepoint is ed25519's, and the one-way map only produces points of the even subgroup (l * 4 * P = identity)
    assert sodium_point(R.BASEPOINT) == nb.crypto_scalarmult_ed25519_base_noclamp(sod_bytes(1))
    for pt in pts:
        back = R.ristretto_decode(R.ristretto_encode(pt))
        assert sodium_point(back) == sodium_point(pt)
        assert nb.crypto_core_ed25519_is_valid_point(sodium_point(pt))
    assert n >= 40


def test_c_oracle_points_vs_libsodium(oracle_lib):
    rnd = random.Random(12)
    pts = random_points(rnd, 200)
    scal = [s for s in EDGE_SCALARS if s % L]
    scal += [rnd.randrange(1, L) for _ in range(1000 - len(scal))]
    for i, s in enumerate(scal):
        pt = pts[i % len(pts)]
        rc, got = oracle_lib.call_bytes("bpo_scalarmult", 32, sod_bytes(s), R.ristretto_encode(pt))
        assert rc == 0 and got == R.ristretto_encode(sodium_mul(s, pt)), i


def sodium_msm(scalars, pts):
    acc = None
    for s, pt in zip(scalars, pts):
        if s % L == 0:
            continue
        t = nb.crypto_scalarmult_ed25519_noclamp(sod_bytes(s), sodium_point(pt))
        acc = t if acc is None else nb.crypto_core_ed25519_add(acc, t)
    return R.ristretto_encode(ed_decode(acc))


def test_c_oracle_msm_and_generators_vs_libsodium(oracle_lib):
    """multiscalar multiplication over the first generators of the chain G: C oracle = libsodium's sum"""
    n = 96
    g = np.zeros((n, 32), np.uint8)
    oracle_lib.lib().bpo_gens_compressed(0, n, g.ctypes.data_as(CO.u8p))
    pts = [R.ristretto_decode(bytes(x)) for x in g]
    rnd = random.Random(13)
    scal = [rnd.randrange(L) for _ in range(n)]
    scal[3], scal[4], scal[5] = 0, 1, L - 1
    sb = np.frombuffer(b"".join(sod_bytes(s) for s in scal), np.uint8)
    out = np.zeros(32, np.uint8)
    rc = oracle_lib.lib().bpo_msm(n, sb.ctypes.data_as(CO.u8p), g.ctypes.data_as(CO.u8p), out.ctypes.data_as(CO.u8p))
    assert rc == 0 and out.tobytes() == sodium_msm(scal, pts)


# ------------------------------------------------------------------------------------------------ the kernel bodies (through the C-ABI)
# The same checks run against the host-emulation build of the kernel bodies here (small sizes) and, GPU-marked in
# tests/test_gpu.py, against the product library on the device at full sizes.
@pytest.fixture(scope="module")
def api(emul_so):
    from bulletproofs_r1cs_gadgets_b200 import api as a
    a._lib = None
    a.load(emul_so)
    yield a
    a._lib = None


def check_device_scalar_field(api, count):
    """bp_selftest_device: sc_mul (4), sc_invert (1), wide reduction (2) against libsodium"""
    rnd = random.Random(21)
    xs, ys = scalar_cases(rnd, count), scalar_cases(random.Random(22), count)
    rnd.shuffle(ys)
    for x, y in zip(xs, ys):
        bx, by = sod_bytes(x), sod_bytes(y)
        assert selftest(api, 4, bx + by, 32) == nb.crypto_core_ed25519_scalar_mul(bx, by)
        if x % L:
            assert selftest(api, 1, bx, 32) == nb.crypto_core_ed25519_scalar_invert(bx)
    assert selftest(api, 1, bytes(32), 32) == bytes(32)
    wides = [bytes(64), b"\xff" * 64, (L).to_bytes(64, "little"), (L * L).to_bytes(64, "little")]
    wides += [bytes(rnd.randrange(256) for _ in range(64)) for _ in range(count)]
    for w in wides:
        assert selftest(api, 2, w, 32) == nb.crypto_core_ed25519_scalar_reduce(w)


def check_device_msm(api, gens, n, call):
    """bp_msm_gens_device = libsodium's sum of scalar multiples of the generators the library exports (which also pins the
    generator chain's points as curve points libsodium accepts)"""
    g = gens.export(0, n)
    pts = [R.ristretto_decode(bytes(x)) for x in g]
    assert all(p is not None for p in pts)
    rnd = random.Random(23 + n)
    scal = [rnd.randrange(L) for _ in range(n)]
    scal[0], scal[1], scal[2] = 0, 1, L - 1
    rc, got = call(api, gens, api.scalars_to_array(scal))
    assert rc == 0 and got == sodium_msm(scal, pts)


def check_device_from_uniform_and_commit(api, gens, count):
    """one-way map (selftest 5) against the pinned Python oracle; Pedersen commitments v*B + r*B_blinding against libsodium"""
    rnd = random.Random(24)
    for _ in range(count):
        u = bytes(rnd.randrange(256) for _ in range(64))
        assert selftest(api, 5, u, 32) == R.ristretto_encode(R.from_uniform_bytes(u))
    pc = R.PedersenGens()
    for _ in range(count):
        v, r = rnd.randrange(1, L), rnd.randrange(1, L)
        assert gens.commit(v, r) == sodium_msm([v, r], [pc.B, pc.B_blinding])


def _msm_host_pointers(api, gens, arr):
    out = np.zeros(32, np.uint8)
    rc = api.load().bp_msm_gens_device(gens._h, arr.shape[0], arr.ctypes.data_as(api.u8p), out.ctypes.data_as(api.u8p), None)
    return rc, out.tobytes()


def test_kernel_bodies_scalar_field_vs_libsodium(api):
    check_device_scalar_field(api, 200)


def test_kernel_bodies_msm_and_commit_vs_libsodium(api):
    gens = api.Gens(256)
    check_device_msm(api, gens, 120, _msm_host_pointers)
    check_device_from_uniform_and_commit(api, gens, 6)
