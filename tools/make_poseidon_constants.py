#!/usr/bin/env python3
"""Derive the Poseidon constants blob from the reference's constants file.

Runs only in the build container (needs /root/reference).  The reference stores 36 MDS
entries and 960 round constants as big-endian hex strings (src/poseidon_constants.rs:1-10)
but loads them with Scalar::from_bytes_mod_order, i.e. little-endian, reduced mod l
(src/scalar_utils.rs:232-237).  The blob holds the values AS USED: 996 canonical 32-byte
little-endian scalars, MDS row-major first, then the round constants.
"""
import hashlib, re, sys
L = 2**252 + 27742317777372353535851937790883648493
src = open("/root/reference/src/poseidon_constants.rs", "rb").read()
assert hashlib.sha256(src).hexdigest() == "c9d320eb8b41e39f4e35badd0f03debfddbc44b5ffa368fc04e501295a01af5e"
hexes = re.findall(rb'"0x([0-9a-fA-F]{64})"', src)
assert len(hexes) == 36 + 960, len(hexes)
out = bytearray()
for h in hexes:
    v = int.from_bytes(bytes.fromhex(h.decode()), "little") % L
    out += v.to_bytes(32, "little")
path = sys.argv[1] if len(sys.argv) > 1 else "bulletproofs_r1cs_gadgets_b200/data/poseidon_constants.bin"
open(path, "wb").write(out)
print(path, len(out), hashlib.sha256(out).hexdigest())
