"""Host-side sparse Merkle trees of the reference: the witness SOURCE of the membership circuits (the step before the
hot path; the benchmark draws synthetic paths instead).  Hashing is the library's native Poseidon
(bp_poseidon_hash_2 / bp_poseidon_hash_4), always with the inverse S-box as in the reference.

VanillaSparseMerkleTree mirrors reference src/gadget_vsmt_2.rs:27-166 (HashMap-backed binary tree: `new` precomputes the
empty-subtree hashes, `update`, `get` with the sibling path root -> leaf, `verify_proof`); `depth` is a parameter here
(the reference hard-codes TreeDepth = 253)."""
import ctypes as C

import numpy as np

from . import api


class VanillaSparseMerkleTree:
    def __init__(self, hash_params, depth=253):  # gadget_vsmt_2.rs:36-61
        self.depth, self.hash_params = depth, hash_params
        self.db = {}
        self.empty_tree_hashes = [0]
        for i in range(1, depth + 1):
            prev = self.empty_tree_hashes[i - 1]
            new = self._h(prev, prev)
            self.db[new] = (prev, prev)
            self.empty_tree_hashes.append(new)
        self.root = self.empty_tree_hashes[depth]

    def _h(self, left, right):
        return self.hash_params.hash_2(left, right, api.SBOX_INVERSE)

    def _bits(self, idx):
        return [(idx >> i) & 1 for i in range(self.depth)]  # ScalarBits::from_scalar(&idx, depth), LSB first

    def update(self, idx, val):  # gadget_vsmt_2.rs:63-98
        sidenodes = []
        self.get(idx, sidenodes)
        cur_val = val % api.L
        for bit in self._bits(idx):
            side = sidenodes.pop()
            if bit:
                h = self._h(side, cur_val); self.db[h] = (side, cur_val)
            else:
                h = self._h(cur_val, side); self.db[h] = (cur_val, side)
            cur_val = h
        self.root = cur_val
        return cur_val

    def get(self, idx, proof=None):  # gadget_vsmt_2.rs:101-131; proof (a list) receives the siblings root -> leaf
        cur = self.root
        for bit in reversed(self._bits(idx)):  # most significant bit first
            left, right = self.db[cur]
            if bit:
                cur = right
                if proof is not None: proof.append(left)
            else:
                cur = left
                if proof is not None: proof.append(right)
        return cur

    def verify_proof(self, idx, val, proof, root=None):  # gadget_vsmt_2.rs:134-161
        cur = val % api.L
        for i, bit in enumerate(self._bits(idx)):
            side = proof[self.depth - 1 - i]
            cur = self._h(side, cur) if bit else self._h(cur, side)
        return cur == (self.root if root is None else root)


class DeviceVsmt2:
    """Batched VanillaSparseMerkleTree on the GPU (bp_vsmt2_* of include/bp_b200.h, kernels in csrc/tree_kernels.h):
    `update_batch` = many `update`s in one level-by-level pass, `get_batch` = many `get`s, `witness_rows` = the committed
    values of the membership circuit for a batch of leaves (reference src/gadget_vsmt_2.rs:63-131,296-330).  Indices are
    integers below 2**depth, depth <= 253 (the reference's TreeDepth; 256-bit keys through the *_wide entry points).  Same
    method names and orientation as the host mirror above."""

    def __init__(self, hash_params, depth=32, sbox=api.SBOX_INVERSE):
        self.depth, self.hash_params = depth, hash_params
        self._h = C.c_void_p()
        api._check(api.load().bp_vsmt2_new(hash_params._h, C.c_uint32(depth), C.c_int32(sbox), C.byref(self._h)), "vsmt2_new")

    def __del__(self):
        if api is not None and api._lib is not None and getattr(self, "_h", None):
            api._lib.bp_vsmt2_free.restype = None
            api._lib.bp_vsmt2_free(self._h)
            self._h = None

    @staticmethod
    def _idx(idx):
        """integers -> uint8 [count][32], little-endian 256-bit indices"""
        idx = [int(i) for i in (idx.tolist() if isinstance(idx, np.ndarray) else idx)]
        if not idx:
            return np.zeros((0, 32), dtype=np.uint8)
        return np.frombuffer(b"".join(i.to_bytes(32, "little") for i in idx), dtype=np.uint8).reshape(-1, 32).copy()

    @property
    def root(self):
        out = (C.c_uint8 * 32)()
        api._check(api.load().bp_vsmt2_root(self._h, out), "vsmt2_root")
        return int.from_bytes(bytes(out), "little")

    @property
    def num_nodes(self):
        f = api.load().bp_vsmt2_num_nodes
        f.restype = C.c_uint64
        return int(f(self._h))

    @property
    def empty_tree_hashes(self):
        out = (C.c_uint8 * (32 * (self.depth + 1)))()
        api._check(api.load().bp_vsmt2_empty_hashes(self._h, out), "vsmt2_empty_hashes")
        b = bytes(out)
        return [int.from_bytes(b[32 * i:32 * i + 32], "little") for i in range(self.depth + 1)]

    def update_batch(self, idx, vals):
        """idx: integers; vals: integers or uint8 [count][32].  Returns the new root."""
        idx = self._idx(idx)
        if not isinstance(vals, np.ndarray):
            vals = api.scalars_to_array(vals)
        vals = api._np_u8(vals).reshape(-1, 32)
        assert len(vals) == len(idx)
        root = (C.c_uint8 * 32)()
        api._check(api.load().bp_vsmt2_update_batch_wide(self._h, C.c_uint32(len(idx)), idx.ctypes.data_as(api.u8p),
                                                         vals.ctypes.data_as(api.u8p), root), "vsmt2_update_batch_wide")
        return int.from_bytes(bytes(root), "little")

    def update(self, idx, val):
        return self.update_batch([idx], [val])

    def get_batch(self, idx):
        """returns (leaves uint8 [count][32], proofs uint8 [count][depth][32] with the siblings root -> leaf)"""
        idx = self._idx(idx)
        leaves = np.zeros((len(idx), 32), dtype=np.uint8)
        proofs = np.zeros((len(idx), self.depth, 32), dtype=np.uint8)
        api._check(api.load().bp_vsmt2_get_batch_wide(self._h, C.c_uint32(len(idx)), idx.ctypes.data_as(api.u8p),
                                                      leaves.ctypes.data_as(api.u8p), proofs.ctypes.data_as(api.u8p)), "vsmt2_get_batch_wide")
        return leaves, proofs

    def get(self, idx, proof=None):
        leaves, proofs = self.get_batch([idx])
        if proof is not None:
            proof.extend(int.from_bytes(proofs[0, i].tobytes(), "little") for i in range(self.depth))
        return int.from_bytes(leaves[0].tobytes(), "little")

    def witness_rows(self, idx):
        """(v uint8 [count][2*depth+5][32], pub uint8 [count][1][32]): inputs of Circuit.prove_batch for workloads.Vsmt2's circuit"""
        idx = self._idx(idx)
        v = np.zeros((len(idx), 2 * self.depth + 5, 32), dtype=np.uint8)
        pub = np.zeros((len(idx), 1, 32), dtype=np.uint8)
        api._check(api.load().bp_vsmt2_witness_batch_wide(self._h, C.c_uint32(len(idx)), idx.ctypes.data_as(api.u8p),
                                                          v.ctypes.data_as(api.u8p), pub.ctypes.data_as(api.u8p)), "vsmt2_witness_batch_wide")
        return v, pub

    def witness_rows_device(self, d_idx, d_v, d_pub=None, stream=None):
        """device-resident variant: d_idx torch int64/uint64 [count] (depth <= 63) or uint8 [count][32] (256-bit little-endian indices),
        d_v torch uint8 [count][2*depth+5][32], d_pub uint8 [count][1][32]"""
        wide = d_idx.dim() == 2
        count = int(d_idx.shape[0])
        fn = api.load().bp_vsmt2_witness_batch_wide_device if wide else api.load().bp_vsmt2_witness_batch_device
        api._check(fn(self._h, C.c_uint32(count), C.c_void_p(d_idx.data_ptr()), C.c_void_p(d_v.data_ptr()),
                      C.c_void_p(d_pub.data_ptr()) if d_pub is not None else None, C.c_void_p(stream) if stream else None), "vsmt2_witness_batch_device")


def poseidon_hash_2_batch(hash_params, xl, xr, sbox=api.SBOX_INVERSE):
    """Poseidon_hash_2 of many pairs on the device; xl, xr: integers or uint8 [count][32]; returns uint8 [count][32]"""
    xl = api._np_u8(xl if isinstance(xl, np.ndarray) else api.scalars_to_array(xl)).reshape(-1, 32)
    xr = api._np_u8(xr if isinstance(xr, np.ndarray) else api.scalars_to_array(xr)).reshape(-1, 32)
    out = np.zeros_like(xl)
    api._check(api.load().bp_poseidon_hash_2_batch(hash_params._h, C.c_int32(sbox), C.c_uint32(len(xl)), xl.ctypes.data_as(api.u8p),
                                                   xr.ctypes.data_as(api.u8p), out.ctypes.data_as(api.u8p)), "poseidon_hash_2_batch")
    return out
