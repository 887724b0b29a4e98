"""Host-side sparse Merkle trees of the reference: the witness SOURCE of the membership circuits (the step before the
hot path; the benchmark draws synthetic paths instead).  Hashing is the library's native Poseidon
(bp_poseidon_hash_2 / bp_poseidon_hash_4), always with the inverse S-box as in the reference.

VanillaSparseMerkleTree mirrors reference src/gadget_vsmt_2.rs:27-166 (HashMap-backed binary tree: `new` precomputes the
empty-subtree hashes, `update`, `get` with the sibling path root -> leaf, `verify_proof`); `depth` is a parameter here
(the reference hard-codes TreeDepth = 253)."""
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
