"""ORACLE (test infrastructure, never on the product path): restatement of the reference's VanillaSparseMerkleTree
(/root/reference/src/gadget_vsmt_2.rs:27-166) -- a content-addressed map hash -> (left, right), one key at a time -- used to
check the device-side batched tree (bulletproofs_r1cs_gadgets_b200/csrc/tree.cu).  `hash2(left, right) -> int` is the
oracle's own Poseidon_hash_2: gadgets_pyref.poseidon_hash_2 (big-int Python) or c_oracle's bpo_poseidon_hash2.
Parity unpinned against the Rust bytes for the same reason as the rest of oracle/ (no Rust toolchain, no golden vectors
upstream: the reference's tree test only checks get-after-update and verify_proof, src/gadget_vsmt_2.rs:223-259)."""

L = 2 ** 252 + 27742317777372353535851937790883648493


class VanillaSparseMerkleTree:
    def __init__(self, hash2, depth=253):  # gadget_vsmt_2.rs:36-61
        self.depth, self.hash2 = depth, hash2
        self.db = {}
        self.empty_tree_hashes = [0]
        for i in range(1, depth + 1):  # :41-50
            prev = self.empty_tree_hashes[-1]
            new = hash2(prev, prev)
            self.db[new] = (prev, prev)
            self.empty_tree_hashes.append(new)
        self.root = self.empty_tree_hashes[depth]

    def update(self, idx, val):  # :63-98
        sidenodes = []
        self.get(idx, sidenodes)
        cur_idx, cur_val = idx, val % L
        for _ in range(self.depth):
            side_elem = sidenodes.pop()
            if cur_idx & 1:  # LSB set: new value on the right (:75-80)
                pair = (side_elem, cur_val)
            else:
                pair = (cur_val, side_elem)
            h = self.hash2(*pair)
            self.db[h] = pair
            cur_idx >>= 1
            cur_val = h
        self.root = cur_val
        return cur_val

    def get(self, idx, proof=None):  # :101-131, siblings pushed root -> leaf
        cur_node = self.root
        for i in range(self.depth):
            left, right = self.db[cur_node]
            if (idx >> (self.depth - 1 - i)) & 1:  # MSB first (:110)
                cur_node, sib = right, left
            else:
                cur_node, sib = left, right
            if proof is not None:
                proof.append(sib)
        return cur_node

    def verify_proof(self, idx, val, proof, root=None):  # :134-161
        cur_idx, cur_val = idx, val % L
        for i in range(self.depth):
            side = proof[self.depth - 1 - i]
            cur_val = self.hash2(side, cur_val) if cur_idx & 1 else self.hash2(cur_val, side)
            cur_idx >>= 1
        return cur_val == (self.root if root is None else root)


def c_oracle_hash2(c_oracle, inverse=1):
    """hash2 backed by oracle/bp_oracle.c (bpo_poseidon_hash2; call c_oracle.poseidon_set_params first)"""
    import ctypes as C
    import numpy as np
    lib = c_oracle.lib()

    def h(left, right):
        out = np.zeros(32, dtype=np.uint8)
        a = (C.c_uint8 * 32).from_buffer_copy((left % L).to_bytes(32, "little"))
        b = (C.c_uint8 * 32).from_buffer_copy((right % L).to_bytes(32, "little"))
        lib.bpo_poseidon_hash2(a, b, C.c_int(inverse), out.ctypes.data_as(C.POINTER(C.c_uint8)))
        return int.from_bytes(out.tobytes(), "little")
    return h
