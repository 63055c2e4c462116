"""Poseidon hashing and Merkle trees: mirror of plonky2/src/hash/{poseidon,hashing,merkle_tree,
merkle_proofs}.rs for F = GoldilocksField, executed by the CUDA engine."""
import numpy as np

from . import _ffi
from .polynomial import GOLDILOCKS_ORDER, log2_strict

SPONGE_RATE = 8      # poseidon.rs:22
SPONGE_CAPACITY = 4  # poseidon.rs:23
SPONGE_WIDTH = 12    # poseidon.rs:24
NUM_HASH_OUT_ELTS = 4


class HashOut:
    """hash_types.rs:22-24: four field elements."""

    def __init__(self, elements):
        self.elements = _ffi.as_u64(elements).reshape(4)

    def __eq__(self, other):
        return np.array_equal(self.elements, other.elements)

    def to_bytes(self):
        """hash_types.rs:83-101: canonical little-endian u64s."""
        return self.elements.astype("<u8").tobytes()

    def __repr__(self):
        return "HashOut(" + ", ".join(hex(int(x)) for x in self.elements) + ")"


class PoseidonPermutation:
    """poseidon.rs:637-702 / hashing.rs:63-95 PlonkyPermutation: sponge state container."""

    RATE = SPONGE_RATE
    WIDTH = SPONGE_WIDTH

    def __init__(self, elts=()):
        self.state = np.zeros(SPONGE_WIDTH, dtype=np.uint64)
        self.set_from_iter(elts, 0)

    def set_elt(self, elt, idx):
        self.state[idx] = elt

    def set_from_slice(self, elts, start_idx):
        elts = _ffi.as_u64(elts)
        self.state[start_idx : start_idx + elts.shape[0]] = elts

    def set_from_iter(self, elts, start_idx):
        for i, e in zip(range(start_idx, SPONGE_WIDTH), elts):
            self.state[i] = e

    def permute(self):
        s = self.state.reshape(1, 12).copy()
        _ffi.check(_ffi.lib().pcs_poseidon_permute(_ffi.ptr(s), 1))
        self.state = s[0]

    def squeeze(self):
        return self.state[: self.RATE]


def poseidon(states):
    """Poseidon::poseidon (poseidon.rs:599) on a batch [n][12]."""
    s = _ffi.as_u64(states, copy=True).reshape(-1, 12)
    _ffi.check(_ffi.lib().pcs_poseidon_permute(_ffi.ptr(s), s.shape[0]))
    return s


class PoseidonHash:
    """poseidon.rs:706-719 `impl Hasher<F> for PoseidonHash`."""

    HASH_SIZE = 4 * 8
    Permutation = PoseidonPermutation

    @staticmethod
    def hash_or_noop_batch(rows):
        r = _ffi.as_u64(rows)
        assert r.ndim == 2
        out = np.empty((r.shape[0], 4), dtype=np.uint64)
        _ffi.check(_ffi.lib().pcs_hash_or_noop(_ffi.ptr(r), r.shape[0], r.shape[1], _ffi.ptr(out)))
        return out

    @staticmethod
    def hash_or_noop(inputs):
        """config.rs:55-66"""
        return HashOut(PoseidonHash.hash_or_noop_batch(_ffi.as_u64(inputs).reshape(1, -1))[0])

    @staticmethod
    def hash_no_pad(inputs):
        """hashing.rs:144 hash_n_to_hash_no_pad.  Inputs of <= 4 elements ARE hashed here (only
        hash_or_noop skips them), so run the sponge explicitly through the permutation kernel."""
        x = _ffi.as_u64(inputs).reshape(-1)
        if x.shape[0] > 4:
            return PoseidonHash.hash_or_noop(x)
        perm = PoseidonPermutation()
        if x.shape[0]:
            perm.set_from_slice(x, 0)
            perm.permute()
        return HashOut(perm.squeeze()[:4] % np.uint64(GOLDILOCKS_ORDER))

    @staticmethod
    def two_to_one_batch(left, right):
        l, r = _ffi.as_u64(left).reshape(-1, 4), _ffi.as_u64(right).reshape(-1, 4)
        out = np.empty_like(l)
        _ffi.check(_ffi.lib().pcs_two_to_one(_ffi.ptr(l), _ffi.ptr(r), l.shape[0], _ffi.ptr(out)))
        return out

    @staticmethod
    def two_to_one(left, right):
        """poseidon.rs:716-718 -> compress (hashing.rs:98-115)"""
        return HashOut(PoseidonHash.two_to_one_batch(left.elements, right.elements)[0])


class MerkleCap:
    """merkle_tree.rs:18: `MerkleCap(pub Vec<H::Hash>)` as a [2^h][4] array."""

    def __init__(self, hashes):
        self.hashes = _ffi.as_u64(hashes).reshape(-1, 4)

    def __len__(self):
        return self.hashes.shape[0]

    def height(self):
        return log2_strict(len(self))

    def flatten(self):
        return self.hashes.reshape(-1)

    def __eq__(self, other):
        return np.array_equal(self.hashes, other.hashes)


class MerkleProof:
    """merkle_proofs.rs:17-21: siblings bottom-up."""

    def __init__(self, siblings):
        self.siblings = _ffi.as_u64(siblings).reshape(-1, 4)

    def __len__(self):
        return self.siblings.shape[0]


def verify_merkle_proof_to_cap(leaf_data, leaf_index, merkle_cap, proof):
    """merkle_proofs.rs:54-77.  Raises ValueError("Invalid Merkle proof.") like the reference's ensure!."""
    index = int(leaf_index)
    current = PoseidonHash.hash_or_noop(leaf_data)
    for sib in proof.siblings:
        bit = index & 1
        index >>= 1
        s = HashOut(sib)
        current = PoseidonHash.two_to_one(s, current) if bit else PoseidonHash.two_to_one(current, s)
    if not np.array_equal(current.elements, merkle_cap.hashes[index]):
        raise ValueError("Invalid Merkle proof.")


class MerkleTree:
    """merkle_tree.rs:39-55.  Fields `leaves`, `digests`, `cap` are public like the reference's."""

    def __init__(self, leaves, digests, cap):
        self.leaves = leaves
        self.digests = digests
        self.cap = cap

    @classmethod
    def new(cls, leaves, cap_height):
        """merkle_tree.rs:135-166; leaves: [n][len] uint64 rows."""
        lv = _ffi.as_u64(leaves)
        if lv.ndim != 2:
            raise ValueError("leaves must be a 2-D array of equal-length rows")
        n = lv.shape[0]
        log2_leaves_len = log2_strict(n)
        if cap_height > log2_leaves_len:
            raise ValueError(f"cap_height={cap_height} should be at most log2(leaves.len())={log2_leaves_len}")
        digests = np.empty((2 * (n - (1 << cap_height)), 4), dtype=np.uint64)
        cap = np.empty((1 << cap_height, 4), dtype=np.uint64)
        _ffi.check(_ffi.lib().pcs_merkle_build(_ffi.ptr(lv), n, lv.shape[1], cap_height, _ffi.ptr(digests), _ffi.ptr(cap)))
        return cls(lv, digests, MerkleCap(cap))

    def get(self, i):
        """:168"""
        return self.leaves[i]

    def prove(self, leaf_index):
        """:173-207 (host index arithmetic over the reference digest layout)."""
        cap_height = log2_strict(len(self.cap))
        num_layers = log2_strict(self.leaves.shape[0]) - cap_height
        assert leaf_index >> (cap_height + num_layers) == 0
        tree_index = leaf_index >> num_layers
        tree_len = self.digests.shape[0] >> cap_height
        digest_tree = self.digests[tree_len * tree_index : tree_len * (tree_index + 1)]
        pair_index = leaf_index & ((1 << num_layers) - 1)
        siblings = []
        for i in range(num_layers):
            parity = pair_index & 1
            pair_index >>= 1
            siblings_index = (pair_index << (i + 1)) + (1 << i) - 1
            siblings.append(digest_tree[2 * siblings_index + (1 - parity)])
        return MerkleProof(np.array(siblings, dtype=np.uint64).reshape(-1, 4))
