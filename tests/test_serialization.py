"""On-wire formats (plonky2_demo_b200/serialization.py) against byte strings assembled by hand from the reference's
writer definitions (plonky2/src/util/serialization/mod.rs), plus read/write round trips.  Host-only: no GPU."""
import struct

import numpy as np
import pytest

from helpers import P, splitmix64_stream
from plonky2_demo_b200.fri_prover import FriInitialTreeProof, FriProof, FriQueryRound, FriQueryStep
from plonky2_demo_b200.hashing import MerkleCap, MerkleProof, MerkleTree
from plonky2_demo_b200.polynomial import PolynomialCoeffs
from plonky2_demo_b200.serialization import Buffer, HostPolynomialBatch, fri_proof_to_bytes, polynomial_batch_to_bytes


def u64(*xs):
    return b"".join(struct.pack("<Q", x) for x in xs)


def test_primitives_match_the_reference_writers():
    b = Buffer()
    b.write_bool(True)                    # write_u8(u8::from(x))                       :1179-1181
    b.write_usize(0x0102030405060708)     # (x as u64).to_le_bytes()                    :1220-1222
    b.write_field(P + 5)                  # to_canonical_u64().to_le_bytes()            :1237-1242
    b.write_field_vec([1, P - 1, P])      # no length prefix                            :1246-1254
    b.write_field_ext_vec([[2, 3], [P + 1, 4]])
    b.write_hash([9, 8, 7, P + 6])        # HashOut::to_bytes: 4 canonical LE u64       hash_types.rs:83-101
    b.write_hash_vec([[1, 2, 3, 4]])      # length, then hashes                         :1352-1363
    want = (b"\x01" + bytes([8, 7, 6, 5, 4, 3, 2, 1]) + u64(5) + u64(1, P - 1, 0) + u64(2, 3, 1, 4) + u64(9, 8, 7, 6)
            + u64(1) + u64(1, 2, 3, 4))
    assert b.bytes() == want


def test_merkle_tree_and_batch_layout():
    leaves = np.array([[1, 2, 3], [4, 5, P + 6]], dtype=np.uint64)
    digests = np.array([[10, 11, 12, 13], [14, 15, 16, 17]], dtype=np.uint64)
    cap = MerkleCap(np.array([[20, 21, 22, 23]], dtype=np.uint64))
    tree = MerkleTree(leaves, digests, cap)
    b = Buffer()
    b.write_merkle_tree(tree)
    # leaves.len(); per leaf: len + elements; digests: len + hashes; cap.height(); cap hashes      :1390-1405
    want_tree = (u64(2) + u64(3, 1, 2, 3) + u64(3, 4, 5, 6) + u64(2) + u64(10, 11, 12, 13, 14, 15, 16, 17) + u64(0)
                 + u64(20, 21, 22, 23))
    assert b.bytes() == want_tree
    t2 = Buffer(want_tree).read_merkle_tree()
    assert np.array_equal(t2.leaves, leaves % np.uint64(P)) and np.array_equal(t2.digests, digests) and t2.cap == cap

    batch = HostPolynomialBatch([PolynomialCoeffs([7, 8]), PolynomialCoeffs([9, P + 1])], tree, 1, 0, False)
    # polynomials.len(); per polynomial: len + coeffs; merkle tree; degree_log; rate_bits; blinding  :1715-1734
    want = u64(2) + u64(2, 7, 8) + u64(2, 9, 1) + want_tree + u64(1) + u64(0) + b"\x00"
    assert polynomial_batch_to_bytes(batch) == want
    r = Buffer(want).read_polynomial_batch()
    assert [p.coeffs.tolist() for p in r.polynomials] == [[7, 8], [9, 1]]
    assert (r.degree_log, r.rate_bits, r.blinding) == (1, 0, False)
    assert np.array_equal(r.merkle_tree.leaves, t2.leaves)
    with pytest.raises(EOFError):
        Buffer(want[:-3]).read_polynomial_batch()


def test_fri_proof_layout_and_roundtrip():
    rng = np.random.default_rng(3)

    def h(n):
        return splitmix64_stream(int(rng.integers(1 << 30)), 4 * n).reshape(n, 4)

    caps = [MerkleCap(h(2)), MerkleCap(h(2))]
    rounds = []
    for _ in range(3):
        initial = [(splitmix64_stream(int(rng.integers(1 << 30)), ln), MerkleProof(h(5))) for ln in (4, 7)]
        steps = [FriQueryStep(splitmix64_stream(int(rng.integers(1 << 30)), 2 << ar).reshape(-1, 2), MerkleProof(h(3 - i)))
                 for i, ar in enumerate((2, 1))]
        rounds.append(FriQueryRound(FriInitialTreeProof(initial), steps))
    final_poly = splitmix64_stream(99, 8).reshape(4, 2)
    proof = FriProof(caps, rounds, final_poly, 12345)
    data = fri_proof_to_bytes(proof)
    # caps (no lengths) | per round: per oracle evals (no length) + u8 path length + path, per step evals + path |
    # final_poly (no length) | pow_witness                                                        :1568-1582, :1532-1548
    per_round = (4 + 7) * 8 + 2 * (1 + 5 * 32) + (8 * 8 + 1 + 3 * 32) + (4 * 8 + 1 + 2 * 32)
    assert len(data) == 2 * 2 * 32 + 3 * per_round + 4 * 16 + 8
    assert data[:32] == caps[0].hashes[0].astype("<u8").tobytes()
    off = 2 * 2 * 32
    assert data[off:off + 32] == rounds[0].initial_trees_proof.evals_proofs[0][0].astype("<u8").tobytes()
    assert data[off + 32] == 5                                    # Merkle path length as ONE byte        :1443-1457
    assert data[-8:] == struct.pack("<Q", 12345)
    back = Buffer(data).read_fri_proof([4, 7], [2, 1], 1, 3, 4)
    assert back.pow_witness == 12345 and np.array_equal(back.final_poly, final_poly)
    for c, c2 in zip(caps, back.commit_phase_merkle_caps):
        assert c == c2
    for r, r2 in zip(rounds, back.query_round_proofs):
        for (v, p), (v2, p2) in zip(r.initial_trees_proof.evals_proofs, r2.initial_trees_proof.evals_proofs):
            assert np.array_equal(v, v2) and np.array_equal(p.siblings, p2.siblings)
        for s, s2 in zip(r.steps, r2.steps):
            assert np.array_equal(s.evals, s2.evals) and np.array_equal(s.merkle_proof.siblings, s2.merkle_proof.siblings)
    assert fri_proof_to_bytes(back) == data
    with pytest.raises(OverflowError):
        Buffer().write_merkle_proof(MerkleProof(np.zeros((256, 4), dtype=np.uint64)))
