"""On-wire formats of the path's data structures: mirror of the `Write` / `Read` methods of
plonky2/src/util/serialization/mod.rs for PolynomialBatch, MerkleTree, MerkleCap, MerkleProof and FriProof
(SURVEY 8f N4), so that prover data built by this engine can be loaded by the stock CPU prover and vice versa.

Format facts (serialization/mod.rs): every integer is little-endian; `usize` is written as u64 (:1220-1222); a field
element is its canonical u64 (:1237-1242); an extension element is its D base elements (:1258-1266); a hash is its 4
field elements (`HashOut::to_bytes`, hash_types.rs:83-101); vectors carry NO length unless the writer adds one.

Host-side byte shuffling only (numpy): nothing here computes field arithmetic.
"""
import io
import struct

import numpy as np

from .fri_prover import FriInitialTreeProof, FriProof, FriQueryRound, FriQueryStep
from .hashing import MerkleCap, MerkleProof, MerkleTree
from .polynomial import GOLDILOCKS_ORDER, PolynomialCoeffs

_P = np.uint64(GOLDILOCKS_ORDER)


def _canon(a):
    a = np.ascontiguousarray(np.asarray(a, dtype=np.uint64))
    return np.where(a >= _P, a - _P, a).astype("<u8")


class Buffer:
    """serialization/mod.rs `Buffer` + `impl Write for Vec<u8>`: a growable byte sink / positioned reader."""

    def __init__(self, data=b""):
        self._w = io.BytesIO()
        self._r = memoryview(bytes(data))
        self._pos = 0

    # ---- Write -----------------------------------------------------------------------------------------------------
    def write_all(self, b):
        self._w.write(b)

    def bytes(self):
        return self._w.getvalue()

    def write_bool(self, x):                       # :1179-1181
        self.write_u8(1 if x else 0)

    def write_u8(self, x):                         # :1202-1204
        self.write_all(struct.pack("<B", x))

    def write_u32(self, x):                        # :1214-1216
        self.write_all(struct.pack("<I", x))

    def write_usize(self, x):                      # :1220-1222
        self.write_all(struct.pack("<Q", x))

    def write_field(self, x):                      # :1237-1242
        self.write_all(struct.pack("<Q", int(x) % GOLDILOCKS_ORDER))

    def write_field_vec(self, v):                  # :1246-1254 (no length)
        self.write_all(_canon(v).tobytes())

    def write_field_ext_vec(self, v):              # :1270-1278: [n][2] -> a0 b0 a1 b1 ...
        self.write_all(_canon(np.asarray(v, dtype=np.uint64).reshape(-1, 2)).tobytes())

    def write_hash(self, h):                       # :1332-1338
        self.write_all(_canon(np.asarray(getattr(h, "elements", h), dtype=np.uint64)).reshape(4).tobytes())

    def write_hash_vec(self, v):                   # :1352-1363 (with length)
        v = np.asarray(v, dtype=np.uint64).reshape(-1, 4)
        self.write_usize(v.shape[0])
        self.write_all(_canon(v).tobytes())

    def write_merkle_cap(self, cap):               # :1367-1376 (no length: the reader knows cap_height)
        hashes = cap.hashes if isinstance(cap, MerkleCap) else np.asarray(cap, dtype=np.uint64)
        self.write_all(_canon(hashes.reshape(-1, 4)).tobytes())

    def write_merkle_tree(self, tree):             # :1390-1405
        leaves = tree.leaves[:] if not isinstance(tree.leaves, np.ndarray) else tree.leaves
        leaves = np.asarray(leaves, dtype=np.uint64)
        n, ln = leaves.shape
        self.write_usize(n)
        rows = np.empty((n, ln + 1), dtype="<u8")  # every leaf: its length, then its elements
        rows[:, 0] = ln
        rows[:, 1:] = _canon(leaves)
        self.write_all(rows.tobytes())
        self.write_hash_vec(tree.digests)
        self.write_usize(tree.cap.height())
        self.write_merkle_cap(tree.cap)

    def write_polynomial_batch(self, batch):       # :1715-1734
        polys = batch.polynomials
        self.write_usize(len(polys))
        if polys:
            m = np.stack([np.asarray(p.coeffs, dtype=np.uint64) for p in polys])
            rows = np.empty((m.shape[0], m.shape[1] + 1), dtype="<u8")
            rows[:, 0] = m.shape[1]
            rows[:, 1:] = _canon(m)
            self.write_all(rows.tobytes())
        self.write_merkle_tree(batch.merkle_tree)
        self.write_usize(batch.degree_log)
        self.write_usize(batch.rate_bits)
        self.write_bool(batch.blinding)

    def write_merkle_proof(self, p):               # :1443-1457: u8 length, then the siblings
        sib = np.asarray(p.siblings, dtype=np.uint64).reshape(-1, 4)
        if sib.shape[0] > 255:
            raise OverflowError("Merkle proof length must fit in u8.")
        self.write_u8(sib.shape[0])
        self.write_all(_canon(sib).tobytes())

    def write_fri_initial_proof(self, fitp):       # :1477-1490
        for v, p in fitp.evals_proofs:
            self.write_field_vec(v)
            self.write_merkle_proof(p)

    def write_fri_query_step(self, fqs):           # :1508-1518
        self.write_field_ext_vec(fqs.evals)
        self.write_merkle_proof(fqs.merkle_proof)

    def write_fri_query_rounds(self, fqrs):        # :1532-1548
        for fqr in fqrs:
            self.write_fri_initial_proof(fqr.initial_trees_proof)
            for fqs in fqr.steps:
                self.write_fri_query_step(fqs)

    def write_fri_proof(self, fp):                 # :1568-1582
        for cap in fp.commit_phase_merkle_caps:
            self.write_merkle_cap(cap)
        self.write_fri_query_rounds(fp.query_round_proofs)
        self.write_field_ext_vec(fp.final_poly)
        self.write_field(fp.pow_witness)

    # ---- Read ------------------------------------------------------------------------------------------------------
    def _take(self, n):
        if self._pos + n > len(self._r):
            raise EOFError("IoError: unexpected end of buffer")     # serialization/mod.rs IoError
        b = self._r[self._pos:self._pos + n]
        self._pos += n
        return b

    def read_bool(self):
        return bool(self.read_u8())

    def read_u8(self):
        return self._take(1)[0]

    def read_usize(self):                          # :136-140
        return struct.unpack("<Q", self._take(8))[0]

    def read_field(self):                          # :156-164 (the reader does not reject non-canonical values)
        return struct.unpack("<Q", self._take(8))[0]

    def read_field_vec(self, length):
        return np.frombuffer(self._take(8 * length), dtype="<u8").astype(np.uint64)

    def read_field_ext_vec(self, length):
        return self.read_field_vec(2 * length).reshape(length, 2)

    def read_hash_vec(self, length):               # :271-279
        return self.read_field_vec(4 * length).reshape(length, 4)

    def read_merkle_cap(self, cap_height):         # :283-294
        return MerkleCap(self.read_hash_vec(1 << cap_height))

    def read_merkle_tree(self):                    # :309-330
        n = self.read_usize()
        if n == 0:
            leaves = np.empty((0, 0), dtype=np.uint64)
        else:
            ln = struct.unpack("<Q", self._r[self._pos:self._pos + 8])[0]
            rows = np.frombuffer(self._take(8 * n * (ln + 1)), dtype="<u8").reshape(n, ln + 1)
            if not (rows[:, 0] == ln).all():
                raise ValueError("ragged leaves: this engine only holds trees whose leaves have one length")
            leaves = rows[:, 1:].astype(np.uint64)
        digests = self.read_hash_vec(self.read_usize())
        cap_height = self.read_usize()
        cap = self.read_merkle_cap(cap_height)
        return MerkleTree(leaves, digests, cap)

    def read_polynomial_batch(self):               # :711-737 -> a host-side record with the reference's fields
        n_polys = self.read_usize()
        polys = []
        for _ in range(n_polys):
            plen = self.read_usize()
            polys.append(PolynomialCoeffs(self.read_field_vec(plen)))
        tree = self.read_merkle_tree()
        degree_log, rate_bits, blinding = self.read_usize(), self.read_usize(), self.read_bool()
        return HostPolynomialBatch(polys, tree, degree_log, rate_bits, blinding)

    def read_merkle_proof(self):                   # :396-407
        return MerkleProof(self.read_hash_vec(self.read_u8()))

    def read_fri_proof(self, initial_leaf_lens, reduction_arity_bits, cap_height, num_query_rounds, final_poly_len):
        """:555-578 / :506-530 / :422-457.  The reference derives these sizes from CommonCircuitData: the leaf length of
        every initial oracle (incl. salts), the arities, cap_height, num_query_rounds, fri_params.final_poly_len()."""
        caps = [self.read_merkle_cap(cap_height) for _ in reduction_arity_bits]
        rounds = []
        for _ in range(num_query_rounds):
            initial = [(self.read_field_vec(ln), self.read_merkle_proof()) for ln in initial_leaf_lens]
            steps = [FriQueryStep(self.read_field_ext_vec(1 << ar), self.read_merkle_proof()) for ar in reduction_arity_bits]
            rounds.append(FriQueryRound(FriInitialTreeProof(initial), steps))
        final_poly = self.read_field_ext_vec(final_poly_len)
        return FriProof(caps, rounds, final_poly, self.read_field())


class HostPolynomialBatch:
    """What read_polynomial_batch returns: the reference's public fields (oracle.rs:30-37) on the host."""

    def __init__(self, polynomials, merkle_tree, degree_log, rate_bits, blinding):
        self.polynomials, self.merkle_tree = polynomials, merkle_tree
        self.degree_log, self.rate_bits, self.blinding = degree_log, rate_bits, blinding


def polynomial_batch_to_bytes(batch):
    b = Buffer()
    b.write_polynomial_batch(batch)
    return b.bytes()


def fri_proof_to_bytes(proof):
    b = Buffer()
    b.write_fri_proof(proof)
    return b.bytes()
