"""FRI opening proof: mirror of plonky2/src/fri/{prover,structure,proof}.rs, iop/challenger.rs and
PolynomialBatch::prove_openings (fri/oracle.rs:162-219), executed by the CUDA engine.

Everything that touches polynomial data runs on the device and stays there: the openings (eval_commitment), the
alpha-combination / quotients of prove_openings, the extension-field coset LDE, every commit-phase Merkle tree, the
folds, the proof-of-work search and the query-phase row / path gathers.  The host keeps the transcript (Challenger,
a 12-element sponge whose permutation is the engine's Poseidon kernel) and the few hundred field elements of the proof.

Extension elements (F::Extension, D = 2) are pairs (a, b) = a + b X, X^2 = 7; arrays of them are uint64 [n][2].
"""
import ctypes as C
from dataclasses import dataclass, field
from typing import List, Tuple

import numpy as np

from . import _ffi
from .fri import PolynomialBatch, _DeviceMerkleTree, fri_proof_of_work
from .hashing import MerkleCap, MerkleProof, PoseidonPermutation
from .polynomial import GOLDILOCKS_ORDER, log2_strict

P = GOLDILOCKS_ORDER
COSET_SHIFT = 7   # F::coset_shift() = MULTIPLICATIVE_GROUP_GENERATOR (types.rs:437, goldilocks_field.rs:76)
POWER_OF_TWO_GENERATOR = 1753635133440165772   # goldilocks_field.rs:84


def primitive_root_of_unity(n_log):
    """types.rs:268-272"""
    assert n_log <= 32
    g = POWER_OF_TWO_GENERATOR
    for _ in range(n_log, 32):
        g = g * g % P
    return g


def ext_mul(x, y):
    """quadratic.rs:184-192 (host scalars: challenges and opening points only)"""
    return ((x[0] * y[0] + 7 * x[1] * y[1]) % P, (x[0] * y[1] + x[1] * y[0]) % P)


def _ext_arg(x):
    return (C.c_uint64 * 2)(int(x[0]) % P, int(x[1]) % P)


# ---- transcript ---------------------------------------------------------------------------------------------------
class Challenger:
    """iop/challenger.rs:16-160: duplex sponge in overwrite mode over the engine's Poseidon permutation."""

    def __init__(self):
        self.sponge_state = PoseidonPermutation()
        self.input_buffer: List[int] = []
        self.output_buffer: List[int] = []

    def observe_element(self, element):
        self.output_buffer = []
        self.input_buffer.append(int(element) % P)
        if len(self.input_buffer) == PoseidonPermutation.RATE:
            self.duplexing()

    def observe_elements(self, elements):
        for e in elements:
            self.observe_element(e)

    def observe_extension_element(self, element):
        self.observe_elements([element[0], element[1]])

    def observe_extension_elements(self, elements):
        for e in elements:
            self.observe_extension_element(e)

    def observe_hash(self, h):
        self.observe_elements(getattr(h, "elements", h))

    def observe_cap(self, cap):
        hashes = cap.hashes if isinstance(cap, MerkleCap) else np.asarray(cap).reshape(-1, 4)
        for h in hashes:
            self.observe_hash(h)

    def get_challenge(self):
        if self.input_buffer or not self.output_buffer:
            self.duplexing()
        return self.output_buffer.pop()

    def get_n_challenges(self, n):
        return [self.get_challenge() for _ in range(n)]

    def get_extension_challenge(self):
        a, b = self.get_n_challenges(2)
        return (a, b)

    def duplexing(self):
        assert len(self.input_buffer) <= PoseidonPermutation.RATE
        self.sponge_state.set_from_iter(self.input_buffer, 0)
        self.input_buffer = []
        self.sponge_state.permute()
        self.output_buffer = [int(x) for x in self.sponge_state.squeeze()]

    def clone(self):
        c = Challenger()
        c.sponge_state.state = self.sponge_state.state.copy()
        c.input_buffer = list(self.input_buffer)
        c.output_buffer = list(self.output_buffer)
        return c


# ---- instance description (fri/structure.rs) ----------------------------------------------------------------------
@dataclass
class FriPolynomialInfo:
    oracle_index: int
    polynomial_index: int

    @staticmethod
    def from_range(oracle_index, polynomial_indices):
        return [FriPolynomialInfo(oracle_index, j) for j in polynomial_indices]


@dataclass
class FriOracleInfo:
    num_polys: int
    blinding: bool


@dataclass
class FriBatchInfo:
    point: Tuple[int, int]
    polynomials: List[FriPolynomialInfo]


@dataclass
class FriInstanceInfo:
    oracles: List[FriOracleInfo]
    batches: List[FriBatchInfo]


# ---- proof (fri/proof.rs) -----------------------------------------------------------------------------------------
@dataclass
class FriQueryStep:
    evals: np.ndarray            # [arity][2]
    merkle_proof: MerkleProof


@dataclass
class FriInitialTreeProof:
    evals_proofs: List[Tuple[np.ndarray, MerkleProof]]


@dataclass
class FriQueryRound:
    initial_trees_proof: FriInitialTreeProof
    steps: List[FriQueryStep]


@dataclass
class FriProof:
    commit_phase_merkle_caps: List[MerkleCap]
    query_round_proofs: List[FriQueryRound]
    final_poly: np.ndarray       # PolynomialCoeffs<F::Extension>, [len][2]
    pow_witness: int
    fri_query_indices: List[int] = field(default_factory=list)   # not serialised by the reference; kept for tests


# ---- device-resident extension polynomial ---------------------------------------------------------------------------
class ExtensionPolynomial:
    """PolynomialCoeffs<F::Extension> held on the device (pcs_ext_poly)."""

    def __init__(self, handle):
        self._h = handle

    @classmethod
    def from_coeffs(cls, coeffs):
        c = _ffi.as_u64(coeffs).reshape(-1, 2)
        h = C.c_void_p()
        _ffi.check(_ffi.lib().pcs_ext_poly_new(_ffi.ptr(c), c.shape[0], C.byref(h)))
        return cls(h)

    def __len__(self):
        n = C.c_size_t()
        _ffi.check(_ffi.lib().pcs_ext_poly_len(self._h, C.byref(n)))
        return int(n.value)

    @property
    def coeffs(self):
        out = np.empty((len(self), 2), dtype=np.uint64)
        _ffi.check(_ffi.lib().pcs_ext_poly_read(self._h, _ffi.ptr(out)))
        return out

    def lde_coset_fft(self, rate_bits, shift=COSET_SHIFT):
        """self.lde(rate_bits).coset_fft(shift.into()): values in natural order, [len << rate_bits][2]."""
        out = np.empty((len(self) << rate_bits, 2), dtype=np.uint64)
        _ffi.check(_ffi.lib().pcs_ext_coset_lde(self._h, rate_bits, int(shift) % P, _ffi.ptr(out)))
        return out

    def commit_layer(self, rate_bits, shift, arity_bits, cap_height):
        """One tree of fri_committed_trees (prover.rs:81-87) -> device-resident MerkleTree view."""
        n_leaves = (len(self) << rate_bits) >> arity_bits
        if cap_height > log2_strict(n_leaves):
            raise ValueError(f"cap_height={cap_height} should be at most log2(leaves.len())={log2_strict(n_leaves)}")
        cap = np.empty((1 << cap_height, 4), dtype=np.uint64)
        h = C.c_void_p()
        _ffi.check(_ffi.lib().pcs_fri_commit_layer(self._h, rate_bits, int(shift) % P, arity_bits, cap_height,
                                                   _ffi.ptr(cap), C.byref(h)))
        t = PolynomialBatch()
        t._h, t._cap, t._coeffs_host = h, cap, []
        t.degree_log, t.rate_bits, t.blinding, t.cap_height = log2_strict(n_leaves), 0, False, cap_height
        t.n_polys, t.salt_w, t.n_leaves = 2 << arity_bits, 0, n_leaves
        t.n_digests = 2 * (n_leaves - (1 << cap_height))
        t.merkle_tree = _DeviceMerkleTree(t)
        return t

    def fold(self, arity_bits, beta):
        """prover.rs:93-101, in place."""
        _ffi.check(_ffi.lib().pcs_fri_fold(self._h, arity_bits, _ext_arg(beta)))

    def free(self):
        if self._h is not None:
            _ffi.lib().pcs_ext_poly_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


# ---- openings (plonk/proof.rs:316-322) ----------------------------------------------------------------------------
def _order_after_torch(batch):
    """A sharded batch's collectives complete on torch's stream: make the engine's stream wait for them."""
    eng = getattr(batch, "engine", None)
    if eng is not None and hasattr(eng, "_order_after_torch"):
        eng._order_after_torch()


def _poly_ptr_array(addresses):
    arr = (_ffi.u64p * len(addresses))()
    for i, a in enumerate(addresses):
        arr[i] = C.cast(C.c_void_p(int(a)), _ffi.u64p)
    return arr


def eval_commitment(z, batch):
    """c.polynomials.par_iter().map(|p| p.to_extension().eval(z)) -> [w][2], on the device.  `batch`: a PolynomialBatch, or a
    ShardedPolynomialBatch (every rank holds every coefficient after the exchange and evaluates all of them)."""
    out = np.empty((batch.n_polys, 2), dtype=np.uint64)
    if hasattr(batch, "poly_device_ptrs"):
        _order_after_torch(batch)
        ptrs = _poly_ptr_array(batch.poly_device_ptrs())
        _ffi.check(_ffi.lib().pcs_eval_ext_dev(ptrs, batch.n_polys, batch.degree_log, _ext_arg(z), _ffi.ptr(out)))
    else:
        _ffi.check(_ffi.lib().pcs_batch_eval_ext(batch._h, _ext_arg(z), _ffi.ptr(out)))
    return out


# ---- prover (fri/oracle.rs:162-219, fri/prover.rs) ------------------------------------------------------------------
def final_poly(instance, oracles, alpha):
    """The polynomial FRI runs on (oracle.rs:171-200), before .lde(rate_bits)."""
    total = sum(len(b.polynomials) for b in instance.batches)
    points = np.array([[int(b.point[0]) % P, int(b.point[1]) % P] for b in instance.batches], dtype=np.uint64)
    lens = (C.c_size_t * len(instance.batches))(*[len(b.polynomials) for b in instance.batches])
    oi = np.fromiter((p.oracle_index for b in instance.batches for p in b.polynomials), dtype=np.uint32, count=total)
    pi = np.fromiter((p.polynomial_index for b in instance.batches for p in b.polynomials), dtype=np.uint32, count=total)
    h = C.c_void_p()
    if any(hasattr(o, "poly_device_ptrs") for o in oracles):
        # sharded commitments: the coefficients live in the exchange buffers, not in a batch -> address them directly
        per_oracle = []
        for o in oracles:
            if hasattr(o, "poly_device_ptrs"):
                _order_after_torch(o)
                per_oracle.append(o.poly_device_ptrs())
            else:
                base = _ffi.lib().pcs_batch_coeffs_dev(o._h)
                if not base:
                    raise _ffi.PcsError(-5, "coefficients were not kept (PCS_KEEP_COEFFS)")
                per_oracle.append([base + 8 * (j << o.degree_log) for j in range(o.n_polys)])
        if len({o.degree_log for o in oracles}) != 1:
            raise _ffi.PcsError(-5, "oracles of different degrees")
        ptrs = _poly_ptr_array([per_oracle[int(o)][int(j)] for o, j in zip(oi, pi)])
        _ffi.check(_ffi.lib().pcs_fri_final_poly_dev(ptrs, oracles[0].degree_log, len(instance.batches), _ffi.ptr(points),
                                                     lens, _ext_arg(alpha), C.byref(h)))
        return ExtensionPolynomial(h)
    handles = (C.c_void_p * len(oracles))(*[o._h for o in oracles])
    _ffi.check(_ffi.lib().pcs_fri_final_poly(handles, len(oracles), len(instance.batches), _ffi.ptr(points), lens,
                                             oi.ctypes.data_as(C.POINTER(C.c_uint32)),
                                             pi.ctypes.data_as(C.POINTER(C.c_uint32)), _ext_arg(alpha), C.byref(h)))
    return ExtensionPolynomial(h)


def fri_committed_trees(poly, challenger, fri_params):
    """prover.rs:69-112.  `poly`: the d non-zero coefficients of lde_polynomial_coeffs (the LDE padding is implicit);
    consumed (folded in place).  Returns (trees, final_coeffs [len][2])."""
    trees = []
    shift = COSET_SHIFT
    rate_bits, cap_height = fri_params.config.rate_bits, fri_params.config.cap_height
    for arity_bits in fri_params.reduction_arity_bits:
        tree = poly.commit_layer(rate_bits, shift, arity_bits, cap_height)
        challenger.observe_cap(tree.merkle_tree.cap)
        trees.append(tree)
        beta = challenger.get_extension_challenge()
        poly.fold(arity_bits, beta)
        shift = pow(shift, 1 << arity_bits, P)
    final_coeffs = poly.coeffs   # coeffs.truncate(len >> rate_bits): exactly the coefficients held
    challenger.observe_extension_elements([(int(a), int(b)) for a, b in final_coeffs])
    return trees, final_coeffs


def fri_prover_query_rounds(initial_merkle_trees, trees, indices, fri_params):
    """prover.rs:162-216 for all query indices at once: every tree serves its rows and paths in one device round trip
    each (pcs_batch_get_rows / pcs_batch_prove_many) instead of one per query."""
    n_q = len(indices)
    init_rows = [t.get_rows(indices) for t in initial_merkle_trees]
    init_paths = [t.prove_many(indices) for t in initial_merkle_trees]
    step_rows, step_paths = [], []
    cur = list(indices)
    for i, tree in enumerate(trees):
        arity_bits = fri_params.reduction_arity_bits[i]
        cur = [x >> arity_bits for x in cur]
        step_rows.append(tree.get_rows(cur))
        step_paths.append(tree.prove_many(cur))
    rounds = []
    for q in range(n_q):
        initial = [(init_rows[k][q], init_paths[k][q]) for k in range(len(initial_merkle_trees))]
        steps = [FriQueryStep(evals=step_rows[i][q].reshape(-1, 2), merkle_proof=step_paths[i][q])   # unflatten
                 for i in range(len(trees))]
        rounds.append(FriQueryRound(FriInitialTreeProof(initial), steps))
    return rounds


def fri_proof(initial_merkle_trees, poly, challenger, fri_params):
    """prover.rs:20-67.  `initial_merkle_trees`: the committed batches (their trees serve the query phase)."""
    n = len(poly) << fri_params.config.rate_bits
    trees, final_coeffs = fri_committed_trees(poly, challenger, fri_params)
    # PoW phase (prover.rs:115-160): search on the device, then replay through the transcript
    pow_witness = fri_proof_of_work(challenger.sponge_state.state, challenger.input_buffer, fri_params.config)
    challenger.observe_element(pow_witness)
    pow_response = challenger.get_challenge()
    assert pow_response < (1 << (64 - fri_params.config.proof_of_work_bits)) or fri_params.config.proof_of_work_bits == 0
    # query phase (prover.rs:162-181)
    indices = [r % n for r in challenger.get_n_challenges(fri_params.config.num_query_rounds)]
    rounds = fri_prover_query_rounds(initial_merkle_trees, trees, indices, fri_params)
    proof = FriProof(commit_phase_merkle_caps=[t.merkle_tree.cap for t in trees], query_round_proofs=rounds,
                     final_poly=final_coeffs, pow_witness=pow_witness, fri_query_indices=indices)
    for t in trees:
        t.free()
    return proof


def prove_openings(instance, oracles, challenger, fri_params, timing=None):
    """PolynomialBatch::prove_openings (oracle.rs:162-219): a batch opening proof for `instance` over the committed
    `oracles` (which must hold their coefficients on the device: from_values, or from_coeffs(keep_coeffs=True)).

    The oracles may be ShardedPolynomialBatch objects (one commitment spread over several GPUs): the call is
    then COLLECTIVE.  Every rank holds every coefficient after the commit's exchange and runs the same deterministic
    transcript, final polynomial, commit-phase trees and folds (they are 1/135th of the commitment's work, so they are
    replicated rather than sharded); only the query phase communicates: the rows and Merkle paths of the sharded initial
    trees are fetched from the rank that owns the leaf.  Every rank returns the same FriProof."""
    alpha = challenger.get_extension_challenge()
    poly = final_poly(instance, oracles, alpha)
    assert len(poly) == 1 << fri_params.degree_bits
    try:
        return fri_proof(oracles, poly, challenger, fri_params)
    finally:
        poly.free()


PolynomialBatch.prove_openings = staticmethod(prove_openings)


# ---- the PLONK prover's use of the above (plonk/proof.rs:286-372, plonk/circuit_data.rs:431-595) ------------------------
@dataclass
class PlonkOpeningShape:
    """The few numbers of CommonCircuitData that fix which polynomial is opened where (no lookups: the demo has none):
    oracle 0 = constants + sigmas, 1 = wires, 2 = Zs + partial products, 3 = quotient chunks."""
    degree_bits: int
    num_constants: int
    num_routed_wires: int
    num_wires: int
    num_challenges: int
    num_partial_products: int
    quotient_degree_factor: int

    def constants_range(self):
        return range(0, self.num_constants)                                            # circuit_data.rs:431-433

    def sigmas_range(self):
        return range(self.num_constants, self.num_constants + self.num_routed_wires)   # :436-438

    def zs_range(self):
        return range(0, self.num_challenges)                                           # :441-443

    def partial_products_range(self):
        return range(self.num_challenges, (self.num_partial_products + 1) * self.num_challenges)   # :446-448

    def oracle_widths(self):
        return [self.num_constants + self.num_routed_wires, self.num_wires,
                self.num_challenges * (1 + self.num_partial_products), self.num_challenges * self.quotient_degree_factor]

    def get_fri_instance(self, zeta):
        """circuit_data.rs:461-481: everything at zeta, the Z polynomials also at g * zeta."""
        g = primitive_root_of_unity(self.degree_bits)
        widths = self.oracle_widths()
        all_polys = [p for k, w in enumerate(widths) for p in FriPolynomialInfo.from_range(k, range(w))]   # fri_all_polys :586-595
        return FriInstanceInfo(
            oracles=[FriOracleInfo(w, False) for w in widths],
            batches=[FriBatchInfo(zeta, all_polys),
                     FriBatchInfo(ext_mul((g, 0), zeta), FriPolynomialInfo.from_range(2, self.zs_range()))])


@dataclass
class OpeningSet:
    """plonk/proof.rs:286-304: the purported values of every committed polynomial at zeta (and of the Zs at g * zeta)."""
    constants: np.ndarray
    plonk_sigmas: np.ndarray
    wires: np.ndarray
    plonk_zs: np.ndarray
    plonk_zs_next: np.ndarray
    partial_products: np.ndarray
    quotient_polys: np.ndarray

    @classmethod
    def new(cls, zeta, g, constants_sigmas_commitment, wires_commitment, zs_partial_products_commitment,
            quotient_polys_commitment, shape):
        """proof.rs:307-344: five batched evaluations on the device (eval_commitment), sliced by the ranges."""
        cs = eval_commitment(zeta, constants_sigmas_commitment)
        zs = eval_commitment(zeta, zs_partial_products_commitment)
        zs_next = eval_commitment(ext_mul((int(g), 0), zeta), zs_partial_products_commitment)   # g: base-field generator
        z0, z1 = shape.zs_range().start, shape.zs_range().stop
        p0, p1 = shape.partial_products_range().start, shape.partial_products_range().stop
        return cls(constants=cs[: shape.num_constants], plonk_sigmas=cs[shape.num_constants: shape.num_constants + shape.num_routed_wires],
                   wires=eval_commitment(zeta, wires_commitment), plonk_zs=zs[z0:z1], plonk_zs_next=zs_next[z0:z1],
                   partial_products=zs[p0:p1], quotient_polys=eval_commitment(zeta, quotient_polys_commitment))

    def to_fri_openings(self):
        """proof.rs:345-372 (no lookups): FriOpenings as two lists of extension values, in the instance's order."""
        zeta_batch = np.concatenate([self.constants, self.plonk_sigmas, self.wires, self.plonk_zs, self.partial_products,
                                     self.quotient_polys])
        return [[(int(a), int(b)) for a, b in zeta_batch], [(int(a), int(b)) for a, b in self.plonk_zs_next]]
