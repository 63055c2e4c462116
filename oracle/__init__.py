"""ORACLE -- test infrastructure, NOT the product.

ctypes/numpy front-end of ``oracle/liboracle.so`` (the C restatement of the reference's CPU
commit path, see oracle.c).  Allowed importers: ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py`` (cpu_baseline leg and ``--impl reference``).  The product package
``plonky2_demo_b200`` must never import this module (tests/test_no_oracle_in_product.py checks).

Parity status: "port" (the Rust reference cannot be built here); pinned by the reference's own
golden numbers in tests/golden/reference_kats.json -- see oracle.c header.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")
_lib = None

P = 0xFFFFFFFF00000001
u64p = C.POINTER(C.c_uint64)


def build(force=False):
    """Compile liboracle.so with the committed Makefile (gcc, OpenMP)."""
    if force or not os.path.exists(_LIB_PATH) or any(
        os.path.getmtime(os.path.join(_HERE, f)) > os.path.getmtime(_LIB_PATH)
        for f in ("oracle.c", "fri.c", "gl64.h", "poseidon_constants.h", "Makefile")
    ):
        subprocess.check_call(["make", "-C", _HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        try:
            _lib = C.CDLL(_LIB_PATH)
        except OSError:
            build(force=True)
            _lib = C.CDLL(_LIB_PATH)
        L = _lib
        sz, u32, u64, i32 = C.c_size_t, C.c_uint, C.c_uint64, C.c_int
        L.ref_reverse_index_bits.argtypes = [u64p, sz]
        L.ref_reverse_index_bits.restype = None
        L.ref_fft_batch.argtypes = [u64p, sz, u32, i32, u32]
        L.ref_coset_lde_batch.argtypes = [u64p, sz, u32, u32, u64, u64p]
        L.ref_poseidon_permute.argtypes = [u64p, sz, i32]
        L.ref_hash_or_noop.argtypes = [u64p, sz, sz, u64p]
        L.ref_two_to_one.argtypes = [u64p, u64p, sz, u64p]
        L.ref_merkle_build.argtypes = [u64p, sz, sz, u32, u64p, u64p]
        L.ref_merkle_prove.argtypes = [u64p, sz, u32, sz, u64p]
        L.ref_merkle_verify.argtypes = [u64p, sz, sz, u64p, u64p, u32]
        L.ref_transpose_bitrev.argtypes = [u64p, sz, sz, u64p]
        L.ref_commit_from_coeffs.argtypes = [u64p, sz, u32, u32, u32, u64p, sz, u64p, u64p, u64p]
        L.ref_commit_from_values.argtypes = [u64p, sz, u32, u32, u32, u64p, sz, u64p, u64p, u64p]
        for f in ("ref_gl_add", "ref_gl_sub", "ref_gl_mul", "ref_gl_pow"):
            getattr(L, f).argtypes = [u64, u64]
            getattr(L, f).restype = u64
        L.ref_gl_inverse_2exp.argtypes = [u32]
        L.ref_gl_inverse_2exp.restype = u64
        L.ref_gl_primitive_root_of_unity.argtypes = [u32]
        L.ref_gl_primitive_root_of_unity.restype = u64
        L.ref_poly_eval.argtypes = [u64p, sz, u64]
        L.ref_poly_eval.restype = u64
        L.ref_num_threads.restype = i32
        L.ref_set_num_threads.argtypes = [i32]
        L.ref_set_num_threads.restype = None
    return _lib


def _p(a):
    return a.ctypes.data_as(u64p) if a is not None else None


def _arr(a, copy=False):
    a = np.array(a, dtype=np.uint64, copy=True) if copy else np.ascontiguousarray(a, dtype=np.uint64)
    return a


def num_threads():
    return lib().ref_num_threads()


def set_num_threads(n):
    """Override OMP_NUM_THREADS (torchrun sets it to 1 for its children)."""
    lib().ref_set_num_threads(int(n))
    return num_threads()


def use_all_cores():
    """All host cores this process may run on, whatever OMP_NUM_THREADS says."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    return set_num_threads(n)


# ---- field --------------------------------------------------------------------------------
def gl_add(a, b):
    return lib().ref_gl_add(a, b)


def gl_sub(a, b):
    return lib().ref_gl_sub(a, b)


def gl_mul(a, b):
    return lib().ref_gl_mul(a, b)


def gl_pow(a, e):
    return lib().ref_gl_pow(a, e)


def inverse_2exp(e):
    return lib().ref_gl_inverse_2exp(e)


def primitive_root_of_unity(lg):
    return lib().ref_gl_primitive_root_of_unity(lg)


def poly_eval(coeffs, x):
    c = _arr(coeffs)
    return lib().ref_poly_eval(_p(c), c.size, x)


# ---- permutations / fft --------------------------------------------------------------------
def reverse_index_bits(a):
    a = _arr(a, copy=True)
    lib().ref_reverse_index_bits(_p(a), a.size)
    return a


def fft(polys, inverse=False, zero_factor=0):
    """fft_with_options / ifft_with_options on [w][n] (natural order)."""
    a = _arr(polys, copy=True)
    w, n = a.shape
    lg = n.bit_length() - 1
    assert 1 << lg == n
    assert lib().ref_fft_batch(_p(a), w, lg, int(inverse), zero_factor) == 0
    return a


def coset_lde(coeffs, rate_bits, shift=7):
    """PolynomialBatch::lde_values: [w][d] -> [w][N] natural order."""
    c = _arr(coeffs)
    w, d = c.shape
    lg_d = d.bit_length() - 1
    assert 1 << lg_d == d
    out = np.empty((w, d << rate_bits), dtype=np.uint64)
    assert lib().ref_coset_lde_batch(_p(c), w, lg_d, rate_bits, shift, _p(out)) == 0
    return out


# ---- hashing -------------------------------------------------------------------------------
def poseidon(states, naive=False):
    s = _arr(states, copy=True).reshape(-1, 12)
    lib().ref_poseidon_permute(_p(s), s.shape[0], int(naive))
    return s


def hash_or_noop(rows):
    r = _arr(rows)
    n, ln = r.shape
    out = np.empty((n, 4), dtype=np.uint64)
    lib().ref_hash_or_noop(_p(r), n, ln, _p(out))
    return out


def two_to_one(left, right):
    l, r = _arr(left).reshape(-1, 4), _arr(right).reshape(-1, 4)
    out = np.empty_like(l)
    lib().ref_two_to_one(_p(l), _p(r), l.shape[0], _p(out))
    return out


def merkle_build(leaves, cap_height):
    """MerkleTree::new -> (digests [2(n-2^cap)][4], cap [2^cap][4]); raises like the reference panics."""
    lv = _arr(leaves)
    n, ln = lv.shape
    if n & (n - 1) or n == 0:
        raise ValueError(f"Not a power of two: {n}")
    lg = n.bit_length() - 1
    if cap_height > lg:
        raise ValueError(f"cap_height={cap_height} should be at most log2(leaves.len())={lg}")
    nd = 2 * (n - (1 << cap_height))
    digests = np.empty((nd, 4), dtype=np.uint64)
    cap = np.empty((1 << cap_height, 4), dtype=np.uint64)
    rc = lib().ref_merkle_build(_p(lv), n, ln, cap_height, _p(digests), _p(cap))
    assert rc == 0, rc
    return digests, cap


def merkle_prove(digests, n_leaves, cap_height, leaf_index):
    dg = _arr(digests)
    lg = n_leaves.bit_length() - 1
    sib = np.empty((max(lg - cap_height, 0), 4), dtype=np.uint64)
    k = lib().ref_merkle_prove(_p(dg), n_leaves, cap_height, leaf_index, _p(sib))
    assert k == sib.shape[0]
    return sib


def merkle_verify(leaf, leaf_index, cap, siblings):
    lf, cp, sb = _arr(leaf), _arr(cap), _arr(siblings)
    return bool(lib().ref_merkle_verify(_p(lf), lf.size, leaf_index, _p(cp), _p(sb), sb.shape[0] if sb.size else 0))


def transpose_bitrev(lde):
    a = _arr(lde)
    w, n = a.shape
    out = np.empty((n, w), dtype=np.uint64)
    assert lib().ref_transpose_bitrev(_p(a), w, n, _p(out)) == 0
    return out


def commit_from_coeffs(coeffs, rate_bits, cap_height, salts=None):
    """PolynomialBatch::from_coeffs -> dict(leaves [N][w+s], digests, cap)."""
    c = _arr(coeffs)
    w, d = c.shape
    lg_d = d.bit_length() - 1
    assert 1 << lg_d == d
    n = d << rate_bits
    sw = 0
    if salts is not None:
        salts = _arr(salts)
        sw = salts.shape[0]
        assert salts.shape[1] == n
    if cap_height > lg_d + rate_bits:
        raise ValueError(f"cap_height={cap_height} should be at most log2(leaves.len())={lg_d + rate_bits}")
    leaves = np.empty((n, w + sw), dtype=np.uint64)
    digests = np.empty((2 * (n - (1 << cap_height)), 4), dtype=np.uint64)
    cap = np.empty((1 << cap_height, 4), dtype=np.uint64)
    rc = lib().ref_commit_from_coeffs(_p(c), w, lg_d, rate_bits, cap_height, _p(salts), sw, _p(leaves), _p(digests), _p(cap))
    assert rc == 0, rc
    return {"leaves": leaves, "digests": digests, "cap": cap}


def commit_from_values(values, rate_bits, cap_height, salts=None):
    """PolynomialBatch::from_values -> dict(coeffs, leaves, digests, cap)."""
    v = _arr(values, copy=True)
    w, d = v.shape
    lg_d = d.bit_length() - 1
    n = d << rate_bits
    sw = 0
    if salts is not None:
        salts = _arr(salts)
        sw = salts.shape[0]
    leaves = np.empty((n, w + sw), dtype=np.uint64)
    digests = np.empty((2 * (n - (1 << cap_height)), 4), dtype=np.uint64)
    cap = np.empty((1 << cap_height, 4), dtype=np.uint64)
    rc = lib().ref_commit_from_values(_p(v), w, lg_d, rate_bits, cap_height, _p(salts), sw, _p(leaves), _p(digests), _p(cap))
    assert rc == 0, rc
    return {"coeffs": v, "leaves": leaves, "digests": digests, "cap": cap}
