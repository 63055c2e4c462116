"""CPU tests: pin the oracle (oracle/oracle.c) against every golden number the reference's own
tests hold for the hot path, and restate the reference's property tests."""
import random

import numpy as np
import pytest

import oracle
from helpers import P, brev, field_grid, seeded_polys, splitmix64_stream


def test_field_constants(golden):
    f = golden["kats"]["field"]
    assert f["order"] == P == oracle.P
    # goldilocks_field.rs:76-87: generator 7 has full order; POWER_OF_TWO_GENERATOR = 7^((p-1)/2^32)
    assert oracle.gl_pow(f["multiplicative_group_generator"], (P - 1) >> f["two_adicity"]) == f["power_of_two_generator"]
    assert oracle.primitive_root_of_unity(32) == f["power_of_two_generator"]
    for lg in range(0, 33):
        w = oracle.primitive_root_of_unity(lg)
        assert pow(w, 1 << lg, P) == 1
        if lg:
            assert pow(w, 1 << (lg - 1), P) == P - 1
    for e in range(0, 33):
        assert oracle.inverse_2exp(e) * pow(2, e, P) % P == 1


def test_field_grid_vs_bigint():
    # prime_field_testing.rs:78-125: add/sub/mul on the boundary grid, checked against big-int %
    g = field_grid()
    for a in g:
        for b in g:
            assert oracle.gl_add(a, b) == (a + b) % P
            assert oracle.gl_sub(a, b) == (a - b) % P
            assert oracle.gl_mul(a, b) == (a * b) % P


def test_field_noncanonical_inputs():
    # goldilocks_field.rs:199-258: double wrap-around corner cases
    nc = [P, P + 1, (1 << 64) - 1, (1 << 64) - 2, P - 1, 0, 1, 0xFFFFFFFF, 0xFFFFFFFE]
    for a in nc:
        for b in nc:
            assert oracle.gl_add(a, b) == (a + b) % P
            assert oracle.gl_sub(a, b) == (a - b) % P
            assert oracle.gl_mul(a, b) == (a * b) % P


def test_reverse_index_bits_golden(golden):
    tab = golden["kats"]["reverse_index_bits_256"]
    out = oracle.reverse_index_bits(np.arange(256, dtype=np.uint64))
    assert out.tolist() == tab
    assert oracle.reverse_index_bits(np.array([10, 20, 30, 40], dtype=np.uint64)).tolist() == [10, 30, 20, 40]


def test_poseidon_kats(golden):
    for kv in golden["kats"]["poseidon12_kats"] + golden["kats"]["poseidon12_extra_bigint"]:
        x = np.array(kv["input"], dtype=np.uint64)
        assert oracle.poseidon(x)[0].tolist() == kv["output"]
        assert oracle.poseidon(x, naive=True)[0].tolist() == kv["output"]
        assert oracle.poseidon(x, naive=2)[0].tolist() == kv["output"]  # plain circulant MDS, no fast form anywhere


def test_poseidon_consistency_and_noncanonical():
    # poseidon.rs:777-790 consistency (fast == naive), extended to random + non-canonical states
    rng = np.random.default_rng(7)
    x = rng.integers(0, 1 << 64, size=(4096, 12), dtype=np.uint64)
    # extreme halves for the add/shift MDS network (x86-64 form, poseidon_goldilocks.rs:217-248) against the
    # plain circulant sums (poseidon.rs:178-198)
    x[:64] = np.uint64(0xFFFFFFFFFFFFFFFF)
    x[64:128] = np.uint64(0xFFFFFFFF00000000)
    x[128:192] = np.uint64(0x00000000FFFFFFFF)
    x[192:256, ::2] = np.uint64(0xFFFFFFFFFFFFFFFF)
    x[192:256, 1::2] = 0
    a, b = oracle.poseidon(x), oracle.poseidon(x, naive=True)
    assert np.array_equal(a, b)
    assert np.array_equal(a, oracle.poseidon(x, naive=2))
    xc = np.where(x >= np.uint64(P), x - np.uint64(P), x)
    assert np.array_equal(oracle.poseidon(xc), a)
    assert (a < np.uint64(P)).all()


def _naive_dft(coeffs, lg_n, shift=1):
    n = 1 << lg_n
    w = oracle.primitive_root_of_unity(lg_n)
    return [oracle.poly_eval(coeffs, shift * pow(w, i, P) % P) for i in range(n)]


def test_fft_and_ifft():
    # fft.rs:219-253: degree 200 padded to 256
    rnd = random.Random(3)
    coeffs = np.array([rnd.randrange(P) for _ in range(200)] + [0] * 56, dtype=np.uint64)
    vals = oracle.fft(coeffs[None, :])[0]
    assert vals.tolist() == _naive_dft(coeffs, 8)
    back = oracle.fft(vals[None, :], inverse=True)[0]
    assert np.array_equal(back, coeffs)
    # zero_factor equivalence: fft(lde(r)) == fft_with_options(lde(r), Some(r))
    for r in range(4):
        padded = np.concatenate([coeffs, np.zeros((256 << r) - 256, dtype=np.uint64)])
        a = oracle.fft(padded[None, :])[0]
        b = oracle.fft(padded[None, :], zero_factor=r)[0]
        assert np.array_equal(a, b)


@pytest.mark.parametrize("lg_n", [0, 1, 2, 3, 5])
def test_fft_tiny(lg_n):
    rnd = random.Random(lg_n)
    c = np.array([rnd.randrange(P) for _ in range(1 << lg_n)], dtype=np.uint64)
    assert oracle.fft(c[None, :])[0].tolist() == _naive_dft(c, lg_n)
    assert np.array_equal(oracle.fft(oracle.fft(c[None, :]), inverse=True)[0], c)


def test_coset_lde_is_naive_eval_and_block_identity():
    # polynomial/mod.rs:478-497 test_coset_fft + SURVEY 8a identity
    lg_d, r = 5, 3
    d, n = 1 << lg_d, 1 << (lg_d + r)
    polys = seeded_polys(3, d)
    lde = oracle.coset_lde(polys, r)
    for j in range(3):
        assert lde[j].tolist() == _naive_dft(polys[j], lg_d + r, shift=7)
    leaves = oracle.transpose_bitrev(lde)
    w_n = oracle.primitive_root_of_unity(lg_d + r)
    w_d = oracle.primitive_root_of_unity(lg_d)
    for c in range(1 << r):
        for t in range(d):
            x = 7 * pow(w_n, brev(c, r), P) * pow(w_d, brev(t, lg_d), P) % P
            for j in range(3):
                assert int(leaves[c * d + t][j]) == oracle.poly_eval(polys[j], x)


def test_hash_or_noop_semantics():
    rng = np.random.default_rng(5)
    for ln in [1, 2, 3, 4]:
        rows = rng.integers(0, 1 << 64, size=(8, ln), dtype=np.uint64)
        out = oracle.hash_or_noop(rows)
        exp = np.zeros((8, 4), dtype=np.uint64)
        exp[:, :ln] = np.where(rows >= np.uint64(P), rows - np.uint64(P), rows)
        assert np.array_equal(out, exp)
    # overwrite-mode sponge: hashing.rs:119-142
    for ln in [5, 7, 8, 9, 16, 17, 135]:
        rows = rng.integers(0, P, size=(4, ln), dtype=np.uint64)
        out = oracle.hash_or_noop(rows)
        for k in range(4):
            st = np.zeros(12, dtype=np.uint64)
            for off in range(0, ln, 8):
                chunk = rows[k, off : off + 8]
                st[: chunk.size] = chunk
                st = oracle.poseidon(st)[0]
            assert np.array_equal(out[k], st[:4])
    l, r = rng.integers(0, P, size=(4, 4), dtype=np.uint64), rng.integers(0, P, size=(4, 4), dtype=np.uint64)
    exp = oracle.poseidon(np.concatenate([l, r, np.zeros((4, 4), dtype=np.uint64)], axis=1))[:, :4]
    assert np.array_equal(oracle.two_to_one(l, r), exp)


def _digest_index(level, k):
    # closed form of merkle_tree.rs:197-201
    return 2 * (((k >> 1) << (level + 1)) + (1 << level) - 1) + (k & 1)


@pytest.mark.parametrize("log_n,leaf_len,cap_height", [(8, 7, 1), (8, 7, 8), (8, 7, 0), (6, 135, 4), (3, 3, 2), (0, 5, 0), (1, 9, 0)])
def test_merkle_trees(log_n, leaf_len, cap_height):
    # merkle_tree.rs:223-281: every proof verifies against the cap
    n = 1 << log_n
    leaves = splitmix64_stream(99 + log_n, n * leaf_len).reshape(n, leaf_len)
    digests, cap = oracle.merkle_build(leaves, cap_height)
    assert digests.shape == (2 * (n - (1 << cap_height)), 4)
    for i in range(n):
        sib = oracle.merkle_prove(digests, n, cap_height, i)
        assert oracle.merkle_verify(leaves[i], i, cap, sib)
    # layout: level-l node k of subtree s at the closed-form index; roots only in cap
    ns = n >> cap_height
    level = oracle.hash_or_noop(leaves)
    lvl = 0
    while level.shape[0] > (1 << cap_height):
        per = ns >> lvl
        for s in range(1 << cap_height):
            sub = digests[s * 2 * (ns - 1) : (s + 1) * 2 * (ns - 1)]
            for k in range(per):
                assert np.array_equal(sub[_digest_index(lvl, k)], level[s * per + k])
        level = oracle.two_to_one(level[0::2], level[1::2])
        lvl += 1
    assert np.array_equal(level, cap)


def test_merkle_cap_height_too_big():
    leaves = splitmix64_stream(1, 8 * 5).reshape(8, 5)
    with pytest.raises(ValueError):
        oracle.merkle_build(leaves, 4)
    with pytest.raises(ValueError):
        oracle.merkle_build(leaves[:6], 0)


def test_commit_from_values_roundtrip():
    lg_d, r, cap = 4, 2, 1
    vals = seeded_polys(5, 1 << lg_d, base_seed=77)
    out = oracle.commit_from_values(vals, r, cap)
    assert np.array_equal(oracle.fft(out["coeffs"]), vals)
    ref = oracle.commit_from_coeffs(out["coeffs"], r, cap)
    for k in ("leaves", "digests", "cap"):
        assert np.array_equal(out[k], ref[k])
    # get_lde_values(index, step): leaves[brev(index*step)] is the natural-order LDE column
    lde = oracle.coset_lde(out["coeffs"], r)
    n = 1 << (lg_d + r)
    for i in range(n):
        assert np.array_equal(out["leaves"][brev(i, lg_d + r)], lde[:, i])
    # salts are appended as extra columns (oracle.rs:119-123)
    salts = splitmix64_stream(5, 4 * n).reshape(4, n)
    salted = oracle.commit_from_coeffs(out["coeffs"], r, cap, salts=salts)
    assert salted["leaves"].shape == (n, 9)
    assert np.array_equal(salted["leaves"][:, :5], ref["leaves"])
    for i in range(n):
        assert np.array_equal(salted["leaves"][brev(i, lg_d + r), 5:], salts[:, i])
