"""Sharded commitment (plonky2_demo_b200/sharded.py): partition plan on CPU, the collective logic under
gloo with world_size 2 (the per-rank compute is replaced by the ORACLE -- test-only engine), and, on a GPU,
every shard of a commitment computed by the CUDA engine and assembled against the unsharded reference."""
import os
import sys

import numpy as np
import pytest

import oracle
from helpers import P, seeded_polys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class OracleShardEngine:
    """Test double of CudaShardEngine: same interface, CPU oracle arithmetic on torch CPU tensors."""

    def __init__(self):
        self.batches = {}

    @staticmethod
    def _u64(t):
        return t.numpy().view(np.uint64)

    def intt_local(self, values):
        if values.shape[0]:
            self._u64(values)[:] = oracle.fft(self._u64(values), inverse=True)

    def commit_shard(self, coeffs, w, plan, rank):
        c = np.ascontiguousarray(self._u64(coeffs)[:w])
        leaves = oracle.transpose_bitrev(oracle.coset_lde(c, plan.rate_bits))
        lo, hi = plan.leaf_range(rank)
        digests, cap = oracle.merkle_build(leaves[lo:hi], plan.local_cap_height)
        h = len(self.batches) + 1
        self.batches[h] = (np.ascontiguousarray(leaves[lo:hi]), digests, plan)
        return h, cap

    def shard_begin(self, plan, rank, salt_w=0):
        h = len(self.batches) + 1
        self.batches[h] = {"plan": plan, "rank": rank, "coeffs": np.zeros((plan.w, 1 << plan.lg_d), dtype=np.uint64), "seen": 0,
                           "salt_w": salt_w, "salts": None}
        return h

    def shard_extend(self, h, poly_first, coeffs, count):
        b = self.batches[h]
        b["coeffs"][poly_first : poly_first + count] = self._u64(coeffs)[:count]
        b["seen"] += count

    # ---- rows shards (all-to-all partition) and salt rows ----
    def lde_full(self, coeffs, rate_bits):
        import torch

        c = np.ascontiguousarray(self._u64(coeffs))
        if c.shape[0] == 0:
            return torch.empty((0, c.shape[1] << rate_bits), dtype=torch.int64)
        leaves = oracle.transpose_bitrev(oracle.coset_lde(c, rate_bits))          # [N][w_loc], leaf order
        return torch.from_numpy(np.ascontiguousarray(leaves.T).view(np.int64).copy())

    def rows_begin(self, width, salt_w, lg_n_local, local_cap_height):
        import torch

        h = len(self.batches) + 1
        self.batches[h] = {"rows": torch.zeros((width, 1 << lg_n_local), dtype=torch.int64), "lch": local_cap_height, "salt_w": salt_w}
        return h

    def rows_buffer(self, h, rows, n_local):
        return self.batches[h]["rows"][:rows]

    def set_rows(self, h, row_first, rows, canonical):
        b = self.batches[h]
        if "rows" in b:
            b["rows"][row_first : row_first + rows.shape[0]] = rows
        else:
            assert row_first == b["plan"].w and rows.shape[0] == b["salt_w"]
            b["salts"] = self._u64(rows).copy()

    def to_device(self, host_rows, like):
        import torch

        return torch.from_numpy(np.ascontiguousarray(host_rows).view(np.int64).copy())

    def shard_finish(self, h, plan):
        import torch

        b = self.batches.pop(h)
        if "rows" in b:
            leaves = np.ascontiguousarray(self._u64(b["rows"]).T) % np.uint64(P)
            digests, cap = oracle.merkle_build(leaves, b["lch"])
            self.batches[h] = (leaves, digests, plan)
            return cap
        assert b["seen"] == plan.w
        c = b["coeffs"]
        leaves = oracle.transpose_bitrev(oracle.coset_lde(c, plan.rate_bits))
        lo, hi = plan.leaf_range(b["rank"])
        leaves = np.ascontiguousarray(leaves[lo:hi])
        if b["salt_w"]:
            assert b["salts"] is not None
            leaves = np.ascontiguousarray(np.concatenate([leaves, (b["salts"] % np.uint64(P)).T], axis=1))
        digests, cap = oracle.merkle_build(leaves, plan.local_cap_height)
        self.batches[h] = (leaves, digests, plan)
        return cap

    def two_to_one(self, l, r):
        return oracle.two_to_one(np.ascontiguousarray(l), np.ascontiguousarray(r))

    def get_rows(self, h, idx, width):
        return self.batches[h][0][np.asarray(idx, dtype=np.int64)]

    def prove(self, h, i, n):
        leaves, digests, plan = self.batches[h]
        return oracle.merkle_prove(digests, leaves.shape[0], plan.local_cap_height, int(i)).reshape(n, 4)

    def digests(self, h, n):
        return self.batches[h][1]

    def free(self, h):
        self.batches.pop(h, None)


# ---------------------------------------------------------------------------------------------
# plan (pure host logic)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("w,lg_d,r,cap,world", [(135, 20, 3, 4, 8), (135, 20, 3, 4, 2), (20, 5, 3, 4, 8), (7, 4, 2, 0, 4), (5, 3, 3, 1, 8), (3, 2, 1, 3, 1)])
def test_shard_plan_partitions(w, lg_d, r, cap, world):
    from plonky2_demo_b200.sharded import ShardPlan

    p = ShardPlan(w, lg_d, r, cap, world)
    # polynomials: contiguous blocks covering [0, w) exactly once, padding only at the end
    got = [p.poly_range(k) for k in range(world)]
    assert got[0][0] == 0 and got[-1][1] == w
    assert all(a[1] == b[0] for a, b in zip(got, got[1:]))
    assert all(hi - lo <= p.w_max for lo, hi in got)
    full = [k for k, (lo, hi) in enumerate(got) if hi - lo == p.w_max]
    assert full == list(range(len(full)))  # full blocks first => gathered rows [0, w) are contiguous
    # leaves / cosets: contiguous, disjoint, whole coset blocks
    assert [p.leaf_range(k) for k in range(world)] == [(k * p.n_leaves // world, (k + 1) * p.n_leaves // world) for k in range(world)]
    assert p.coset_first(world - 1) + (1 << p.lg_cosets) == 1 << r
    assert p.local_leaves == (1 << lg_d) << p.lg_cosets
    assert (p.local_cap_len() * world) >> p.top_levels == 1 << cap


@pytest.mark.parametrize("w,world,chunks", [(135, 8, 4), (135, 2, 3), (20, 8, 4), (7, 4, 2), (5, 2, 9)])
def test_shard_plan_chunked_distribution(w, world, chunks):
    from plonky2_demo_b200.sharded import ShardPlan

    p = ShardPlan(w, 4, 3, 0, world, chunks)
    owned = sorted(j for k in range(world) for j in p.local_polys(k))
    assert owned == list(range(w))                       # every polynomial exactly once
    for c in range(p.chunks):
        lo, hi = p.chunk_range(c)
        m = p.chunk_rows(c)
        # gathered chunk = ranks' sub-blocks padded to m rows: valid rows are contiguous from the start
        got = [j for k in range(world) for j in range(*p.poly_ranges(k)[c])]
        assert got == list(range(lo, hi))
        sizes = [p.poly_ranges(k)[c][1] - p.poly_ranges(k)[c][0] for k in range(world)]
        full = [k for k, sz in enumerate(sizes) if sz == m]
        assert full == list(range(len(full))) and all(sz <= m for sz in sizes)


def test_shard_plan_rejects():
    from plonky2_demo_b200.sharded import ShardPlan

    with pytest.raises(ValueError, match="power of two"):
        ShardPlan(4, 4, 3, 0, 3)
    p = ShardPlan(4, 4, 1, 0, 4)                      # more ranks than coset blocks: only the all-to-all partition applies
    assert not p.coset_partition and p.local_leaves == 8 and p.lg_local == 3
    with pytest.raises(ValueError, match="world <= 2\\^rate_bits"):
        p.coset_first(1)
    with pytest.raises(ValueError, match="cannot shard 32 leaves over 64 ranks"):
        ShardPlan(4, 4, 1, 0, 64)
    with pytest.raises(ValueError, match="cap_height=9 should be at most"):
        ShardPlan(4, 4, 3, 9, 2)


@pytest.mark.parametrize("w,lg_d,r,cap,world", [(9, 4, 3, 4, 8), (9, 4, 3, 1, 8), (5, 3, 2, 0, 4), (6, 5, 1, 3, 2)])
def test_single_process_shards_assemble_to_reference(w, lg_d, r, cap, world):
    """All ranks' shards computed in one process (oracle engine): cap / digests / proofs == unsharded."""
    import torch

    from plonky2_demo_b200.sharded import ShardPlan

    coeffs = seeded_polys(w, 1 << lg_d, base_seed=77)
    ref = oracle.commit_from_coeffs(coeffs, r, cap)
    plan = ShardPlan(w, lg_d, r, cap, world)
    eng = OracleShardEngine()
    t = torch.from_numpy(coeffs.view(np.int64).copy())
    hs, caps = zip(*[eng.commit_shard(t, w, plan, k) for k in range(world)])
    cap_full = plan.assemble_cap(np.stack(caps), eng.two_to_one)
    assert np.array_equal(cap_full, ref["cap"])
    if plan.top_levels == 0:
        dig = np.concatenate([eng.digests(h, 0) for h in hs])
        assert np.array_equal(dig, ref["digests"])
    roots = np.stack(caps).reshape(world, -1, 4)[:, 0]
    for leaf in [0, 1, plan.local_leaves - 1, plan.local_leaves, plan.n_leaves - 1]:
        k = plan.owner_of_leaf(leaf)
        n_local = lg_d + plan.lg_cosets - plan.local_cap_height
        sib = eng.prove(hs[k], leaf - plan.leaf_range(k)[0], n_local) if n_local else np.zeros((0, 4), np.uint64)
        if plan.top_levels:
            sib = plan.extend_proof(k, sib, roots, eng.two_to_one)
        assert np.array_equal(sib.reshape(-1, 4), oracle.merkle_prove(ref["digests"], plan.n_leaves, cap, leaf).reshape(-1, 4)) or plan.top_levels
        assert oracle.merkle_verify(ref["leaves"][leaf], leaf, ref["cap"], sib)


# ---------------------------------------------------------------------------------------------
# collectives under gloo, world_size 2
# ---------------------------------------------------------------------------------------------
def _gloo_worker(rank, world, port, w, lg_d, r, cap, from_values, q, chunks=1, exchange="auto", blind=False):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist

    from plonky2_demo_b200.sharded import ShardedPolynomialBatch, ShardPlan

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        coeffs = seeded_polys(w, 1 << lg_d, base_seed=5)
        salts = seeded_polys(4, 1 << (lg_d + r), base_seed=0x5A1) if blind else None
        ref = oracle.commit_from_coeffs(coeffs, r, cap, salts=salts)
        plan = ShardPlan(w, lg_d, r, cap, world, chunks)
        src = oracle.fft(coeffs) if from_values else coeffs
        local = torch.from_numpy(np.ascontiguousarray(src[plan.local_polys(rank)]).view(np.int64).copy())
        eng = OracleShardEngine()
        if from_values:
            b = ShardedPolynomialBatch.from_values(local, w, r, cap, engine=eng)
        else:
            b = ShardedPolynomialBatch.from_coeffs(local, w, r, cap, engine=eng, chunks=chunks, exchange=exchange, salts=salts)
        ok = np.array_equal(b.cap, ref["cap"])
        if b.exchange.startswith("alltoall") != (exchange == "alltoall" or not plan.coset_partition):
            ok = False
        if chunks == 1 and not b.exchange.startswith("alltoall") and not blind:
            ok &= np.array_equal(b._coeffs.numpy().view(np.uint64)[:w], coeffs)
        leaves = [0, plan.local_leaves - 1, plan.local_leaves, plan.n_leaves - 1]
        ok &= np.array_equal(b.get_rows(leaves), ref["leaves"][leaves])
        ok &= np.array_equal(b.get_lde_values(3, 2), ref["leaves"][int(format(6, f"0{lg_d + r}b")[::-1], 2)][:w])
        for leaf in leaves:
            pr = b.prove(leaf)
            ok &= bool(oracle.merkle_verify(ref["leaves"][leaf], leaf, ref["cap"], pr.siblings))
        # the query phase's batched form: one collective for all paths, same siblings as the single-leaf calls
        many = b.prove_many(leaves)
        for leaf, pr in zip(leaves, many):
            ok &= np.array_equal(pr.siblings, b.prove(leaf).siblings)
            ok &= bool(oracle.merkle_verify(ref["leaves"][leaf], leaf, ref["cap"], pr.siblings))
        if plan.top_levels == 0:
            per = ref["digests"].shape[0] // world
            ok &= np.array_equal(b.local_digests(), ref["digests"][rank * per:(rank + 1) * per])
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("w,lg_d,r,cap,from_values,chunks,exchange,blind", [
    (9, 5, 3, 4, False, 1, "auto", False), (7, 4, 1, 0, False, 1, "auto", False), (9, 5, 3, 4, True, 1, "auto", False),
    (9, 5, 3, 4, False, 3, "auto", False),
    (9, 5, 3, 4, False, 1, "alltoall", False),          # the north-star's split: all-to-all of LDE rows
    (7, 4, 0, 0, False, 1, "auto", False),              # rate_bits 0: more ranks than coset blocks -> all-to-all, half a coset per rank
    (5, 3, 1, 3, False, 1, "alltoall", True),           # blinding salts through the all-to-all partition (ragged polynomial blocks)
    (9, 5, 3, 4, False, 1, "auto", True), (9, 5, 2, 1, False, 2, "auto", True),   # salts through the coset partition, one-shot and streaming
])
def test_gloo_world2_sharded_commit(w, lg_d, r, cap, from_values, chunks, exchange, blind):
    _run_gloo_world(2, w, lg_d, r, cap, from_values, chunks, exchange, blind)


def _run_gloo_world(world, w, lg_d, r, cap, from_values, chunks, exchange, blind):
    import socket

    import torch.multiprocessing as mp

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gloo_worker, args=(k, world, port, w, lg_d, r, cap, from_values, q, chunks, exchange, blind)) for k in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(k, True) for k in range(world)]


@pytest.mark.parametrize("w,lg_d,r,cap,exchange,blind", [
    (6, 4, 1, 2, "auto", False),        # 4 ranks, 2 coset blocks: auto picks the all-to-all, half a coset per rank, cap subtrees split evenly
    (3, 3, 2, 0, "alltoall", True),     # fewer polynomials than ranks (an empty block on the last rank), one root per rank + top levels, salts
    (9, 4, 3, 4, "allgather", False),   # the coset partition on 4 ranks for comparison
])
def test_gloo_world4_partitions(w, lg_d, r, cap, exchange, blind):
    _run_gloo_world(4, w, lg_d, r, cap, False, 1, exchange, blind)


# ---------------------------------------------------------------------------------------------
# the CUDA shard commit (every shard on one GPU) against the unsharded commit
# ---------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("w,lg_d,r,cap,world", [(135, 10, 3, 4, 8), (20, 12, 3, 4, 4), (9, 7, 3, 1, 8), (5, 3, 2, 0, 2), (17, 14, 3, 4, 2), (6, 9, 1, 3, 1)])
def test_cuda_shards_assemble_to_reference(w, lg_d, r, cap, world):
    import torch

    import plonky2_demo_b200 as p
    from plonky2_demo_b200.sharded import CudaShardEngine, ShardPlan

    p.init(0)
    coeffs = seeded_polys(w, 1 << lg_d, base_seed=31)
    ref = oracle.commit_from_coeffs(coeffs, r, cap)
    plan = ShardPlan(w, lg_d, r, cap, world)
    eng = CudaShardEngine()
    t = torch.from_numpy(coeffs.view(np.int64).copy()).cuda()
    hs, caps = zip(*[eng.commit_shard(t, w, plan, k) for k in range(world)])
    assert np.array_equal(plan.assemble_cap(np.stack(caps), eng.two_to_one), ref["cap"])
    n_dig = 2 * (plan.local_leaves - plan.local_cap_len())
    if plan.top_levels == 0:
        assert np.array_equal(np.concatenate([eng.digests(h, n_dig) for h in hs]), ref["digests"])
    rng = np.random.default_rng(1)
    for k in range(world):
        idx = rng.integers(0, plan.local_leaves, size=16)
        assert np.array_equal(eng.get_rows(hs[k], idx, w), ref["leaves"][idx + plan.leaf_range(k)[0]])
    for h in hs:
        eng.free(h)


@pytest.mark.gpu
@pytest.mark.parametrize("w,lg_d,r,cap,world,groups", [(135, 10, 3, 4, 8, [40, 40, 55]), (9, 12, 3, 2, 2, [1, 8]), (20, 13, 2, 0, 4, [7, 6, 7])])
def test_cuda_streaming_shard_commit(w, lg_d, r, cap, world, groups):
    """pcs_shard_begin / extend / finish: polynomials supplied in groups (contiguous and scattered device pointers)."""
    import ctypes as C

    import torch

    import plonky2_demo_b200 as p
    from plonky2_demo_b200 import _ffi
    from plonky2_demo_b200.sharded import ShardPlan

    p.init(0)
    L = _ffi.lib()
    coeffs = seeded_polys(w, 1 << lg_d, base_seed=17)
    ref = oracle.commit_from_coeffs(coeffs, r, cap)
    plan = ShardPlan(w, lg_d, r, cap, world)
    t = torch.from_numpy(coeffs.view(np.int64).copy()).cuda()
    scattered = [torch.from_numpy(coeffs[j].view(np.int64).copy()).cuda() for j in range(w)]
    caps = []
    for k in range(world):
        h = C.c_void_p()
        _ffi.check(L.pcs_shard_begin(w, 0, lg_d, r, plan.coset_first(k), plan.lg_cosets, plan.local_cap_height, C.byref(h)))
        first = 0
        for gi, cnt in enumerate(groups):
            if gi % 2 == 0:
                ptrs = _ffi.dev_ptr_array(t.data_ptr() + 8 * first * (1 << lg_d), cnt, 1 << lg_d)
            else:
                ptrs = (_ffi.u64p * cnt)(*[C.cast(C.c_void_p(scattered[first + i].data_ptr()), _ffi.u64p) for i in range(cnt)])
            _ffi.check(L.pcs_shard_extend(h, first, cnt, ptrs))
            first += cnt
        assert first == w
        cap_k = np.empty((plan.local_cap_len(), 4), dtype=np.uint64)
        _ffi.check(L.pcs_shard_finish(h, _ffi.ptr(cap_k)))
        assert L.pcs_shard_extend(h, 0, 1, _ffi.dev_ptr_array(t.data_ptr(), 1, 1 << lg_d)) != 0   # finished batches are closed
        caps.append(cap_k)
        if plan.top_levels == 0:
            n_dig = 2 * (plan.local_leaves - plan.local_cap_len())
            dig = np.empty((n_dig, 4), dtype=np.uint64)
            _ffi.check(L.pcs_batch_digests(h, _ffi.ptr(dig)))
            per = ref["digests"].shape[0] // world
            assert np.array_equal(dig, ref["digests"][k * per:(k + 1) * per])
        L.pcs_batch_free(h)
    from plonky2_demo_b200.sharded import CudaShardEngine

    assert np.array_equal(plan.assemble_cap(np.stack(caps), CudaShardEngine().two_to_one), ref["cap"])


@pytest.mark.gpu
def test_cuda_commit_from_scattered_device_pointers():
    """PCS_DEVICE_PTRS with separately allocated polynomials: the first NTT pass follows the pointer table (the path a
    shard uses to read peer GPUs' coefficient blocks in place); result == contiguous commit == oracle."""
    import ctypes as C

    import torch

    import plonky2_demo_b200 as p
    from plonky2_demo_b200 import _ffi

    p.init(0)
    for w, lg_d, r, cap in [(9, 12, 3, 4), (135, 10, 3, 4), (3, 1, 2, 0), (20, 15, 1, 2)]:
        coeffs = seeded_polys(w, 1 << lg_d, base_seed=91)
        ref = oracle.commit_from_coeffs(coeffs, r, cap)
        rows = [torch.from_numpy(coeffs[j].view(np.int64).copy()).cuda() for j in reversed(range(w))][::-1]  # scattered allocations
        ptrs = (_ffi.u64p * w)(*[C.cast(C.c_void_p(t.data_ptr()), _ffi.u64p) for t in rows])
        cap_out = np.empty((1 << cap, 4), dtype=np.uint64)
        h = C.c_void_p()
        _ffi.check(_ffi.lib().pcs_commit_from_coeffs(ptrs, w, lg_d, r, cap, None, 0, _ffi.PCS_DEVICE_PTRS, _ffi.ptr(cap_out), C.byref(h)))
        assert np.array_equal(cap_out, ref["cap"])
        dig = np.empty_like(ref["digests"])
        _ffi.check(_ffi.lib().pcs_batch_digests(h, _ffi.ptr(dig)))
        assert np.array_equal(dig, ref["digests"])
        _ffi.lib().pcs_batch_free(h)


@pytest.mark.gpu
def test_cuda_intt_dev_matches_oracle():
    import torch

    import plonky2_demo_b200 as p
    from plonky2_demo_b200.sharded import CudaShardEngine

    p.init(0)
    vals = seeded_polys(5, 1 << 11, base_seed=3)
    t = torch.from_numpy(vals.view(np.int64).copy()).cuda()
    CudaShardEngine().intt_local(t)
    p.synchronize()
    assert np.array_equal(t.cpu().numpy().view(np.uint64), oracle.fft(vals, inverse=True))


# ---------------------------------------------------------------------------------------------
# opening proof over sharded commitments (world = 1 here; tests/harness/sharded_opening_check.py runs it on 2+ GPUs)
# ---------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("chunks", [1, 3])
def test_cuda_opening_proof_over_sharded_oracles(chunks):
    import torch

    import plonky2_demo_b200 as p
    from oracle import fri_ref as fr
    from plonky2_demo_b200 import fri_prover as fp
    from plonky2_demo_b200.sharded import ShardedPolynomialBatch

    p.init(0)
    lg_d, r, cap, widths = 9, 2, 2, [7, 12, 3]
    coeffs = [seeded_polys(w, 1 << lg_d, base_seed=77 + k) for k, w in enumerate(widths)]
    plain = [p.PolynomialBatch.from_coeffs(c, r, False, cap, keep_coeffs=True) for c in coeffs]
    # chunks > 1 only takes effect with a process group; world == 1 exercises the single-tensor layout either way
    sharded = [ShardedPolynomialBatch.from_coeffs(torch.from_numpy(c.view(np.int64).copy()).cuda(), c.shape[0], r, cap,
                                                  partitioned=False, chunks=chunks) for c in coeffs]
    for a, b in zip(plain, sharded):
        assert np.array_equal(a.merkle_tree.cap.hashes, b.cap)
    zeta = (123456789, 987654321)
    for a, b, c in zip(plain, sharded, coeffs):
        want = fr.eval_base_polys_ext(c, zeta)
        assert np.array_equal(fp.eval_commitment(zeta, a), want)
        assert np.array_equal(fp.eval_commitment(zeta, b), want)
    all_polys = [(k, j) for k, w in enumerate(widths) for j in range(w)]
    zeta_next = fr.ext_mul((fr.primitive_root_of_unity(lg_d), 0), zeta)
    batches = [(zeta, all_polys), (zeta_next, [(2, 0), (2, 1)])]
    inst = fp.FriInstanceInfo([fp.FriOracleInfo(w, False) for w in widths],
                              [fp.FriBatchInfo(pt, [fp.FriPolynomialInfo(o, j) for o, j in polys]) for pt, polys in batches])
    cfg = p.FriConfig(r, cap, 6, p.FriReductionStrategy.Fixed([3, 2]), 5)
    params = cfg.fri_params(lg_d, False)

    def transcript():
        ch = fp.Challenger()
        for b in sharded:
            ch.observe_cap(b.cap)
        return ch

    alpha = (5, 6)
    f1, f2 = fp.final_poly(inst, plain, alpha), fp.final_poly(inst, sharded, alpha)
    mixed = fp.final_poly(inst, [plain[0], sharded[1], plain[2]], alpha)
    want = fr.final_poly(coeffs, batches, alpha)
    assert np.array_equal(f1.coeffs, want) and np.array_equal(f2.coeffs, want) and np.array_equal(mixed.coeffs, want)
    pa = fp.prove_openings(inst, plain, transcript(), params)
    pb = fp.prove_openings(inst, sharded, transcript(), params)
    assert pa.pow_witness == pb.pow_witness and pa.fri_query_indices == pb.fri_query_indices
    assert np.array_equal(pa.final_poly, pb.final_poly)
    for ca, cb in zip(pa.commit_phase_merkle_caps, pb.commit_phase_merkle_caps):
        assert ca == cb
    for ra, rb in zip(pa.query_round_proofs, pb.query_round_proofs):
        for (va, ma), (vb, mb) in zip(ra.initial_trees_proof.evals_proofs, rb.initial_trees_proof.evals_proofs):
            assert np.array_equal(va, vb) and np.array_equal(ma.siblings, mb.siblings)
        for sa, sb in zip(ra.steps, rb.steps):
            assert np.array_equal(sa.evals, sb.evals) and np.array_equal(sa.merkle_proof.siblings, sb.merkle_proof.siblings)
    for b in plain + sharded:
        b.free()
    for f in (f1, f2, mixed):
        f.free()


# ---------------------------------------------------------------------------------------------
# opening proof over a commitment sharded across two gloo ranks (CPU): the product's host glue (prove_openings, fri_proof,
# fri_committed_trees, the batched query phase) and the collectives of ShardedPolynomialBatch run for real; the device entry
# points they call are replaced by oracle-backed stand-ins, as OracleShardEngine does for the commit.
# ---------------------------------------------------------------------------------------------
def _gloo_opening_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist

    from oracle import fri_ref as fr
    from plonky2_demo_b200 import fri_prover as fp
    from plonky2_demo_b200.fri import FriConfig, FriReductionStrategy
    from plonky2_demo_b200.hashing import MerkleCap, MerkleProof, PoseidonPermutation
    from plonky2_demo_b200.sharded import ShardedPolynomialBatch, ShardPlan

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lg_d, r, cap, widths, arities, pow_bits, n_q = 6, 2, 2, [5, 3], [2, 2], 4, 5
        coeffs = [seeded_polys(w, 1 << lg_d, base_seed=900 + k) for k, w in enumerate(widths)]

        # ---- stand-ins for the device entry points ----
        PoseidonPermutation.permute = lambda self: setattr(self, "state", oracle.poseidon(self.state)[0])

        class _Tree:
            def __init__(self, leaves, cap_height):
                self.leaves, self.cap_height = leaves, cap_height
                self.digests, cap_ = oracle.merkle_build(leaves, cap_height)
                self.merkle_tree = type("T", (), {"cap": MerkleCap(cap_)})()

            def get_rows(self, idx):
                return self.leaves[np.asarray(list(idx), dtype=np.int64)]

            def prove_many(self, idx):
                return [MerkleProof(oracle.merkle_prove(self.digests, self.leaves.shape[0], self.cap_height, int(i))) for i in idx]

            def free(self):
                pass

        class _ExtPoly:
            def __init__(self, c):
                self.c = c

            def __len__(self):
                return self.c.shape[0]

            @property
            def coeffs(self):
                return self.c

            def commit_layer(self, rate_bits, shift, arity_bits, cap_height):
                vals = fr.ext_coset_lde(self.c, rate_bits, shift)
                n = vals.shape[0]
                idx = oracle.reverse_index_bits(np.arange(n, dtype=np.uint64)).astype(np.int64)
                return _Tree(np.ascontiguousarray(vals[idx].reshape(n >> arity_bits, 2 << arity_bits)), cap_height)

            def fold(self, arity_bits, beta):
                self.c = fr.fri_fold(self.c, 1 << arity_bits, beta)

            def free(self):
                pass

        def _final_poly(instance, oracles, alpha):
            full = [o._coeffs.numpy().view(np.uint64)[: o.n_polys] for o in oracles]      # replicated by the exchange
            batches = [(b.point, [(p.oracle_index, p.polynomial_index) for p in b.polynomials]) for b in instance.batches]
            return _ExtPoly(fr.final_poly(full, batches, alpha))

        def _pow(state, buf, cfg):
            ch = fr.Challenger()
            ch.sponge_state, ch.input_buffer = [int(x) for x in state], [int(x) for x in buf]
            return fr.fri_proof_of_work(ch, cfg.proof_of_work_bits)

        fp.final_poly, fp.fri_proof_of_work = _final_poly, _pow

        # ---- the sharded commitments (oracle-backed shard engine, real gloo collectives) ----
        eng = OracleShardEngine()
        sharded = []
        for c in coeffs:
            plan = ShardPlan(c.shape[0], lg_d, r, cap, world, 1)
            local = torch.from_numpy(np.ascontiguousarray(c[plan.local_polys(rank)]).view(np.int64).copy())
            sharded.append(ShardedPolynomialBatch.from_coeffs(local, c.shape[0], r, cap, engine=eng))
        zeta = (123456789, 987654321)
        zeta_next = fr.ext_mul((fr.primitive_root_of_unity(lg_d), 0), zeta)
        batches = [(zeta, [(k, j) for k, w in enumerate(widths) for j in range(w)]), (zeta_next, [(1, 0), (1, 2)])]
        inst = fp.FriInstanceInfo([fp.FriOracleInfo(w, False) for w in widths],
                                  [fp.FriBatchInfo(pt, [fp.FriPolynomialInfo(o, j) for o, j in polys]) for pt, polys in batches])
        openings = [[tuple(int(x) for x in fr.eval_base_polys_ext(coeffs[o][j:j + 1], pt)[0]) for o, j in polys]
                    for pt, polys in batches]
        params = FriConfig(r, cap, pow_bits, FriReductionStrategy.Fixed(arities), n_q).fri_params(lg_d, False)

        def transcript(cls):
            ch = cls()
            for b in sharded:
                ch.observe_cap(b.cap)
            for vals in openings:
                ch.observe_extension_elements(vals)
            return ch

        proof = fp.prove_openings(inst, sharded, transcript(fp.Challenger), params)      # collective
        # reference: the restated prover over the unsharded commitments
        cpu = []
        for c in coeffs:
            o = oracle.commit_from_coeffs(c, r, cap)
            o["coeffs"], o["cap_height"] = c, cap
            cpu.append(o)
        want = fr.prove_openings(cpu, batches, transcript(fr.Challenger), r, cap, arities, pow_bits, n_q)
        ok = all(np.array_equal(b.cap, o["cap"]) for b, o in zip(sharded, cpu))
        ok &= proof.pow_witness == want["pow_witness"] and proof.fri_query_indices == want["_indices"]
        ok &= bool(np.array_equal(proof.final_poly, want["final_poly"]))
        ok &= all(np.array_equal(c.hashes, wc) for c, wc in zip(proof.commit_phase_merkle_caps, want["commit_phase_merkle_caps"]))
        for rr, wr in zip(proof.query_round_proofs, want["query_round_proofs"]):
            for (ev, mp), (wev, wmp) in zip(rr.initial_trees_proof.evals_proofs, wr["initial_trees_proof"]):
                ok &= bool(np.array_equal(ev, wev)) and bool(np.array_equal(np.asarray(mp.siblings).reshape(-1, 4), wmp))
            for s, ws in zip(rr.steps, wr["steps"]):
                ok &= bool(np.array_equal(s.evals, ws["evals"]))
                ok &= bool(np.array_equal(np.asarray(s.merkle_proof.siblings).reshape(-1, 4), ws["merkle_proof"]))
        as_oracle = {
            "commit_phase_merkle_caps": [c.hashes for c in proof.commit_phase_merkle_caps], "final_poly": proof.final_poly,
            "pow_witness": proof.pow_witness,
            "query_round_proofs": [
                {"initial_trees_proof": [(ev, np.asarray(mp.siblings).reshape(-1, 4)) for ev, mp in rr.initial_trees_proof.evals_proofs],
                 "steps": [{"evals": s.evals, "merkle_proof": np.asarray(s.merkle_proof.siblings).reshape(-1, 4)} for s in rr.steps]}
                for rr in proof.query_round_proofs]}
        ok &= bool(fr.verify_fri_proof(batches, openings, transcript(fr.Challenger), [o["cap"] for o in cpu], as_oracle, r, cap,
                                       arities, pow_bits, n_q, lg_d))
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_gloo_world2_opening_proof_over_sharded_commitments():
    import socket

    import torch.multiprocessing as mp

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gloo_opening_worker, args=(k, 2, port, q)) for k in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]


# ---------------------------------------------------------------------------------------------
# host helpers of the round-2 partitions (pure CPU)
# ---------------------------------------------------------------------------------------------
def test_chunk_specs_cover_every_polynomial_with_rate_aligned_boundaries():
    """streaming groups: every spec covers [0, w) exactly once; all boundaries but the last are multiples of the sponge rate 8
    (a group can only be hashed up to a multiple of 8 columns); "host:<rho>" groups obey b[k+2] <= b[1] + b[k+1] / rho"""
    from plonky2_demo_b200.sharded import ShardPlan

    for w in (1, 7, 8, 9, 16, 17, 84, 135, 1000):
        for spec in (1, 2, 3, 4, 9, "host:0.75", "host:0.4", "host:0.16", [w]):
            p = ShardPlan(w, 6, 3, 0, 4, spec)
            assert p.bounds[0] == 0 and p.bounds[-1] == w and p.bounds == sorted(set(p.bounds)), (w, spec, p.bounds)
            assert sum(p.sizes) == w and p.chunks == len(p.sizes)
            if w > 16 and p.chunks > 1:
                assert all(b % 8 == 0 for b in p.bounds[1:-1]), (w, spec, p.bounds)
            if isinstance(spec, str) and p.chunks > 2:
                rho = float(spec.split(":")[1])
                b = p.bounds
                assert all(b[k + 2] <= b[1] + b[k + 1] / rho + 8 for k in range(len(b) - 2)), (w, spec, b)
            # every rank's polynomials, chunk by chunk, tile the chunk
            for c in range(p.chunks):
                lo, hi = p.chunk_range(c)
                got = [j for k in range(4) for j in range(*p.poly_ranges(k)[c])]
                assert got == list(range(lo, hi))
    with pytest.raises(ValueError, match="must add up"):
        ShardPlan(10, 4, 3, 0, 2, [4, 4])


@pytest.mark.parametrize("lg_n,first,count", [(6, 0, 64), (6, 16, 16), (9, 384, 128), (3, 5, 3), (0, 0, 1)])
def test_salt_leaf_slice_is_the_row_bit_reversal(lg_n, first, count):
    """a rank's salt rows = the leaf-order slice of the natural-order salt columns: leaves[i] = column[reverse_bits(i)]
    (reverse_index_bits_in_place on the rows, oracle.rs:84), checked against the oracle's transpose + bit-reversal"""
    from plonky2_demo_b200.sharded import reverse_bits_array, salt_leaf_slice

    n = 1 << lg_n
    salts = seeded_polys(4, n, base_seed=0x5A17)
    leaves = oracle.transpose_bitrev(salts)                     # [n][4] in leaf order
    got = salt_leaf_slice(salts, lg_n, first, count)
    assert got.shape == (4, count)
    assert np.array_equal(got.T, leaves[first:first + count])
    idx = np.arange(n, dtype=np.uint64)
    assert np.array_equal(reverse_bits_array(reverse_bits_array(idx, lg_n), lg_n), idx)
