"""One commitment spread over the GPUs of a single box: one process per GPU, torch.distributed for the plumbing.

The reference has no distributed path (one process, rayon); this is the north-star's
"large commitments are sharded across a single 8xB200 box".  The partition follows from the leaf-order
identity of the LDE (oracle.rs:83-84): after the row bit-reversal, leaves
[c*d, (c+1)*d) are the evaluations of ALL polynomials on ONE coset 7*w_N^{brev(c)}*<w_d>.  So

  * rank r of G owns the coset blocks [r*2^rate_bits/G, (r+1)*2^rate_bits/G) = a contiguous leaf range,
    hence whole cap subtrees and a contiguous slice of the reference's `digests` (merkle_tree.rs:43-46);
  * it needs every polynomial's d coefficients (W*d elements), never another rank's LDE output.

Two partitions are built (SURVEY 8e lists both; results are bit-identical):
  "coset"    (default when world <= 2^rate_bits): described above -- the coefficients travel (all-gather, or peer loads
             fused into the first NTT pass), every rank extends its own cosets;
  "alltoall" (the north-star's split; any power-of-two world up to the number of leaves, blinding included): the
             polynomials are partitioned, every rank extends ITS polynomials over all cosets, and one NCCL all-to-all over
             NVLink carries the LDE rows to the rank that owns their leaf range (rows land directly in that rank's leaf
             buffer, poly-major, ready for hashing).  Leaf ranges need not be whole cosets, so world may exceed 2^rate_bits.
Bytes received per rank: coset = (G-1)/G * W*d*8, alltoall = (G-1)/G * W*N*8/G -- equal at G = 2^rate_bits, the coset
partition is cheaper below it.

Exchange steps (the only collectives on the path):
  1. all-gather of the coefficients when the input is polynomial-partitioned (rank r holds a block of the
     W polynomials, e.g. after a per-rank IFFT in from_values): W*d*8 bytes in total, 8x less than the
     all-to-all of LDE rows a polynomial-partitioned LDE would need (9 GB at 135 x 2^20, rate 3);
  2. all-gather of the local caps (2^cap_height * 32 bytes in total).
When G > 2^cap_height the ranks' single roots are combined on every rank with two_to_one (log2 G - cap_height
levels, a handful of permutations).

The per-rank compute (IFFT of the local block, LDE of the local cosets, leaf hashing, node levels) is the
CUDA engine (`pcs_ntt_dev`, `pcs_commit_shard_from_coeffs`).  The engine object is injectable so that the
partition / collective logic can be exercised on CPU (gloo, world_size 2) in tests; the product never
substitutes it.
"""
import ctypes as C
import os

import numpy as np

from . import _ffi
from .polynomial import log2_strict, reverse_bits


class _null_ctx:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


class ShardPlan:
    """Who owns what for (w polys, d = 2^lg_d, rate_bits, cap_height) over `world` ranks."""

    def __init__(self, w, lg_d, rate_bits, cap_height, world, chunks=1):
        if world < 1 or world & (world - 1):
            raise ValueError(f"world size must be a power of two, got {world}")
        lg_w = log2_strict(world)
        if lg_w > lg_d + rate_bits:
            raise ValueError(f"cannot shard {1 << (lg_d + rate_bits)} leaves over {world} ranks")
        if cap_height > lg_d + rate_bits:
            raise ValueError(f"cap_height={cap_height} should be at most log2(leaves.len())={lg_d + rate_bits}")
        self.w, self.lg_d, self.rate_bits, self.cap_height, self.world = w, lg_d, rate_bits, cap_height, world
        self.lg_world = lg_w
        # whole coset blocks per rank: the coefficient exchange ("coset" partition) needs it; the all-to-all of LDE rows
        # works for any power-of-two world
        self.coset_partition = lg_w <= rate_bits
        self.lg_cosets = rate_bits - lg_w                     # coset blocks per rank (log2; negative: a fraction of a block)
        self.lg_local = lg_d + rate_bits - lg_w               # log2(leaves per rank)
        self.local_cap_height = max(cap_height - lg_w, 0)
        self.top_levels = max(lg_w - cap_height, 0)           # levels combined from the ranks' roots
        self.n_leaves = 1 << (lg_d + rate_bits)
        self.local_leaves = self.n_leaves >> lg_w
        self.w_max = -(-w // world)                           # block distribution of polynomials
        # streaming exchange: the polynomials are cut into contiguous groups and EVERY group is block-distributed over the
        # ranks, so that group c can be gathered, extended AND hashed (streaming sponge: group boundaries are multiples of the
        # sponge rate 8) while group c+1 is still moving.  `chunks` = a count (first group small -- it is the only transfer
        # nothing hides --, the rest in equal multiples of 8) or an explicit list of group sizes.
        if isinstance(chunks, str):
            # "host:<rho>": inputs stream from pinned host memory; rho = transfer time / compute time per polynomial.  The copies
            # run back to back, so group k+1 has landed before group k is finished iff b[k+2] <= b[1] + b[k+1] / rho: the
            # groups grow geometrically and only the first transfer (8 polynomials) is exposed.
            rho = float(chunks.split(":")[1]) if ":" in chunks else 0.75
            b = [0, min(8, w)]
            while b[-1] < w:
                nxt = int(8 + b[-1] / rho) // 8 * 8
                b.append(min(max(nxt, b[-1] + 8), w))
            if len(b) > 2 and w - b[-2] < 8:
                del b[-2]
            chunks = [hi - lo for lo, hi in zip(b, b[1:])]
        if isinstance(chunks, (list, tuple)):
            sizes = [int(x) for x in chunks if int(x) > 0]
            if sum(sizes) != w:
                raise ValueError(f"chunk sizes {sizes} must add up to {w} polynomials")
        else:
            k = max(1, min(int(chunks), w))
            if k == 1 or w <= 16:
                k = min(k, w)
                base = -(-w // k)
                sizes = [min(base, w - i * base) for i in range(k) if w - i * base > 0]
            else:
                first = 8
                per = -(-(w - first) // (k - 1))
                per = -(-per // 8) * 8
                sizes, left = [first], w - first
                while left > 0:
                    sizes.append(min(per, left))
                    left -= sizes[-1]
        self.bounds = [0]
        for x in sizes:
            self.bounds.append(self.bounds[-1] + x)
        self.chunks = len(sizes)
        self.sizes = sizes
        self.w_chunk = max(sizes)

    def coset_first(self, rank):
        if not self.coset_partition:
            raise ValueError(f"cannot shard 2^{self.rate_bits} coset blocks over {self.world} ranks (need world <= 2^rate_bits); "
                             "use exchange='alltoall'")
        return rank << self.lg_cosets

    def leaf_range(self, rank):
        return rank * self.local_leaves, (rank + 1) * self.local_leaves

    def owner_of_leaf(self, leaf):
        return leaf // self.local_leaves

    def poly_range(self, rank):
        """chunks == 1: the contiguous block of polynomials rank `rank` holds."""
        lo = min(rank * self.w_max, self.w)
        return lo, min(lo + self.w_max, self.w)

    def chunk_range(self, c):
        return self.bounds[c], self.bounds[c + 1]

    def chunk_rows(self, c):
        """rows per rank in the gather of chunk c (padded)"""
        lo, hi = self.chunk_range(c)
        return -(-(hi - lo) // self.world)

    def poly_ranges(self, rank):
        """[(lo, hi)] per chunk: the polynomials rank `rank` holds, in the order of its local rows."""
        out = []
        for c in range(self.chunks):
            lo, hi = self.chunk_range(c)
            m = self.chunk_rows(c)
            a = min(lo + rank * m, hi)
            out.append((a, min(a + m, hi)))
        return out

    def local_polys(self, rank):
        return [j for lo, hi in self.poly_ranges(rank) for j in range(lo, hi)]

    def local_cap_len(self):
        return 1 << self.local_cap_height

    def assemble_cap(self, local_caps, two_to_one):
        """local_caps: [world][2^local_cap_height][4] in rank order -> the reference's cap [2^cap_height][4]."""
        caps = np.asarray(local_caps, dtype=np.uint64).reshape(-1, 4)
        for _ in range(self.top_levels):
            caps = two_to_one(caps[0::2], caps[1::2])
        return np.ascontiguousarray(caps)

    def extend_proof(self, rank, local_siblings, roots, two_to_one):
        """Siblings above the local root when world > 2^cap_height (roots: [world][4])."""
        sib = [np.asarray(local_siblings, dtype=np.uint64).reshape(-1, 4)]
        level = np.asarray(roots, dtype=np.uint64).reshape(-1, 4)
        idx = rank
        for _ in range(self.top_levels):
            sib.append(level[idx ^ 1].reshape(1, 4))
            level = two_to_one(level[0::2], level[1::2])
            idx >>= 1
        return np.concatenate(sib, axis=0)


def reverse_bits_array(idx, bits):
    """vectorised reverse_bits (plonky2/src/util/mod.rs:30-38) of a uint64 numpy array"""
    idx = np.asarray(idx, dtype=np.uint64)
    out = np.zeros_like(idx)
    for b in range(bits):
        out |= ((idx >> np.uint64(b)) & np.uint64(1)) << np.uint64(bits - 1 - b)
    return out


def salt_leaf_slice(salts, lg_n, first, count):
    """salts: [salt_w][N] in natural LDE order (oracle.rs:119-123) -> rows [first, first+count) in LEAF order
    (reverse_index_bits_in_place, oracle.rs:84)."""
    src = reverse_bits_array(np.arange(first, first + count, dtype=np.uint64), lg_n).astype(np.int64)
    return np.ascontiguousarray(np.asarray(salts, dtype=np.uint64)[:, src])


class _DeviceView:
    """__cuda_array_interface__ over raw device memory owned by the engine (a shard's LDE rows)"""

    def __init__(self, ptr, n_elems):
        self.__cuda_array_interface__ = {"shape": (int(n_elems),), "typestr": "<i8", "data": (int(ptr), False), "version": 2}


class CudaShardEngine:
    """Per-rank compute through the C ABI (no CPU path)."""

    def __init__(self):
        # the engine's current context must be this rank's GPU: a rank that forgot pcs.init(local_rank) would otherwise
        # run on device 0 with another GPU's pointers
        try:
            import torch
            if torch.cuda.is_available():
                L = _ffi.lib()
                if L.pcs_device() < 0:
                    _ffi.check(L.pcs_init(torch.cuda.current_device(), None))
                if L.pcs_device() != torch.cuda.current_device():
                    raise RuntimeError(f"the engine is bound to GPU {L.pcs_device()} but torch's current device is "
                                       f"{torch.cuda.current_device()}: call plonky2_demo_b200.init(local_rank[, stream]) first")
        except ImportError:
            pass

    @staticmethod
    def _order_after_torch():
        """The collectives complete on torch's current stream, the engine enqueues on pcs_stream(): when the two
        differ (the caller did not hand torch's stream to pcs_init) make the engine's work wait for torch's."""
        import torch

        cur = torch.cuda.current_stream()
        if _ffi.lib().pcs_stream() != cur.cuda_stream:
            cur.synchronize()

    def intt_local(self, values):
        """values: torch CUDA int64 tensor [w_local][d], transformed in place (values -> coefficients)."""
        if values.shape[0] == 0:
            return
        self._order_after_torch()
        _ffi.check(_ffi.lib().pcs_ntt_dev(C.c_void_p(values.data_ptr()), values.shape[0], log2_strict(values.shape[1]), 1))
        import torch

        if _ffi.lib().pcs_stream() != torch.cuda.current_stream().cuda_stream:
            _ffi.check(_ffi.lib().pcs_synchronize())   # the all-gather that follows runs on torch's stream

    def commit_shard(self, coeffs, w, plan, rank, poly_ptrs=None):
        """coeffs: torch CUDA int64 tensor [>= w][d] (contiguous), or poly_ptrs: w raw device pointers (possibly
        into peer GPUs' memory).  Returns (handle, local cap [2^lch][4])."""
        L = _ffi.lib()
        self._order_after_torch()
        if poly_ptrs is not None:
            ptrs = (_ffi.u64p * w)(*[C.cast(C.c_void_p(int(a)), _ffi.u64p) for a in poly_ptrs])
        else:
            ptrs = _ffi.dev_ptr_array(coeffs.data_ptr(), w, coeffs.shape[1])
        cap = np.empty((plan.local_cap_len(), 4), dtype=np.uint64)
        h = C.c_void_p()
        _ffi.check(L.pcs_commit_shard_from_coeffs(ptrs, w, plan.lg_d, plan.rate_bits, plan.coset_first(rank), plan.lg_cosets,
                                                  plan.local_cap_height, None, 0, _ffi.PCS_DEVICE_PTRS, _ffi.ptr(cap), C.byref(h)))
        return h, cap

    def shard_begin(self, plan, rank, salt_w=0):
        h = C.c_void_p()
        self._order_after_torch()
        _ffi.check(_ffi.lib().pcs_shard_begin(plan.w, salt_w, plan.lg_d, plan.rate_bits, plan.coset_first(rank), plan.lg_cosets,
                                              plan.local_cap_height, C.byref(h)))
        return h

    # ---- the all-to-all partition: LDE rows computed by the ranks that hold the polynomials ----
    def lde_full(self, coeffs, rate_bits):
        """coeffs: torch CUDA int64 [w_loc][d] -> [w_loc][N] in leaf order (pcs_coset_lde_dev, all cosets)"""
        import torch

        w_loc, d = int(coeffs.shape[0]), int(coeffs.shape[1])
        out = torch.empty((w_loc, d << rate_bits), dtype=torch.int64, device=coeffs.device)
        if w_loc:
            self._order_after_torch()
            _ffi.check(_ffi.lib().pcs_coset_lde_dev(_ffi.dev_ptr_array(coeffs.data_ptr(), w_loc, d), w_loc, log2_strict(d), rate_bits, 7,
                                                    C.c_void_p(out.data_ptr())))
            self._order_torch_after_engine()
        return out

    @staticmethod
    def _order_torch_after_engine():
        import torch

        if _ffi.lib().pcs_stream() != torch.cuda.current_stream().cuda_stream:
            _ffi.check(_ffi.lib().pcs_synchronize())

    def rows_begin(self, width, salt_w, lg_n_local, local_cap_height):
        h = C.c_void_p()
        self._order_after_torch()
        _ffi.check(_ffi.lib().pcs_shard_begin_rows(width, salt_w, lg_n_local, local_cap_height, C.byref(h)))
        self._order_torch_after_engine()      # the buffer exists before torch / NCCL write into it
        return h

    def rows_buffer(self, handle, rows, n_local):
        """torch view [rows][n_local] of the shard's first `rows` LDE rows (written in place by the all-to-all)"""
        import torch

        base = _ffi.lib().pcs_batch_lde_dev(handle)
        return torch.as_tensor(_DeviceView(base, rows * n_local), device="cuda").view(rows, n_local)

    def set_rows(self, handle, row_first, rows, canonical):
        """rows: torch CUDA int64 [count][n_local], leaf order"""
        self._order_after_torch()
        count, n_local = int(rows.shape[0]), int(rows.shape[1])
        _ffi.check(_ffi.lib().pcs_shard_set_rows(handle, row_first, count, _ffi.dev_ptr_array(rows.data_ptr(), count, n_local),
                                                 1 if canonical else 0))

    def to_device(self, host_rows, like):
        import torch

        return torch.from_numpy(np.ascontiguousarray(host_rows).view(np.int64)).to(like.device)

    def shard_extend(self, handle, poly_first, coeffs, count):
        """coeffs: torch CUDA int64 tensor [>= count][d] holding polynomials poly_first .. poly_first+count-1"""
        self._order_after_torch()
        ptrs = _ffi.dev_ptr_array(coeffs.data_ptr(), count, coeffs.shape[1])
        _ffi.check(_ffi.lib().pcs_shard_extend(handle, poly_first, count, ptrs))

    def shard_finish(self, handle, plan):
        cap = np.empty((plan.local_cap_len(), 4), dtype=np.uint64)
        self._order_after_torch()
        _ffi.check(_ffi.lib().pcs_shard_finish(handle, _ffi.ptr(cap)))
        return cap

    def shard_finish_dev(self, handle, plan):
        """pcs_shard_finish without the host round trip: the local cap stays on the device (torch view) for the cap all-gather"""
        import torch

        self._order_after_torch()
        _ffi.check(_ffi.lib().pcs_shard_finish(handle, None))
        self._order_torch_after_engine()
        base = _ffi.lib().pcs_batch_cap_dev(handle)
        return torch.as_tensor(_DeviceView(base, plan.local_cap_len() * 4), device="cuda").view(-1, 4)

    def two_to_one(self, left, right):
        from .hashing import PoseidonHash

        return PoseidonHash.two_to_one_batch(np.ascontiguousarray(left), np.ascontiguousarray(right))

    def get_rows(self, handle, local_indices, width):
        idx = np.ascontiguousarray(np.asarray(local_indices, dtype=np.uint64))
        out = np.empty((idx.shape[0], width), dtype=np.uint64)
        _ffi.check(_ffi.lib().pcs_batch_get_rows(handle, _ffi.ptr(idx), idx.shape[0], _ffi.ptr(out)))
        return out

    def prove(self, handle, local_index, n_siblings):
        sib = np.empty((n_siblings, 4), dtype=np.uint64)
        _ffi.check(_ffi.lib().pcs_batch_prove(handle, int(local_index), _ffi.ptr(sib)))
        return sib

    def digests(self, handle, n_digests):
        out = np.empty((n_digests, 4), dtype=np.uint64)
        _ffi.check(_ffi.lib().pcs_batch_digests(handle, _ffi.ptr(out)))
        return out

    def free(self, handle):
        _ffi.lib().pcs_batch_free(handle)


class ShardedPolynomialBatch:
    """PolynomialBatch (oracle.rs:30-37) whose leaves / digests are spread over the ranks of a process group.

    `cap` is replicated on every rank and equals the reference's `merkle_tree.cap` bit for bit;
    rank r holds leaves [r*N/G, (r+1)*N/G) and the matching slice of `digests`.
    """

    def __init__(self):
        self._h = None

    _side_streams = {}

    @classmethod
    def _side_stream(cls, dev):
        """one copy / collective stream per device, kept for the life of the process"""
        import torch

        k = str(dev)
        if k not in cls._side_streams:
            cls._side_streams[k] = torch.cuda.Stream(device=dev)
        return cls._side_streams[k]

    # ---- peer-memory exchange ---------------------------------------------------------------------------
    _symm_cache = {}

    @classmethod
    def _symm_block(cls, group, w_max, d, device):
        """One symmetric-memory buffer [w_max][d] per (group, shape): every rank's buffer is mapped into every
        process (torch.distributed._symmetric_memory: cuMem + NVLink peer mapping), rendezvous happens once."""
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem

        key = (id(group), w_max, d, str(device))
        if key not in cls._symm_cache:
            buf = symm_mem.empty((w_max, d), dtype=torch.int64, device=device)
            hdl = symm_mem.rendezvous(buf, group if group is not None else dist.group.WORLD)
            cls._symm_cache[key] = (buf, hdl)
        return cls._symm_cache[key]

    # ---- constructors (collective: every rank of `group` calls them) ---------------------------------
    @classmethod
    def from_coeffs(cls, local_coeffs, n_polys, rate_bits, cap_height, group=None, engine=None, partitioned=True,
                    exchange="auto", chunks=1, blinding=False, salts=None, gather_coeffs=False):
        """oracle.rs:68-98 over a process group.

        partitioned=True : `local_coeffs` is this rank's block [poly_range(rank)][d] of the n_polys polynomials
                           (torch int64 tensor on the rank's device, bit pattern = u64);
        partitioned=False: every rank already holds all [n_polys][d] coefficients.
        chunks           : > 1 = streaming exchange (needs partitioned=True): the polynomials are cut into `chunks`
                           groups, each block-distributed over the ranks (ShardPlan.poly_ranges); `local_coeffs`
                           holds this rank's rows in chunk order and may live in PINNED HOST memory.  Chunk c+1 is
                           copied / all-gathered on a side stream while the LDE of chunk c runs.
        exchange         : how the data a rank hashes reaches it --
            "allgather": coset partition; NCCL all-gather of the coefficients into a local [W][d] matrix, then the LDE;
            "peer"     : coset partition; NO collective: every block sits in symmetric memory and the first NTT pass of each
                         rank loads the other ranks' coefficients straight from their HBM over NVLink (fused exchange +
                         compute, transfer hidden behind the butterflies);
            "alltoall" : polynomial partition (the north-star's split): every rank extends its own polynomials over all
                         cosets and ONE NCCL all-to-all carries the LDE rows to the rank that hashes their leaf range.
                         Works for any power-of-two world (also > 2^rate_bits) and needs partitioned=True, chunks=1;
            "auto"     : "allgather" when world <= 2^rate_bits, else "alltoall".  Measured on 8 x B200 (135 x 2^20, one coset
                         per rank, profiles/r01_scaling_v5.md): the peer-reading LDE costs 3.41 ms against 2.72 ms + ~1.4 ms
                         of all-gather, but the two cross-rank barriers and the symmetric-buffer copy it needs cost more
                         than that on the host side across PROCESSES (21.9 ms vs 18.3 ms per commitment), so the
                         collective stays the default here; inside one process (pcs_multi_*) the peer form is the default.
        blinding / salts : SALT_SIZE = 4 extra leaf columns (oracle.rs:26,119-123).  `salts` = [4][N] in natural LDE order,
                           the same array on every rank (reproducible commitments; a rank uses only its leaf range); with
                           blinding=True and salts=None every rank draws the salt values of its own leaves from the OS.
        gather_coeffs    : "alltoall" only: also all-gather the coefficients so that every rank can serve the opening
                           proof (poly_device_ptrs); the coset partition replicates them anyway.
        """
        import torch
        import torch.distributed as dist

        world = dist.get_world_size(group) if dist.is_initialized() else 1
        rank = dist.get_rank(group) if dist.is_initialized() else 0
        engine = engine or CudaShardEngine()
        d = int(local_coeffs.shape[1])
        plan = ShardPlan(n_polys, log2_strict(d), rate_bits, cap_height, world, chunks if (partitioned and world > 1) else 1)
        if exchange == "auto":
            exchange = "allgather" if plan.coset_partition else "alltoall"
        if exchange not in ("allgather", "peer", "alltoall"):
            raise ValueError(f"unknown exchange {exchange!r}")
        # ---- blinding: this rank's slice [4][local_leaves] of the salt columns, in leaf order ----
        salt_rows = None
        if salts is not None or blinding:
            lo_leaf, hi_leaf = plan.leaf_range(rank)
            if salts is not None:
                sal = np.asarray(salts, dtype=np.uint64)
                if sal.shape != (4, plan.n_leaves):
                    raise ValueError(f"salts must be [4][{plan.n_leaves}] (SALT_SIZE columns of N values, oracle.rs:26), got {sal.shape}")
                salt_rows = salt_leaf_slice(sal, plan.lg_d + rate_bits, lo_leaf, hi_leaf - lo_leaf)
            else:
                # F::rand_vec (types.rs:33-35, OsRng): every rank draws the values of its own leaves
                rng = np.random.default_rng(int.from_bytes(os.urandom(16), "little"))
                salt_rows = rng.integers(0, 0xFFFFFFFF00000001, size=(4, hi_leaf - lo_leaf), dtype=np.uint64)
        if exchange == "alltoall" and world > 1:
            if not partitioned or plan.chunks > 1:
                raise ValueError("exchange='alltoall' needs partitioned=True and chunks=1")
            return cls._from_coeffs_alltoall(local_coeffs, plan, rank, world, group, engine, salt_rows, gather_coeffs)
        if not plan.coset_partition:
            plan.coset_first(rank)    # raises: the coset partition needs world <= 2^rate_bits
        if plan.chunks > 1:
            return cls._from_coeffs_streaming(local_coeffs, plan, rank, world, group, engine, salt_rows)
        if salt_rows is not None:
            # one-shot coset partition with salts: go through begin / extend / set_rows / finish
            return cls._from_coeffs_streaming(local_coeffs, plan, rank, world, group, engine, salt_rows, one_shot_exchange=exchange,
                                              partitioned=partitioned)
        poly_ptrs = None
        if partitioned and world > 1 and exchange == "peer":
            lo, hi = plan.poly_range(rank)
            if local_coeffs.shape[0] != hi - lo:
                raise ValueError(f"rank {rank} must hold polynomials [{lo}, {hi}), got {local_coeffs.shape[0]} rows")
            buf, hdl = cls._symm_block(group, plan.w_max, d, local_coeffs.device)
            buf[: hi - lo].copy_(local_coeffs)
            hdl.barrier()                      # every rank's block is in place (stream-ordered)
            poly_ptrs = []
            for q in range(world):
                qlo, qhi = plan.poly_range(q)
                poly_ptrs += [int(hdl.buffer_ptrs[q]) + 8 * d * i for i in range(qhi - qlo)]
            full = local_coeffs
        elif partitioned and world > 1:
            lo, hi = plan.poly_range(rank)
            if local_coeffs.shape[0] != hi - lo:
                raise ValueError(f"rank {rank} must hold polynomials [{lo}, {hi}), got {local_coeffs.shape[0]} rows")
            # exchange step 1: all-gather of the coefficient blocks (padded to w_max rows per rank; the
            # block distribution leaves padding only after the last polynomial, so rows [0, w) stay contiguous)
            full = torch.empty((world * plan.w_max, d), dtype=torch.int64, device=local_coeffs.device)
            if hi - lo == plan.w_max:
                mine = local_coeffs.contiguous()
            else:
                mine = torch.zeros((plan.w_max, d), dtype=torch.int64, device=local_coeffs.device)
                mine[: hi - lo] = local_coeffs
            dist.all_gather_into_tensor(full, mine, group=group)
        else:
            full = local_coeffs.contiguous()
        self = cls()
        self.plan, self.rank, self.world, self.group, self.engine = plan, rank, world, group, engine
        self.degree_log, self.rate_bits, self.blinding, self.cap_height = plan.lg_d, rate_bits, False, cap_height
        self.n_polys, self.salt_w = n_polys, 0
        self._coeffs = full          # PolynomialBatch.polynomials (replicated; only the local block with exchange="peer")
        if poly_ptrs is not None:
            self._h, local_cap = engine.commit_shard(None, n_polys, plan, rank, poly_ptrs=poly_ptrs)
            hdl.barrier()                      # all ranks are done reading before any block is overwritten
        else:
            self._h, local_cap = engine.commit_shard(full, n_polys, plan, rank)
        self.exchange = exchange if (partitioned and world > 1) else "none"
        # exchange step 2: all-gather of the local caps
        if world > 1:
            mine = torch.from_numpy(local_cap.view(np.int64)).to(full.device)
            allc = torch.empty((world * mine.shape[0], 4), dtype=torch.int64, device=full.device)
            dist.all_gather_into_tensor(allc, mine, group=group)
            caps = allc.cpu().numpy().view(np.uint64).reshape(world, -1, 4)
        else:
            caps = local_cap[None]
        self._local_caps = caps
        self.cap = plan.assemble_cap(caps, engine.two_to_one)
        return self

    @classmethod
    def _from_coeffs_streaming(cls, local_coeffs, plan, rank, world, group, engine, salt_rows=None, one_shot_exchange=None,
                               partitioned=True):
        """chunked exchange overlapped with the LDE: side stream = (H2D of chunk c) -> all-gather of chunk c;
        main stream = LDE of chunk c as soon as its gather has landed; then leaf hashing and the subtrees."""
        import torch
        import torch.distributed as dist

        d = int(local_coeffs.shape[1])
        ranges = plan.poly_ranges(rank)
        if not partitioned or world == 1:
            # every rank holds all coefficients already: one group, nothing to gather
            self = cls()
            self.plan, self.rank, self.world, self.group, self.engine = plan, rank, world, group, engine
            self.degree_log, self.rate_bits, self.blinding, self.cap_height = plan.lg_d, plan.rate_bits, salt_rows is not None, plan.cap_height
            self.n_polys, self.salt_w = plan.w, 0 if salt_rows is None else 4
            full = local_coeffs.contiguous()
            self._h = engine.shard_begin(plan, rank, self.salt_w)
            engine.shard_extend(self._h, 0, full, plan.w)
            if salt_rows is not None:
                engine.set_rows(self._h, plan.w, engine.to_device(salt_rows, full), False)
            local_cap = engine.shard_finish(self._h, plan)
            self._coeffs = full
            return cls._finish_caps(self, local_cap, full.device, world, group, engine, "none")
        if local_coeffs.shape[0] != sum(hi - lo for lo, hi in ranges):
            raise ValueError(f"rank {rank} must hold polynomials {ranges} ({sum(hi - lo for lo, hi in ranges)} rows), "
                             f"got {local_coeffs.shape[0]}")
        cuda = torch.cuda.is_available() and dist.get_backend(group) == "nccl"
        dev = torch.device("cuda", torch.cuda.current_device()) if cuda else local_coeffs.device
        main = torch.cuda.current_stream() if cuda else None
        side = cls._side_stream(dev) if cuda else None
        if cuda:
            side.wait_stream(main)
        gathered, events, row = [], [], 0
        for c in range(plan.chunks):
            lo, hi = ranges[c]
            m = plan.chunk_rows(c)
            # buffers come from the MAIN stream's allocator pool (reused from commit to commit) and are handed to the
            # side stream with record_stream
            mine = torch.zeros((m, d), dtype=torch.int64, device=dev) if hi - lo < m else torch.empty((m, d), dtype=torch.int64, device=dev)
            g = torch.empty((world * m, d), dtype=torch.int64, device=dev)
            if cuda:
                side.wait_stream(main)       # the zero fill above runs on the main stream
                mine.record_stream(side)
                g.record_stream(side)
            with (torch.cuda.stream(side) if cuda else _null_ctx()):
                mine[: hi - lo].copy_(local_coeffs[row : row + hi - lo], non_blocking=True)
                dist.all_gather_into_tensor(g, mine, group=group)
                ev = torch.cuda.Event() if cuda else None
                if cuda:
                    ev.record(side)
            row += hi - lo
            gathered.append((g, mine))
            events.append(ev)
        self = cls()
        self.plan, self.rank, self.world, self.group, self.engine = plan, rank, world, group, engine
        self.degree_log, self.rate_bits, self.blinding, self.cap_height = plan.lg_d, plan.rate_bits, salt_rows is not None, plan.cap_height
        self.n_polys, self.salt_w = plan.w, 0 if salt_rows is None else 4
        self._h = engine.shard_begin(plan, rank, self.salt_w)
        for c in range(plan.chunks):
            clo, chi = plan.chunk_range(c)
            if cuda:
                main.wait_event(events[c])
            engine.shard_extend(self._h, clo, gathered[c][0], chi - clo)
        if salt_rows is not None:
            engine.set_rows(self._h, plan.w, engine.to_device(salt_rows, gathered[0][0]), False)
        if cuda and hasattr(engine, "shard_finish_dev"):
            local_cap = engine.shard_finish_dev(self._h, plan)      # device tensor: gathered without a host round trip
        else:
            local_cap = engine.shard_finish(self._h, plan)
        self._coeffs = gathered          # keeps the gathered chunks alive ([chunk][world*m][d]); polynomials of chunk c = rows [0, len_c)
        self._dev = dev
        return cls._finish_caps(self, local_cap, dev, world, group, engine,
                                f"allgather, {plan.chunks} chunk{'s' if plan.chunks > 1 else ''} overlapped with the LDE")

    @staticmethod
    def _finish_caps(self, local_cap, dev, world, group, engine, exchange):
        """exchange step 2: all-gather of the local caps, top levels on every rank"""
        import torch
        import torch.distributed as dist

        if world > 1:
            mine = local_cap if isinstance(local_cap, torch.Tensor) else torch.from_numpy(local_cap.view(np.int64)).to(dev)
            allc = torch.empty((world * mine.shape[0], 4), dtype=torch.int64, device=dev)
            dist.all_gather_into_tensor(allc, mine, group=group)
            caps = allc.cpu().numpy().view(np.uint64).reshape(world, -1, 4)
        else:
            if isinstance(local_cap, torch.Tensor):
                local_cap = local_cap.cpu().numpy().view(np.uint64)
            caps = local_cap[None]
        self._local_caps = caps
        self.cap = self.plan.assemble_cap(caps, engine.two_to_one)
        self.exchange = exchange
        return self

    @classmethod
    def _from_coeffs_alltoall(cls, local_coeffs, plan, rank, world, group, engine, salt_rows, gather_coeffs):
        """The north-star's split: polynomial-partitioned LDE, then an all-to-all of LDE rows into leaf-range shards.
        Rank r extends its block [poly_range(r)][d] over ALL cosets (leaf order), sends rows [w_r][q N/G : (q+1) N/G] to rank
        q, and receives every polynomial's slice of its own leaf range straight into its shard's leaf buffer [W][N/G]
        (poly-major: the layout the hash kernel reads).  Then leaf hashing, the local subtree(s), cap all-gather."""
        import torch
        import torch.distributed as dist

        lo, hi = plan.poly_range(rank)
        w_loc, n_loc = hi - lo, plan.local_leaves
        if local_coeffs.shape[0] != w_loc:
            raise ValueError(f"rank {rank} must hold polynomials [{lo}, {hi}), got {local_coeffs.shape[0]} rows")
        dev = local_coeffs.device
        self = cls()
        self.plan, self.rank, self.world, self.group, self.engine = plan, rank, world, group, engine
        self.degree_log, self.rate_bits, self.blinding, self.cap_height = plan.lg_d, plan.rate_bits, salt_rows is not None, plan.cap_height
        self.n_polys, self.salt_w = plan.w, 0 if salt_rows is None else 4
        lde = engine.lde_full(local_coeffs.contiguous(), plan.rate_bits)                       # [w_loc][N], leaf order
        send = lde.view(w_loc, world, n_loc).permute(1, 0, 2).contiguous()                     # [dest][w_loc][n_loc]
        del lde
        self._h = engine.rows_begin(plan.w + self.salt_w, self.salt_w, plan.lg_local, plan.local_cap_height)
        recv = engine.rows_buffer(self._h, plan.w, n_loc)                                      # the shard's own rows [W][n_loc]
        in_splits = [w_loc * n_loc] * world
        out_splits = [(plan.poly_range(q)[1] - plan.poly_range(q)[0]) * n_loc for q in range(world)]
        dist.all_to_all_single(recv.view(-1), send.view(-1), out_splits, in_splits, group=group)
        del send
        if salt_rows is not None:
            engine.set_rows(self._h, plan.w, engine.to_device(salt_rows, local_coeffs), False)
        local_cap = engine.shard_finish(self._h, plan)
        if gather_coeffs:
            full = torch.empty((world * plan.w_max, 1 << plan.lg_d), dtype=torch.int64, device=dev)
            mine = local_coeffs.contiguous()
            if w_loc < plan.w_max:
                mine = torch.zeros((plan.w_max, 1 << plan.lg_d), dtype=torch.int64, device=dev)
                mine[:w_loc] = local_coeffs
            dist.all_gather_into_tensor(full, mine, group=group)
            self._coeffs = full
        else:
            self._coeffs = local_coeffs
        self._dev = dev
        return cls._finish_caps(self, local_cap, dev, world, group, engine, "alltoall of LDE rows (polynomial-partitioned LDE)")

    @classmethod
    def from_values(cls, local_values, n_polys, rate_bits, cap_height, group=None, engine=None):
        """oracle.rs:43-65 over a process group: each rank IFFTs its own block of the polynomials
        (values are overwritten by coefficients), then from_coeffs."""
        engine = engine or CudaShardEngine()
        engine.intt_local(local_values)
        return cls.from_coeffs(local_values, n_polys, rate_bits, cap_height, group, engine, partitioned=True)

    # ---- accessors ---------------------------------------------------------------------------------
    def _device(self):
        d = getattr(self, "_dev", None)
        return d if d is not None else self._coeffs.device

    @property
    def n_local_digests(self):
        return 2 * (self.plan.local_leaves - self.plan.local_cap_len())

    def local_digests(self):
        """This rank's contiguous slice of the reference's `digests` (world <= 2^cap_height)."""
        return self.engine.digests(self._h, self.n_local_digests)

    def get_rows(self, leaf_indices):
        """merkle_tree.leaves[i] for global leaf indices; every rank gets every row (collective)."""
        import torch
        import torch.distributed as dist

        idx = np.asarray(list(leaf_indices), dtype=np.int64)
        lo, hi = self.plan.leaf_range(self.rank)
        mine = (idx >= lo) & (idx < hi)
        width = self.n_polys + getattr(self, "salt_w", 0)
        rows = np.zeros((idx.shape[0], width), dtype=np.uint64)
        if mine.any():
            rows[mine] = self.engine.get_rows(self._h, idx[mine] - lo, width)
        if self.world > 1:
            t = torch.from_numpy(rows.view(np.int64)).to(self._device())
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)   # disjoint owners: sum == select
            rows = t.cpu().numpy().view(np.uint64)
        return rows

    def get_lde_values(self, index, step):
        """oracle.rs:128-133"""
        leaf = reverse_bits(index * step, self.degree_log + self.rate_bits)
        return self.get_rows([leaf])[0][: self.n_polys]      # minus the salt columns (oracle.rs:131-132)

    def prove(self, leaf_index):
        """MerkleTree::prove (merkle_tree.rs:173-207) for a global leaf index (collective)."""
        import torch
        import torch.distributed as dist

        plan = self.plan
        owner = plan.owner_of_leaf(leaf_index)
        n_local = log2_strict(plan.local_leaves) - plan.local_cap_height
        n_total = n_local + plan.top_levels
        sib = np.zeros((n_total, 4), dtype=np.uint64)
        if owner == self.rank:
            local = self.engine.prove(self._h, leaf_index - plan.leaf_range(owner)[0], n_local) if n_local else np.zeros((0, 4), np.uint64)
            sib[:] = plan.extend_proof(owner, local, self._local_caps.reshape(self.world, -1, 4)[:, 0], self.engine.two_to_one) \
                if plan.top_levels else local
        if self.world > 1:
            t = torch.from_numpy(sib.view(np.int64)).to(self._device())
            dist.broadcast(t, src=dist.get_global_rank(self.group, owner) if self.group is not None else owner, group=self.group)
            sib = t.cpu().numpy().view(np.uint64)
        from .hashing import MerkleProof

        return MerkleProof(sib)

    def prove_many(self, leaf_indices):
        """MerkleTree::prove for several global leaf indices in ONE collective (the FRI query phase, fri/prover.rs:162-216):
        every rank fills in the paths of the leaves it owns, a sum all-reduce (disjoint owners) replicates them."""
        import torch
        import torch.distributed as dist
        from .hashing import MerkleProof

        plan = self.plan
        idx = [int(i) for i in leaf_indices]
        n_local = log2_strict(plan.local_leaves) - plan.local_cap_height
        n_total = n_local + plan.top_levels
        sib = np.zeros((len(idx), n_total, 4), dtype=np.uint64)
        for k, leaf in enumerate(idx):
            owner = plan.owner_of_leaf(leaf)
            if owner != self.rank:
                continue
            local = self.engine.prove(self._h, leaf - plan.leaf_range(owner)[0], n_local) if n_local else np.zeros((0, 4), np.uint64)
            sib[k] = plan.extend_proof(owner, local, self._local_caps.reshape(self.world, -1, 4)[:, 0], self.engine.two_to_one) \
                if plan.top_levels else local
        if self.world > 1 and sib.size:
            t = torch.from_numpy(sib.view(np.int64)).to(self._device())
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
            sib = t.cpu().numpy().view(np.uint64)
        return [MerkleProof(s) for s in sib]

    def poly_device_ptrs(self):
        """Device address of every polynomial's coefficient vector on THIS rank (PolynomialBatch.polynomials, replicated by
        the coefficient exchange): what the opening proof reads (fri_prover.eval_commitment / final_poly)."""
        d = 1 << self.degree_log
        if isinstance(self._coeffs, list):          # streaming exchange: chunk c holds its polynomials in rows [0, len_c)
            ptrs = []
            for c, (g, _mine) in enumerate(self._coeffs):
                lo, hi = self.plan.chunk_range(c)
                ptrs += [g.data_ptr() + 8 * d * i for i in range(hi - lo)]
            return ptrs
        if self._coeffs is None or self._coeffs.shape[0] < self.n_polys:
            raise ValueError("this rank does not hold every polynomial's coefficients (exchange='peer' keeps only its block)")
        return [self._coeffs.data_ptr() + 8 * d * j for j in range(self.n_polys)]

    def free(self):
        if self._h is not None:
            self.engine.free(self._h)
            self._h = None
        self._coeffs = None   # release the gathered coefficient matrix as well

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass
