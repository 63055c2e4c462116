#!/usr/bin/env python3
"""Verification harness (test infrastructure: imports the CPU oracle as the checker).

The sharded commitment over a REAL NCCL world (one rank per GPU) against the CPU oracle, every exchange mode:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tests/harness/nccl_commit_check.py

tests/test_gpu_nccl.py runs it under pytest (skipped on a 1-GPU box).  Every rank checks the replicated cap, rows and Merkle
paths of sampled leaves (served across ranks), its slice of the digests, and get_lde_values; exit code != 0 on any mismatch."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    import oracle
    import plonky2_demo_b200 as p
    from helpers import brev, canon, seeded_polys
    from plonky2_demo_b200.sharded import ShardedPolynomialBatch, ShardPlan

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    p.init(local, stream.cuda_stream)
    lg_w = world.bit_length() - 1
    # (w, lg_d, rate_bits, cap_height, mode, chunks, blinding)
    cases = [
        (135, 10, 3, 4, "allgather", 1, False), (135, 10, 3, 4, "allgather", 3, False), (20, 12, 3, 2, "peer", 1, False),
        (135, 10, 3, 4, "alltoall", 1, False), (9, 7, 3, 0, "alltoall", 1, True), (33, 11, 3, 4, "allgather", 2, True),
        (7, 9, max(lg_w - 1, 0), 1, "auto", 1, False),        # more ranks than coset blocks: auto -> all-to-all, sub-coset leaf ranges
        (64, 13, 3, 4, "from_values", 1, False), (17, 14, 3, 4, "host_streaming", 4, False),
    ]
    results = []
    ok_all = True
    for (w, lg_d, r, cap, mode, chunks, blind) in cases:
        d, n = 1 << lg_d, 1 << (lg_d + r)
        coeffs = seeded_polys(w, d, base_seed=0xC0DE + w + lg_d)
        coeffs[0, :2] = [0xFFFFFFFF00000001 + 3, (1 << 64) - 1]          # non-canonical inputs
        salts = seeded_polys(4, n, base_seed=0x5A17 + w) if blind else None
        ref = oracle.commit_from_coeffs(coeffs, r, cap, salts=salts)
        streaming = chunks if mode in ("allgather", "host_streaming") else 1
        plan = ShardPlan(w, lg_d, r, cap, world, streaming)
        mine = np.ascontiguousarray(coeffs[plan.local_polys(rank)]) if plan.local_polys(rank) else np.empty((0, d), np.uint64)
        if mode == "from_values":
            vals = oracle.fft(coeffs)
            t = torch.from_numpy(np.ascontiguousarray(vals[plan.local_polys(rank)]).view(np.int64)).to(dev)
            b = ShardedPolynomialBatch.from_values(t, w, r, cap)
        elif mode == "host_streaming":
            t = torch.from_numpy(mine.view(np.int64)).pin_memory()        # pinned HOST block: chunk c+1 crosses PCIe under the LDE of chunk c
            b = ShardedPolynomialBatch.from_coeffs(t, w, r, cap, partitioned=True, chunks=chunks)
        else:
            t = torch.from_numpy(mine.view(np.int64)).to(dev)
            b = ShardedPolynomialBatch.from_coeffs(t, w, r, cap, partitioned=True, exchange=mode, chunks=streaming, salts=salts)
        ok = bool(np.array_equal(b.cap, ref["cap"]))
        rng = np.random.default_rng(5)
        leaves = sorted(set([0, plan.local_leaves - 1, plan.local_leaves % n, n - 1] + [int(x) for x in rng.integers(0, n, size=6)]))
        ok &= bool(np.array_equal(b.get_rows(leaves), canon(ref["leaves"][leaves])))
        for leaf, pr in zip(leaves, b.prove_many(leaves)):
            ok &= bool(np.array_equal(np.asarray(pr.siblings).reshape(-1, 4), oracle.merkle_prove(ref["digests"], n, cap, leaf).reshape(-1, 4)))
        ok &= bool(np.array_equal(np.asarray(b.prove(leaves[1]).siblings).reshape(-1, 4), oracle.merkle_prove(ref["digests"], n, cap, leaves[1]).reshape(-1, 4)))
        ok &= bool(np.array_equal(b.get_lde_values(5 % n, 1), canon(ref["leaves"][brev(5 % n, lg_d + r)][:w])))
        if plan.top_levels == 0:
            per = ref["digests"].shape[0] // world
            ok &= bool(np.array_equal(b.local_digests(), ref["digests"][rank * per:(rank + 1) * per]))
        flag = torch.tensor([int(ok)], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        results.append({"case": [w, lg_d, r, cap, mode, chunks, blind], "exchange": b.exchange, "ok_on_all_ranks": bool(flag.item())})
        ok_all &= bool(flag.item())
        b.free()
    if rank == 0:
        print(json.dumps({"world": world, "all_ok": ok_all, "cases": results}))
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok_all else 1)


if __name__ == "__main__":
    main()
