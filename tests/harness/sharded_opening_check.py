#!/usr/bin/env python3
"""Verification harness (test infrastructure: imports the CPU oracle as the checker).

Opening proof over commitments SHARDED across GPUs (SURVEY 8e + 8f): run under torchrun, one rank per GPU,

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tests/harness/sharded_opening_check.py

Every rank commits its block of each oracle's polynomials (streaming coefficient exchange), then all ranks run
fri_prover.prove_openings collectively; rank 0 compares the proof field by field with the CPU oracle's prover and has the
restated verifier accept it; all ranks check that they hold the same proof."""
import hashlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    import oracle
    import plonky2_demo_b200 as p
    from helpers import seeded_polys
    from oracle import fri_ref as fr
    from plonky2_demo_b200 import fri_prover as fp
    from plonky2_demo_b200.serialization import fri_proof_to_bytes
    from plonky2_demo_b200.sharded import ShardedPolynomialBatch, ShardPlan

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    p.init(local)
    lg_d = int(os.environ.get("CHECK_LG_D", "12"))
    r, cap, widths, chunks = 3, 4, [20, 33, 6, 4], 2
    cfg = p.FriConfig(r, cap, 8, p.FriReductionStrategy.ConstantArityBits(4, 5), 7)
    params = cfg.fri_params(lg_d, False)
    coeffs = [seeded_polys(w, 1 << lg_d, base_seed=4000 + k) for k, w in enumerate(widths)]
    sharded = []
    for c in coeffs:
        plan = ShardPlan(c.shape[0], lg_d, r, cap, world, chunks)
        mine = np.ascontiguousarray(c[plan.local_polys(rank)])
        t = torch.from_numpy(mine.view(np.int64)).to(dev)
        sharded.append(ShardedPolynomialBatch.from_coeffs(t, c.shape[0], r, cap, partitioned=True, chunks=chunks))
    zeta = (0x1234567, 0x89ABCDE)
    zeta_next = fr.ext_mul((fr.primitive_root_of_unity(lg_d), 0), zeta)
    all_polys = [(k, j) for k, w in enumerate(widths) for j in range(w)]
    batches = [(zeta, all_polys), (zeta_next, [(2, 0), (2, 1)])]
    inst = fp.FriInstanceInfo([fp.FriOracleInfo(w, False) for w in widths],
                              [fp.FriBatchInfo(pt, [fp.FriPolynomialInfo(o, j) for o, j in polys]) for pt, polys in batches])
    at = {pt: [fp.eval_commitment(pt, b) for b in sharded] for pt, _ in batches}
    openings = [[tuple(int(x) for x in at[pt][o][j]) for o, j in polys] for pt, polys in batches]

    def transcript(cls):
        ch = cls()
        for b in sharded:
            ch.observe_cap(b.cap)
        for vals in openings:
            ch.observe_extension_elements(vals)
        return ch

    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    proof = fp.prove_openings(inst, sharded, transcript(fp.Challenger), params)
    torch.cuda.synchronize()
    ms = 1e3 * (time.perf_counter() - t0)
    digest = hashlib.sha256(fri_proof_to_bytes(proof)).digest()
    t = torch.tensor(list(digest), dtype=torch.uint8, device=dev)
    all_d = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(all_d, t)
    same_on_all_ranks = all(bool((x == all_d[0]).all()) for x in all_d)
    out = {"world": world, "lg_d": lg_d, "widths": widths, "same_proof_on_all_ranks": same_on_all_ranks, "prove_openings_ms": ms}
    if rank == 0:
        cpu = []
        for c in coeffs:
            o = oracle.commit_from_coeffs(c, r, cap)
            o["coeffs"], o["cap_height"] = c, cap
            cpu.append(o)
        out["caps_equal"] = all(np.array_equal(b.cap, o["cap"]) for b, o in zip(sharded, cpu))
        out["openings_equal"] = all(
            v == tuple(int(x) for x in fr.eval_base_polys_ext(cpu[o]["coeffs"][j:j + 1], pt)[0])
            for (pt, polys), vals in zip(batches, openings) for (o, j), v in zip(polys, vals))
        och = transcript(fr.Challenger)
        vch = och.clone()
        want = fr.prove_openings(cpu, batches, och, r, cap, params.reduction_arity_bits, cfg.proof_of_work_bits, cfg.num_query_rounds)
        eq = (all(np.array_equal(c.hashes, wc) for c, wc in zip(proof.commit_phase_merkle_caps, want["commit_phase_merkle_caps"]))
              and np.array_equal(proof.final_poly, want["final_poly"]) and proof.pow_witness == want["pow_witness"]
              and proof.fri_query_indices == want["_indices"])
        for rr, wr in zip(proof.query_round_proofs, want["query_round_proofs"]):
            for (ev, mp), (wev, wmp) in zip(rr.initial_trees_proof.evals_proofs, wr["initial_trees_proof"]):
                eq = eq and np.array_equal(ev, wev) and np.array_equal(np.asarray(mp.siblings).reshape(-1, 4), wmp)
            for s, ws in zip(rr.steps, wr["steps"]):
                eq = eq and np.array_equal(s.evals, ws["evals"]) and np.array_equal(np.asarray(s.merkle_proof.siblings).reshape(-1, 4), ws["merkle_proof"])
        out["proof_equal_to_cpu_oracle"] = bool(eq)
        as_oracle = {
            "commit_phase_merkle_caps": [c.hashes for c in proof.commit_phase_merkle_caps], "final_poly": proof.final_poly,
            "pow_witness": proof.pow_witness,
            "query_round_proofs": [
                {"initial_trees_proof": [(ev, np.asarray(mp.siblings).reshape(-1, 4)) for ev, mp in rr.initial_trees_proof.evals_proofs],
                 "steps": [{"evals": s.evals, "merkle_proof": np.asarray(s.merkle_proof.siblings).reshape(-1, 4)} for s in rr.steps]}
                for rr in proof.query_round_proofs]}
        out["verifier_accepts"] = bool(fr.verify_fri_proof(batches, openings, vch, [o["cap"] for o in cpu], as_oracle, r, cap,
                                                           params.reduction_arity_bits, cfg.proof_of_work_bits,
                                                           cfg.num_query_rounds, lg_d))
        print(json.dumps(out))
        assert out["caps_equal"] and out["openings_equal"] and eq and out["verifier_accepts"] and same_on_all_ranks
    for b in sharded:
        b.free()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
