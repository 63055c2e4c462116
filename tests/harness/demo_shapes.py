#!/usr/bin/env python3
"""Verification harness (test infrastructure: imports the CPU oracle as the checker).

BASELINE.json configs[0]/[1]: the commitments of the matmul demo (plonky2/src/bin/matrix_mul.rs) replayed by SHAPE.

The Rust prover cannot be built in this image, so "prove ms" itself is not measurable here; what this engine replaces
inside build()+prove() are the PolynomialBatch commits (circuit_builder.rs:1021; plonk/prover.rs:145,212,260) and the
FRI commit-phase trees (fri/prover.rs:81-87).  For m = 2 (degree 2^3) and m = 64 (degree 2^15) this script commits
seeded polynomials of exactly those shapes through the C ABI with HOST buffers (pinned-free numpy arrays, like the
Rust Vecs), checks caps against the CPU oracle and prints GPU vs CPU-port wall times (ms, median of 5).

    python tests/harness/demo_shapes.py            # needs a B200
"""
import ctypes as C
import json
import os
import statistics
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

import oracle  # noqa: E402
import plonky2_demo_b200 as pcs  # noqa: E402
from helpers import seeded_polys, splitmix64_stream  # noqa: E402

# (name, polys, from_values) -- SURVEY 8a: build: ~84 constants+sigmas; prove: 135 wires, 20 Z/partial products (values), 16 quotient chunks (coeffs)
COMMITS = [("constants_sigmas", 84, True), ("wires", 135, True), ("zs_partial_products", 20, True), ("quotient_chunks", 16, False)]


def med(f, reps=5):
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        f()
        ts.append(1e3 * (time.perf_counter() - t0))
    return statistics.median(ts)


def opening_proof(m, lg_d):
    """OpeningSet::new + PolynomialBatch::prove_openings (plonk/proof.rs:316-344, fri/oracle.rs:162-219) over the four
    committed oracles of one prove(): everything opened at zeta, the 2 Z polynomials also at g*zeta
    (circuit_data.rs:461-481), standard_recursion_config.  GPU (engine + host transcript) against the CPU oracle
    (C arithmetic + Python glue), proofs compared field by field."""
    from oracle import fri_ref as fr
    from plonky2_demo_b200 import fri_prover as fp

    cfg = pcs.CircuitConfig.standard_recursion_config().fri_config
    params = cfg.fri_params(lg_d, False)
    coeffs = [seeded_polys(w, 1 << lg_d, base_seed=7000 * m + w) for _, w, _ in COMMITS]
    gpu = [pcs.PolynomialBatch.from_coeffs(c, cfg.rate_bits, False, cfg.cap_height, keep_coeffs=True) for c in coeffs]
    cpu = []
    for c in coeffs:
        o = oracle.commit_from_coeffs(c, cfg.rate_bits, cfg.cap_height)
        o["coeffs"], o["cap_height"] = c, cfg.cap_height
        cpu.append(o)
    zeta = (0x0123456789ABCDEF % fp.P, 0x0FEDCBA987654321 % fp.P)
    zeta_next = fp.ext_mul((fp.primitive_root_of_unity(lg_d), 0), zeta)
    all_polys = [(k, j) for k, (_, w, _) in enumerate(COMMITS) for j in range(w)]
    batches = [(zeta, all_polys), (zeta_next, [(2, 0), (2, 1)])]
    inst = fp.FriInstanceInfo(
        oracles=[fp.FriOracleInfo(w, False) for _, w, _ in COMMITS],
        batches=[fp.FriBatchInfo(pt, [fp.FriPolynomialInfo(o, j) for o, j in polys]) for pt, polys in batches])

    def gpu_openings():
        at_zeta = [fp.eval_commitment(zeta, b) for b in gpu]
        at_next = fp.eval_commitment(zeta_next, gpu[2])
        return [[tuple(int(x) for x in at_zeta[o][j]) for o, j in all_polys], [tuple(int(x) for x in at_next[j]) for j in (0, 1)]]

    openings = gpu_openings()

    def transcript(cls):
        ch = cls()
        for o in cpu:
            ch.observe_cap(o["cap"])
        for vals in openings:
            ch.observe_extension_elements(vals)
        return ch

    t0 = time.perf_counter()
    want = fr.prove_openings(cpu, batches, transcript(fr.Challenger), cfg.rate_bits, cfg.cap_height, params.reduction_arity_bits,
                             cfg.proof_of_work_bits, cfg.num_query_rounds)
    cpu_ms = 1e3 * (time.perf_counter() - t0)
    got = fp.prove_openings(inst, gpu, transcript(fp.Challenger), params)
    same = (all(np.array_equal(c.hashes, wc) for c, wc in zip(got.commit_phase_merkle_caps, want["commit_phase_merkle_caps"]))
            and np.array_equal(got.final_poly, want["final_poly"]) and got.pow_witness == want["pow_witness"]
            and got.fri_query_indices == want["_indices"]
            and all(np.array_equal(s.evals, ws["evals"]) for r, wr in zip(got.query_round_proofs, want["query_round_proofs"])
                    for s, ws in zip(r.steps, wr["steps"])))
    base_transcript = transcript(fp.Challenger)   # prove() has observed caps and openings before prove_openings starts
    rec = {"polys_at_zeta": len(all_polys), "polys_at_zeta_next": 2, "arities": params.reduction_arity_bits,
           "proof_equal_to_cpu_oracle": bool(same),
           "gpu_openings_ms": med(gpu_openings, 3),
           "gpu_prove_openings_ms": med(lambda: fp.prove_openings(inst, gpu, base_transcript.clone(), params), 5),
           "cpu_oracle_prove_openings_ms": cpu_ms,
           "note": "CPU figure = oracle/fri.c arithmetic (OpenMP where the reference uses rayon) + Python protocol glue, one run"}
    for b in gpu:
        b.free()
    return rec


def quotient_phase(m, lg_d, r=3, quotient_degree_bits=3):
    """what compute_quotient_polys (plonk/prover.rs:576-744) asks of the committed batches when they are
    device-resident.  It walks all degree * 2^quotient_degree_bits points in batches of 32 (:574,615), per batch one
    get_lde_values_packed per oracle (constants+sigmas, wires, Z / partial products; oracle.rs:137-159), and ends with one
    coset_ifft per challenge (:739-743).  Measured: (a) ONE bulk pcs_batch_lde_natural per oracle (rows land in host memory in
    the order the loop walks them), (b) the per-batch-of-32 synchronous form, extrapolated from 64 calls, (c) the coset_ifft of
    the 2 quotient polynomials through the host entry point and on the device."""
    import torch

    from plonky2_demo_b200 import _ffi

    step = 1 << (r - quotient_degree_bits)
    n_points = (1 << lg_d) << quotient_degree_bits
    rec = {"points": n_points, "step": step, "oracles": []}
    for name, w, _ in COMMITS[:3]:
        x = seeded_polys(w, 1 << lg_d, base_seed=3000 * m + w)
        b = pcs.PolynomialBatch.from_coeffs(x, r, False, 4)
        bulk = med(lambda: b.lde_values_natural(0, step, n_points), 3)
        rows = b.lde_values_natural(0, step, n_points)
        ok = np.array_equal(rows[:64], np.stack([b.get_lde_values(i, step) for i in range(64)]))
        calls = min(64, n_points // 32)
        t0 = time.perf_counter()
        for k in range(calls):
            b.get_lde_values_packed(32 * k, step, 32)
        per_call = 1e3 * (time.perf_counter() - t0) / calls
        rec["oracles"].append({"name": name, "polys": w, "bulk_rows_ms": bulk, "bulk_GBps": n_points * w * 8 / (bulk * 1e-3) / 1e9,
                               "rows_equal_get_lde_values": bool(ok), "per_32_points_call_ms": per_call,
                               "per_32_points_total_ms_extrapolated": per_call * (n_points // 32)})
        b.free()
    lg_q = lg_d + quotient_degree_bits
    vals = seeded_polys(2, 1 << lg_q, base_seed=99 + m)
    want = None
    def host():
        v = vals.copy()
        _ffi.check(_ffi.lib().pcs_coset_intt(_ffi.ptr(v), 2, lg_q, 7))
        return v
    want = host()
    t = torch.from_numpy(vals.view(np.int64).copy()).cuda()
    def dev():
        t2 = t.clone()
        _ffi.check(_ffi.lib().pcs_coset_intt_dev(C.c_void_p(t2.data_ptr()), 2, lg_q, 7))
        _ffi.check(_ffi.lib().pcs_synchronize())
        return t2
    same = np.array_equal(dev().cpu().numpy().view(np.uint64), want)
    cpu_ms = med(lambda: oracle.fft(vals, inverse=True), 3)
    rec["coset_ifft"] = {"polys": 2, "lg_n": lg_q, "host_entry_ms": med(host, 5), "device_entry_ms": med(dev, 5), "device_equals_host": bool(same),
                         "cpu_port_ifft_ms": cpu_ms}
    rec["bulk_rows_ms_total"] = sum(o["bulk_rows_ms"] for o in rec["oracles"])
    rec["per_32_points_ms_total_extrapolated"] = sum(o["per_32_points_total_ms_extrapolated"] for o in rec["oracles"])
    return rec


def main():
    pcs.init(0)
    out = {"rate_bits": 3, "cap_height": 4, "cpu_threads": oracle.num_threads(), "configs": []}
    for m, lg_d, fri_layers in [(2, 3, []), (64, 15, [14, 10, 6])]:
        rec = {"m": m, "degree_bits": lg_d, "commits": [], "fri_layer_trees": []}
        for name, w, from_values in COMMITS:
            x = seeded_polys(w, 1 << lg_d, base_seed=1000 * m + w)
            if from_values:
                gpu = lambda: pcs.PolynomialBatch.from_values(x, 3, False, 4).free()
                cpu = lambda: oracle.commit_from_values(x, 3, 4)
                b = pcs.PolynomialBatch.from_values(x, 3, False, 4)
                ok = np.array_equal(b.merkle_tree.cap.hashes, oracle.commit_from_values(x, 3, 4)["cap"])
            else:
                gpu = lambda: pcs.PolynomialBatch.from_coeffs(x, 3, False, 4).free()
                cpu = lambda: oracle.commit_from_coeffs(x, 3, 4)
                b = pcs.PolynomialBatch.from_coeffs(x, 3, False, 4)
                ok = np.array_equal(b.merkle_tree.cap.hashes, oracle.commit_from_coeffs(x, 3, 4)["cap"])
            b.free()
            rec["commits"].append({"name": name, "polys": w, "from_values": from_values, "cap_equal": bool(ok),
                                   "gpu_ms": med(gpu), "cpu_port_ms": med(cpu, 3)})
        for log_n in fri_layers:
            leaves = splitmix64_stream(log_n, (1 << log_n) * 32).reshape(1 << log_n, 32)
            t = pcs.MerkleTree.new(leaves, 4)
            ok = np.array_equal(t.cap.hashes, oracle.merkle_build(leaves, 4)[1])
            rec["fri_layer_trees"].append({"log_leaves": log_n, "leaf_len": 32, "cap_equal": bool(ok),
                                           "gpu_ms": med(lambda: pcs.MerkleTree.new(leaves, 4)),
                                           "cpu_port_ms": med(lambda: oracle.merkle_build(leaves, 4), 3)})
        rec["opening_proof"] = opening_proof(m, lg_d)
        rec["quotient_phase"] = quotient_phase(m, lg_d)
        rec["gpu_ms_total"] = sum(c["gpu_ms"] for c in rec["commits"]) + sum(c["gpu_ms"] for c in rec["fri_layer_trees"])
        rec["cpu_port_ms_total"] = sum(c["cpu_port_ms"] for c in rec["commits"]) + sum(c["cpu_port_ms"] for c in rec["fri_layer_trees"])
        out["configs"].append(rec)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
