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
        rec["gpu_ms_total"] = sum(c["gpu_ms"] for c in rec["commits"]) + sum(c["gpu_ms"] for c in rec["fri_layer_trees"])
        rec["cpu_port_ms_total"] = sum(c["cpu_port_ms"] for c in rec["commits"]) + sum(c["cpu_port_ms"] for c in rec["fri_layer_trees"])
        out["configs"].append(rec)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
