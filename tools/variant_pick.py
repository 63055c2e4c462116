"""Read the A/B bench lines of tools/gpu_variant_round.sh (gpurun_out/ab_<tag>.json), pick the fastest leaf-hash form and the
fastest NTT form among the runs whose parity check is green, and install build/variants/libpcs_<hash>_<ntt>.so as the
in-tree libpcs.so (the full `-m gpu` suite then runs on exactly that binary).  Prints the choice as JSON."""
import glob
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")


def main():
    rows = {}
    for f in sorted(glob.glob(os.path.join(OUT, "ab_*.json"))):
        tag = os.path.basename(f)[3:-5]
        try:
            line = [l for l in open(f) if l.startswith("{")][-1]
            d = json.loads(line)
            ok = all(v is True for k, v in d["parity_check"].items() if k.endswith("sample_size"))
            rows[tag] = {"ok": ok, "commit_ms": d["ms_per_step"], "lde_ms": d["phase_ms"]["FFT + blinding"],
                         "leaf_ms": d["phase_ms"]["leaf hashing"], "node_ms": d["phase_ms"]["node levels"]}
        except Exception as e:  # a variant that crashed or printed nothing is simply not a candidate
            rows[tag] = {"ok": False, "error": repr(e)}
    good = {t: r for t, r in rows.items() if r.get("ok")}
    choice = {"rows": rows, "hash": "base", "ntt": "n0", "tag": "base"}
    # the round-2 build's phase times on this pool's B200s (profiles/r02_bench_1gpu_final.json) stand in when its own run is missing
    base = good.get("base") or {"lde_ms": 20.17, "leaf_ms": 97.96, "node_ms": 5.98, "stored": True}
    choice["base_used"] = base
    # hash form = first letter of the tag, NTT form = suffix
    best_h, best_h_ms = "base", base["leaf_ms"] + base["node_ms"]
    best_n, best_n_ms = "n0", base["lde_ms"]
    for t, r in good.items():
        if t == "base":
            continue
        h, n = t.split("_")
        if r["leaf_ms"] + r["node_ms"] < 0.99 * best_h_ms:
            best_h, best_h_ms = h, r["leaf_ms"] + r["node_ms"]
        if n != "n0" and r["lde_ms"] < 0.995 * best_n_ms:
            best_n, best_n_ms = n, r["lde_ms"]
    if best_h == "base" and best_n != "n0":
        best_n = "n0"          # no prebuilt library pairs the old hash forms with a new NTT form
    tag = "base" if best_h == "base" else f"{best_h}_{best_n}"
    choice.update(hash=best_h, ntt=best_n, tag=tag)
    lib = os.path.join(ROOT, "build", "variants", f"libpcs_{choice['tag']}.so")
    if not os.path.exists(lib):
        choice.update(tag="base")
        lib = os.path.join(ROOT, "build", "variants", "libpcs_base.so")
    shutil.copyfile(lib, os.path.join(ROOT, "plonky2_demo_b200", "libpcs.so"))
    choice["installed"] = os.path.relpath(lib, ROOT)
    json.dump(choice, sys.stdout, indent=1)
    print()


if __name__ == "__main__":
    main()
