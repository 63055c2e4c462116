"""Stand-alone timing of the opening-proof kernels at the headline shape: the `fri_opening` block of
bench.py without the rest of the bench, so that an ncu launch list / full capture of k_eval_ext and
k_reduce_polys_base stays short.

    python tools/fri_bench.py [--width 135 --lg-d 20 --steps 3]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--width", type=int, default=135)
    ap.add_argument("--lg-d", type=int, default=20)
    ap.add_argument("--rate-bits", type=int, default=3)
    ap.add_argument("--cap-height", type=int, default=4)
    ap.add_argument("--steps", type=int, default=3)
    a = ap.parse_args()
    import torch

    import bench
    import plonky2_demo_b200 as pcs
    from plonky2_demo_b200 import _ffi

    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    pcs.init(0, stream.cuda_stream)
    d = 1 << a.lg_d
    gen = torch.Generator(device=dev).manual_seed(1)
    coeffs = torch.randint(-(1 << 63), (1 << 63) - 1, (a.width, d), dtype=torch.int64, device=dev, generator=gen)
    ptrs = _ffi.dev_ptr_array(coeffs.data_ptr(), a.width, d)
    peaks, _ = bench.measured_peaks()
    out = bench.fri_opening_bench(pcs, a.width, a.lg_d, a.rate_bits, a.cap_height, ptrs, a.steps, float(peaks["hbm_gbs"]))
    print(json.dumps(out))
    pcs.shutdown()


if __name__ == "__main__":
    main()
