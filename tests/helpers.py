"""Shared test helpers: seeded inputs (SURVEY 8d generator) and big-int field arithmetic."""
import numpy as np

P = 0xFFFFFFFF00000001
MASK = (1 << 64) - 1


def splitmix64_stream(seed, n):
    """splitmix64(seed) stream with rejection < p => canonical uniform elements (vectorised)."""
    out = np.empty(0, dtype=np.uint64)
    state = np.uint64(seed)
    with np.errstate(over="ignore"):
        while out.size < n:
            m = n - out.size + 16
            idx = np.arange(1, m + 1, dtype=np.uint64)
            z = state + idx * np.uint64(0x9E3779B97F4A7C15)
            state = z[-1]
            z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
            z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
            z = z ^ (z >> np.uint64(31))
            out = np.concatenate([out, z[z < np.uint64(P)]])
    return out[:n]


def seeded_polys(w, d, base_seed=0x5EED0000):
    return np.stack([splitmix64_stream(base_seed + j, d) for j in range(w)])


def field_grid():
    """field/src/prime_field_testing.rs:7-17 test_inputs(p) (restated)."""
    vals = list(range(0, 10))
    for k in (31, 32, 63):
        vals += list(range((1 << k) - 10, (1 << k) + 11))
    vals += list(range(P - 10, P))
    return sorted(set(v for v in vals if v < P))


def brev(i, bits):
    return int(format(i, f"0{bits}b")[::-1], 2) if bits else 0


def canon(a):
    a = np.asarray(a, dtype=np.uint64)
    return np.where(a >= np.uint64(P), a - np.uint64(P), a)
