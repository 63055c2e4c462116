"""Exactness of the FP64 MDS network (csrc/poseidon.cuh: mds_net_d, renorm_d, fold_d).

Every value the network computes is an integer linear form of its 12 inputs; a double holds it exactly iff its
magnitude stays below 2^53.  This test replays the network's operation sequence on coefficient vectors, takes
the largest L1 norm of any intermediate (incl. the inner results of nested FMAs) and checks the magnitude
chain of the two-rounds-between-renormalisations schedule.  It also checks, on big integers, that the network
equals the MDS matrix and that the renormalisation preserves the value mod p and lands in the stated ranges."""
import random

import numpy as np

P = 0xFFFFFFFF00000001
CIRC = [17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20]  # poseidon_goldilocks.rs:24
DIAG0 = 8                                                # poseidon_goldilocks.rs:25


class Rec:
    """runs the network on arbitrary ring elements and records every intermediate"""

    def __init__(self):
        self.seen = []

    def t(self, v):
        self.seen.append(v)
        return v

    def fma(self, a, c, b):
        return self.t(a * c + b)


def mds_net(s, R):
    """mirror of mds_net_d (same operations, same order)"""
    t, fma = R.t, R.fma
    A, B, Pp, Q = [0] * 3, [0] * 3, [0] * 3, [0] * 3
    for j in range(3):
        u, v = t(s[j] + s[j + 6]), t(s[j + 3] + s[j + 9])
        A[j], B[j] = t(u + v), t(u - v)
        Pp[j], Q[j] = t(s[j] - s[j + 6]), t(s[j + 3] - s[j + 9])
    tt = t(t(A[0] + A[1]) + A[2])
    Ya = [t(tt + A[2]), t(tt + A[0]), t(tt + A[1])]
    Yb = [fma(B[1], -2, fma(B[2], 8, -B[0])), fma(B[2], -2, fma(B[0], -8, -B[1])), fma(B[1], -8, fma(B[0], 2, -B[2]))]
    re, im = [0] * 3, [0] * 3
    re[0] = fma(Q[2], 4, t(fma(Q[1], -16, t(fma(Pp[0], 2, -Q[0]) + Pp[1])) + Pp[2]))
    im[0] = fma(Pp[2], -4, t(fma(Pp[1], 16, t(fma(Q[0], 2, Pp[0]) + Q[1])) + Q[2]))
    re[1] = fma(Q[2], -16, t(fma(Pp[1], 2, t(fma(Pp[0], -4, Q[0]) - Q[1])) + Pp[2]))
    im[1] = fma(Pp[2], 16, t(fma(Q[1], 2, t(fma(Q[0], -4, -Pp[0]) + Pp[1])) + Q[2]))
    re[2] = fma(Pp[2], 2, t(fma(Pp[1], -4, t(fma(Pp[0], 16, Q[0]) + Q[1])) - Q[2]))
    im[2] = fma(Q[2], 2, t(fma(Q[1], -4, t(fma(Q[0], 16, -Pp[0]) - Pp[1])) + Pp[2]))
    y = [0] * 12
    for j in range(3):
        e1, e2 = fma(Ya[j], 16, Yb[j]), fma(Ya[j], 16, -Yb[j])
        y[j], y[j + 3], y[j + 6], y[j + 9] = t(e1 + re[j]), t(e2 + im[j]), t(e1 - re[j]), t(e2 - im[j])
    y[0] = fma(s[0], DIAG0, y[0])
    return y


def test_network_is_the_mds_matrix():
    rnd = random.Random(1)
    for _ in range(50):
        s = [rnd.randrange(-(1 << 40), 1 << 41) for _ in range(12)]
        want = [sum(CIRC[i] * s[(r + i) % 12] for i in range(12)) + (DIAG0 * s[0] if r == 0 else 0) for r in range(12)]
        assert mds_net(s, Rec()) == want


def test_intermediates_fit_a_double_for_two_rounds():
    R = Rec()
    basis = [np.eye(12, dtype=np.int64)[i] for i in range(12)]
    y = mds_net(basis, R)
    l1 = max(int(np.abs(v).sum()) for v in R.seen)
    row = max(int(np.abs(v).sum()) for v in y)
    assert row == sum(CIRC) + DIAG0 == 264
    assert l1 <= 700, l1
    b0 = (1 << 33) + (1 << 19)          # normalised halves (renorm_d), lane 0 halves are < 2^32
    assert l1 * b0 < 1 << 53            # round A exact
    b1 = row * b0                        # round A outputs
    assert l1 * b1 < 1 << 53            # round B exact
    b2 = row * b1                        # round B outputs
    assert b2 + (1 << 52) + (1 << 32) < 1 << 53   # + biased constant / + 2^52 for the integer fold
    assert b2 >> 32 < 1 << 18           # l1, h1 of renorm_d / fold_halves_biased


def renorm(lo, hi):
    l1m, h1m = (lo >> 32) - 1, (hi >> 32) - 1
    return lo - l1m * (1 << 32) - h1m, hi - h1m * ((1 << 32) - 1) + l1m


def test_renormalisation_preserves_value_and_ranges():
    rnd = random.Random(2)
    cases = [(0, 0), (1, 1), ((1 << 50) - 1, (1 << 50) - 1), ((1 << 32) - 1, 0), (0, (1 << 32) - 1)]
    cases += [(rnd.randrange(1 << 50), rnd.randrange(1 << 50)) for _ in range(2000)]
    for lo, hi in cases:
        lo2, hi2 = renorm(lo, hi)
        assert (lo2 + (hi2 << 32) - lo - (hi << 32)) % P == 0
        assert (1 << 32) - (1 << 18) <= lo2 <= (1 << 33) and (1 << 32) - 2 <= hi2 < (1 << 33) + (1 << 19)
