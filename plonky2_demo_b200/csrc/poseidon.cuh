// Poseidon-12 permutation over Goldilocks, one permutation per thread, state in registers.
//
// GPU counterpart of Poseidon::poseidon (plonky2/src/hash/poseidon.rs:599-609):
//   4 full rounds -> 22 partial rounds -> 4 full rounds, S-box x^7, MDS = circ(17,15,41,16,2,28,
//   13,13,39,18,34,20) + diag(8,0,..) (poseidon_goldilocks.rs:24-25).
// Output is bit-identical (after canonicalisation) to the reference for every input; the
// reference's own `consistency` test (poseidon.rs:777-790) licenses any algebraically equal
// evaluation order, and tests/ pin this kernel to the reference KATs (poseidon_goldilocks.rs:461-482).
//
// Design for the B200 integer pipes (no tensor cores: 64-bit modular arithmetic).  Measured rates
// (profiles/r01_intpipe.md): IADD3 128/clk/SM, IMAD/LOP3/SHF 64, IMAD.WIDE 32 -- so the wide
// multiplier is the scarce unit and is reserved for the S-boxes (4 IMAD.WIDE per modmul):
//  * MDS layer = add/shift network, zero multiplies.  Each lane is split in 32-bit halves; the
//    12-point circulant product of a half is computed exactly in int64 through the factorisation
//    x^12 - 1 = prod over zeta in {1,-1,i,-i} of (x^3 - zeta) with x^4 = zeta (three 4-point integer
//    DFTs, three 3x3 twisted products whose constants are (2+i), (-4-i), (16-i), [16,32,16],
//    [-1,-8,2] -- the MDS matrix was designed to have these -- and three inverse DFTs).  Same
//    mathematics as the reference's x86 path (poseidon_goldilocks.rs:217-248,307-441), derived
//    and verified independently in tests/golden/make_golden.py:mds_network_check.
//    The next round's constant is folded into the network's additions (3-input IADD3), so
//    constant_layer costs nothing; one 96-bit fold per lane brings the result back to 64 bits.
//  * partial rounds use the dense network too, with the constants pushed through the linear layer
//    (FAST_PARTIAL_FIRST_ROUND_CONSTANT once, then the scalar FAST_PARTIAL_ROUND_CONSTANTS[r] times
//    the first MDS column); the sparse "fast" matrices of poseidon.rs:584-596 would need 23 full
//    64x64 multiplies per round on the scarce pipe.
#pragma once
#include "gl64.cuh"
#include "poseidon_constants.cuh"

namespace pcs {

#ifndef PCS_MDS_FP64
#define PCS_MDS_FP64 1
#endif
#ifndef PCS_SBOX_GROUP
#define PCS_SBOX_GROUP 12
#endif

constexpr int SPONGE_WIDTH = 12;
constexpr int SPONGE_RATE = 8;

__device__ __forceinline__ uint64_t sbox7(uint64_t x) {
    uint64_t x2 = gl::sqr(x);
    uint64_t x4 = gl::sqr(x2);
    uint64_t x3 = gl::mul(x, x2);
    return gl::mul(x3, x4);
}

// y = circ-correlation(s) + 8*s0*e0 + add, for 12 non-negative 32-bit inputs; exact in int64
// (|intermediates| < 2^40, results < 2^43 + add).
__device__ __forceinline__ void mds_half(const uint32_t (&s)[12], const uint32_t (&add)[12], uint64_t (&y)[12]) {
    int64_t A[3], B[3], P[3], Q[3];
#pragma unroll
    for (int j = 0; j < 3; j++) {
        int64_t s0 = s[j], s3 = s[j + 3], s6 = s[j + 6], s9 = s[j + 9];
        int64_t u = s0 + s6, v = s3 + s9;
        A[j] = u + v;    // S_j(1)
        B[j] = u - v;    // S_j(-1)
        P[j] = s0 - s6;  // Re S_j(i)
        Q[j] = s3 - s9;  // Im S_j(i)
    }
    // zeta = 1: cyclic 3-product with [16, 32, 16]
    int64_t t = A[0] + A[1] + A[2];
    int64_t Ya[3] = {(t + A[2]) << 4, (t + A[0]) << 4, (t + A[1]) << 4};
    // zeta = -1: negacyclic 3-product with [-1, -8, 2]
    int64_t Yb[3] = {(B[2] << 3) - (B[1] << 1) - B[0],
                     -(B[0] << 3) - B[1] - (B[2] << 1),
                     (B[0] << 1) - (B[1] << 3) - B[2]};
    // zeta = i: i-twisted 3-product with (2+i), (-4-i), (16-i)
    int64_t re[3], im[3];
    re[0] = (P[0] << 1) - Q[0] + P[1] - (Q[1] << 4) + P[2] + (Q[2] << 2);
    im[0] = P[0] + (Q[0] << 1) + (P[1] << 4) + Q[1] - (P[2] << 2) + Q[2];
    re[1] = -(P[0] << 2) + Q[0] + (P[1] << 1) - Q[1] + P[2] - (Q[2] << 4);
    im[1] = -P[0] - (Q[0] << 2) + P[1] + (Q[1] << 1) + (P[2] << 4) + Q[2];
    re[2] = (P[0] << 4) + Q[0] - (P[1] << 2) + Q[1] + (P[2] << 1) - Q[2];
    im[2] = -P[0] + (Q[0] << 4) - P[1] - (Q[1] << 2) + P[2] + (Q[2] << 1);
#pragma unroll
    for (int j = 0; j < 3; j++) {
        int64_t e1 = Ya[j] + Yb[j], e2 = Ya[j] - Yb[j];
        y[j] = (uint64_t)(e1 + re[j] + (int64_t)add[j]);
        y[j + 3] = (uint64_t)(e2 + im[j] + (int64_t)add[j + 3]);
        y[j + 6] = (uint64_t)(e1 - re[j] + (int64_t)add[j + 6]);
        y[j + 9] = (uint64_t)(e2 - im[j] + (int64_t)add[j + 9]);
    }
    y[0] += (uint64_t)s[0] << 3;  // MDS_MATRIX_DIAG[0] = 8
}

// s <- MDS * s + rc   (rc = nullptr: no constant; lane0_only: rc[0] is added to lane 0 only)
template <int RC_MODE /*0 none, 1 all 12 lanes, 2 lane 0 only*/>
__device__ __forceinline__ void mds_layer(uint64_t (&s)[12], const uint64_t* __restrict__ rc) {
    uint32_t lo[12], hi[12], alo[12], ahi[12];
#pragma unroll
    for (int i = 0; i < 12; i++) {
        lo[i] = (uint32_t)s[i];
        hi[i] = (uint32_t)(s[i] >> 32);
        uint64_t c = (RC_MODE == 1 || (RC_MODE == 2 && i == 0)) ? rc[i] : 0;
        alo[i] = (uint32_t)c;
        ahi[i] = (uint32_t)(c >> 32);
    }
    uint64_t L[12], H[12];
    mds_half(lo, alo, L);
    mds_half(hi, ahi, H);
#pragma unroll
    for (int r = 0; r < 12; r++) s[r] = gl::fold_halves(L[r], H[r]);
}

// ------------------------------------------------------------------------------------------------
// The same network on the FP64 pipe.
//
// B200 has a full-rate FP64 pipe (measured 61 DFMA/clk/SM, co-issuing with the integer pipes at
// 121 instr/clk/SM, profiles/r01_fp64pipe.md) that integer code leaves idle, while the integer
// version of the network saturates the ALU pipe (profiles/r01_poseidon_unified.md).  Every value in
// the network is an integer of magnitude < 2^45 and every operation is an add, a subtract or a
// multiplication by 2, 4, 8 or 16 -- all EXACT in IEEE double (53-bit significand), in any
// evaluation order and rounding mode.  One DADD/DFMA replaces the IADD3 + IADD3.X pair of a 64-bit
// integer add, and fused shift-adds are free.
//   in : a 32-bit half x enters as the double with bit pattern 0x43300000_xxxxxxxx = 2^52 + x
//        (no conversion instruction); differences of two such values are exact, and
//        (2^52 + x) - 2^52 recovers x where a sum needs it.
//   out: the last addition adds the pre-biased round constant 2^52 + c (ROUND_ADD_D), so the result
//        2^52 + y has y's low 32 bits in its low word and y >> 32 in its high word - 0x43300000.
// ------------------------------------------------------------------------------------------------
constexpr double TWO52 = 4503599627370496.0;

__device__ __forceinline__ void mds_half_fp64(const uint32_t (&s)[12], const double* __restrict__ addb /*stride 2*/,
                                              uint32_t (&ylo)[12], uint32_t (&yhi)[12]) {
    double A[3], B[3], P[3], Q[3];
    double s0d;
#pragma unroll
    for (int j = 0; j < 3; j++) {
        double v0 = __hiloint2double(0x43300000, (int)s[j]), v3 = __hiloint2double(0x43300000, (int)s[j + 3]);
        double v6 = __hiloint2double(0x43300000, (int)s[j + 6]), v9 = __hiloint2double(0x43300000, (int)s[j + 9]);
        double d6 = v6 - TWO52, d9 = v9 - TWO52;  // s6, s9 as exact doubles
        P[j] = v0 - v6;                           // s0 - s6
        Q[j] = v3 - v9;                           // s3 - s9
        double u = fma(d6, 2.0, P[j]);            // s0 + s6
        double v = fma(d9, 2.0, Q[j]);            // s3 + s9
        A[j] = u + v;
        B[j] = u - v;
        if (j == 0) s0d = P[0] + d6;              // s0 itself, for the diagonal term
    }
    double t = A[0] + A[1] + A[2];
    double Ya[3] = {t + A[2], t + A[0], t + A[1]};  // times 16, applied below
    double Yb[3] = {fma(B[1], -2.0, fma(B[2], 8.0, -B[0])),
                    fma(B[2], -2.0, fma(B[0], -8.0, -B[1])),
                    fma(B[1], -8.0, fma(B[0], 2.0, -B[2]))};
    double re[3], im[3];
    re[0] = fma(Q[2], 4.0, fma(Q[1], -16.0, fma(P[0], 2.0, -Q[0]) + P[1]) + P[2]);
    im[0] = fma(P[2], -4.0, fma(P[1], 16.0, fma(Q[0], 2.0, P[0]) + Q[1]) + Q[2]);
    re[1] = fma(Q[2], -16.0, fma(P[1], 2.0, fma(P[0], -4.0, Q[0]) - Q[1]) + P[2]);
    im[1] = fma(P[2], 16.0, fma(Q[1], 2.0, fma(Q[0], -4.0, -P[0]) + P[1]) + Q[2]);
    re[2] = fma(P[2], 2.0, fma(P[1], -4.0, fma(P[0], 16.0, Q[0]) + Q[1]) - Q[2]);
    im[2] = fma(Q[2], 2.0, fma(Q[1], -4.0, fma(Q[0], 16.0, -P[0]) - P[1]) + P[2]);
#pragma unroll
    for (int j = 0; j < 3; j++) {
        double e1 = fma(Ya[j], 16.0, Yb[j]), e2 = fma(Ya[j], 16.0, -Yb[j]);
        double y0 = e1 + re[j];
        if (j == 0) y0 = fma(s0d, 8.0, y0);       // MDS_MATRIX_DIAG[0] = 8
        double r0 = y0 + addb[2 * j];
        double r3 = (e2 + im[j]) + addb[2 * (j + 3)];
        double r6 = (e1 - re[j]) + addb[2 * (j + 6)];
        double r9 = (e2 - im[j]) + addb[2 * (j + 9)];
        ylo[j] = (uint32_t)__double2loint(r0);     yhi[j] = (uint32_t)__double2hiint(r0);
        ylo[j + 3] = (uint32_t)__double2loint(r3); yhi[j + 3] = (uint32_t)__double2hiint(r3);
        ylo[j + 6] = (uint32_t)__double2loint(r6); yhi[j + 6] = (uint32_t)__double2hiint(r6);
        ylo[j + 9] = (uint32_t)__double2loint(r9); yhi[j + 9] = (uint32_t)__double2hiint(r9);
    }
}

// s <- MDS * s + ROUND_ADD[r]  on the FP64 pipe; rcd = &ROUND_ADD_D[24 * r]
__device__ __forceinline__ void mds_layer_fp64(uint64_t (&s)[12], const uint64_t* __restrict__ rcd) {
    const double* addb = reinterpret_cast<const double*>(rcd);
    uint32_t lo[12], hi[12];
#pragma unroll
    for (int i = 0; i < 12; i++) {
        lo[i] = (uint32_t)s[i];
        hi[i] = (uint32_t)(s[i] >> 32);
    }
    uint32_t l0[12], l1[12], h0[12], h1[12];
    mds_half_fp64(lo, addb, l0, l1);
    mds_half_fp64(hi, addb + 1, h0, h1);
#pragma unroll
    for (int r = 0; r < 12; r++) s[r] = gl::fold_halves_biased(l0[r], l1[r], h0[r], h1[r]);
}

// ------------------------------------------------------------------------------------------------
// Partial rounds with lanes 1..11 RESIDENT in the FP64 network's domain.
//
// In the 22 partial rounds only lane 0 passes the S-box, so only lane 0 has to exist as a 64-bit integer.
// Lanes 1..11 stay as pairs of exact doubles (lo, hi), value = lo + hi * 2^32 (mod p), and go from one
// round's network output straight into the next round's network input: no 96-bit fold, no int<->double
// re-biasing for 11 of the 12 lanes.  Magnitudes grow by the MDS row sum (264 < 2^8.05) per round, and the
// network's widest intermediate is bounded by 2^8.5 x its inputs (L1 norms checked symbolically in
// tests/test_mds_network_bounds.py), so TWO rounds fit the 53-bit significand:
//     normalised inputs < 2^33.01 -> round A outputs < 2^41.1 -> round B intermediates < 2^49.7, outputs < 2^49.2.
// After round B the lanes are renormalised in 8 FP64 operations each (floor by the 2^52 trick with
// round-down FMAs, 2^64 = 2^32 - 1 folded back, plus one multiple of p so that both halves stay positive):
//     lo'' = lo - (l1-1) 2^32 - (h1-1)                 in [2^32 - 2^18, 2^33]
//     hi'' = hi - (h1-1)(2^32 - 1) + (l1-1)            in [2^32 - 2, 2^33 + 2^19),      l1 = lo >> 32, h1 = hi >> 32
//     lo'' + hi'' 2^32 = lo + hi 2^32 + p - h1 (2^64 - 2^32 + 1)  ==  lo + hi 2^32   (mod p).
// The scalar round constant is added to lane 0 right after its S-box (poseidon.rs:584-596 does the same in
// the "fast" form), so the network carries no per-lane constants in these rounds.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double u32_as_double(uint32_t x) { return __hiloint2double(0x43300000, (int)x) - TWO52; }

// y = circ-correlation(s) + 8*s0*e0 on exact doubles (no constants).  ZERO0: lane 0 is taken as 0 (its
// contribution is added afterwards, see partial_round_d).
template <bool ZERO0>
__device__ __forceinline__ void mds_net_d(const double (&s)[12], double (&y)[12]) {
    double A[3], B[3], P[3], Q[3];
#pragma unroll
    for (int j = 0; j < 3; j++) {
        const bool z = ZERO0 && j == 0;
        double u = z ? s[6] : s[j] + s[j + 6], v = s[j + 3] + s[j + 9];
        A[j] = u + v;
        B[j] = u - v;
        P[j] = z ? -s[6] : s[j] - s[j + 6];
        Q[j] = s[j + 3] - s[j + 9];
    }
    double t = A[0] + A[1] + A[2];
    double Ya[3] = {t + A[2], t + A[0], t + A[1]};  // times 16, applied below
    double Yb[3] = {fma(B[1], -2.0, fma(B[2], 8.0, -B[0])),
                    fma(B[2], -2.0, fma(B[0], -8.0, -B[1])),
                    fma(B[1], -8.0, fma(B[0], 2.0, -B[2]))};
    double re[3], im[3];
    re[0] = fma(Q[2], 4.0, fma(Q[1], -16.0, fma(P[0], 2.0, -Q[0]) + P[1]) + P[2]);
    im[0] = fma(P[2], -4.0, fma(P[1], 16.0, fma(Q[0], 2.0, P[0]) + Q[1]) + Q[2]);
    re[1] = fma(Q[2], -16.0, fma(P[1], 2.0, fma(P[0], -4.0, Q[0]) - Q[1]) + P[2]);
    im[1] = fma(P[2], 16.0, fma(Q[1], 2.0, fma(Q[0], -4.0, -P[0]) + P[1]) + Q[2]);
    re[2] = fma(P[2], 2.0, fma(P[1], -4.0, fma(P[0], 16.0, Q[0]) + Q[1]) - Q[2]);
    im[2] = fma(Q[2], 2.0, fma(Q[1], -4.0, fma(Q[0], 16.0, -P[0]) - P[1]) + P[2]);
#pragma unroll
    for (int j = 0; j < 3; j++) {
        double e1 = fma(Ya[j], 16.0, Yb[j]), e2 = fma(Ya[j], 16.0, -Yb[j]);
        y[j] = e1 + re[j];
        y[j + 3] = e2 + im[j];
        y[j + 6] = e1 - re[j];
        y[j + 9] = e2 - im[j];
    }
    if (!ZERO0) y[0] = fma(s[0], 8.0, y[0]);  // MDS_MATRIX_DIAG[0] = 8
}

// (lo, hi) < 2^52, positive  ->  equivalent pair with both halves in [2^32 - 2^18, 2^33 + 2^19)
__device__ __forceinline__ void renorm_d(double& lo, double& hi) {
    const double M1 = TWO52 + 1.0, INV32 = 1.0 / 4294967296.0;
    double l1m = __fma_rd(lo, INV32, TWO52) - M1;   // (lo >> 32) - 1
    double h1m = __fma_rd(hi, INV32, TWO52) - M1;   // (hi >> 32) - 1
    double lo2 = fma(l1m, -4294967296.0, lo) - h1m;
    double hi2 = fma(h1m, -4294967295.0, hi) + l1m;
    lo = lo2;
    hi = hi2;
}

// lane value (lo + hi 2^32, both < 2^52 and >= 0) as a loose u64
__device__ __forceinline__ uint64_t fold_d(double lo, double hi) {
    double bl = lo + TWO52, bh = hi + TWO52;
    return gl::fold_halves_biased((uint32_t)__double2loint(bl), (uint32_t)__double2hiint(bl),
                                  (uint32_t)__double2loint(bh), (uint32_t)__double2hiint(bh));
}

// One partial round: lane 0 = x0 (integer), lanes 1..11 = (dl, dh); (dl, dh) receive M * state.
// M s = M (0, s1..s11) + x0' * M[:,0]: the bulk network does not depend on this round's S-box, so the FP64
// work of lanes 1..11 and the (serial, latency-bound) integer S-box chain of lane 0 are independent
// instruction streams that the scheduler interleaves; lane 0 then enters through 12 FMAs per half.
#ifndef PCS_PARTIAL_SPLIT
#define PCS_PARTIAL_SPLIT 0   // measured: 203.6 vs 199.2 clk per permutation with the split (the extra 24 FMAs cost more than the ILP gains)
#endif
__device__ __forceinline__ void partial_round_d(uint64_t& x0, double (&dl)[12], double (&dh)[12], uint64_t rc) {
    double yl[12], yh[12];
#if PCS_PARTIAL_SPLIT
    mds_net_d<true>(dl, yl);
    mds_net_d<true>(dh, yh);
    x0 = gl::add_lc_p(sbox7(x0), rc);
    const double al = u32_as_double((uint32_t)x0), ah = u32_as_double((uint32_t)(x0 >> 32));
    // M[i][0] = MDS_MATRIX_CIRC[(12 - i) % 12] (+ MDS_MATRIX_DIAG[0] for i = 0)   poseidon_goldilocks.rs:24-25
    constexpr double COL0[12] = {25.0, 20.0, 34.0, 18.0, 39.0, 13.0, 13.0, 28.0, 2.0, 16.0, 41.0, 15.0};
#pragma unroll
    for (int i = 0; i < 12; i++) {
        dl[i] = fma(al, COL0[i], yl[i]);
        dh[i] = fma(ah, COL0[i], yh[i]);
    }
#else
    x0 = gl::add_lc_p(sbox7(x0), rc);
    dl[0] = u32_as_double((uint32_t)x0);
    dh[0] = u32_as_double((uint32_t)(x0 >> 32));
    mds_net_d<false>(dl, yl);
    mds_net_d<false>(dh, yh);
#pragma unroll
    for (int i = 0; i < 12; i++) {
        dl[i] = yl[i];
        dh[i] = yh[i];
    }
#endif
}

#ifndef PCS_PARTIAL_FP64
#define PCS_PARTIAL_FP64 1
#endif

// The permutation.  Input lanes: any u64 (loose); output lanes: canonical (CANON_OUT) or loose -- a sponge feeds the state
// straight into the next permutation and canonicalises only the digest it finally emits.
//
// ONE round loop for all 30 rounds (the body is ~22 KB of SASS and must stay resident in the SM's
// instruction cache: with separate full/partial loop bodies the kernel was instruction-fetch bound,
// profiles/r01_leafhash_v2.md).  Every round is  "s-box ; s = M s + ROUND_ADD[r]"  where the s-box
// covers all lanes in the 8 full rounds and lane 0 only in the 22 partial rounds (warp-uniform
// branch) and ROUND_ADD[r] is the constant needed by the NEXT s-box layer moved behind the linear
// layer (tests/golden/make_golden.py:round_addends); it is absorbed by 3-input adds of the network.
__device__ __forceinline__ void full_round(uint64_t (&s)[12], int r) {
    using namespace pconst;
#if PCS_SBOX_GROUP == 12
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = sbox7(s[i]);
#else
    // S-box layer as a rolled loop over groups of lanes (the state rotates by one group per pass): smaller code for
    // the instruction cache at the price of 2 register moves per lane per pass
#pragma unroll 1
    for (int g = 0; g < 12 / PCS_SBOX_GROUP; g++) {
        uint64_t t[PCS_SBOX_GROUP];
#pragma unroll
        for (int i = 0; i < PCS_SBOX_GROUP; i++) t[i] = sbox7(s[i]);
#pragma unroll
        for (int i = 0; i < 12 - PCS_SBOX_GROUP; i++) s[i] = s[i + PCS_SBOX_GROUP];
#pragma unroll
        for (int i = 0; i < PCS_SBOX_GROUP; i++) s[12 - PCS_SBOX_GROUP + i] = t[i];
    }
#endif
#if PCS_MDS_FP64
    mds_layer_fp64(s, &ROUND_ADD_D[24 * r]);
#else
    mds_layer<1>(s, &ROUND_ADD[12 * r]);
#endif
}

template <bool CANON_OUT = true>
__device__ __forceinline__ void poseidon12(uint64_t (&s)[12]) {
    using namespace pconst;
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = gl::add_lc_p(s[i], RC[i]);
#if PCS_PARTIAL_FP64 && PCS_MDS_FP64
    // rounds 0..3 (full), 4..25 (partial, FP64-resident lanes), 26..29 (full): the two full-round groups share
    // ONE loop body (phase 0 and phase 1) so that the code stays small enough for the instruction cache
#pragma unroll 1
    for (int phase = 0; phase < 2; phase++) {
#pragma unroll 1
        for (int r = 0; r < 4; r++) full_round(s, 26 * phase + r);
        if (phase == 0) {
            uint64_t x0 = s[0];
            double dl[12], dh[12];
#pragma unroll
            for (int i = 1; i < 12; i++) {
                dl[i] = u32_as_double((uint32_t)s[i]);
                dh[i] = u32_as_double((uint32_t)(s[i] >> 32));
            }
#pragma unroll 1
            for (int k = 0; k < 11; k++) {
                partial_round_d(x0, dl, dh, PARTIAL_RC[2 * k]);        // inputs normalised
                x0 = fold_d(dl[0], dh[0]);
                partial_round_d(x0, dl, dh, PARTIAL_RC[2 * k + 1]);    // inputs up to 2^41
                if (k < 10) {
                    x0 = fold_d(dl[0], dh[0]);
#pragma unroll
                    for (int i = 1; i < 12; i++) renorm_d(dl[i], dh[i]);
                }
            }
            // leave the FP64 domain: + RC[26] (the next full round's constants), all lanes back to integers
            const double* rcd = reinterpret_cast<const double*>(&ROUND_ADD_D[24 * 25]);  // 2^52 + halves of RC[26..]
#pragma unroll
            for (int i = 0; i < 12; i++) {
                double bl = dl[i] + rcd[2 * i], bh = dh[i] + rcd[2 * i + 1];
                s[i] = gl::fold_halves_biased((uint32_t)__double2loint(bl), (uint32_t)__double2hiint(bl),
                                              (uint32_t)__double2loint(bh), (uint32_t)__double2hiint(bh));
            }
        }
    }
#else
#pragma unroll 1
    for (int r = 0; r < 30; r++) {
        if (r < 4 || r >= 26) {
#pragma unroll
            for (int i = 0; i < 12; i++) s[i] = sbox7(s[i]);
        } else {
            s[0] = sbox7(s[0]);
        }
#if PCS_MDS_FP64
        mds_layer_fp64(s, &ROUND_ADD_D[24 * r]);
#else
        mds_layer<1>(s, &ROUND_ADD[12 * r]);
#endif
    }
#endif
    if (CANON_OUT) {
#pragma unroll
        for (int i = 0; i < 12; i++) s[i] = gl::canon(s[i]);
    }
}

}  // namespace pcs
