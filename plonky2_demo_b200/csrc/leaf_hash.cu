// Sponge leaf hashing over poly-major (column-major) LDE data: one thread per leaf, every load
// coalesced across the warp (thread i reads element i of each column).
//
// Reference: hash_or_noop (plonky2/src/plonk/config.rs:55-66) -> hash_n_to_m_no_pad
// (plonky2/src/hash/hashing.rs:119-142): overwrite-mode sponge, rate 8, no padding; rows of <= 4
// elements are copied (canonicalised, zero padded), not hashed.  Digests land directly in the
// reference's interleaved layout (merkle_tree.rs:43-51), leaf level of fill_subtree (:69-96).
//
// Bound: integer pipes, NOT HBM: a 135-element leaf = 17 permutations per 1080 B read.
#include "hash_common.cuh"
#include "poseidon_coop.cuh"

namespace pcs {

// tree_mode: digest of leaf i goes to its tree slot (or cap); otherwise to out[4*i].
__global__ void __launch_bounds__(HASH_THREADS)
k_hash_cols(const uint64_t* __restrict__ cols, size_t col_stride, uint32_t width, size_t first, size_t n, int tree_mode,
            unsigned lg_sub, uint64_t* __restrict__ digests, uint64_t* __restrict__ cap) {
    size_t i = first + (size_t)blockIdx.x * blockDim.x + threadIdx.x;   // leaves [first, n) of the tree
    if (i >= n) return;
    uint64_t d[4];
    const uint64_t* p = cols + i;
    if (width <= 4) {  // hash_or_noop: not hashed
#pragma unroll
        for (int k = 0; k < 4; k++) d[k] = (uint32_t)k < width ? gl::canon(p[(size_t)k * col_stride]) : 0;
    } else {
        uint64_t s[12];
#pragma unroll
        for (int k = 0; k < 12; k++) s[k] = 0;
#pragma unroll 1
        for (uint32_t j = 0; j < width; j += 8) {
            // a short last chunk overwrites only the first (width - j) lanes (hashing.rs:128-131)
#pragma unroll
            // no canonicalisation on the way in or between permutations: the permutation takes loose values (any u64
            // congruent to the element) and the sponge state never leaves the registers; only the digest is canonical
            for (int k = 0; k < 8; k++)
                if (j + k < width) s[k] = __ldg(p + (size_t)(j + k) * col_stride);
            poseidon12<false>(s);
        }
#pragma unroll
        for (int k = 0; k < 4; k++) d[k] = gl::canon(s[k]);
    }
    uint64_t* slot = tree_mode ? digest_slot(digests, cap, lg_sub, 0, i) : digests + 4 * i;
    ulonglong2* o = reinterpret_cast<ulonglong2*>(slot);
    o[0] = make_ulonglong2(d[0], d[1]);
    o[1] = make_ulonglong2(d[2], d[3]);
}

// The same sponge absorbed in GROUPS of columns (a streaming commit: group g is hashed while group g+1 is still crossing
// PCIe / NVLink).  The overwrite-mode sponge is sequential in the columns, so a group boundary only has to fall on a multiple
// of the rate (8): the 12-lane state of every leaf is parked in HBM between groups ([12][n], lane-major so that loads and
// stores stay coalesced; 96 B per leaf against 64 B per absorbed chunk -- irrelevant for a kernel at 1 % of the HBM roofline).
//   FIRST: the state starts at zero (hashing.rs:124);  LAST: the digest (lanes 0..3, canonical) goes to the leaf's tree slot.
template <bool FIRST, bool LAST>
__global__ void __launch_bounds__(HASH_THREADS)
k_hash_cols_stream(const uint64_t* __restrict__ cols, size_t col_stride, uint32_t width, size_t n, uint64_t* __restrict__ state,
                   unsigned lg_sub, uint64_t* __restrict__ digests, uint64_t* __restrict__ cap) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t s[12];
#pragma unroll
    for (int k = 0; k < 12; k++) s[k] = FIRST ? 0 : state[(size_t)k * n + i];
    const uint64_t* p = cols + i;
#pragma unroll 1
    for (uint32_t j = 0; j < width; j += 8) {
#pragma unroll
        for (int k = 0; k < 8; k++)
            if (j + k < width) s[k] = __ldg(p + (size_t)(j + k) * col_stride);
        poseidon12<false>(s);
    }
    if (LAST) {
        ulonglong2* o = reinterpret_cast<ulonglong2*>(digest_slot(digests, cap, lg_sub, 0, i));
        o[0] = make_ulonglong2(gl::canon(s[0]), gl::canon(s[1]));
        o[1] = make_ulonglong2(gl::canon(s[2]), gl::canon(s[3]));
    } else {
#pragma unroll
        for (int k = 0; k < 12; k++) state[(size_t)k * n + i] = s[k];
    }
}

// Latency form for small trees (<= COOP_MAX_PERMS leaves): one LEAF per half-warp, the 12 state lanes on 12 threads.
__global__ void __launch_bounds__(128)
k_hash_cols_coop(const uint64_t* __restrict__ cols, size_t col_stride, uint32_t width, size_t first, size_t n, int tree_mode,
                 unsigned lg_sub, uint64_t* __restrict__ digests, uint64_t* __restrict__ cap) {
    __shared__ uint64_t rc[360];
    coop_load_rc(rc);
    const size_t leaf = first + ((size_t)blockIdx.x * blockDim.x + threadIdx.x) / COOP_LANES;
    const unsigned i = threadIdx.x & (COOP_LANES - 1);
    const bool active = leaf < n;                       // uniform over the half-warp; idle half-warps still run the shuffles
    const uint64_t* p = cols + (active ? leaf : first);
    uint64_t s = 0;
    if (width <= 4) {                                   // hash_or_noop: not hashed
        s = i < width ? gl::canon(p[(size_t)i * col_stride]) : 0;
    } else {
#pragma unroll 1
        for (uint32_t j = 0; j < width; j += 8) {
            if (i < 8 && j + i < width) s = __ldg(p + (size_t)(j + i) * col_stride);
            s = poseidon12_coop(s, i, rc);
        }
        s = gl::canon(s);
    }
    if (active && i < 4) {
        uint64_t* slot = tree_mode ? digest_slot(digests, cap, lg_sub, 0, leaf) : digests + 4 * leaf;
        slot[i] = s;
    }
}

cudaError_t launch_leaf_hash_group(const uint64_t* cols, size_t col_stride, uint32_t width, size_t n_leaves, bool first, bool last,
                                   uint64_t* state, unsigned lg_sub, uint64_t* digests, uint64_t* cap, cudaStream_t st) {
    if (n_leaves == 0) return cudaSuccess;
    const unsigned grid = grid_for(n_leaves, HASH_THREADS);
    if (first && last)
        k_hash_cols_stream<true, true><<<grid, HASH_THREADS, 0, st>>>(cols, col_stride, width, n_leaves, state, lg_sub, digests, cap);
    else if (first)
        k_hash_cols_stream<true, false><<<grid, HASH_THREADS, 0, st>>>(cols, col_stride, width, n_leaves, state, lg_sub, digests, cap);
    else if (last)
        k_hash_cols_stream<false, true><<<grid, HASH_THREADS, 0, st>>>(cols, col_stride, width, n_leaves, state, lg_sub, digests, cap);
    else
        k_hash_cols_stream<false, false><<<grid, HASH_THREADS, 0, st>>>(cols, col_stride, width, n_leaves, state, lg_sub, digests, cap);
    return cudaGetLastError();
}

cudaError_t launch_leaf_hash_cols(const uint64_t* cols, size_t col_stride, uint32_t width, size_t n_leaves,
                                  unsigned lg_sub, uint64_t* digests, uint64_t* cap, cudaStream_t st,
                                  size_t first_leaf, size_t leaf_count) {
    if (leaf_count == (size_t)-1) leaf_count = n_leaves - first_leaf;
    if (leaf_count == 0) return cudaSuccess;
    if (leaf_count <= COOP_MAX_PERMS)
        k_hash_cols_coop<<<grid_for(leaf_count * COOP_LANES, 128), 128, 0, st>>>(cols, col_stride, width, first_leaf,
                                                                                first_leaf + leaf_count, 1, lg_sub, digests, cap);
    else
        k_hash_cols<<<grid_for(leaf_count, HASH_THREADS), HASH_THREADS, 0, st>>>(cols, col_stride, width, first_leaf,
                                                                                first_leaf + leaf_count, 1, lg_sub, digests, cap);
    return cudaGetLastError();
}

cudaError_t launch_hash_cols_plain(const uint64_t* cols, size_t col_stride, uint32_t width, size_t n,
                                   uint64_t* out, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    if (n <= COOP_MAX_PERMS)
        k_hash_cols_coop<<<grid_for(n * COOP_LANES, 128), 128, 0, st>>>(cols, col_stride, width, 0, n, 0, 0, out, nullptr);
    else
        k_hash_cols<<<grid_for(n, HASH_THREADS), HASH_THREADS, 0, st>>>(cols, col_stride, width, 0, n, 0, 0, out, nullptr);
    return cudaGetLastError();
}

}  // namespace pcs
