// Merkle node levels (two_to_one) and path extraction.
//
// Reference: compress (plonky2/src/hash/hashing.rs:98-115): state = [l0..3, r0..3, 0,0,0,0],
// permute, take 4.  fill_subtree (merkle_tree.rs:69-96) recurses with rayon::join; here each
// tree level is one launch with one thread per node, writing into the reference digest layout.
// prove(): merkle_tree.rs:173-207.
#include "hash_common.cuh"
#include "poseidon_coop.cuh"

namespace pcs {

// tree_mode: node k of `level` from its two children in the tree layout.
// plain mode: out[k] = two_to_one(l[k], r[k])  (a = l, b = r, out = c).
__global__ void __launch_bounds__(HASH_THREADS)
k_compress(uint64_t* __restrict__ a, uint64_t* __restrict__ b, uint64_t* __restrict__ c, int tree_mode,
           unsigned lg_sub, unsigned level, size_t n) {
    size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const ulonglong2 *l, *r;
    uint64_t* out;
    if (tree_mode) {
        // children 2k, 2k+1 of level-1 are stored side by side (64 contiguous bytes)
        l = reinterpret_cast<const ulonglong2*>(digest_slot(a, b, lg_sub, level - 1, 2 * k));
        r = l + 2;
        out = digest_slot(a, b, lg_sub, level, k);
    } else {
        l = reinterpret_cast<const ulonglong2*>(a + 4 * k);
        r = reinterpret_cast<const ulonglong2*>(b + 4 * k);
        out = c + 4 * k;
    }
    uint64_t s[12];
    ulonglong2 v0 = l[0], v1 = l[1], v2 = r[0], v3 = r[1];
    s[0] = v0.x; s[1] = v0.y; s[2] = v1.x; s[3] = v1.y;   // loose inputs are fine: only the digest is canonicalised
    s[4] = v2.x; s[5] = v2.y; s[6] = v3.x; s[7] = v3.y;
    s[8] = s[9] = s[10] = s[11] = 0;
    poseidon12<false>(s);
    ulonglong2* o = reinterpret_cast<ulonglong2*>(out);
    o[0] = make_ulonglong2(gl::canon(s[0]), gl::canon(s[1]));
    o[1] = make_ulonglong2(gl::canon(s[2]), gl::canon(s[3]));
}

// The TOP of every cap subtree in one launch: levels [level_first, lg_sub] with at most 256 nodes per subtree at level_first.
// One CTA per subtree; a level's nodes are written to the digest array by this CTA and read back by it for the next level
// (__syncthreads orders the global accesses inside the block).  Replaces up to 9 latency-bound launches of k_compress --
// what is left of the tree once the levels are too small to fill the GPU (0.3-0.4 ms of a sharded commit, most of a small one).
__global__ void __launch_bounds__(256)
k_compress_top(uint64_t* __restrict__ digests, uint64_t* __restrict__ cap, unsigned lg_sub, unsigned level_first) {
    const size_t sub = blockIdx.x;
    for (unsigned level = level_first; level <= lg_sub; level++) {
        const unsigned per = 1u << (lg_sub - level);      // nodes of this subtree at this level
        if (threadIdx.x < per) {
            const size_t k = (sub << (lg_sub - level)) + threadIdx.x;
            const ulonglong2* l = reinterpret_cast<const ulonglong2*>(digest_slot(digests, cap, lg_sub, level - 1, 2 * k));
            const ulonglong2* r = l + 2;
            ulonglong2 v0 = l[0], v1 = l[1], v2 = r[0], v3 = r[1];
            uint64_t s[12];
            s[0] = v0.x; s[1] = v0.y; s[2] = v1.x; s[3] = v1.y;
            s[4] = v2.x; s[5] = v2.y; s[6] = v3.x; s[7] = v3.y;
            s[8] = s[9] = s[10] = s[11] = 0;
            poseidon12<false>(s);
            ulonglong2* o = reinterpret_cast<ulonglong2*>(digest_slot(digests, cap, lg_sub, level, k));
            o[0] = make_ulonglong2(gl::canon(s[0]), gl::canon(s[1]));
            o[1] = make_ulonglong2(gl::canon(s[2]), gl::canon(s[3]));
        }
        __syncthreads();
    }
}

// Latency forms (<= COOP_MAX_PERMS nodes in a level): one NODE per half-warp.
// compress of (l, r): lanes 0..3 hold l, lanes 4..7 hold r, the rest 0 (hashing.rs:98-115)
__device__ __forceinline__ uint64_t coop_compress(const uint64_t* l, const uint64_t* r, unsigned i, const uint64_t* rc) {
    uint64_t s = i < 4 ? l[i] : (i < 8 ? r[i - 4] : 0);
    return gl::canon(poseidon12_coop(s, i, rc));
}

__global__ void __launch_bounds__(128)
k_compress_coop(uint64_t* __restrict__ a, uint64_t* __restrict__ b, uint64_t* __restrict__ c, int tree_mode, unsigned lg_sub,
                unsigned level, size_t n) {
    __shared__ uint64_t rc[360];
    coop_load_rc(rc);
    const size_t k0 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) / COOP_LANES;
    const unsigned i = threadIdx.x & (COOP_LANES - 1);
    const bool active = k0 < n;
    const size_t k = active ? k0 : 0;
    const uint64_t *l, *r;
    uint64_t* out;
    if (tree_mode) {
        l = digest_slot(a, b, lg_sub, level - 1, 2 * k);
        r = l + 4;
        out = digest_slot(a, b, lg_sub, level, k);
    } else {
        l = a + 4 * k;
        r = b + 4 * k;
        out = c + 4 * k;
    }
    const uint64_t s = coop_compress(l, r, i, rc);
    if (active && i < 4) out[i] = s;
}

// The top of every cap subtree in one launch, latency form: levels [level_first, lg_sub], <= 64 nodes per subtree at
// level_first, one CTA (1024 threads = 64 half-warps) per subtree.
__global__ void __launch_bounds__(1024)
k_compress_top_coop(uint64_t* __restrict__ digests, uint64_t* __restrict__ cap, unsigned lg_sub, unsigned level_first) {
    __shared__ uint64_t rc[360];
    coop_load_rc(rc);
    const size_t sub = blockIdx.x;
    const unsigned g = threadIdx.x / COOP_LANES, i = threadIdx.x & (COOP_LANES - 1);
    for (unsigned level = level_first; level <= lg_sub; level++) {
        const unsigned per = 1u << (lg_sub - level);
        if ((g & ~1u) < per) {        // whole warps stay together: both half-warps of a warp run the shuffles
            const bool active = g < per;
            const size_t k = (sub << (lg_sub - level)) + (active ? g : 0);
            const uint64_t* l = digest_slot(digests, cap, lg_sub, level - 1, 2 * k);
            const uint64_t s = coop_compress(l, l + 4, i, rc);
            if (active && i < 4) digest_slot(digests, cap, lg_sub, level, k)[i] = s;
        }
        __syncthreads();
    }
}

cudaError_t launch_node_top(uint64_t* digests, uint64_t* cap, unsigned lg_sub, unsigned level_first, size_t n_subtrees,
                            cudaStream_t st) {
    if (n_subtrees == 0 || level_first > lg_sub) return cudaSuccess;
    if (lg_sub - level_first > 8) return cudaErrorInvalidValue;
    if (lg_sub - level_first <= 6) {    // <= 64 nodes per subtree: the latency form
        const unsigned nodes = 1u << (lg_sub - level_first);
        const unsigned threads = nodes * COOP_LANES < 32 ? 32 : nodes * COOP_LANES;
        k_compress_top_coop<<<(unsigned)n_subtrees, threads, 0, st>>>(digests, cap, lg_sub, level_first);
        return cudaGetLastError();
    }
    const unsigned threads = 1u << (lg_sub - level_first) < 32 ? 32 : 1u << (lg_sub - level_first);
    k_compress_top<<<(unsigned)n_subtrees, threads, 0, st>>>(digests, cap, lg_sub, level_first);
    return cudaGetLastError();
}

// MerkleTree::prove: one thread per layer copies the sibling digest.
__global__ void k_prove(const uint64_t* __restrict__ digests, unsigned lg_sub, size_t leaf_index,
                        uint64_t* __restrict__ siblings) {
    unsigned i = threadIdx.x;
    if (i >= lg_sub) return;
    size_t s = leaf_index >> lg_sub;
    size_t q = leaf_index & (((size_t)1 << lg_sub) - 1);
    size_t node = (q >> i) ^ 1;  // sibling at level i
    size_t sub_len = 2 * (((size_t)1 << lg_sub) - 1);
    size_t idx = 2 * (((node >> 1) << (i + 1)) + ((size_t)1 << i) - 1) + (node & 1);
    const uint64_t* src = digests + (s * sub_len + idx) * 4;
#pragma unroll
    for (int k = 0; k < 4; k++) siblings[4 * i + k] = src[k];
}

cudaError_t launch_node_level(uint64_t* digests, uint64_t* cap, unsigned lg_sub, unsigned level, size_t n_nodes,
                              cudaStream_t st) {
    if (n_nodes == 0) return cudaSuccess;
    if (n_nodes <= COOP_MAX_PERMS)
        k_compress_coop<<<grid_for(n_nodes * COOP_LANES, 128), 128, 0, st>>>(digests, cap, nullptr, 1, lg_sub, level, n_nodes);
    else
        k_compress<<<grid_for(n_nodes, HASH_THREADS), HASH_THREADS, 0, st>>>(digests, cap, nullptr, 1, lg_sub, level, n_nodes);
    return cudaGetLastError();
}

cudaError_t launch_two_to_one(const uint64_t* l, const uint64_t* r, size_t n, uint64_t* out, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    if (n <= COOP_MAX_PERMS)
        k_compress_coop<<<grid_for(n * COOP_LANES, 128), 128, 0, st>>>(const_cast<uint64_t*>(l), const_cast<uint64_t*>(r), out, 0, 0, 0, n);
    else
        k_compress<<<grid_for(n, HASH_THREADS), HASH_THREADS, 0, st>>>(const_cast<uint64_t*>(l), const_cast<uint64_t*>(r),
                                                                      out, 0, 0, 0, n);
    return cudaGetLastError();
}

// The same for many leaves: block b copies the path of leaf_indices[b] to siblings[b][lg_sub][4].
__global__ void k_prove_many(const uint64_t* __restrict__ digests, unsigned lg_sub, const uint64_t* __restrict__ leaf_indices,
                             uint64_t* __restrict__ siblings) {
    unsigned i = threadIdx.x;
    if (i >= lg_sub) return;
    size_t leaf_index = leaf_indices[blockIdx.x];
    size_t s = leaf_index >> lg_sub;
    size_t q = leaf_index & (((size_t)1 << lg_sub) - 1);
    size_t node = (q >> i) ^ 1;
    size_t sub_len = 2 * (((size_t)1 << lg_sub) - 1);
    size_t idx = 2 * (((node >> 1) << (i + 1)) + ((size_t)1 << i) - 1) + (node & 1);
    const uint64_t* src = digests + (s * sub_len + idx) * 4;
    uint64_t* dst = siblings + ((size_t)blockIdx.x * lg_sub + i) * 4;
#pragma unroll
    for (int k = 0; k < 4; k++) dst[k] = src[k];
}

cudaError_t launch_prove_many(const uint64_t* digests, unsigned lg_sub, const uint64_t* leaf_indices_dev, size_t n,
                              uint64_t* siblings, cudaStream_t st) {
    if (lg_sub == 0 || n == 0) return cudaSuccess;
    k_prove_many<<<(unsigned)n, 64, 0, st>>>(digests, lg_sub, leaf_indices_dev, siblings);
    return cudaGetLastError();
}

cudaError_t launch_prove(const uint64_t* digests, unsigned lg_sub, size_t leaf_index, uint64_t* siblings,
                         cudaStream_t st) {
    if (lg_sub == 0) return cudaSuccess;
    k_prove<<<1, 64, 0, st>>>(digests, lg_sub, leaf_index, siblings);
    return cudaGetLastError();
}

}  // namespace pcs
