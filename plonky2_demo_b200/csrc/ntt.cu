// Batched Goldilocks NTT / coset low-degree extension for sm_100a.
//
// What the reference computes (per polynomial, on the CPU):
//   PolynomialCoeffs::lde          field/src/polynomial/mod.rs:201-203   zero-pad d -> N = d*2^r
//   coset_fft_with_options         field/src/polynomial/mod.rs:282-295   c_i *= 7^i, then
//   fft_classic (radix-2 DIT)      field/src/fft.rs:169-206              bit-reverse + lg N stages
//   transpose + reverse_index_bits plonky2/src/fri/oracle.rs:83-84       -> leaf order
//   ifft_with_options              field/src/fft.rs:72-95
//
// What this file computes instead (same values, different algorithm -- SURVEY 8a identity):
//   the length-N transform with N - d trailing zero coefficients, run as a natural-order-in /
//   bit-reversed-order-out Cooley-Tukey NTT whose twiddles depend only on the butterfly BLOCK:
//       stage sigma, block B:  (u, v) <- (u + w v, u - w v),
//       w = shift^(N / 2^(sigma+1)) * omega_N^((N / 2^(sigma+1)) * brev_sigma(B))
//   The first r stages see v = 0 and are plain copies, so coset c in [0, 2^r) is an independent
//   length-d transform of the SAME coefficients whose output IS the contiguous leaf range
//   [c*d, (c+1)*d): no zero padding, no 7^i pass, no transpose, no bit-reversal pass.
//   The coset shift is absorbed into the twiddles (zero extra multiplies per element).
//
// Decomposition: lg d stages are split into passes of <= 10 stages.  A pass handles, per CTA, a
// tile of R = 2^nb "rows" (the nb index bits it transforms) x C "columns" (index bits below the
// pass, or different polynomials) staged in shared memory.  Inside a tile every twiddle factors as
//       w(i, q) = gamma_i(G) * psi[q],   gamma_{i-1} = gamma_i^2,
// with G = (coset, high index bits) fixed per tile: one table lookup (Gamma[G]), nb-1 squarings and
// 2^nb - 1 products per TILE give all twiddles, shared by all C columns.
//
// Roofline: HBM traffic is 8 B/elem per pass per direction, but a butterfly costs ~30 integer
// instructions (4 IMAD.WIDE + reduction + add + sub), so the kernel is bound by the integer pipes
// unless the tile stages are register-blocked; see DESIGN.md.
#include <map>
#include <tuple>
#include <vector>

#include "common.cuh"
#include "gl64.cuh"

namespace pcs {

constexpr int NTT_MAX_NB = 10;          // stages per pass
constexpr int NTT_TILE_LOG = 13;        // 8192 elements (64 KiB) per tile
constexpr int NTT_THREADS = 512;

struct NttPass {
    unsigned s0, nb;   // stages [s0, s0+nb)
    unsigned L;        // index bits below the pass
    uint64_t* gamma;   // [2^(r+s0)]  Gamma[G] = (shift * root_N^brev(G))^(2^L)
    uint64_t* psi;     // [2^(nb-1)]  psi[q] = omega_{2^nb}^{brev_{nb-1}(q)}
};

struct NttPlan {
    unsigned lg_d, r;
    bool inverse;
    uint64_t shift;
    uint64_t scale;  // applied in the last pass (1/d for the inverse transform, else 1)
    std::vector<NttPass> passes;
};

// -------------------------------------------------------------------------------------------------
// plan construction (device-side table fill)
// -------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t brev32(uint32_t x, unsigned bits) { return bits ? __brev(x) >> (32 - bits) : 0; }

__global__ void k_fill_gamma(uint64_t* gamma, unsigned gbits, uint64_t root_n, uint64_t shift, unsigned L) {
    uint32_t G = blockIdx.x * blockDim.x + threadIdx.x;
    if (G >= (1u << gbits)) return;
    uint64_t beta = gl::mul(shift, gl::pow(root_n, brev32(G, gbits)));
    for (unsigned i = 0; i < L; i++) beta = gl::sqr(beta);
    gamma[G] = beta;
}

__global__ void k_fill_psi(uint64_t* psi, unsigned nb, uint64_t root_2nb) {
    uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= (1u << (nb - 1))) return;
    psi[q] = gl::pow(root_2nb, brev32(q, nb - 1));
}

static uint64_t h_mul(uint64_t a, uint64_t b) {
    unsigned __int128 x = (unsigned __int128)a * b;
    return (uint64_t)(x % gl::P);
}
static uint64_t h_pow(uint64_t b, uint64_t e) {
    uint64_t acc = 1;
    while (e) {
        if (e & 1) acc = h_mul(acc, b);
        b = h_mul(b, b);
        e >>= 1;
    }
    return acc;
}

static std::map<std::tuple<unsigned, unsigned, bool, uint64_t>, NttPlan*> g_plans;

static std::vector<unsigned> split_stages(unsigned lg_d) {
    // equal-ish passes of at most NTT_MAX_NB stages; the FIRST pass gets the remainder so the
    // contiguous last pass is the biggest.
    std::vector<unsigned> v;
    if (lg_d == 0) return v;
    unsigned n = (lg_d + NTT_MAX_NB - 1) / NTT_MAX_NB;
    unsigned base = lg_d / n, extra = lg_d % n;
    for (unsigned i = 0; i < n; i++) v.push_back(base + (i >= n - extra ? 1 : 0));
    return v;
}

NttPlan* ntt_plan_get(unsigned lg_d, unsigned r, bool inverse, uint64_t shift, cudaStream_t st) {
    auto key = std::make_tuple(lg_d, r, inverse, shift);
    auto it = g_plans.find(key);
    if (it != g_plans.end()) return it->second;
    unsigned lg_n = lg_d + r;
    // primitive_root_of_unity(lg_n) = POWER_OF_TWO_GENERATOR^(2^(32-lg_n))   types.rs:268-272
    uint64_t root = h_pow(1753635133440165772ULL, 1ULL << (32 - lg_n));
    if (inverse) root = h_pow(root, gl::P - 2);
    NttPlan* p = new NttPlan();
    p->lg_d = lg_d;
    p->r = r;
    p->inverse = inverse;
    p->shift = shift % gl::P;
    p->scale = inverse ? gl::P - ((gl::P - 1) >> lg_d) : 1;  // inverse_2exp, types.rs:227-262
    unsigned s0 = 0;
    for (unsigned nb : split_stages(lg_d)) {
        NttPass ps;
        ps.s0 = s0;
        ps.nb = nb;
        ps.L = lg_d - s0 - nb;
        unsigned gbits = r + s0;
        if (cudaMalloc(&ps.gamma, sizeof(uint64_t) << gbits) != cudaSuccess) return nullptr;
        if (cudaMalloc(&ps.psi, sizeof(uint64_t) << (nb - 1)) != cudaSuccess) return nullptr;
        size_t ng = (size_t)1 << gbits;
        k_fill_gamma<<<(unsigned)((ng + 255) / 256), 256, 0, st>>>(ps.gamma, gbits, root, p->shift, ps.L);
        uint64_t root_2nb = h_pow(root, 1ULL << (lg_n - nb));
        size_t nq = (size_t)1 << (nb - 1);
        k_fill_psi<<<(unsigned)((nq + 255) / 256), 256, 0, st>>>(ps.psi, nb, root_2nb);
        p->passes.push_back(ps);
        s0 += nb;
    }
    g_plans[key] = p;
    return p;
}

void ntt_plans_free() {
    for (auto& kv : g_plans) {
        for (auto& ps : kv.second->passes) {
            cudaFree(ps.gamma);
            cudaFree(ps.psi);
        }
        delete kv.second;
    }
    g_plans.clear();
}

// -------------------------------------------------------------------------------------------------
// the pass kernel
// -------------------------------------------------------------------------------------------------
struct PassArgs {
    const uint64_t* in;
    uint64_t* out;
    size_t in_poly_stride, out_poly_stride;
    size_t in_coset_stride, out_coset_stride;  // in: 0 for the first pass (all cosets read the coefficients)
    const uint64_t* gamma;
    const uint64_t* psi;
    uint32_t n_polys;
    uint32_t r, s0, nb, L, lg_d;
    uint32_t lc;        // log2(columns per tile)
    uint32_t lcl;       // log2(columns per tile taken from the low index bits) = min(L, lc)
    uint32_t col_groups_per_poly_block;  // 2^(L - lcl)
    uint64_t scale;     // multiply outputs (last pass of the inverse transform), 1 = none
    int canon_in;       // canonicalise inputs (first pass reads caller data)
};

// Tile element (row rho, column col) lives at smem[rho * pitch + col], pitch odd.
__global__ void __launch_bounds__(NTT_THREADS) k_ntt_pass(const PassArgs a) {
    extern __shared__ uint64_t smem[];
    const uint32_t R = 1u << a.nb, C = 1u << a.lc;
    const uint32_t pitch = C | 1u;
    uint64_t* tw = smem;            // [R] (index 0 unused)
    uint64_t* tile = smem + R;      // [R][pitch]

    // ---- which tile ----
    // 1-D grid, G fastest: neighbouring CTAs are the 2^r cosets of the same input tile (L2 reuse)
    const uint32_t G = blockIdx.x & ((1u << (a.r + a.s0)) - 1);  // (coset, high bits) prefix, r + s0 bits
    const uint32_t c = G >> a.s0, H = G & ((1u << a.s0) - 1);
    const uint32_t cg = blockIdx.x >> (a.r + a.s0);               // column group
    const uint32_t lgroup = cg % a.col_groups_per_poly_block;   // which slice of the low index bits
    const uint32_t pgroup = cg / a.col_groups_per_poly_block;   // which block of polynomials
    const uint32_t polys_per_tile = 1u << (a.lc - a.lcl);
    const uint32_t l_base = lgroup << a.lcl;
    const uint32_t poly_base = pgroup * polys_per_tile;
    const size_t row_stride = (size_t)1 << a.L;
    const size_t h_off = (size_t)H << (a.lg_d - a.s0);

    // ---- twiddles: tw[2^i + q] = gamma_i * psi[q] ----
    {
        uint64_t g = a.gamma[G];  // gamma_{nb-1}
        // thread t computes entries of every stage where q = t is in range
        for (int i = (int)a.nb - 1; i >= 0; i--) {
            for (uint32_t q = threadIdx.x; q < (1u << i); q += blockDim.x) tw[(1u << i) + q] = gl::mul(g, a.psi[q]);
            g = gl::sqr(g);
        }
    }

    // ---- load ----
    const bool row_major_threads = (a.lcl == 0);  // columns are polynomials: make consecutive threads walk rows
    for (uint32_t e = threadIdx.x; e < R * C; e += blockDim.x) {
        uint32_t rho, col;
        if (row_major_threads) { rho = e & (R - 1); col = e >> a.nb; }
        else { col = e & (C - 1); rho = e >> a.lc; }
        uint32_t lcol = col & ((1u << a.lcl) - 1), pcol = col >> a.lcl;
        uint32_t poly = poly_base + pcol;
        uint64_t v = 0;
        if (poly < a.n_polys) {
            size_t idx = (size_t)poly * a.in_poly_stride + (size_t)c * a.in_coset_stride + h_off +
                         (size_t)rho * row_stride + l_base + lcol;
            v = a.in[idx];
            if (a.canon_in) v = gl::canon(v);
        }
        tile[rho * pitch + col] = v;
    }
    __syncthreads();

    // ---- nb radix-2 stages in shared memory ----
    for (uint32_t i = 0; i < a.nb; i++) {
        const uint32_t lg_half = a.nb - 1 - i, half = 1u << lg_half;
        for (uint32_t e = threadIdx.x; e < (R >> 1) * C; e += blockDim.x) {
            uint32_t col, pr;
            if (row_major_threads) { pr = e & ((R >> 1) - 1); col = e >> (a.nb - 1); }
            else { col = e & (C - 1); pr = e >> a.lc; }
            uint32_t q = pr >> lg_half, lowb = pr & (half - 1);
            uint32_t r0 = (q << (lg_half + 1)) + lowb, r1 = r0 + half;
            uint64_t w = tw[(1u << i) + q];
            uint64_t u = tile[r0 * pitch + col];
            uint64_t v = gl::mul(tile[r1 * pitch + col], w);
            tile[r0 * pitch + col] = gl::add(u, v);
            tile[r1 * pitch + col] = gl::sub(u, v);
        }
        __syncthreads();
    }

    // ---- store ----
    for (uint32_t e = threadIdx.x; e < R * C; e += blockDim.x) {
        uint32_t rho, col;
        if (row_major_threads) { rho = e & (R - 1); col = e >> a.nb; }
        else { col = e & (C - 1); rho = e >> a.lc; }
        uint32_t lcol = col & ((1u << a.lcl) - 1), pcol = col >> a.lcl;
        uint32_t poly = poly_base + pcol;
        if (poly < a.n_polys) {
            uint64_t v = tile[rho * pitch + col];
            if (a.scale != 1) v = gl::mul(v, a.scale);
            size_t idx = (size_t)poly * a.out_poly_stride + (size_t)c * a.out_coset_stride + h_off +
                         (size_t)rho * row_stride + l_base + lcol;
            a.out[idx] = v;
        }
    }
}

// d == 1: every output of the coset LDE equals the single coefficient.
__global__ void k_broadcast_const(const uint64_t* in, size_t in_stride, uint64_t* out, size_t out_stride, size_t w,
                                  size_t n, uint64_t scale) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= w * n) return;
    size_t j = i / n, k = i % n;
    out[j * out_stride + k] = gl::mul(gl::canon(in[j * in_stride]), scale);
}

static cudaError_t run_passes(const NttPlan* plan, const uint64_t* in, size_t in_stride, uint64_t* out,
                              size_t out_stride, size_t w, cudaStream_t st) {
    const unsigned lg_d = plan->lg_d, r = plan->r;
    const size_t d = (size_t)1 << lg_d;
    if (w == 0) return cudaSuccess;
    if (lg_d == 0) {
        size_t n = (size_t)1 << r;
        k_broadcast_const<<<(unsigned)((w * n + 255) / 256), 256, 0, st>>>(in, in_stride, out, out_stride, w, n,
                                                                          plan->scale);
        return cudaGetLastError();
    }
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(k_ntt_pass, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    for (size_t pi = 0; pi < plan->passes.size(); pi++) {
        const NttPass& ps = plan->passes[pi];
        PassArgs a;
        bool first = pi == 0, last = pi + 1 == plan->passes.size();
        a.in = first ? in : out;
        a.out = out;
        a.in_poly_stride = first ? in_stride : out_stride;
        a.out_poly_stride = out_stride;
        a.in_coset_stride = first ? 0 : d;
        a.out_coset_stride = d;
        a.gamma = ps.gamma;
        a.psi = ps.psi;
        a.n_polys = (uint32_t)w;
        a.r = r; a.s0 = ps.s0; a.nb = ps.nb; a.L = ps.L; a.lg_d = lg_d;
        a.lc = NTT_TILE_LOG - ps.nb;
        a.lcl = ps.L < a.lc ? ps.L : a.lc;
        a.col_groups_per_poly_block = 1u << (ps.L - a.lcl);
        a.scale = last ? plan->scale : 1;
        a.canon_in = first ? 1 : 0;
        uint32_t polys_per_tile = 1u << (a.lc - a.lcl);
        uint32_t pgroups = (uint32_t)((w + polys_per_tile - 1) / polys_per_tile);
        size_t n_blocks = ((size_t)a.col_groups_per_poly_block * pgroups) << (r + ps.s0);
        if (n_blocks > 0x7fffffffULL) return cudaErrorInvalidConfiguration;
        dim3 grid((unsigned)n_blocks);
        uint32_t R = 1u << ps.nb, C = 1u << a.lc;
        size_t smem = (size_t)(R + R * (C | 1u)) * sizeof(uint64_t);
        k_ntt_pass<<<grid, NTT_THREADS, smem, st>>>(a);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

cudaError_t ntt_lde(const NttPlan* plan, const uint64_t* coeffs, size_t in_stride, uint64_t* out, size_t out_stride,
                    size_t w, cudaStream_t st) {
    return run_passes(plan, coeffs, in_stride, out, out_stride, w, st);
}

cudaError_t ntt_inverse_bitrev(const NttPlan* plan, const uint64_t* values, size_t in_stride, uint64_t* out,
                               size_t out_stride, size_t w, cudaStream_t st) {
    return run_passes(plan, values, in_stride, out, out_stride, w, st);
}

// -------------------------------------------------------------------------------------------------
// data-movement helpers
// -------------------------------------------------------------------------------------------------
__global__ void k_bitrev_permute(const uint64_t* __restrict__ in, size_t in_stride, uint64_t* __restrict__ out,
                                 size_t out_stride, unsigned lg_n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t j = blockIdx.y;
    if (i >= ((size_t)1 << lg_n)) return;
    size_t bi = lg_n ? (size_t)(__brevll((unsigned long long)i) >> (64 - lg_n)) : 0;
    out[j * out_stride + bi] = in[j * in_stride + i];
}

cudaError_t launch_bitrev_permute(const uint64_t* in, size_t in_stride, uint64_t* out, size_t out_stride, size_t w,
                                  unsigned lg_n, cudaStream_t st) {
    if (w == 0) return cudaSuccess;
    size_t n = (size_t)1 << lg_n;
    for (size_t j0 = 0; j0 < w; j0 += 65535) {
        size_t wj = w - j0 < 65535 ? w - j0 : 65535;
        dim3 grid((unsigned)((n + 255) / 256), (unsigned)wj);
        k_bitrev_permute<<<grid, 256, 0, st>>>(in + j0 * in_stride, in_stride, out + j0 * out_stride, out_stride, lg_n);
    }
    return cudaGetLastError();
}

// out[c][r] = in[r][c]; 32x32 tiles through shared memory, both sides coalesced.
__global__ void k_transpose(const uint64_t* __restrict__ in, size_t in_pitch, uint64_t* __restrict__ out,
                            size_t out_pitch, size_t rows, size_t cols) {
    __shared__ uint64_t t[32][33];
    size_t c0 = (size_t)blockIdx.x * 32, r0 = (size_t)blockIdx.y * 32;
    for (int k = threadIdx.y; k < 32; k += blockDim.y) {
        size_t r = r0 + k, c = c0 + threadIdx.x;
        if (r < rows && c < cols) t[k][threadIdx.x] = in[r * in_pitch + c];
    }
    __syncthreads();
    for (int k = threadIdx.y; k < 32; k += blockDim.y) {
        size_t c = c0 + k, r = r0 + threadIdx.x;
        if (r < rows && c < cols) out[c * out_pitch + r] = t[threadIdx.x][k];
    }
}

cudaError_t launch_transpose(const uint64_t* in, size_t in_pitch, uint64_t* out, size_t out_pitch, size_t rows,
                             size_t cols, cudaStream_t st) {
    if (rows == 0 || cols == 0) return cudaSuccess;
    // grid.y limit 65535: put the long dimension on x
    size_t bx = (cols + 31) / 32, by = (rows + 31) / 32;
    if (by > 65535) {
        // process in row slabs
        size_t slab_rows = (size_t)65535 * 32;
        for (size_t rr = 0; rr < rows; rr += slab_rows) {
            size_t nr = rows - rr < slab_rows ? rows - rr : slab_rows;
            dim3 grid((unsigned)bx, (unsigned)((nr + 31) / 32));
            k_transpose<<<grid, dim3(32, 8), 0, st>>>(in + rr * in_pitch, in_pitch, out + rr, out_pitch, nr, cols);
        }
    } else {
        dim3 grid((unsigned)bx, (unsigned)by);
        k_transpose<<<grid, dim3(32, 8), 0, st>>>(in, in_pitch, out, out_pitch, rows, cols);
    }
    return cudaGetLastError();
}

__global__ void k_gather_rows(const uint64_t* __restrict__ cols, size_t col_stride, uint32_t width,
                              const uint64_t* __restrict__ idx, size_t n_idx, uint64_t* __restrict__ out) {
    size_t k = blockIdx.x;
    if (k >= n_idx) return;
    size_t leaf = idx[k];
    for (uint32_t j = threadIdx.x; j < width; j += blockDim.x) out[k * width + j] = cols[(size_t)j * col_stride + leaf];
}

cudaError_t launch_gather_rows(const uint64_t* cols, size_t col_stride, uint32_t width, const uint64_t* idx,
                               size_t n_idx, uint64_t* out, cudaStream_t st) {
    if (n_idx == 0 || width == 0) return cudaSuccess;
    k_gather_rows<<<(unsigned)n_idx, 128, 0, st>>>(cols, col_stride, width, idx, n_idx, out);
    return cudaGetLastError();
}

__global__ void k_canon(uint64_t* data, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) data[i] = gl::canon(data[i]);
}

cudaError_t launch_canonicalize(uint64_t* data, size_t n, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    k_canon<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(data, n);
    return cudaGetLastError();
}

}  // namespace pcs
