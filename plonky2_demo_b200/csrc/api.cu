// C ABI (include/pcs.h) of the polynomial-commitment engine: context, the fused
// PolynomialBatch::from_coeffs / from_values path (plonky2/src/fri/oracle.rs:43-98) and the
// accessors the reference's consumers need.  No CPU fallback anywhere in this file.
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

#include "common.cuh"

namespace pcs {

static thread_local std::string g_err;
void set_error(const std::string& msg) { g_err = msg; }

struct PendingEvents { cudaEvent_t ev[6]; std::vector<cudaEvent_t> absorb_ev; };

struct Ctx {
    std::vector<PendingEvents> pending;   // events of freed batches, not yet folded into totals
    float totals[5] = {0, 0, 0, 0, 0};
    unsigned n_totals = 0;
    bool init = false;
    int device = -1;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    // host-input pipeline: H2D copies run on their own stream, one event per chunk of polynomials,
    // so that the LDE of chunk k overlaps the PCIe transfer of chunk k+1
    // small host inputs / outputs (polynomials of a few KB): gathered through pinned staging so that W tiny copies
    // become one (a cudaMemcpyAsync call costs more than moving 64 bytes)
    void* pin_in = nullptr;  size_t pin_in_bytes = 0;
    void* pin_out = nullptr; size_t pin_out_bytes = 0;
    cudaStream_t copy_stream = nullptr;
    cudaStream_t d2h_stream = nullptr;    // coefficients going back to the host while the LDE / hashing continue
    std::vector<cudaEvent_t> d2h_ev;
    cudaEvent_t ev_sync = nullptr;
    std::vector<cudaEvent_t> chunk_ev;
    // PAGEABLE host inputs (Rust Vecs, numpy arrays): gathered by a few host threads into a ring of pinned slots and
    // sent with one DMA per slot; a plain cudaMemcpyAsync from pageable memory stages single-threaded inside the driver
    void* ring[3] = {nullptr, nullptr, nullptr};
    cudaEvent_t ring_ev[3] = {nullptr, nullptr, nullptr};   // the DMA out of the slot has finished
    bool ring_busy[3] = {false, false, false};
    unsigned next_slot = 0;
};
// One context per CUDA device (created by pcs_init / pcs_multi_init, destroyed by pcs_shutdown).  The single-device
// entry points work on the calling thread's CURRENT context (the device of its last pcs_init, device 0 by default); a
// batch remembers the context that made it, so accessors and pcs_batch_free act on the right device and stream whatever
// is current.  One call in flight per context; different contexts may be driven from different host threads.
constexpr int MAX_DEVICES = 64;
static Ctx g_none;                                   // never initialised: what a thread sees before pcs_init
static Ctx* g_ctxs[MAX_DEVICES] = {nullptr};
static std::mutex g_ctx_mutex;
static thread_local Ctx* g_cur = &g_none;
#define g_ctx (*g_cur)

// make `c` current on this thread (device included); returns the previous one
static Ctx* ctx_enter(Ctx* c) {
    Ctx* prev = g_cur;
    if (c && c != g_cur) {
        g_cur = c;
        cudaSetDevice(c->device);
    }
    return prev;
}
// fold completed/pending event sets into the running totals (caller has synchronised the stream)
static void drain_pending() {
    for (auto& pe : g_ctx.pending) {
        bool ok = true;
        float ms[5];
        for (int i = 0; i < 5 && ok; i++) ok = cudaEventElapsedTime(&ms[i], pe.ev[i], pe.ev[i + 1]) == cudaSuccess;
        // leaf hashing that ran group by group inside the "FFT + blinding" phase (streaming sponge) is booked as leaf hashing
        for (size_t k = 0; k + 1 < pe.absorb_ev.size() && ok; k += 2) {
            float t = 0;
            ok = cudaEventElapsedTime(&t, pe.absorb_ev[k], pe.absorb_ev[k + 1]) == cudaSuccess;
            ms[1] -= t;
            ms[3] += t;
        }
        if (ok) {
            for (int i = 0; i < 5; i++) g_ctx.totals[i] += ms[i];
            g_ctx.n_totals++;
        }
        for (auto& e : pe.ev) cudaEventDestroy(e);
        for (auto& e : pe.absorb_ev) cudaEventDestroy(e);
    }
    g_ctx.pending.clear();
    cudaGetLastError();
}

int fail(int code, const std::string& msg) {
    set_error(msg);
    return code;
}

BatchScope::BatchScope(const pcs_batch* b) : prev(g_cur) {
    if (b && b->ctx) ctx_enter((Ctx*)b->ctx);
}
BatchScope::~BatchScope() { ctx_enter((Ctx*)prev); }

void* cur_ctx() { return g_cur; }
CtxScope::CtxScope(void* ctx) : prev(g_cur) {
    if (ctx) ctx_enter((Ctx*)ctx);
}
CtxScope::~CtxScope() { ctx_enter((Ctx*)prev); }

pcs_batch* batch_new() {
    pcs_batch* b = new pcs_batch();
    b->ctx = g_cur;
    return b;
}

#define PCS_NEED_INIT()                                                            \
    do {                                                                           \
        if (!g_ctx.init) {                                                         \
            int _rc = pcs_init(-1, nullptr);                                       \
            if (_rc) return _rc;                                                   \
        } else {                                                                   \
            cudaSetDevice(g_ctx.device); /* the caller (torch ...) may have switched this thread's device */ \
        }                                                                          \
    } while (0)

static int node_levels_dev(size_t n, unsigned lg_sub, uint64_t* digests, uint64_t* cap, cudaStream_t st) {
    // one launch per level while a level still fills the GPU (more than 256 nodes per cap subtree), then the top of every
    // subtree in ONE launch
    // (levels of <= 8192 nodes run in the latency form, one node per half-warp: node_hash.cu; the top launch then starts
    // once a subtree has <= 64 nodes, before that it needs <= 256)
    unsigned level = 1;
    for (; level <= lg_sub; level++) {
        const size_t nodes = n >> level;
        const unsigned left = lg_sub - level;                     // log2(nodes per subtree at this level)
        if (nodes <= 8192 ? left <= 6 : left <= 8) break;
        PCS_CUDA(launch_node_level(digests, cap, lg_sub, level, nodes, st));
    }
    if (level <= lg_sub) PCS_CUDA(launch_node_top(digests, cap, lg_sub, level, n >> lg_sub, st));
    return PCS_OK;
}

// leaf digests + all node levels; cols = [width][n] poly-major
int build_tree_dev(const uint64_t* cols, size_t n, size_t width, unsigned lg_n, unsigned cap_height,
                          uint64_t* digests, uint64_t* cap, cudaStream_t st, cudaEvent_t after_leaves) {
    unsigned lg_sub = lg_n - cap_height;
    PCS_CUDA(launch_leaf_hash_cols(cols, n, (uint32_t)width, n, lg_sub, digests, cap, st));
    if (after_leaves) PCS_CUDA(cudaEventRecord(after_leaves, st));
    return node_levels_dev(n, lg_sub, digests, cap, st);
}

// Streaming sponge: absorb the LDE columns [b->absorbed, upto) of every leaf (upto a multiple of the rate, or all columns
// with last = true, which also emits the digests).  timed: bracket the launch with an event pair (absorbs that run inside the
// "FFT + blinding" phase are booked as leaf hashing by the timing queries).
static int absorb_columns(pcs_batch* b, size_t upto, bool last, cudaStream_t st, bool timed) {
    if (upto < b->absorbed) return fail(PCS_ERR_ARG, "internal: sponge cannot go backwards");
    if (upto == b->absorbed && !last) return PCS_OK;
    const bool first = b->absorbed == 0;
    const unsigned lg_sub = b->lg_d + b->rate_bits - b->cap_height;
    if (!(first && last) && !b->sponge) PCS_CUDA(cudaMallocAsync((void**)&b->sponge, 12 * b->n * 8, st));
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (timed) {
        PCS_CUDA(cudaEventCreate(&e0));
        PCS_CUDA(cudaEventCreate(&e1));
        b->absorb_ev.push_back(e0);
        b->absorb_ev.push_back(e1);
        PCS_CUDA(cudaEventRecord(e0, st));
    }
    PCS_CUDA(launch_leaf_hash_group(b->lde + b->absorbed * b->n, b->n, (uint32_t)(upto - b->absorbed), b->n, first, last, b->sponge,
                                    lg_sub, b->digests, b->cap, st));
    if (timed) PCS_CUDA(cudaEventRecord(e1, st));
    b->absorbed = upto;
    if (last && b->sponge) {
        cudaFreeAsync(b->sponge, st);
        b->sponge = nullptr;
    }
    return PCS_OK;
}

}  // namespace pcs

using namespace pcs;

extern "C" {

const char* pcs_last_error(void) { return g_err.c_str(); }

static int ctx_create(int device, void* stream, Ctx** out) {
    PCS_CUDA(cudaSetDevice(device));
    Ctx* c = new Ctx();
    c->device = device;
    if (stream) {
        c->stream = (cudaStream_t)stream;
        c->own_stream = false;
    } else {
        cudaError_t e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
        if (e != cudaSuccess) {
            delete c;
            return fail(PCS_ERR_CUDA, std::string("cudaStreamCreateWithFlags: ") + cudaGetErrorString(e));
        }
        c->own_stream = true;
    }
    // keep freed blocks cached in the pool: commits reuse multi-GB buffers
    cudaMemPool_t pool;
    uint64_t thr = UINT64_MAX;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess)
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    cudaGetLastError();
    c->init = true;
    *out = c;
    return PCS_OK;
}

int pcs_init(int device, void* stream) {
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(PCS_ERR_CUDA, std::string("no CUDA device: the engine has no CPU fallback (") +
                                      cudaGetErrorString(e) + ")");
    if (device < 0) {
        if (g_ctx.init) return PCS_OK;
        device = 0;
    }
    if (device >= count || device >= MAX_DEVICES)
        return fail(PCS_ERR_ARG, "device " + std::to_string(device) + " out of range (" + std::to_string(count) + " visible)");
    std::lock_guard<std::mutex> lock(g_ctx_mutex);
    Ctx* c = g_ctxs[device];
    if (!c) {
        int rc = ctx_create(device, stream, &c);
        if (rc) return rc;
        g_ctxs[device] = c;
    } else if (stream && stream != (void*)c->stream) {
        // re-bind the stream only; work already enqueued on the old stream must be finished first (twiddle tables,
        // batches) because nothing orders the new stream behind it
        cudaSetDevice(device);
        cudaStreamSynchronize(c->stream);
        if (c->own_stream) cudaStreamDestroy(c->stream);
        c->stream = (cudaStream_t)stream;
        c->own_stream = false;
    }
    ctx_enter(c);                      // current for this thread; other devices' contexts (and their batches) stay alive
    PCS_CUDA(cudaSetDevice(device));
    return PCS_OK;
}

int pcs_device(void) { return g_ctx.init ? g_ctx.device : -1; }

static void ctx_destroy(Ctx* c) {
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    g_cur = c;
    drain_pending();
    ntt_plans_free();                  // this device's twiddle tables
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    if (c->d2h_stream) cudaStreamDestroy(c->d2h_stream);
    for (auto& e : c->d2h_ev) cudaEventDestroy(e);
    if (c->pin_in) cudaFreeHost(c->pin_in);
    if (c->pin_out) cudaFreeHost(c->pin_out);
    if (c->ev_sync) cudaEventDestroy(c->ev_sync);
    for (auto& e : c->chunk_ev) cudaEventDestroy(e);
    for (auto& r : c->ring)
        if (r) cudaFreeHost(r);
    for (auto& e : c->ring_ev)
        if (e) cudaEventDestroy(e);
    if (c->own_stream) cudaStreamDestroy(c->stream);
    delete c;
}

void pcs_shutdown(void) {
    std::lock_guard<std::mutex> lock(g_ctx_mutex);
    multi_shutdown_locked();
    for (auto& c : g_ctxs)
        if (c) {
            ctx_destroy(c);
            c = nullptr;
        }
    g_cur = &g_none;
    cudaGetLastError();
}

void* pcs_stream(void) { return g_ctx.init ? (void*)g_ctx.stream : nullptr; }

int pcs_synchronize(void) {
    PCS_NEED_INIT();
    PCS_CUDA(cudaStreamSynchronize(g_ctx.stream));
    return PCS_OK;
}

// ------------------------------------------------------------------------------------------------
// primitives
// ------------------------------------------------------------------------------------------------
int pcs_field_op(int op, const uint64_t* a, const uint64_t* b, size_t n, uint64_t* out) {
    PCS_NEED_INIT();
    if (n == 0) return PCS_OK;
    if (!a || !out) return fail(PCS_ERR_ARG, "NULL pointer");
    if (op < PCS_OP_ADD || op > PCS_OP_MUL_2EXP) return fail(PCS_ERR_ARG, "unknown field operation");
    const bool binary = op != PCS_OP_SQUARE && op != PCS_OP_CANON && op != PCS_OP_NEG;
    if (binary && !b) return fail(PCS_ERR_ARG, "second operand is NULL");
    cudaStream_t st = g_ctx.stream;
    DevBuf da, db, dout;
    PCS_CUDA(da.alloc(n * 8, st));
    PCS_CUDA(dout.alloc(n * 8, st));
    PCS_CUDA(cudaMemcpyAsync(da.p, a, n * 8, cudaMemcpyHostToDevice, st));
    if (binary) {
        PCS_CUDA(db.alloc(n * 8, st));
        PCS_CUDA(cudaMemcpyAsync(db.p, b, n * 8, cudaMemcpyHostToDevice, st));
    }
    PCS_CUDA(launch_field_op(op, da.u64(), binary ? db.u64() : nullptr, n, dout.u64(), st));
    PCS_CUDA(cudaMemcpyAsync(out, dout.p, n * 8, cudaMemcpyDeviceToHost, st));
    PCS_CUDA(cudaStreamSynchronize(st));
    return PCS_OK;
}

int pcs_poseidon_permute(uint64_t* states, size_t n) {
    PCS_NEED_INIT();
    if (n == 0) return PCS_OK;
    if (!states) return fail(PCS_ERR_ARG, "states is NULL");
    cudaStream_t st = g_ctx.stream;
    DevBuf d;
    PCS_CUDA(d.alloc(n * 12 * 8, st));
    PCS_CUDA(cudaMemcpyAsync(d.p, states, n * 12 * 8, cudaMemcpyHostToDevice, st));
    PCS_CUDA(launch_permute(d.u64(), n, st));
    PCS_CUDA(cudaMemcpyAsync(states, d.p, n * 12 * 8, cudaMemcpyDeviceToHost, st));
    PCS_CUDA(cudaStreamSynchronize(st));
    return PCS_OK;
}

int pcs_pow_grind(const uint64_t* state, unsigned witness_pos, unsigned min_leading_zeros, uint64_t* witness) {
    PCS_NEED_INIT();
    if (!state || !witness) return fail(PCS_ERR_ARG, "NULL pointer");
    if (witness_pos >= 12) return fail(PCS_ERR_ARG, "witness position outside the sponge state");
    if (min_leading_zeros > 40) return fail(PCS_ERR_ARG, "proof-of-work difficulty above 40 bits is not supported");
    cudaStream_t st = g_ctx.stream;
    DevBuf d_state, d_best;
    PCS_CUDA(d_state.alloc(12 * 8, st));
    PCS_CUDA(d_best.alloc(8, st));
    PCS_CUDA(cudaMemcpyAsync(d_state.p, state, 12 * 8, cudaMemcpyHostToDevice, st));
    const uint64_t P = 0xFFFFFFFF00000001ULL, none = ~0ULL;
    PCS_CUDA(cudaMemcpyAsync(d_best.p, &none, 8, cudaMemcpyHostToDevice, st));
    // batches of four times the expected number of tries (2^min_lz): one launch finds the witness with probability
    // 1 - e^-4; at least 2^14 and at most 2^26 candidates per launch.  Batches are searched in order and a batch returns its
    // smallest hit, so the result is the smallest witness whatever the batch size.
    const unsigned lg_batch = min_leading_zeros + 2 < 14 ? 14 : (min_leading_zeros + 2 > 26 ? 26 : min_leading_zeros + 2);
    uint64_t batch = (uint64_t)1 << lg_batch;
    for (uint64_t base = 0; base < P; base += batch) {          // candidates 0 ..= p - 1 (prover.rs:141)
        uint64_t n = P - base < batch ? P - base : batch;
        PCS_CUDA(launch_pow_search(d_state.u64(), witness_pos, min_leading_zeros, base, n, (unsigned long long*)d_best.p, st));
        uint64_t best;
        PCS_CUDA(cudaMemcpyAsync(&best, d_best.p, 8, cudaMemcpyDeviceToHost, st));
        PCS_CUDA(cudaStreamSynchronize(st));
        if (best != none) {
            *witness = best;
            return PCS_OK;
        }
    }
    return fail(PCS_ERR_ARG, "Proof of work failed. This is highly unlikely!");
}

int pcs_hash_or_noop(const uint64_t* rows, size_t n, size_t len, uint64_t* out) {
    PCS_NEED_INIT();
    if (n == 0) return PCS_OK;
    if (!rows || !out) return fail(PCS_ERR_ARG, "NULL pointer");
    if (len == 0) {  // hash_or_noop of an empty slice: all-zero HashOut (config.rs:56-62)
        memset(out, 0, n * 32);
        return PCS_OK;
    }
    cudaStream_t st = g_ctx.stream;
    DevBuf d_rows, d_cols, d_out;
    PCS_CUDA(d_rows.alloc(n * len * 8, st));
    PCS_CUDA(d_cols.alloc(n * len * 8, st));
    PCS_CUDA(d_out.alloc(n * 32, st));
    PCS_CUDA(cudaMemcpyAsync(d_rows.p, rows, n * len * 8, cudaMemcpyHostToDevice, st));
    PCS_CUDA(launch_transpose(d_rows.u64(), len, d_cols.u64(), n, n, len, st));
    PCS_CUDA(launch_hash_cols_plain(d_cols.u64(), n, (uint32_t)len, n, d_out.u64(), st));
    PCS_CUDA(cudaMemcpyAsync(out, d_out.p, n * 32, cudaMemcpyDeviceToHost, st));
    PCS_CUDA(cudaStreamSynchronize(st));
    return PCS_OK;
}

int pcs_two_to_one(const uint64_t* left, const uint64_t* right, size_t n, uint64_t* out) {
    PCS_NEED_INIT();
    if (n == 0) return PCS_OK;
    if (!left || !right || !out) return fail(PCS_ERR_ARG, "NULL pointer");
    cudaStream_t st = g_ctx.stream;
    DevBuf l, r, o;
    PCS_CUDA(l.alloc(n * 32, st));
    PCS_CUDA(r.alloc(n * 32, st));
    PCS_CUDA(o.alloc(n * 32, st));
    PCS_CUDA(cudaMemcpyAsync(l.p, left, n * 32, cudaMemcpyHostToDevice, st));
    PCS_CUDA(cudaMemcpyAsync(r.p, right, n * 32, cudaMemcpyHostToDevice, st));
    PCS_CUDA(launch_two_to_one(l.u64(), r.u64(), n, o.u64(), st));
    PCS_CUDA(cudaMemcpyAsync(out, o.p, n * 32, cudaMemcpyDeviceToHost, st));
    PCS_CUDA(cudaStreamSynchronize(st));
    return PCS_OK;
}

int pcs_ntt(uint64_t* polys, size_t w, unsigned lg_n, int inverse) {
    PCS_NEED_INIT();
    if (w == 0) return PCS_OK;
    if (!polys) return fail(PCS_ERR_ARG, "polys is NULL");
    if (lg_n > 32) return fail(PCS_ERR_TWO_ADICITY, "n_log <= TWO_ADICITY violated");
    cudaStream_t st = g_ctx.stream;
    size_t n = (size_t)1 << lg_n;
    NttPlan* plan = ntt_plan_get(lg_n, 0, inverse != 0, 1, st);
    if (!plan) return fail(PCS_ERR_ALLOC, "twiddle table allocation failed");
    DevBuf a, b;
    PCS_CUDA(a.alloc(w * n * 8, st));
    PCS_CUDA(b.alloc(w * n * 8, st));
    PCS_CUDA(cudaMemcpyAsync(a.p, polys, w * n * 8, cudaMemcpyHostToDevice, st));
    PCS_CUDA(ntt_lde(plan, a.u64(), n, b.u64(), n, w, st));            // natural -> bit-reversed
    PCS_CUDA(launch_bitrev_permute(b.u64(), n, a.u64(), n, w, lg_n, st, ntt_plan_scale(plan)));  // -> natural (+ 1/n)
    PCS_CUDA(cudaMemcpyAsync(polys, a.p, w * n * 8, cudaMemcpyDeviceToHost, st));
    PCS_CUDA(cudaStreamSynchronize(st));
    return PCS_OK;
}

static uint64_t host_inverse(uint64_t a) {  // a^(p-2) mod p
    const uint64_t P = 0xFFFFFFFF00000001ULL;
    unsigned __int128 acc = 1, b = a % P;
    for (uint64_t e = P - 2; e; e >>= 1) {
        if (e & 1) acc = acc * b % P;
        b = b * b % P;
    }
    return (uint64_t)acc;
}

int pcs_coset_intt(uint64_t* values, size_t w, unsigned lg_n, uint64_t shift) {
    PCS_NEED_INIT();
    if (w == 0) return PCS_OK;
    if (!values) return fail(PCS_ERR_ARG, "values is NULL");
    if (lg_n > 32) return fail(PCS_ERR_TWO_ADICITY, "n_log <= TWO_ADICITY violated");
    if (shift % 0xFFFFFFFF00000001ULL == 0) return fail(PCS_ERR_ARG, "shift must be non-zero");
    cudaStream_t st = g_ctx.stream;
    size_t n = (size_t)1 << lg_n;
    NttPlan* plan = ntt_plan_get(lg_n, 0, true, 1, st);
    if (!plan) return fail(PCS_ERR_ALLOC, "twiddle table allocation failed");
    DevBuf a, b;
    PCS_CUDA(a.alloc(w * n * 8, st));
    PCS_CUDA(b.alloc(w * n * 8, st));
    PCS_CUDA(cudaMemcpyAsync(a.p, values, w * n * 8, cudaMemcpyHostToDevice, st));
    PCS_CUDA(ntt_lde(plan, a.u64(), n, b.u64(), n, w, st));
    PCS_CUDA(launch_bitrev_permute(b.u64(), n, a.u64(), n, w, lg_n, st, ntt_plan_scale(plan)));
    PCS_CUDA(launch_mul_powers(a.u64(), n, w, n, host_inverse(shift), st));   // c_i *= shift^-i  (mod.rs:68-72)
    PCS_CUDA(cudaMemcpyAsync(values, a.p, w * n * 8, cudaMemcpyDeviceToHost, st));
    PCS_CUDA(cudaStreamSynchronize(st));
    return PCS_OK;
}

int pcs_coset_intt_dev(uint64_t* values_dev, size_t w, unsigned lg_n, uint64_t shift) {
    PCS_NEED_INIT();
    if (w == 0) return PCS_OK;
    if (!values_dev) return fail(PCS_ERR_ARG, "values is NULL");
    if (lg_n > 32) return fail(PCS_ERR_TWO_ADICITY, "n_log <= TWO_ADICITY violated");
    if (shift % 0xFFFFFFFF00000001ULL == 0) return fail(PCS_ERR_ARG, "shift must be non-zero");
    cudaStream_t st = g_ctx.stream;
    size_t n = (size_t)1 << lg_n;
    NttPlan* plan = ntt_plan_get(lg_n, 0, true, 1, st);
    if (!plan) return fail(PCS_ERR_ALLOC, "twiddle table allocation failed");
    DevBuf b;
    PCS_CUDA(b.alloc(w * n * 8, st));
    PCS_CUDA(ntt_lde(plan, values_dev, n, b.u64(), n, w, st));                                             // -> bit-reversed
    PCS_CUDA(launch_bitrev_permute(b.u64(), n, values_dev, n, w, lg_n, st, ntt_plan_scale(plan)));         // -> natural, * 1/n
    PCS_CUDA(launch_mul_powers(values_dev, n, w, n, host_inverse(shift), st));                             // c_i *= shift^-i
    return PCS_OK;   // asynchronous on pcs_stream()
}

int pcs_ntt_dev(uint64_t* polys_dev, size_t w, unsigned lg_n, int inverse) {
    PCS_NEED_INIT();
    if (w == 0) return PCS_OK;
    if (!polys_dev) return fail(PCS_ERR_ARG, "polys is NULL");
    if (lg_n > 32) return fail(PCS_ERR_TWO_ADICITY, "n_log <= TWO_ADICITY violated");
    cudaStream_t st = g_ctx.stream;
    size_t n = (size_t)1 << lg_n;
    NttPlan* plan = ntt_plan_get(lg_n, 0, inverse != 0, 1, st);
    if (!plan) return fail(PCS_ERR_ALLOC, "twiddle table allocation failed");
    DevBuf b;
    PCS_CUDA(b.alloc(w * n * 8, st));
    PCS_CUDA(ntt_lde(plan, polys_dev, n, b.u64(), n, w, st));                                              // -> bit-reversed
    PCS_CUDA(launch_bitrev_permute(b.u64(), n, polys_dev, n, w, lg_n, st, ntt_plan_scale(plan)));          // -> natural
    return PCS_OK;   // asynchronous on pcs_stream()
}

}  // extern "C"

constexpr size_t SMALL_POLY_BYTES = 32 * 1024;     // below this, host polynomials go through pinned staging
constexpr size_t PIN_STAGING_MAX = 64u << 20;

static void* pinned(void*& buf, size_t& cap, size_t bytes) {
    if (bytes > cap) {
        if (buf) cudaFreeHost(buf);
        buf = nullptr;
        cap = 0;
        if (cudaHostAlloc(&buf, bytes, cudaHostAllocDefault) != cudaSuccess) {
            cudaGetLastError();
            return nullptr;
        }
        cap = bytes;
    }
    return buf;
}

// Device -> pageable host memory: D2H into one of two pinned buffers, then a multi-threaded copy into the caller's memory
// (fresh Vecs / numpy arrays fault their pages on first touch: one thread moves ~5 GB/s, four ~15) while the next piece is
// already crossing PCIe.  Synchronises the stream.
static int d2h_pageable(void* dst, const void* src_dev, size_t bytes, cudaStream_t st);

constexpr size_t RING_SLOT_BYTES = 16u << 20;
constexpr unsigned STAGE_THREADS = 4;

static bool is_pageable_host(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return true;
    }
    return a.type == cudaMemoryTypeUnregistered;
}

// copy stream bytes [off, off + len) of the concatenation polys[0] | polys[1] | ... (bytes_each each) to dst
static void gather_bytes(char* dst, const uint64_t* const* polys, size_t bytes_each, size_t off, size_t len) {
    while (len) {
        size_t j = off / bytes_each, o = off % bytes_each;
        size_t n = bytes_each - o < len ? bytes_each - o : len;
        memcpy(dst, (const char*)polys[j] + o, n);
        dst += n;
        off += n;
        len -= n;
    }
}

// Send `count` pageable host polynomials (d elements each) to the contiguous device block `dst` through the pinned
// ring: pieces of <= RING_SLOT_BYTES are gathered by STAGE_THREADS host threads and leave with one DMA each on the
// copy stream; the gather of piece p + 1 overlaps the DMA of piece p.
int pcs::stage_pageable(const uint64_t* const* polys, size_t count, size_t d, uint64_t* dst, cudaStream_t copy_st) {
    const size_t bytes_each = d * 8, total = count * bytes_each;
    for (size_t off = 0; off < total; off += RING_SLOT_BYTES) {
        const size_t len = total - off < RING_SLOT_BYTES ? total - off : RING_SLOT_BYTES;
        const unsigned sl = g_ctx.next_slot++ % 3;
        if (!g_ctx.ring[sl]) {
            PCS_CUDA(cudaHostAlloc(&g_ctx.ring[sl], RING_SLOT_BYTES, cudaHostAllocDefault));
            PCS_CUDA(cudaEventCreateWithFlags(&g_ctx.ring_ev[sl], cudaEventDisableTiming));
        }
        if (g_ctx.ring_busy[sl]) PCS_CUDA(cudaEventSynchronize(g_ctx.ring_ev[sl]));
        char* slot = (char*)g_ctx.ring[sl];
        const unsigned nt = len >= (1u << 20) ? STAGE_THREADS : 1;
        if (nt == 1) {
            gather_bytes(slot, polys, bytes_each, off, len);
        } else {
            std::thread th[STAGE_THREADS];
            for (unsigned t = 0; t < nt; t++) {
                size_t a = len * t / nt, e = len * (t + 1) / nt;
                th[t] = std::thread(gather_bytes, slot + a, polys, bytes_each, off + a, e - a);
            }
            for (unsigned t = 0; t < nt; t++) th[t].join();
        }
        PCS_CUDA(cudaMemcpyAsync((char*)dst + off, slot, len, cudaMemcpyHostToDevice, copy_st));
        PCS_CUDA(cudaEventRecord(g_ctx.ring_ev[sl], copy_st));
        g_ctx.ring_busy[sl] = true;
    }
    return PCS_OK;
}

// gather w separately allocated host/device polynomials into one [w][d] device matrix
static int stage_polys(const uint64_t* const* polys, size_t w, size_t d, bool device_ptrs, uint64_t* dst,
                       cudaStream_t st) {
    for (size_t j = 0; j < w; j++) {
        if (!polys[j]) return fail(PCS_ERR_ARG, "NULL polynomial pointer");
        PCS_CUDA(cudaMemcpyAsync(dst + j * d, polys[j], d * 8, device_ptrs ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st));
    }
    return PCS_OK;
}

static void host_copy_mt(char* dst, const char* src, size_t bytes) {
    const unsigned nt = bytes >= (4u << 20) ? STAGE_THREADS : 1;
    if (nt == 1) {
        memcpy(dst, src, bytes);
        return;
    }
    std::thread th[STAGE_THREADS];
    const size_t per = (bytes / nt + 4095) & ~(size_t)4095;
    for (unsigned t = 0; t < nt; t++) {
        const size_t o = (size_t)t * per, l = o >= bytes ? 0 : (bytes - o < per ? bytes - o : per);
        th[t] = std::thread([=]() { if (l) memcpy(dst + o, src + o, l); });
    }
    for (unsigned t = 0; t < nt; t++) th[t].join();
}

static int d2h_pageable(void* dst, const void* src_dev, size_t bytes, cudaStream_t st) {
    const size_t PIECE = 16u << 20;
    if (bytes <= (1u << 20) || !is_pageable_host(dst)) {
        PCS_CUDA(cudaMemcpyAsync(dst, src_dev, bytes, cudaMemcpyDeviceToHost, st));
        PCS_CUDA(cudaStreamSynchronize(st));
        return PCS_OK;
    }
    if (!pinned(g_ctx.pin_out, g_ctx.pin_out_bytes, 2 * PIECE)) return fail(PCS_ERR_ALLOC, "pinned staging allocation failed");
    cudaEvent_t ev[2];
    PCS_CUDA(cudaEventCreateWithFlags(&ev[0], cudaEventDisableTiming));
    PCS_CUDA(cudaEventCreateWithFlags(&ev[1], cudaEventDisableTiming));
    struct EvGuard { cudaEvent_t* e; ~EvGuard() { cudaEventDestroy(e[0]); cudaEventDestroy(e[1]); } } evg{ev};
    size_t prev_off = 0, prev_len = 0;
    int slot = 0;
    for (size_t off = 0; off < bytes; off += PIECE, slot ^= 1) {
        const size_t len = bytes - off < PIECE ? bytes - off : PIECE;
        PCS_CUDA(cudaMemcpyAsync((char*)g_ctx.pin_out + (size_t)slot * PIECE, (const char*)src_dev + off, len, cudaMemcpyDeviceToHost, st));
        PCS_CUDA(cudaEventRecord(ev[slot], st));
        if (prev_len) {
            PCS_CUDA(cudaEventSynchronize(ev[slot ^ 1]));
            host_copy_mt((char*)dst + prev_off, (char*)g_ctx.pin_out + (size_t)(slot ^ 1) * PIECE, prev_len);
        }
        prev_off = off;
        prev_len = len;
    }
    PCS_CUDA(cudaEventSynchronize(ev[slot ^ 1]));
    host_copy_mt((char*)dst + prev_off, (char*)g_ctx.pin_out + (size_t)(slot ^ 1) * PIECE, prev_len);
    return PCS_OK;
}

extern "C" {

int pcs_coset_lde(const uint64_t* const* coeffs, size_t w, unsigned lg_d, unsigned rate_bits, uint64_t shift,
                  uint64_t* out, int layout) {
    PCS_NEED_INIT();
    if (w == 0) return fail(PCS_ERR_ARG, "empty batch (oracle.rs:103 polynomials[0])");
    if (!coeffs || !out) return fail(PCS_ERR_ARG, "NULL pointer");
    if (lg_d + rate_bits > 32) return fail(PCS_ERR_TWO_ADICITY, "n_log <= TWO_ADICITY violated");
    cudaStream_t st = g_ctx.stream;
    size_t d = (size_t)1 << lg_d, n = d << rate_bits;
    NttPlan* plan = ntt_plan_get(lg_d, rate_bits, false, shift, st);
    if (!plan) return fail(PCS_ERR_ALLOC, "twiddle table allocation failed");
    DevBuf c, lde, o;
    PCS_CUDA(c.alloc(w * d * 8, st));
    PCS_CUDA(lde.alloc(w * n * 8, st));
    PCS_CUDA(o.alloc(w * n * 8, st));
    int rc = stage_polys(coeffs, w, d, false, c.u64(), st);
    if (rc) return rc;
    PCS_CUDA(ntt_lde(plan, c.u64(), d, lde.u64(), n, w, st));
    if (layout == 0)
        PCS_CUDA(launch_bitrev_permute(lde.u64(), n, o.u64(), n, w, lg_d + rate_bits, st));
    else
        PCS_CUDA(launch_transpose(lde.u64(), n, o.u64(), w, w, n, st));
    PCS_CUDA(cudaMemcpyAsync(out, o.p, w * n * 8, cudaMemcpyDeviceToHost, st));
    PCS_CUDA(cudaStreamSynchronize(st));
    return PCS_OK;
}

int pcs_coset_lde_dev(const uint64_t* const* coeffs_dev, size_t w, unsigned lg_d, unsigned rate_bits, uint64_t shift,
                      uint64_t* out_dev) {
    PCS_NEED_INIT();
    if (w == 0) return fail(PCS_ERR_ARG, "empty batch (oracle.rs:103 polynomials[0])");
    if (!coeffs_dev || !out_dev) return fail(PCS_ERR_ARG, "NULL pointer");
    if (lg_d + rate_bits > 32) return fail(PCS_ERR_TWO_ADICITY, "n_log <= TWO_ADICITY violated");
    cudaStream_t st = g_ctx.stream;
    const size_t d = (size_t)1 << lg_d, n = d << rate_bits;
    NttPlan* plan = ntt_plan_get(lg_d, rate_bits, false, shift, st);
    if (!plan) return fail(PCS_ERR_ALLOC, "twiddle table allocation failed");
    bool contiguous = true;
    for (size_t j = 0; j < w; j++) {
        if (!coeffs_dev[j]) return fail(PCS_ERR_ARG, "NULL polynomial pointer");
        contiguous = contiguous && coeffs_dev[j] == coeffs_dev[0] + j * d;
    }
    DevBuf table, staged;
    const uint64_t* src = coeffs_dev[0];
    const uint64_t* const* ptrs = nullptr;
    if (!contiguous) {
        if (lg_d >= 1) {
            PCS_CUDA(table.alloc(w * sizeof(uint64_t*), st));
            PCS_CUDA(cudaMemcpyAsync(table.p, coeffs_dev, w * sizeof(uint64_t*), cudaMemcpyHostToDevice, st));
            ptrs = (const uint64_t* const*)table.p;
        } else {
            PCS_CUDA(staged.alloc(w * d * 8, st));
            int rc = stage_polys(coeffs_dev, w, d, true, staged.u64(), st);
            if (rc) return rc;
            src = staged.u64();
        }
    }
    PCS_CUDA(ntt_lde_cosets(plan, src, d, out_dev, n, w, 0, rate_bits, st, ptrs));
    return PCS_OK;   // asynchronous on pcs_stream()
}

int pcs_merkle_build(const uint64_t* leaves, size_t n, size_t len, unsigned cap_height, uint64_t* digests,
                     uint64_t* cap) {
    PCS_NEED_INIT();
    int lg_n = ilog2_strict(n);
    if (lg_n < 0) return fail(PCS_ERR_NOT_POW2, "Not a power of two: " + std::to_string(n));
    if ((int)cap_height > lg_n)
        return fail(PCS_ERR_CAP_HEIGHT, "cap_height=" + std::to_string(cap_height) +
                                            " should be at most log2(leaves.len())=" + std::to_string(lg_n));
    if (!leaves || !cap || len == 0) return fail(PCS_ERR_ARG, "NULL pointer or empty leaves");
    cudaStream_t st = g_ctx.stream;
    size_t n_cap = (size_t)1 << cap_height, n_dig = 2 * (n - n_cap);
    if (n_dig && !digests) return fail(PCS_ERR_ARG, "digests is NULL");
    DevBuf d_rows, d_cols, d_dig, d_cap;
    PCS_CUDA(d_rows.alloc(n * len * 8, st));
    PCS_CUDA(d_cols.alloc(n * len * 8, st));
    PCS_CUDA(d_dig.alloc(n_dig * 32, st));
    PCS_CUDA(d_cap.alloc(n_cap * 32, st));
    PCS_CUDA(cudaMemcpyAsync(d_rows.p, leaves, n * len * 8, cudaMemcpyHostToDevice, st));
    PCS_CUDA(launch_transpose(d_rows.u64(), len, d_cols.u64(), n, n, len, st));
    int rc = build_tree_dev(d_cols.u64(), n, len, (unsigned)lg_n, cap_height, d_dig.u64(), d_cap.u64(), st, nullptr);
    if (rc) return rc;
    if (n_dig) PCS_CUDA(cudaMemcpyAsync(digests, d_dig.p, n_dig * 32, cudaMemcpyDeviceToHost, st));
    PCS_CUDA(cudaMemcpyAsync(cap, d_cap.p, n_cap * 32, cudaMemcpyDeviceToHost, st));
    PCS_CUDA(cudaStreamSynchronize(st));
    return PCS_OK;
}

// ------------------------------------------------------------------------------------------------
// fused hot path
// ------------------------------------------------------------------------------------------------
void pcs_batch_free(pcs_batch* b) {
    if (!b) return;
    BatchScope scope(b);
    cudaStream_t st = g_ctx.stream;
    if (b->coeffs) cudaFreeAsync(b->coeffs, st);
    if (b->lde) cudaFreeAsync(b->lde, st);
    if (b->digests) cudaFreeAsync(b->digests, st);
    if (b->cap) cudaFreeAsync(b->cap, st);
    if (b->sponge) cudaFreeAsync(b->sponge, st);
    if (b->ev[5] && b->committed) {
        PendingEvents pe;
        for (int i = 0; i < 6; i++) pe.ev[i] = b->ev[i];
        pe.absorb_ev.swap(b->absorb_ev);
        g_ctx.pending.push_back(pe);
        if (g_ctx.pending.size() > 256) {
            cudaStreamSynchronize(st);
            drain_pending();
        }
    } else {
        for (auto& e : b->ev)
            if (e) cudaEventDestroy(e);
    }
    for (auto& e : b->absorb_ev) cudaEventDestroy(e);
    delete b;
}

// shard == true: only the cosets [coset_first, coset_first + 2^lg_cosets) (leaf order) are extended and the
// tree is built over those d << lg_cosets leaves with `cap_height` (the caller's LOCAL cap height); salts are
// then the shard's rows, already in leaf order.
static int commit_common(const uint64_t* const* polys, bool from_values, size_t w, unsigned lg_d, unsigned rate_bits,
                         unsigned cap_height, const uint64_t* const* salts, size_t salt_w, unsigned flags,
                         uint64_t* const* coeffs_out, uint64_t* cap_out, pcs_batch** out, bool shard = false,
                         unsigned coset_first = 0, unsigned lg_cosets = 0) {
    PCS_NEED_INIT();
    if (!out) return fail(PCS_ERR_ARG, "out is NULL");
    *out = nullptr;
    if (w == 0 || !polys) return fail(PCS_ERR_ARG, "empty batch (oracle.rs:76 polynomials[0])");
    if (salt_w && !salts) return fail(PCS_ERR_ARG, "salts is NULL");
    if (lg_d + rate_bits > 32) return fail(PCS_ERR_TWO_ADICITY, "n_log <= TWO_ADICITY violated");
    if (!shard) lg_cosets = rate_bits;
    if (lg_cosets > rate_bits || ((size_t)coset_first + ((size_t)1 << lg_cosets)) > ((size_t)1 << rate_bits))
        return fail(PCS_ERR_ARG, "coset range outside [0, 2^rate_bits)");
    unsigned lg_n = lg_d + lg_cosets;   // leaves of this (shard of the) tree
    if (cap_height > lg_n)
        return fail(PCS_ERR_CAP_HEIGHT, "cap_height=" + std::to_string(cap_height) +
                                            " should be at most log2(leaves.len())=" + std::to_string(lg_n));
    cudaStream_t st = g_ctx.stream;
    const bool dev_ptrs = flags & PCS_DEVICE_PTRS;
    const size_t d = (size_t)1 << lg_d, n = d << lg_cosets, wt = w + salt_w;
    const size_t n_cap = (size_t)1 << cap_height;

    NttPlan* plan = ntt_plan_get(lg_d, rate_bits, false, 7 /* F::coset_shift(), types.rs:437 */, st);
    NttPlan* iplan = from_values ? ntt_plan_get(lg_d, 0, true, 1, st) : nullptr;
    if (!plan || (from_values && !iplan)) return fail(PCS_ERR_ALLOC, "twiddle table allocation failed");

    pcs_batch* b = batch_new();
    b->w = w; b->salt_w = salt_w; b->lg_d = lg_d; b->rate_bits = lg_cosets; b->cap_height = cap_height;
    b->full_rate_bits = rate_bits; b->coset_first = coset_first;
    b->n = n; b->n_digests = 2 * (n - n_cap); b->has_ifft = from_values;
    struct Guard { pcs_batch* b; bool armed = true; ~Guard() { if (armed) pcs_batch_free(b); } } guard{b};
    // Declared after `guard`, so destroyed before it: on every failure exit taken once H2D copies were enqueued on the copy
    // stream, wait for them before the guard returns their destination to the pool (and release the pinned ring slots).
    struct CopyDrain {
        bool armed = false;
        ~CopyDrain() {
            if (g_ctx.d2h_stream) cudaStreamSynchronize(g_ctx.d2h_stream);   // D2H copies out of a buffer the guard is about to free
            if (!armed || !g_ctx.copy_stream) return;
            cudaStreamSynchronize(g_ctx.copy_stream);
            for (auto& busy : g_ctx.ring_busy) busy = false;
        }
    } copy_drain;
    for (auto& e : b->ev) PCS_CUDA(cudaEventCreate(&e));
    PCS_CUDA(cudaMallocAsync((void**)&b->lde, wt * n * 8, st));
    PCS_CUDA(cudaMallocAsync((void**)&b->digests, b->n_digests ? b->n_digests * 32 : 32, st));
    PCS_CUDA(cudaMallocAsync((void**)&b->cap, n_cap * 32, st));

    // ---- inputs -> one contiguous [w][d] device matrix ----
    bool contiguous_dev = dev_ptrs;
    if (dev_ptrs)
        for (size_t j = 0; j < w; j++) contiguous_dev = contiguous_dev && polys[j] == polys[0] + j * d;
    const uint64_t* src = nullptr;   // [w][d] device input (values or coefficients)
    DevBuf ptr_table;                // or: device table of the caller's w polynomial pointers, read in place
    uint64_t* staged = nullptr;
    bool scatter_coeffs = false;
    std::vector<size_t> cb;           // chunk k = polynomials [cb[k], cb[k+1]) of the host-input pipeline
    size_t n_chunks = 0;              // > 0: host inputs arrive chunk by chunk on the copy stream
    bool pageable = false;            // ... through the pinned ring, each chunk staged right before its compute is enqueued
    if (contiguous_dev && !from_values && !(flags & PCS_KEEP_COEFFS)) {
        src = polys[0];
    } else if (dev_ptrs && !from_values && !(flags & PCS_KEEP_COEFFS) && lg_d >= 1) {
        // separately allocated device polynomials (possibly in PEER memory): no staging copy, the first NTT pass
        // follows the pointer table
        for (size_t j = 0; j < w; j++)
            if (!polys[j]) return fail(PCS_ERR_ARG, "NULL polynomial pointer");
        PCS_CUDA(ptr_table.alloc(w * sizeof(uint64_t*), st));
        PCS_CUDA(cudaMemcpyAsync(ptr_table.p, polys, w * sizeof(uint64_t*), cudaMemcpyHostToDevice, st));
    } else {
        PCS_CUDA(cudaMallocAsync((void**)&staged, w * d * 8, st));
        b->coeffs = staged;  // owned by the batch from here on
        if (contiguous_dev)
            PCS_CUDA(cudaMemcpyAsync(staged, polys[0], w * d * 8, cudaMemcpyDeviceToDevice, st));
        else if (dev_ptrs) {
            int rc = stage_polys(polys, w, d, true, staged, st);
            if (rc) return rc;
        } else if (d * 8 <= SMALL_POLY_BYTES && w * d * 8 <= PIN_STAGING_MAX &&
                   pinned(g_ctx.pin_in, g_ctx.pin_in_bytes, w * d * 8)) {
            // small host polynomials: one gather on the host, one H2D copy
            for (size_t j = 0; j < w; j++) {
                if (!polys[j]) return fail(PCS_ERR_ARG, "NULL polynomial pointer");
                memcpy((char*)g_ctx.pin_in + j * d * 8, polys[j], d * 8);
            }
            PCS_CUDA(cudaMemcpyAsync(staged, g_ctx.pin_in, w * d * 8, cudaMemcpyHostToDevice, st));
        } else {
            // host inputs: chunked H2D on the copy stream, one event per chunk
            if (!g_ctx.copy_stream) PCS_CUDA(cudaStreamCreateWithFlags(&g_ctx.copy_stream, cudaStreamNonBlocking));
            if (!g_ctx.ev_sync) PCS_CUDA(cudaEventCreateWithFlags(&g_ctx.ev_sync, cudaEventDisableTiming));
            for (size_t j = 0; j < w; j++)
                if (!polys[j]) return fail(PCS_ERR_ARG, "NULL polynomial pointer");
            pageable = w * d * 8 >= (8u << 20) && is_pageable_host(polys[0]);   // below 8 MB the threads cost more than they save
            // Chunk sizes: the first transfer is the only one nothing hides, so it is small (8 polynomials = one absorb of the
            // sponge, or one 16 MB ring slot's worth for short polynomials); a chunk's compute (LDE + its share of the leaf
            // hashing, ~0.9 ms per polynomial at 2^20) outlasts a transfer 5x its size, so the groups grow 5x and few sponge
            // states have to be parked (each group boundary costs 192 B of HBM traffic per leaf).
            size_t first = 8;
            if (first * d * 8 < RING_SLOT_BYTES) first = (RING_SLOT_BYTES / (d * 8) + 7) / 8 * 8;
            cb.assign(1, 0);
            for (size_t sz = first; cb.back() < w; sz *= 5) cb.push_back(cb.back() + sz < w ? cb.back() + sz : w);
            if (cb.size() > 2 && w - cb[cb.size() - 2] < 8) cb.erase(cb.end() - 2);   // no tiny tail group
            n_chunks = cb.size() - 1;
            while (g_ctx.chunk_ev.size() < n_chunks) {
                cudaEvent_t e;
                PCS_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
                g_ctx.chunk_ev.push_back(e);
            }
            PCS_CUDA(cudaEventRecord(g_ctx.ev_sync, st));               // `staged` exists from here on
            PCS_CUDA(cudaStreamWaitEvent(g_ctx.copy_stream, g_ctx.ev_sync, 0));
            copy_drain.armed = true;
            for (size_t k = 0; k < n_chunks && !pageable; k++) {
                size_t j0 = cb[k], j1 = cb[k + 1];
                int rc = stage_polys(polys + j0, j1 - j0, d, false, staged + j0 * d, g_ctx.copy_stream);
                if (rc) {
                    cudaStreamSynchronize(g_ctx.copy_stream);   // `staged` is released by the guard
                    return rc;
                }
                PCS_CUDA(cudaEventRecord(g_ctx.chunk_ev[k], g_ctx.copy_stream));
            }
        }
        src = staged;
    }
    PCS_CUDA(cudaEventRecord(b->ev[0], st));

    // ---- "IFFT" (oracle.rs:51-55) for polynomials [j0, j1): values (natural order) -> coefficients in `staged` ----
    // scratch for the bit-reversed intermediate = the TAIL of the LDE buffer (the last w*d elements): LDE rows already
    // written for earlier polynomials end at j0*n <= wt*n - w*d + j0*d, so chunk-by-chunk processing never clobbers them
    uint64_t* const ifft_scratch = b->lde + (wt * n - w * d);
    const bool small_out = from_values && coeffs_out && d * 8 <= SMALL_POLY_BYTES && w * d * 8 <= PIN_STAGING_MAX &&
                           pinned(g_ctx.pin_out, g_ctx.pin_out_bytes, w * d * 8);
    // larger outputs (the reference keeps `polynomials`, so a Rust caller always asks for them): every chunk's coefficients
    // leave on their own stream into pinned staging while the LDE and the hashing go on, and are scattered into the caller's
    // (pageable, separately allocated) vectors by a few host threads after the commit
    const bool big_out = from_values && coeffs_out && !small_out && w * d * 8 <= ((size_t)512 << 20) &&
                         pinned(g_ctx.pin_out, g_ctx.pin_out_bytes, w * d * 8);
    size_t n_d2h = 0;
    if (big_out && !g_ctx.d2h_stream) PCS_CUDA(cudaStreamCreateWithFlags(&g_ctx.d2h_stream, cudaStreamNonBlocking));
    auto ifft_range = [&](size_t j0, size_t j1) -> int {
        PCS_CUDA(ntt_inverse_bitrev(iplan, src + j0 * d, d, ifft_scratch + j0 * d, d, j1 - j0, st));
        PCS_CUDA(launch_bitrev_permute(ifft_scratch + j0 * d, d, staged + j0 * d, d, j1 - j0, lg_d, st, ntt_plan_scale(iplan)));
        if (big_out) {
            if (g_ctx.d2h_ev.size() <= n_d2h) {
                cudaEvent_t e;
                PCS_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
                g_ctx.d2h_ev.push_back(e);
            }
            PCS_CUDA(cudaEventRecord(g_ctx.d2h_ev[n_d2h], st));
            PCS_CUDA(cudaStreamWaitEvent(g_ctx.d2h_stream, g_ctx.d2h_ev[n_d2h], 0));
            n_d2h++;
            PCS_CUDA(cudaMemcpyAsync((char*)g_ctx.pin_out + j0 * d * 8, staged + j0 * d, (j1 - j0) * d * 8, cudaMemcpyDeviceToHost,
                                     g_ctx.d2h_stream));
        } else if (coeffs_out && !small_out) {
            for (size_t j = j0; j < j1; j++)
                if (coeffs_out[j])
                    PCS_CUDA(cudaMemcpyAsync(coeffs_out[j], staged + j * d, d * 8, cudaMemcpyDeviceToHost, st));
        }
        return PCS_OK;
    };
    const bool chunked = n_chunks > 0;   // host inputs arriving chunk by chunk: IFFT and LDE follow each chunk
    if (from_values && !chunked) {
        int rc = ifft_range(0, w);
        if (rc) return rc;
    }
    PCS_CUDA(cudaEventRecord(b->ev[1], st));

    // ---- "FFT + blinding" (oracle.rs:100-125), output already in leaf order ----
    if (chunked) {
        for (size_t k = 0; k < n_chunks; k++) {
            size_t j0 = cb[k], j1 = cb[k + 1];
            if (pageable) {
                int rc = stage_pageable(polys + j0, j1 - j0, d, staged + j0 * d, g_ctx.copy_stream);
                if (rc) {
                    cudaStreamSynchronize(g_ctx.copy_stream);
                    return rc;
                }
                PCS_CUDA(cudaEventRecord(g_ctx.chunk_ev[k], g_ctx.copy_stream));
            }
            PCS_CUDA(cudaStreamWaitEvent(st, g_ctx.chunk_ev[k], 0));
            if (from_values) {
                int rc = ifft_range(j0, j1);
                if (rc) return rc;
            }
            PCS_CUDA(ntt_lde_cosets(plan, src + j0 * d, d, b->lde + j0 * n, n, j1 - j0, coset_first, lg_cosets, st));
            // streaming sponge: hash what has landed while the next chunks are still crossing PCIe (the last group is absorbed
            // by "build Merkle tree" below, together with the salt columns)
            b->extended = j1;
            if (wt > 4 && j1 < w) {
                int rc = absorb_columns(b, (j1 / 8) * 8, false, st, true);
                if (rc) return rc;
            }
        }
    } else {
        PCS_CUDA(ntt_lde_cosets(plan, src, d, b->lde, n, w, coset_first, lg_cosets, st,
                                (const uint64_t* const*)ptr_table.p));
    }
    if (small_out) {
        PCS_CUDA(cudaMemcpyAsync(g_ctx.pin_out, staged, w * d * 8, cudaMemcpyDeviceToHost, st));
        scatter_coeffs = true;   // scattered to the caller's vectors after the final synchronisation
    }
    for (size_t k = 0; k < salt_w; k++) {
        if (!salts[k]) return fail(PCS_ERR_ARG, "NULL salt pointer");
        // the caller's salt column k is in natural LDE order like lde_values (oracle.rs:119-123);
        // leaf order = bit-reversed rows (oracle.rs:84)
        DevBuf tmp;
        PCS_CUDA(tmp.alloc(n * 8, st));
        PCS_CUDA(cudaMemcpyAsync(tmp.p, salts[k], n * 8, dev_ptrs ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st));
        PCS_CUDA(launch_canonicalize(tmp.u64(), n, st));
        if (shard)   // already in leaf order
            PCS_CUDA(cudaMemcpyAsync(b->lde + (w + k) * n, tmp.p, n * 8, cudaMemcpyDeviceToDevice, st));
        else
            PCS_CUDA(launch_bitrev_permute(tmp.u64(), n, b->lde + (w + k) * n, n, 1, lg_n, st));
    }
    PCS_CUDA(cudaEventRecord(b->ev[2], st));
    // "transpose LDEs" (oracle.rs:83): fused away -- the hash kernel reads columns directly
    PCS_CUDA(cudaEventRecord(b->ev[3], st));

    // ---- "build Merkle tree" (oracle.rs:85-89) ----
    int rc;
    if (b->absorbed > 0) {
        rc = absorb_columns(b, wt, true, st, false);     // the remaining columns (+ salts) and the digests
        if (rc) return rc;
        PCS_CUDA(cudaEventRecord(b->ev[4], st));
        rc = node_levels_dev(n, lg_n - cap_height, b->digests, b->cap, st);
    } else {
        rc = build_tree_dev(b->lde, n, wt, lg_n, cap_height, b->digests, b->cap, st, b->ev[4]);
    }
    if (rc) return rc;
    PCS_CUDA(cudaEventRecord(b->ev[5], st));

    if (b->coeffs && !from_values && !(flags & PCS_KEEP_COEFFS)) {
        cudaFreeAsync(b->coeffs, st);
        b->coeffs = nullptr;
    }
    if (cap_out) {
        PCS_CUDA(cudaMemcpyAsync(cap_out, b->cap, n_cap * 32, cudaMemcpyDeviceToHost, st));
        PCS_CUDA(cudaStreamSynchronize(st));
    } else if (!dev_ptrs || scatter_coeffs) {
        PCS_CUDA(cudaStreamSynchronize(st));  // host inputs must not be reused before the copies finish
    }
    if (scatter_coeffs)
        for (size_t j = 0; j < w; j++)
            if (coeffs_out[j]) memcpy(coeffs_out[j], (char*)g_ctx.pin_out + j * d * 8, d * 8);
    if (big_out) {
        PCS_CUDA(cudaStreamSynchronize(st));
        PCS_CUDA(cudaStreamSynchronize(g_ctx.d2h_stream));
        const unsigned nt = STAGE_THREADS;
        std::thread th[STAGE_THREADS];
        const char* srcp = (const char*)g_ctx.pin_out;
        for (unsigned t = 0; t < nt; t++)
            th[t] = std::thread([=]() {
                for (size_t j = t; j < w; j += nt)
                    if (coeffs_out[j]) memcpy(coeffs_out[j], srcp + j * d * 8, d * 8);
            });
        for (unsigned t = 0; t < nt; t++) th[t].join();
    }
    guard.armed = false;
    copy_drain.armed = false;   // the main stream waited on every chunk event and has been synchronised (host inputs)
    b->committed = true;
    *out = b;
    return PCS_OK;
}

int pcs_commit_from_coeffs(const uint64_t* const* polys, size_t w, unsigned lg_d, unsigned rate_bits,
                           unsigned cap_height, const uint64_t* const* salts, size_t salt_w, unsigned flags,
                           uint64_t* cap_out, pcs_batch** out) {
    return commit_common(polys, false, w, lg_d, rate_bits, cap_height, salts, salt_w, flags, nullptr, cap_out, out);
}

int pcs_commit_shard_from_coeffs(const uint64_t* const* polys, size_t w, unsigned lg_d, unsigned rate_bits,
                                 unsigned coset_first, unsigned lg_cosets, unsigned local_cap_height,
                                 const uint64_t* const* salts, size_t salt_w, unsigned flags, uint64_t* cap_out,
                                 pcs_batch** out) {
    return commit_common(polys, false, w, lg_d, rate_bits, local_cap_height, salts, salt_w, flags, nullptr, cap_out, out,
                         true, coset_first, lg_cosets);
}

// ---- streaming shard commit: the polynomials arrive in groups (e.g. chunks of an all-gather still in flight) ----
int pcs_shard_begin(size_t w, size_t salt_w, unsigned lg_d, unsigned rate_bits, unsigned coset_first, unsigned lg_cosets,
                    unsigned local_cap_height, pcs_batch** out) {
    PCS_NEED_INIT();
    if (!out) return fail(PCS_ERR_ARG, "out is NULL");
    *out = nullptr;
    if (w == 0) return fail(PCS_ERR_ARG, "empty batch (oracle.rs:76 polynomials[0])");
    if (lg_d + rate_bits > 32) return fail(PCS_ERR_TWO_ADICITY, "n_log <= TWO_ADICITY violated");
    if (lg_cosets > rate_bits || ((size_t)coset_first + ((size_t)1 << lg_cosets)) > ((size_t)1 << rate_bits))
        return fail(PCS_ERR_ARG, "coset range outside [0, 2^rate_bits)");
    unsigned lg_n = lg_d + lg_cosets;
    if (local_cap_height > lg_n)
        return fail(PCS_ERR_CAP_HEIGHT, "cap_height=" + std::to_string(local_cap_height) +
                                            " should be at most log2(leaves.len())=" + std::to_string(lg_n));
    cudaStream_t st = g_ctx.stream;
    if (!ntt_plan_get(lg_d, rate_bits, false, 7, st)) return fail(PCS_ERR_ALLOC, "twiddle table allocation failed");
    const size_t n = (size_t)1 << lg_n, n_cap = (size_t)1 << local_cap_height;
    pcs_batch* b = batch_new();
    b->w = w; b->salt_w = salt_w; b->lg_d = lg_d; b->rate_bits = lg_cosets; b->cap_height = local_cap_height;
    b->full_rate_bits = rate_bits; b->coset_first = coset_first;
    b->n = n; b->n_digests = 2 * (n - n_cap);
    struct Guard { pcs_batch* b; bool armed = true; ~Guard() { if (armed) pcs_batch_free(b); } } guard{b};
    for (auto& e : b->ev) PCS_CUDA(cudaEventCreate(&e));
    PCS_CUDA(cudaMallocAsync((void**)&b->lde, (w + salt_w) * n * 8, st));
    PCS_CUDA(cudaMallocAsync((void**)&b->digests, b->n_digests ? b->n_digests * 32 : 32, st));
    PCS_CUDA(cudaMallocAsync((void**)&b->cap, n_cap * 32, st));
    PCS_CUDA(cudaEventRecord(b->ev[0], st));
    PCS_CUDA(cudaEventRecord(b->ev[1], st));
    guard.armed = false;
    *out = b;
    return PCS_OK;
}

// A tree shard whose LDE rows are computed elsewhere (another GPU's LDE arriving over NVLink: the north-star's
// polynomial-partitioned LDE + all-to-all): `width` rows of 2^lg_n leaves each, filled through pcs_shard_set_rows or written
// in place at pcs_batch_lde_dev() + row * 2^lg_n, then pcs_shard_finish.
int pcs_shard_begin_rows(size_t width, size_t salt_w, unsigned lg_n, unsigned local_cap_height, pcs_batch** out) {
    PCS_NEED_INIT();
    if (!out) return fail(PCS_ERR_ARG, "out is NULL");
    *out = nullptr;
    if (width == 0) return fail(PCS_ERR_ARG, "empty batch (oracle.rs:76 polynomials[0])");
    if (salt_w > width) return fail(PCS_ERR_ARG, "more salt columns than columns");
    if (lg_n > 32) return fail(PCS_ERR_TWO_ADICITY, "n_log <= TWO_ADICITY violated");
    if (local_cap_height > lg_n)
        return fail(PCS_ERR_CAP_HEIGHT, "cap_height=" + std::to_string(local_cap_height) +
                                            " should be at most log2(leaves.len())=" + std::to_string(lg_n));
    cudaStream_t st = g_ctx.stream;
    const size_t n = (size_t)1 << lg_n, n_cap = (size_t)1 << local_cap_height;
    pcs_batch* b = batch_new();
    b->w = width - salt_w; b->salt_w = salt_w; b->lg_d = lg_n; b->rate_bits = 0; b->cap_height = local_cap_height;
    b->full_rate_bits = 0; b->coset_first = 0; b->rows_only = true;
    b->n = n; b->n_digests = 2 * (n - n_cap);
    struct Guard { pcs_batch* b; bool armed = true; ~Guard() { if (armed) pcs_batch_free(b); } } guard{b};
    for (auto& e : b->ev) PCS_CUDA(cudaEventCreate(&e));
    PCS_CUDA(cudaMallocAsync((void**)&b->lde, width * n * 8, st));
    PCS_CUDA(cudaMallocAsync((void**)&b->digests, b->n_digests ? b->n_digests * 32 : 32, st));
    PCS_CUDA(cudaMallocAsync((void**)&b->cap, n_cap * 32, st));
    PCS_CUDA(cudaEventRecord(b->ev[0], st));
    PCS_CUDA(cudaEventRecord(b->ev[1], st));
    guard.armed = false;
    *out = b;
    return PCS_OK;
}

int pcs_shard_set_rows(pcs_batch* b, size_t row_first, size_t count, const uint64_t* const* rows_dev, int canonical) {
    if (!b || (count && !rows_dev)) return fail(PCS_ERR_ARG, "NULL pointer");
    BatchScope scope(b);
    if (b->committed) return fail(PCS_ERR_ARG, "batch already finished");
    const size_t wt = b->w + b->salt_w;
    if (row_first > wt || count > wt - row_first) return fail(PCS_ERR_ARG, "row range out of bounds");
    cudaStream_t st = g_ctx.stream;
    for (size_t k = 0; k < count; k++) {
        if (!rows_dev[k]) return fail(PCS_ERR_ARG, "NULL row pointer");
        uint64_t* dst = b->lde + (row_first + k) * b->n;
        if (rows_dev[k] != dst)
            PCS_CUDA(cudaMemcpyAsync(dst, rows_dev[k], b->n * 8, cudaMemcpyDeviceToDevice, st));
    }
    if (!canonical && count) PCS_CUDA(launch_canonicalize(b->lde + row_first * b->n, count * b->n, st));
    return PCS_OK;   // asynchronous on pcs_stream()
}

int pcs_shard_extend(pcs_batch* b, size_t poly_first, size_t count, const uint64_t* const* polys) {
    BatchScope scope(b);
    if (!b || !polys) return fail(PCS_ERR_ARG, "NULL pointer");
    if (b->committed) return fail(PCS_ERR_ARG, "batch already finished");
    if (b->rows_only) return fail(PCS_ERR_ARG, "this shard takes LDE rows (pcs_shard_set_rows), not polynomials");
    if (poly_first > b->w || count > b->w - poly_first) return fail(PCS_ERR_ARG, "polynomial range out of bounds");
    if (count == 0) return PCS_OK;
    cudaStream_t st = g_ctx.stream;
    const size_t d = (size_t)1 << b->lg_d;
    NttPlan* plan = ntt_plan_get(b->lg_d, b->full_rate_bits, false, 7, st);
    if (!plan) return fail(PCS_ERR_ALLOC, "twiddle table allocation failed");
    bool contiguous = true;
    for (size_t j = 0; j < count; j++) {
        if (!polys[j]) return fail(PCS_ERR_ARG, "NULL polynomial pointer");
        contiguous = contiguous && polys[j] == polys[0] + j * d;
    }
    DevBuf table, staged;
    const uint64_t* src = polys[0];
    const uint64_t* const* ptrs = nullptr;
    if (!contiguous) {
        if (b->lg_d >= 1) {
            PCS_CUDA(table.alloc(count * sizeof(uint64_t*), st));
            PCS_CUDA(cudaMemcpyAsync(table.p, polys, count * sizeof(uint64_t*), cudaMemcpyHostToDevice, st));
            ptrs = (const uint64_t* const*)table.p;
        } else {
            PCS_CUDA(staged.alloc(count * d * 8, st));
            int rc = stage_polys(polys, count, d, true, staged.u64(), st);
            if (rc) return rc;
            src = staged.u64();
        }
    }
    PCS_CUDA(ntt_lde_cosets(plan, src, d, b->lde + poly_first * b->n, b->n, count, b->coset_first, b->rate_bits, st, ptrs));
    // streaming sponge: groups that arrive in order are hashed at once (while the next group's coefficients are still in
    // flight); a shard that gets all its polynomials in one call keeps the single hash kernel of pcs_shard_finish
    if (poly_first == b->extended) b->extended += count;
    const bool whole = poly_first == 0 && count == b->w;
    if (!whole && b->w + b->salt_w > 4 && b->extended < b->w) {
        int rc = absorb_columns(b, (b->extended / 8) * 8, false, st, true);
        if (rc) return rc;
    }
    return PCS_OK;   // asynchronous on pcs_stream()
}

int pcs_shard_finish(pcs_batch* b, uint64_t* cap_out) {
    BatchScope scope(b);
    if (!b) return fail(PCS_ERR_ARG, "NULL pointer");
    if (b->committed) return fail(PCS_ERR_ARG, "batch already finished");
    cudaStream_t st = g_ctx.stream;
    PCS_CUDA(cudaEventRecord(b->ev[2], st));
    PCS_CUDA(cudaEventRecord(b->ev[3], st));
    int rc;
    if (b->absorbed > 0) {
        rc = absorb_columns(b, b->w + b->salt_w, true, st, false);
        if (rc) return rc;
        PCS_CUDA(cudaEventRecord(b->ev[4], st));
        rc = node_levels_dev(b->n, b->lg_d + b->rate_bits - b->cap_height, b->digests, b->cap, st);
    } else {
        rc = build_tree_dev(b->lde, b->n, b->w + b->salt_w, b->lg_d + b->rate_bits, b->cap_height, b->digests, b->cap, st, b->ev[4]);
    }
    if (rc) return rc;
    PCS_CUDA(cudaEventRecord(b->ev[5], st));
    b->committed = true;
    if (cap_out) {
        PCS_CUDA(cudaMemcpyAsync(cap_out, b->cap, ((size_t)32) << b->cap_height, cudaMemcpyDeviceToHost, st));
        PCS_CUDA(cudaStreamSynchronize(st));
    }
    return PCS_OK;
}

int pcs_commit_from_values(const uint64_t* const* values, size_t w, unsigned lg_d, unsigned rate_bits,
                           unsigned cap_height, const uint64_t* const* salts, size_t salt_w, unsigned flags,
                           uint64_t* const* coeffs_out, uint64_t* cap_out, pcs_batch** out) {
    return commit_common(values, true, w, lg_d, rate_bits, cap_height, salts, salt_w, flags, coeffs_out, cap_out, out);
}

int pcs_batch_shape(const pcs_batch* b, size_t* n_leaves, size_t* leaf_len, size_t* n_digests, unsigned* cap_height) {
    if (!b) return fail(PCS_ERR_ARG, "batch is NULL");
    if (n_leaves) *n_leaves = b->n;
    if (leaf_len) *leaf_len = b->w + b->salt_w;
    if (n_digests) *n_digests = b->n_digests;
    if (cap_height) *cap_height = b->cap_height;
    return PCS_OK;
}

int pcs_batch_cap(const pcs_batch* b, uint64_t* cap) {
    BatchScope scope(b);
    if (!b || !cap) return fail(PCS_ERR_ARG, "NULL pointer");
    cudaStream_t st = g_ctx.stream;
    PCS_CUDA(cudaMemcpyAsync(cap, b->cap, ((size_t)32) << b->cap_height, cudaMemcpyDeviceToHost, st));
    PCS_CUDA(cudaStreamSynchronize(st));
    return PCS_OK;
}

int pcs_batch_digests(const pcs_batch* b, uint64_t* digests) {
    BatchScope scope(b);
    if (!b) return fail(PCS_ERR_ARG, "NULL pointer");
    if (b->n_digests == 0) return PCS_OK;
    if (!digests) return fail(PCS_ERR_ARG, "NULL pointer");
    cudaStream_t st = g_ctx.stream;
    PCS_CUDA(cudaMemcpyAsync(digests, b->digests, b->n_digests * 32, cudaMemcpyDeviceToHost, st));
    PCS_CUDA(cudaStreamSynchronize(st));
    return PCS_OK;
}

int pcs_batch_leaves(const pcs_batch* b, size_t first, size_t count, uint64_t* rows) {
    BatchScope scope(b);
    if (!b || !rows) return fail(PCS_ERR_ARG, "NULL pointer");
    if (first > b->n || count > b->n - first) return fail(PCS_ERR_ARG, "leaf range out of bounds");
    if (count == 0) return PCS_OK;
    cudaStream_t st = g_ctx.stream;
    size_t wt = b->w + b->salt_w;
    // transpose in slabs so that the staging buffer stays small
    const size_t slab = (size_t)1 << 18;
    DevBuf tmp;
    PCS_CUDA(tmp.alloc((count < slab ? count : slab) * wt * 8, st));
    for (size_t off = 0; off < count; off += slab) {
        size_t c = count - off < slab ? count - off : slab;
        PCS_CUDA(launch_transpose(b->lde + first + off, b->n, tmp.u64(), wt, wt, c, st));
        int rc = d2h_pageable(rows + off * wt, tmp.p, c * wt * 8, st);
        if (rc) return rc;
    }
    return PCS_OK;
}

int pcs_batch_get_rows(const pcs_batch* b, const uint64_t* leaf_indices, size_t n, uint64_t* rows) {
    BatchScope scope(b);
    if (!b || (n && (!leaf_indices || !rows))) return fail(PCS_ERR_ARG, "NULL pointer");
    if (n == 0) return PCS_OK;
    for (size_t k = 0; k < n; k++)
        if (leaf_indices[k] >= b->n) return fail(PCS_ERR_ARG, "leaf index out of bounds");
    cudaStream_t st = g_ctx.stream;
    size_t wt = b->w + b->salt_w;
    DevBuf idx, o;
    PCS_CUDA(idx.alloc(n * 8, st));
    PCS_CUDA(o.alloc(n * wt * 8, st));
    PCS_CUDA(cudaMemcpyAsync(idx.p, leaf_indices, n * 8, cudaMemcpyHostToDevice, st));
    PCS_CUDA(launch_gather_rows(b->lde, b->n, (uint32_t)wt, idx.u64(), n, o.u64(), st));
    PCS_CUDA(cudaMemcpyAsync(rows, o.p, n * wt * 8, cudaMemcpyDeviceToHost, st));
    PCS_CUDA(cudaStreamSynchronize(st));
    return PCS_OK;
}

int pcs_batch_lde_natural(const pcs_batch* b, size_t index_start, size_t step, size_t count, uint64_t* rows) {
    BatchScope scope(b);
    if (!b || (count && !rows)) return fail(PCS_ERR_ARG, "NULL pointer");
    if (count == 0) return PCS_OK;
    if (step == 0 || (step & (step - 1))) return fail(PCS_ERR_ARG, "step must be a power of two");
    // (index_start + k) * step < N for every k (oracle.rs:129-130 indexes merkle_tree.leaves with the reversed product)
    if (index_start > b->n / step || count > b->n / step - index_start) return fail(PCS_ERR_ARG, "point range out of bounds");
    cudaStream_t st = g_ctx.stream;
    const size_t w = b->w;                       // salt columns are dropped (oracle.rs:131-132)
    const unsigned lg_n = b->lg_d + b->rate_bits;
    // slabs of <= 256 MB: gather on the device, then stream to the caller's (usually pageable) rows
    const size_t SLAB = (size_t)256 << 20;
    size_t ch = SLAB / (w * 8);
    if (ch < 1) ch = 1;
    if (ch > count) ch = count;
    DevBuf tmp;
    PCS_CUDA(tmp.alloc(ch * w * 8, st));
    for (size_t off = 0; off < count; off += ch) {
        const size_t c = count - off < ch ? count - off : ch;
        PCS_CUDA(launch_lde_natural(b->lde, b->n, (uint32_t)w, lg_n, index_start + off, step, c, tmp.u64(), st));
        int rc = d2h_pageable(rows + off * w, tmp.p, c * w * 8, st);
        if (rc) return rc;
    }
    return PCS_OK;
}

int pcs_batch_prove(const pcs_batch* b, size_t leaf_index, uint64_t* siblings) {
    BatchScope scope(b);
    if (!b) return fail(PCS_ERR_ARG, "NULL pointer");
    if (leaf_index >= b->n) return fail(PCS_ERR_ARG, "leaf index out of bounds");
    unsigned lg_sub = b->lg_d + b->rate_bits - b->cap_height;
    if (lg_sub == 0) return PCS_OK;
    if (!siblings) return fail(PCS_ERR_ARG, "NULL pointer");
    cudaStream_t st = g_ctx.stream;
    DevBuf o;
    PCS_CUDA(o.alloc(lg_sub * 32, st));
    PCS_CUDA(launch_prove(b->digests, lg_sub, leaf_index, o.u64(), st));
    PCS_CUDA(cudaMemcpyAsync(siblings, o.p, lg_sub * 32, cudaMemcpyDeviceToHost, st));
    PCS_CUDA(cudaStreamSynchronize(st));
    return PCS_OK;
}

int pcs_batch_prove_many(const pcs_batch* b, const uint64_t* leaf_indices, size_t n, uint64_t* siblings) {
    BatchScope scope(b);
    if (!b || (n && !leaf_indices)) return fail(PCS_ERR_ARG, "NULL pointer");
    for (size_t k = 0; k < n; k++)
        if (leaf_indices[k] >= b->n) return fail(PCS_ERR_ARG, "leaf index out of bounds");
    unsigned lg_sub = b->lg_d + b->rate_bits - b->cap_height;
    if (lg_sub == 0 || n == 0) return PCS_OK;
    if (!siblings) return fail(PCS_ERR_ARG, "NULL pointer");
    cudaStream_t st = g_ctx.stream;
    DevBuf o, idx;
    PCS_CUDA(o.alloc(n * lg_sub * 32, st));
    PCS_CUDA(idx.alloc(n * 8, st));
    PCS_CUDA(cudaMemcpyAsync(idx.p, leaf_indices, n * 8, cudaMemcpyHostToDevice, st));
    PCS_CUDA(launch_prove_many(b->digests, lg_sub, idx.u64(), n, o.u64(), st));
    PCS_CUDA(cudaMemcpyAsync(siblings, o.p, n * lg_sub * 32, cudaMemcpyDeviceToHost, st));
    PCS_CUDA(cudaStreamSynchronize(st));
    return PCS_OK;
}

int pcs_batch_coeffs(const pcs_batch* b, size_t poly, uint64_t* coeffs) {
    BatchScope scope(b);
    if (!b || !coeffs) return fail(PCS_ERR_ARG, "NULL pointer");
    if (!b->coeffs) return fail(PCS_ERR_ARG, "coefficients were not kept (PCS_KEEP_COEFFS)");
    if (poly >= b->w) return fail(PCS_ERR_ARG, "polynomial index out of bounds");
    cudaStream_t st = g_ctx.stream;
    size_t d = (size_t)1 << b->lg_d;
    PCS_CUDA(cudaMemcpyAsync(coeffs, b->coeffs + poly * d, d * 8, cudaMemcpyDeviceToHost, st));
    PCS_CUDA(cudaStreamSynchronize(st));
    return PCS_OK;
}

int pcs_batch_all_coeffs(const pcs_batch* b, uint64_t* coeffs) {
    BatchScope scope(b);
    if (!b || !coeffs) return fail(PCS_ERR_ARG, "NULL pointer");
    if (!b->coeffs) return fail(PCS_ERR_ARG, "coefficients were not kept (PCS_KEEP_COEFFS)");
    cudaStream_t st = g_ctx.stream;
    PCS_CUDA(cudaMemcpyAsync(coeffs, b->coeffs, (b->w << b->lg_d) * 8, cudaMemcpyDeviceToHost, st));
    PCS_CUDA(cudaStreamSynchronize(st));
    return PCS_OK;
}

const uint64_t* pcs_batch_lde_dev(const pcs_batch* b) { return b ? b->lde : nullptr; }
const uint64_t* pcs_batch_coeffs_dev(const pcs_batch* b) { return b ? b->coeffs : nullptr; }
const uint64_t* pcs_batch_digests_dev(const pcs_batch* b) { return b ? b->digests : nullptr; }
const uint64_t* pcs_batch_cap_dev(const pcs_batch* b) { return b ? b->cap : nullptr; }

int pcs_batch_timings(const pcs_batch* b, float ms[5]) {
    BatchScope scope(b);
    if (!b || !ms) return fail(PCS_ERR_ARG, "NULL pointer");
    PCS_CUDA(cudaStreamSynchronize(g_ctx.stream));
    for (int i = 0; i < 5; i++) PCS_CUDA(cudaEventElapsedTime(&ms[i], b->ev[i], b->ev[i + 1]));
    for (size_t k = 0; k + 1 < b->absorb_ev.size(); k += 2) {   // streamed leaf hashing ran inside the LDE phase
        float t = 0;
        PCS_CUDA(cudaEventElapsedTime(&t, b->absorb_ev[k], b->absorb_ev[k + 1]));
        ms[1] -= t;
        ms[3] += t;
    }
    return PCS_OK;
}

int pcs_timing_totals(float ms[5], unsigned* n_commits, int reset) {
    PCS_NEED_INIT();
    PCS_CUDA(cudaStreamSynchronize(g_ctx.stream));
    drain_pending();
    if (ms)
        for (int i = 0; i < 5; i++) ms[i] = g_ctx.totals[i];
    if (n_commits) *n_commits = g_ctx.n_totals;
    if (reset) {
        for (auto& t : g_ctx.totals) t = 0;
        g_ctx.n_totals = 0;
    }
    return PCS_OK;
}

}  // extern "C"
