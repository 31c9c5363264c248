// Kernel (c): batched clause evaluation and unit propagation over the encoder's CNF (src/encoder.rs:435-667), used
// to verify GPU witnesses against the exact clauses the SAT solver receives (Solve::add_cnf,
// crates/repl/src/solver_runner.rs:12).
//
// Layout in HBM
//   clauses      CSR: lits int32[n_lits] (DIMACS-signed, 1-based), offsets u32[n_clauses+1]
//   assignments  BIT-SLICED over the batch: two planes pos/neg, u32[n_vars+1][nbw], bit b of word bw = assignment
//                32*bw+b.  pos = assigned True, neg = assigned False, neither = DontCare/unassigned.  One clause
//                evaluation is one LOP per literal for 32 assignments; consecutive threads take consecutive batch
//                words of the same clause, so plane reads are coalesced 128 B lines when the batch is >= 1024.
//   propagation  assigning a literal is ONE atomicOr on ONE plane (monotone, so the fixpoint is order independent);
//                rounds are synchronous — read the planes of the previous round, OR into a copy — so the number of
//                rounds is a property of the instance and equals the oracle's (oracle/capi.cpp tsso_cnf_propagate).
#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <cstring>

#include "engine.hpp"

struct tss_cnf {
    tss_engine* engine = nullptr;
    void* blob = nullptr;          // ONE stream-ordered allocation from the engine's pool: lits | offsets | pad4 | long_ids | n_long | active
    int32_t* lits = nullptr;
    uint32_t* offsets = nullptr;
    int4* pad4 = nullptr;          // [n_clauses] clauses of 1..4 literals padded with 0 (one 16-byte load per clause); x == 0: see long_ids
    int* active = nullptr;         // [n_clauses] scratch of cnf_complete_kernel: active clauses beyond what its shared memory holds
    int* long_ids = nullptr;       // clauses with more than 4 literals (and empty ones), in no particular order; long_ids[n_clauses] = their number
    int n_clauses = 0, n_vars = 0;
    int64_t n_lits = 0;
};

namespace tss {

// u8 assignments [n][stride] -> planes.  One thread per (var, batch word).
__global__ void cnf_pack_kernel(const uint8_t* __restrict__ a, long long n, int n_vars, int nbw, uint32_t* __restrict__ pos,
                                uint32_t* __restrict__ neg) {
    long long total = (long long)(n_vars + 1) * nbw;
    const long long stride = n_vars + 1;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        int bw = (int)(i / (n_vars + 1)), v = (int)(i % (n_vars + 1));  // consecutive threads -> consecutive vars (coalesced bytes)
        uint32_t p = 0, q = 0;
        for (int b = 0; b < 32; b++) {
            long long idx = (long long)bw * 32 + b;
            if (idx >= n) break;
            uint8_t x = a[idx * stride + v];
            p |= (uint32_t)(x == 1) << b;
            q |= (uint32_t)(x == 0) << b;
        }
        if (v == 0) { p = 0; q = 0; }
        pos[(long long)v * nbw + bw] = p;
        neg[(long long)v * nbw + bw] = q;
    }
}

__global__ void cnf_unpack_kernel(const uint32_t* __restrict__ pos, const uint32_t* __restrict__ neg, long long n, int n_vars,
                                  int nbw, uint8_t* __restrict__ a) {
    long long total = (long long)(n_vars + 1) * nbw;
    const long long stride = n_vars + 1;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        int bw = (int)(i / (n_vars + 1)), v = (int)(i % (n_vars + 1));
        uint32_t p = pos[(long long)v * nbw + bw], q = neg[(long long)v * nbw + bw];
        for (int b = 0; b < 32; b++) {
            long long idx = (long long)bw * 32 + b;
            if (idx >= n) break;
            a[idx * stride + v] = v == 0 ? 2 : (((p >> b) & 1u) ? 1 : (((q >> b) & 1u) ? 0 : 2));  // True wins a forced clash
        }
    }
}

__device__ __forceinline__ uint32_t batch_mask(long long n, int bw) {
    long long rem = n - (long long)bw * 32;
    return rem >= 32 ? 0xffffffffu : (rem <= 0 ? 0u : ((1u << rem) - 1u));
}

// Clause evaluation.  falsified bit = no literal of the clause is True.
// Thread = one word of the batch (32 assignments), blockIdx.y strides over the clauses: the clause's literals are uniform
// across the block (broadcast loads), the plane reads of a warp are 32 consecutive words, and there is no per-element
// index arithmetic (the first version recovered (clause, word) from a flat 64-bit index with a division: 156 instructions
// per clause and word, now ~30).
constexpr int CHECK_WPT = 4;  // words of the batch per thread: four independent plane reads in flight per literal
__global__ void cnf_check_kernel(const int32_t* __restrict__ lits, const uint32_t* __restrict__ offsets, int n_clauses, int nbw,
                                 long long n, const uint32_t* __restrict__ pos, const uint32_t* __restrict__ neg,
                                 int* __restrict__ n_falsified, int* __restrict__ first_falsified, bool chunked) {
    // (with one word per thread a warp had a single plane read in flight: ~54 warps x 128 B per ~600-cycle L2 round trip
    // = 2.9 TB/s, which is what it measured; staging the CSR slice in shared memory did not help, more loads in flight do)
    int bw[CHECK_WPT];
    uint32_t mask[CHECK_WPT];
#pragma unroll
    for (int j = 0; j < CHECK_WPT; j++) {
        bw[j] = (blockIdx.x * CHECK_WPT + j) * blockDim.x + threadIdx.x;
        mask[j] = bw[j] < nbw ? batch_mask(n, bw[j]) : 0u;
        bw[j] = bw[j] < nbw ? bw[j] : 0;   // (masked out: read word 0 instead of branching)
    }
    uint32_t any = 0;
#pragma unroll
    for (int j = 0; j < CHECK_WPT; j++) any |= mask[j];
    if (!any) return;
    // clause slices are CONTIGUOUS (clauses are emitted tile by tile, a variable's occurrences sit within a few tile rows of
    // each other): consecutive clauses of one block re-read the same plane words while they are still in L1
    // (measured, profiles/cnf_ab.py: 0.070 ms against 0.085 ms with strided slices; longer slices / other block orders change nothing)
    const int per = (n_clauses + gridDim.y - 1) / gridDim.y, c_begin = chunked ? blockIdx.y * per : blockIdx.y;
    const int c_end = chunked ? min(n_clauses, c_begin + per) : n_clauses, c_step = chunked ? 1 : gridDim.y;
    for (int c = c_begin; c < c_end; c += c_step) {
        uint32_t sat[CHECK_WPT] = {};
        const uint32_t k1 = offsets[c + 1];
        for (uint32_t k = offsets[c]; k < k1; k++) {
            const int l = lits[k];
            const uint32_t* row = (l > 0 ? pos : neg) + (size_t)(l > 0 ? l : -l) * nbw;
#pragma unroll
            for (int j = 0; j < CHECK_WPT; j++) sat[j] |= row[bw[j]];
        }
#pragma unroll
        for (int j = 0; j < CHECK_WPT; j++) {
            uint32_t bad = ~sat[j] & mask[j];
            while (bad) {
                int b = __ffs(bad) - 1;
                bad &= bad - 1;
                atomicAdd(&n_falsified[bw[j] * 32 + b], 1);
                atomicMin(&first_falsified[bw[j] * 32 + b], c);
            }
        }
    }
}

// One unit-propagation round over all clauses.  SYNCHRONOUS: every clause is evaluated against the planes as they stood at
// the start of the round (pos / neg, read only) and its forced literal is OR-ed into the next state (pos_n / neg_n, which
// the host initialises as a copy), so the outcome of a round does not depend on how the clause slices are scheduled and
// the number of rounds to the fixpoint is a property of the instance (the oracle's tsso_cnf_propagate counts the same).
__global__ void cnf_propagate_kernel(const int32_t* __restrict__ lits, const uint32_t* __restrict__ offsets, int n_clauses, int nbw,
                                     long long n, const uint32_t* __restrict__ pos, const uint32_t* __restrict__ neg,
                                     uint32_t* __restrict__ pos_n, uint32_t* __restrict__ neg_n, int* __restrict__ changed) {
    const int bw = blockIdx.x * blockDim.x + threadIdx.x;   // (same mapping as cnf_check_kernel)
    if (bw >= nbw) return;
    for (int c = blockIdx.y; c < n_clauses; c += gridDim.y) {
        const uint32_t k0 = offsets[c], k1 = offsets[c + 1];
        uint32_t sat = 0, un1 = 0, un2 = 0;
        for (uint32_t k = k0; k < k1; k++) {
            int l = lits[k], v = l > 0 ? l : -l;
            uint32_t p = pos[(long long)v * nbw + bw], q = neg[(long long)v * nbw + bw];
            sat |= l > 0 ? p : q;
            uint32_t u = ~(p | q);
            un2 |= un1 & u;
            un1 |= u;
        }
        uint32_t unit = ~sat & un1 & ~un2 & batch_mask(n, bw);
        if (!unit) continue;
        for (uint32_t k = k0; k < k1; k++) {
            int l = lits[k], v = l > 0 ? l : -l;
            uint32_t u = ~(pos[(long long)v * nbw + bw] | neg[(long long)v * nbw + bw]) & unit;
            if (u) {
                atomicOr(l > 0 ? &pos_n[(long long)v * nbw + bw] : &neg_n[(long long)v * nbw + bw], u);
                unit &= ~u;
                *changed = 1;
            }
        }
    }
}

// ONE witness, one launch (tss_witness_for_cnf): unit propagation to the fixpoint, open variables set False, every clause
// checked — the whole completion of a GPU layout into a model of the uploaded clauses.  The batch kernels above pay a launch
// and two plane copies per round plus a host round trip per eight rounds (210-610 us for one assignment of test/ex1 / ex2,
// r2 TSS_TRACE); here the assignment sits in shared memory as one byte per variable and ONE CTA sweeps the clauses until
// nothing changes.  Propagation is IN PLACE (a clause sees what earlier clauses of the same sweep forced): unit propagation is
// confluent — the fixpoint, and whether it holds a conflict, do not depend on the order — so the result equals the
// synchronous rounds of cnf_propagate_kernel (tests/test_gpu.py compares both); only the number of sweeps is smaller.
// out[0] = a clause left without a true or open literal (conflict) or -1, out[1] = clauses falsified by the completed
// assignment, out[2] = sweeps.
constexpr int COMPLETE_THREADS = 1024;
constexpr size_t COMPLETE_MAX_BYTES = 200 * 1024;   // variables + 1 that fit one CTA's shared memory

// padded copy of the short clauses (built once per upload, on the device): the sweep of cnf_complete_kernel then needs ONE
// independent, coalesced 16-byte load per clause instead of the dependent chain offsets -> literals (the first version walked
// the CSR: 33 us per sweep over the 25 K clauses of test/ex2.toml with the default-8 set, latency of one CTA's dependent L2 reads)
__global__ void cnf_short_kernel(const int32_t* __restrict__ lits, const uint32_t* __restrict__ offsets, int n_clauses, int4* __restrict__ pad4,
                                 int* __restrict__ long_ids) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_clauses) return;
    const uint32_t k0 = offsets[c], len = offsets[c + 1] - k0;
    int4 q = make_int4(0, 0, 0, 0);
    if (len >= 1 && len <= 4) {
        q.x = lits[k0];
        if (len > 1) q.y = lits[k0 + 1];
        if (len > 2) q.z = lits[k0 + 2];
        if (len > 3) q.w = lits[k0 + 3];
    } else {
        long_ids[atomicAdd(&long_ids[n_clauses], 1)] = c;
    }
    pad4[c] = q;
}

// one padded clause against the assignment bytes: 0 = satisfied, 1 = unit (forces `forced`), 2 = every literal false, 3 = open
// in two or more literals
__device__ __forceinline__ int clause4(const uint8_t* a, int l0, int l1, int l2, int l3, int& forced) {
    const uint8_t x0 = a[abs(l0)], x1 = a[abs(l1)], x2 = a[abs(l2)], x3 = a[abs(l3)];
    if (x0 == (l0 > 0) || x1 == (l1 > 0) || x2 == (l2 > 0) || x3 == (l3 > 0)) return 0;
    const int open = (x0 == 2) + (x1 == 2) + (x2 == 2) + (x3 == 2);
    forced = x0 == 2 ? l0 : (x1 == 2 ? l1 : (x2 == 2 ? l2 : l3));
    return open == 1 ? 1 : (open == 0 ? 2 : 3);
}
// the same for a clause of any length from the CSR arrays
__device__ __forceinline__ int clause_csr(const uint8_t* a, const int32_t* __restrict__ lits, const uint32_t* __restrict__ offsets, int c, int& forced) {
    const uint32_t k1 = offsets[c + 1];
    int open = 0;
    for (uint32_t k = offsets[c]; k < k1; k++) {
        const int l = lits[k];
        const uint8_t x = a[abs(l)];
        if (x == (l > 0)) return 0;
        if (x == 2) { open++; forced = l; }
    }
    return open == 1 ? 1 : (open == 0 ? 2 : 3);
}

// Shared memory: [assignment bytes, padded to 16][ids of the first `cap` active clauses][those clauses, padded int4].
// The FIRST sweep visits every clause (one coalesced 16-byte load each) and keeps the ones that are not satisfied yet — `active`:
// only those can ever force a literal or end up falsified, a satisfied clause stays satisfied because assignments are never
// withdrawn — so the later sweeps and the final check walk that list alone, out of shared memory (entries beyond `cap` spill to a
// global list and are re-read from pad4).  For a GPU layout every base variable arrives decided and the list is the cardinality
// network of the limit (2.8 K of the 25 K clauses of test/ex2.toml with the default-8 set): < 1 us per sweep instead of 10 us
// (a full sweep is bound by bank conflicts of the random byte reads on ONE SM's shared memory).
__global__ void __launch_bounds__(COMPLETE_THREADS, 1) cnf_complete_kernel(const int32_t* __restrict__ lits, const uint32_t* __restrict__ offsets,
                                                                            const int4* __restrict__ pad4, const int* __restrict__ long_ids,
                                                                            int* __restrict__ spill, int n_clauses, int n_vars, int cap,
                                                                            uint8_t* __restrict__ a_glob, int* __restrict__ out) {
    extern __shared__ __align__(16) uint8_t smem[];
    uint8_t* a = smem;   // 0 = False, 1 = True, 2 = unassigned; a[0] = 1 makes the padding literal 0 a false, decided literal
    int* act_id = reinterpret_cast<int*>(smem + (((size_t)n_vars + 1 + 15) & ~(size_t)15));
    int4* act_cl = reinterpret_cast<int4*>(act_id + cap);   // (cap is a multiple of 4)
    __shared__ int conflict, falsified, n_active;
    for (int v = threadIdx.x; v <= n_vars; v += blockDim.x) a[v] = v ? a_glob[v] : 1;
    if (threadIdx.x == 0) { conflict = 0x7fffffff; falsified = 0; n_active = 0; }
    const int n_long = long_ids[n_clauses];
    __syncthreads();
    auto act = [&](int c, int r, int forced, int& changed) {
        if (r == 1) { a[abs(forced)] = forced > 0; changed = 1; }   // (two clauses forcing opposite values: one of them ends up falsified, found by the next sweep)
        else if (r == 2) atomicMin(&conflict, c);
    };
    auto keep = [&](int c, int4 q) {
        const int slot = atomicAdd(&n_active, 1);
        if (slot < cap) { act_id[slot] = c; act_cl[slot] = q; }
        else spill[slot - cap] = c;
    };
    // ---- first sweep: every clause
    int changed = 0, forced = 0, sweeps = 1;
#pragma unroll 4
    for (int c = threadIdx.x; c < n_clauses; c += COMPLETE_THREADS) {
        const int4 q = pad4[c];
        if (q.x == 0) continue;   // long or empty: in long_ids
        const int r = clause4(a, q.x, q.y, q.z, q.w, forced);
        if (r == 0) continue;
        act(c, r, forced, changed);
        keep(c, q);
    }
    for (int i = threadIdx.x; i < n_long; i += COMPLETE_THREADS) {
        const int c = long_ids[i], r = clause_csr(a, lits, offsets, c, forced);
        if (r == 0) continue;
        act(c, r, forced, changed);
        keep(c, make_int4(0, 0, 0, 0));
    }
    auto revisit = [&](int i, int& forced) -> int {   // active clause i against the current assignment; returns its clause id in `forced`'s place via act()
        int c;
        int4 q;
        if (i < cap) { c = act_id[i]; q = act_cl[i]; }
        else { c = spill[i - cap]; q = pad4[c]; }
        const int r = q.x ? clause4(a, q.x, q.y, q.z, q.w, forced) : clause_csr(a, lits, offsets, c, forced);
        return r | (c << 2);
    };
    // ---- later sweeps: the clauses that were not satisfied then
    while (__syncthreads_or(changed) && conflict == 0x7fffffff) {   // (the barrier also publishes n_active, the list and the forced values)
        __syncthreads();   // (nobody raises `conflict` for the next sweep before everyone has read it)
        changed = 0;
        const int n = n_active;
        for (int i = threadIdx.x; i < n; i += COMPLETE_THREADS) {
            const int rc = revisit(i, forced);
            act(rc >> 2, rc & 3, forced, changed);
        }
        sweeps++;
    }
    if (conflict == 0x7fffffff) {   // open variables False, then the clauses that were ever unsatisfied against the completed assignment
        for (int v = threadIdx.x + 1; v <= n_vars; v += blockDim.x)
            if (a[v] == 2) a[v] = 0;
        __syncthreads();
        int bad = 0;
        const int n = n_active;
        for (int i = threadIdx.x; i < n; i += COMPLETE_THREADS) bad += (revisit(i, forced) & 3) == 2;
        if (bad) atomicAdd(&falsified, bad);
        __syncthreads();
    }
    for (int v = threadIdx.x + 1; v <= n_vars; v += blockDim.x) a_glob[v] = a[v];
    if (threadIdx.x == 0) { out[0] = conflict == 0x7fffffff ? -1 : conflict; out[1] = falsified; out[2] = sweeps; }
}

__global__ void cnf_relax_kernel(const uint32_t* __restrict__ pos, const uint32_t* __restrict__ neg, uint32_t* __restrict__ pos2,
                                 uint32_t* __restrict__ neg2, long long total) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        uint32_t p = pos[i], q = neg[i], u = ~(p | q);
        pos2[i] = p | u;
        neg2[i] = (q & ~p) | u;  // True wins a forced clash
    }
}

// clause kernels: x = words of the batch (block of 32..128 threads), y = clause slices filling the device
static void clause_geometry(tss_engine* e, int n_clauses, int nbw, dim3& grid, dim3& block, int words_per_thread = 1) {
    const int bx = nbw >= 128 ? 128 : ((nbw + 31) / 32) * 32;
    const int gx = (nbw + bx * words_per_thread - 1) / (bx * words_per_thread);
    long long gy = (long long)e->prop.multiProcessorCount * 16 * 256 / ((long long)gx * bx);
    gy = gy < 1 ? 1 : (gy > n_clauses ? n_clauses : gy);
    if (gy > 65535) gy = 65535;
    grid = dim3((unsigned)gx, (unsigned)(gy < 1 ? 1 : gy), 1);
    block = dim3((unsigned)bx, 1, 1);
}

static bool cnf_chunked() { const char* v = getenv("TSS_CNF_STRIDED"); return !(v && v[0] == '1'); }   // (A/B switch for profiles/cnf_stream.py)

static unsigned grid_for(tss_engine* e, long long total) {
    long long blocks = (total + 255) / 256, cap = (long long)e->prop.multiProcessorCount * 16;
    return (unsigned)(blocks < 1 ? 1 : (blocks > cap ? cap : blocks));
}

}  // namespace tss

// Completion of ONE assignment in one launch (cnf_complete_kernel); TSS_E_UNSUPPORTED when the variables do not fit one CTA's
// shared memory (the caller then takes the batch kernels).  conflict: clause index or -1; n_falsified counts the clauses the
// completed assignment (open variables False) falsifies.
int tss::cnf_complete_single(tss_engine* e, const tss_cnf* c, uint8_t* assignment, int32_t* conflict, int32_t* n_falsified) {
    const size_t bytes = (size_t)c->n_vars + 1;
    if (bytes > COMPLETE_MAX_BYTES || !c->pad4) return TSS_E_UNSUPPORTED;
    TSS_CUDA(e, cudaSetDevice(e->device));
    uint8_t* a_dev = (uint8_t*)e->dev(0, bytes + 16);
    uint8_t* a_pin = (uint8_t*)e->pin(1, bytes + 16);
    if (!a_dev || !a_pin) return TSS_E_CUDA;
    int* out_dev = (int*)(a_dev + ((bytes + 3) & ~(size_t)3));
    int* out_pin = (int*)(a_pin + ((bytes + 3) & ~(size_t)3));
    // shared memory: the assignment bytes, then room for the active clauses (20 bytes each: id + padded literals)
    const size_t a_bytes = (bytes + 15) & ~(size_t)15, smem_cap = (size_t)220 * 1024;
    const int cap = (int)std::min<size_t>(((size_t)c->n_clauses + 3) & ~(size_t)3, ((smem_cap - a_bytes) / 20) & ~(size_t)3);
    const size_t smem = a_bytes + (size_t)20 * cap;
    TSS_CUDA(e, cudaFuncSetAttribute(cnf_complete_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_cap));
    std::memcpy(a_pin, assignment, bytes);
    TSS_CUDA(e, cudaMemcpyAsync(a_dev, a_pin, bytes, cudaMemcpyHostToDevice, e->stream));
    TSS_CUDA(e, cudaEventRecord(e->ev0, e->stream));
    cnf_complete_kernel<<<1, COMPLETE_THREADS, smem, e->stream>>>(c->lits, c->offsets, c->pad4, c->long_ids, c->active, c->n_clauses, c->n_vars, cap, a_dev, out_dev);
    TSS_CHECK_LAUNCH(e);
    TSS_CUDA(e, cudaEventRecord(e->ev1, e->stream));
    e->stats.kernel_launches++;
    TSS_CUDA(e, cudaMemcpyAsync(a_pin, a_dev, ((bytes + 3) & ~(size_t)3) + 12, cudaMemcpyDeviceToHost, e->stream));
    TSS_CUDA(e, cudaStreamSynchronize(e->stream));
    std::memcpy(assignment, a_pin, bytes);
    *conflict = out_pin[0];
    *n_falsified = out_pin[1];
    e->stats.clauses_checked += (uint64_t)c->n_clauses * (uint64_t)(out_pin[2] + 1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e->ev0, e->ev1);
    e->stats.device_ms = ms;
    return TSS_OK;
}

using namespace tss;

extern "C" {

int tss_cnf_upload(tss_engine* e, const int32_t* lits, const uint32_t* offsets, int32_t n_clauses, int32_t n_vars, tss_cnf** out) {
    if (!e) return TSS_E_INVALID;
    if (!out || !offsets || n_clauses < 0 || n_vars < 0) return e->fail(TSS_E_INVALID, "tss_cnf_upload: bad arguments");
    int64_t n_lits = offsets[n_clauses];
    if (n_lits > 0 && !lits) return e->fail(TSS_E_INVALID, "tss_cnf_upload: lits is null");
    for (int64_t k = 0; k < n_lits; k++) {
        int v = lits[k] > 0 ? lits[k] : -lits[k];
        if (v == 0 || v > n_vars) return e->fail(TSS_E_INVALID, "tss_cnf_upload: literal %d out of range (n_vars = %d)", lits[k], n_vars);
    }
    TSS_CUDA(e, cudaSetDevice(e->device));
    // one stream-ordered allocation from the engine's pool (the bound-tightening loop uploads a CNF per iteration: cudaMalloc +
    // cudaFree cost more than the copy), nothing synchronises: the copies and the short-clause build are ordered on the stream
    const bool fused = (size_t)n_vars + 1 <= COMPLETE_MAX_BYTES;
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const size_t o_lits = 0, o_offs = o_lits + al(sizeof(int32_t) * (size_t)(n_lits > 0 ? n_lits : 1)), o_short = o_offs + al(sizeof(uint32_t) * ((size_t)n_clauses + 1)),
                 o_long = o_short + (fused ? al(sizeof(int4) * (size_t)(n_clauses > 0 ? n_clauses : 1)) : 0),
                 o_act = o_long + (fused ? al(sizeof(int) * ((size_t)n_clauses + 1)) : 0), total = o_act + (fused ? al(sizeof(int) * ((size_t)n_clauses + 1)) : 0);
    tss_cnf* c = new tss_cnf();
    c->engine = e; c->n_clauses = n_clauses; c->n_vars = n_vars; c->n_lits = n_lits;
    cudaError_t err = e->pool ? cudaMallocFromPoolAsync(&c->blob, total, e->pool, e->stream) : cudaMallocAsync(&c->blob, total, e->stream);
    if (err == cudaSuccess) {
        c->lits = (int32_t*)((char*)c->blob + o_lits);
        c->offsets = (uint32_t*)((char*)c->blob + o_offs);
        if (fused) { c->pad4 = (int4*)((char*)c->blob + o_short); c->long_ids = (int*)((char*)c->blob + o_long); }
        if (fused) c->active = (int*)((char*)c->blob + o_act);
    }
    if (err == cudaSuccess && n_lits) err = cudaMemcpyAsync(c->lits, lits, sizeof(int32_t) * (size_t)n_lits, cudaMemcpyHostToDevice, e->stream);
    if (err == cudaSuccess) err = cudaMemcpyAsync(c->offsets, offsets, sizeof(uint32_t) * (size_t)(n_clauses + 1), cudaMemcpyHostToDevice, e->stream);
    if (err == cudaSuccess && fused) {
        err = cudaMemsetAsync(c->long_ids + n_clauses, 0, sizeof(int), e->stream);
        if (err == cudaSuccess && n_clauses > 0) {
            cnf_short_kernel<<<(n_clauses + 255) / 256, 256, 0, e->stream>>>(c->lits, c->offsets, n_clauses, c->pad4, c->long_ids);
            err = cudaGetLastError();
            e->stats.kernel_launches++;
        }
    }
    if (err != cudaSuccess) { tss_cnf_destroy(c); return e->fail(TSS_E_CUDA, "tss_cnf_upload: %s", cudaGetErrorString(err)); }
    *out = c;
    return TSS_OK;
}

int tss_cnf_complete(tss_engine* e, const tss_cnf* c, uint8_t* assignment, int32_t* out_conflict, int32_t* out_n_falsified) {
    if (!e) return TSS_E_INVALID;
    if (!c || !assignment || !out_conflict || !out_n_falsified) return e->fail(TSS_E_INVALID, "tss_cnf_complete: bad arguments");
    *out_conflict = -1;
    *out_n_falsified = 0;
    int rc = cnf_complete_single(e, c, assignment, out_conflict, out_n_falsified);
    if (rc != TSS_E_UNSUPPORTED) return rc;
    // more variables than one CTA's shared memory holds: the batch kernels, round by round
    rc = tss_cnf_propagate(e, c, assignment, 1, out_conflict, nullptr);
    if (rc < 0 || *out_conflict >= 0) return rc;
    for (int v = 1; v <= c->n_vars; v++)
        if (assignment[v] == 2) assignment[v] = 0;
    return tss_cnf_check(e, c, assignment, 1, out_n_falsified, nullptr);
}

int tss_cnf_num_vars(const tss_cnf* c) { return c ? c->n_vars : TSS_E_INVALID; }

void tss_cnf_destroy(tss_cnf* c) {
    if (!c) return;
    // stream-ordered free into the engine's pool (no device synchronisation); a handle that outlived its engine frees synchronously
    if (c->blob) {
        if (tss::engine_alive(c->engine)) cudaFreeAsync(c->blob, c->engine->stream);
        else cudaFree(c->blob);
    }
    delete c;
}

// scratch slots: 0 = u8 assignments, 1 = pos plane, 2 = neg plane, 3 = int outputs
static int cnf_stage(tss_engine* e, const tss_cnf* c, const uint8_t* assignments, int64_t n, uint32_t** pos, uint32_t** neg, int* nbw_out) {
    const int nbw = (int)((n + 31) / 32);
    const size_t abytes = (size_t)n * (c->n_vars + 1), pbytes = sizeof(uint32_t) * (size_t)(c->n_vars + 1) * nbw;
    uint8_t* a = (uint8_t*)e->dev(0, abytes);
    *pos = (uint32_t*)e->dev(1, pbytes);
    *neg = (uint32_t*)e->dev(2, pbytes);
    if (!a || !*pos || !*neg) return TSS_E_CUDA;
    TSS_CUDA(e, cudaMemcpyAsync(a, assignments, abytes, cudaMemcpyHostToDevice, e->stream));
    cnf_pack_kernel<<<grid_for(e, (long long)(c->n_vars + 1) * nbw), 256, 0, e->stream>>>(a, n, c->n_vars, nbw, *pos, *neg);
    TSS_CHECK_LAUNCH(e);
    e->stats.kernel_launches++;
    *nbw_out = nbw;
    return TSS_OK;
}

int tss_cnf_check(tss_engine* e, const tss_cnf* c, const uint8_t* assignments, int64_t n, int32_t* out_n_falsified,
                  int32_t* out_first_falsified) {
    if (!e) return TSS_E_INVALID;
    if (!c || !assignments || n < 0 || !out_n_falsified) return e->fail(TSS_E_INVALID, "tss_cnf_check: bad arguments");
    if (n == 0) return TSS_OK;
    TSS_CUDA(e, cudaSetDevice(e->device));
    uint32_t *pos, *neg;
    int nbw;
    int rc = cnf_stage(e, c, assignments, n, &pos, &neg, &nbw);
    if (rc) return rc;
    int* outs = (int*)e->dev(3, sizeof(int) * (size_t)nbw * 64);
    if (!outs) return TSS_E_CUDA;
    int *cnt = outs, *first = outs + (size_t)nbw * 32;
    TSS_CUDA(e, cudaMemsetAsync(cnt, 0, sizeof(int) * (size_t)nbw * 32, e->stream));
    TSS_CUDA(e, cudaMemsetAsync(first, 0x7f, sizeof(int) * (size_t)nbw * 32, e->stream));
    TSS_CUDA(e, cudaEventRecord(e->ev0, e->stream));
    if (c->n_clauses > 0) {
        dim3 cg, cb;
        clause_geometry(e, c->n_clauses, nbw, cg, cb, CHECK_WPT);
        cnf_check_kernel<<<cg, cb, 0, e->stream>>>(c->lits, c->offsets, c->n_clauses, nbw, n, pos, neg, cnt, first, cnf_chunked());
        TSS_CHECK_LAUNCH(e);
        e->stats.kernel_launches++;
    }
    TSS_CUDA(e, cudaEventRecord(e->ev1, e->stream));
    e->stats.clauses_checked += (uint64_t)c->n_clauses * (uint64_t)n;
    TSS_CUDA(e, cudaMemcpyAsync(out_n_falsified, cnt, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, e->stream));
    std::vector<int> tmp;
    if (out_first_falsified) TSS_CUDA(e, cudaMemcpyAsync(out_first_falsified, first, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, e->stream));
    TSS_CUDA(e, cudaStreamSynchronize(e->stream));
    float ms = 0;
    cudaEventElapsedTime(&ms, e->ev0, e->ev1);
    e->stats.device_ms = ms;
    if (out_first_falsified)
        for (int64_t i = 0; i < n; i++)
            if (out_n_falsified[i] == 0) out_first_falsified[i] = -1;
    return TSS_OK;
}

int tss_cnf_propagate(tss_engine* e, const tss_cnf* c, uint8_t* assignments, int64_t n, int32_t* out_conflict, int32_t* out_rounds) {
    if (!e) return TSS_E_INVALID;
    if (!c || !assignments || n < 0) return e->fail(TSS_E_INVALID, "tss_cnf_propagate: bad arguments");
    if (out_rounds) *out_rounds = 0;
    if (n == 0) return TSS_OK;
    TSS_CUDA(e, cudaSetDevice(e->device));
    uint32_t *pos, *neg;
    int nbw;
    int rc = cnf_stage(e, c, assignments, n, &pos, &neg, &nbw);
    if (rc) return rc;
    constexpr int BATCH = 8;  // rounds queued per host round trip, each with its own "changed" flag
    int* outs = (int*)e->dev(3, sizeof(int) * ((size_t)nbw * 64 + BATCH));
    int* flag_host = (int*)e->pin(0, sizeof(int) * BATCH);
    const size_t plane_bytes = sizeof(uint32_t) * (size_t)(c->n_vars + 1) * nbw;
    uint32_t* pos_b = (uint32_t*)e->dev(4, plane_bytes);   // the other half of the double buffer (slots 4, 5; reused by the conflict check below)
    uint32_t* neg_b = (uint32_t*)e->dev(5, plane_bytes);
    if (!outs || !flag_host || !pos_b || !neg_b) return TSS_E_CUDA;
    int *cnt = outs, *first = outs + (size_t)nbw * 32, *changed = outs + (size_t)nbw * 64;
    dim3 cg(1, 1, 1), cb(32, 1, 1);
    if (c->n_clauses > 0) clause_geometry(e, c->n_clauses, nbw, cg, cb);
    int rounds = 0;
    TSS_CUDA(e, cudaEventRecord(e->ev0, e->stream));
    // every productive round assigns >= 1 variable; a round that changes nothing is the fixpoint (later rounds of the same
    // batch are no-ops), so the reported round count is exactly that of a round-by-round loop
    for (bool fix = false; c->n_clauses > 0 && rounds <= c->n_vars && !fix;) {
        if (e->interrupted()) break;
        TSS_CUDA(e, cudaMemsetAsync(changed, 0, sizeof(int) * BATCH, e->stream));
        for (int b = 0; b < BATCH; b++) {
            // next state starts as a copy of the current one; the round ORs the forced literals into it; then they swap roles
            TSS_CUDA(e, cudaMemcpyAsync(pos_b, pos, plane_bytes, cudaMemcpyDeviceToDevice, e->stream));
            TSS_CUDA(e, cudaMemcpyAsync(neg_b, neg, plane_bytes, cudaMemcpyDeviceToDevice, e->stream));
            cnf_propagate_kernel<<<cg, cb, 0, e->stream>>>(c->lits, c->offsets, c->n_clauses, nbw, n, pos, neg, pos_b, neg_b, changed + b);
            TSS_CHECK_LAUNCH(e);
            std::swap(pos, pos_b);
            std::swap(neg, neg_b);
        }
        e->stats.kernel_launches += BATCH;
        TSS_CUDA(e, cudaMemcpyAsync(flag_host, changed, sizeof(int) * BATCH, cudaMemcpyDeviceToHost, e->stream));
        TSS_CUDA(e, cudaStreamSynchronize(e->stream));
        for (int b = 0; b < BATCH && !fix; b++) {
            rounds++;
            e->stats.clauses_checked += (uint64_t)c->n_clauses * (uint64_t)n;
            fix = flag_host[b] == 0;
        }
    }
    if (out_rounds) *out_rounds = rounds;
    // conflicts at the fixpoint = clauses whose literals are all False (True wins a forced clash, so the clause that
    // forced the opposite value shows up as falsified)
    TSS_CUDA(e, cudaMemsetAsync(cnt, 0, sizeof(int) * (size_t)nbw * 32, e->stream));
    TSS_CUDA(e, cudaMemsetAsync(first, 0x7f, sizeof(int) * (size_t)nbw * 32, e->stream));
    if (c->n_clauses > 0) {
        // a clause is in conflict only if it has no unassigned literal: evaluate with "unassigned counts as True"
        // by checking against pos|~assigned is wrong for negatives, so run the check on dedicated planes:
        // reuse cnf_check_kernel with pos' = pos | unassigned, neg' = neg | unassigned
        uint32_t *pos2 = pos_b, *neg2 = neg_b;   // the idle half of the double buffer
        cnf_relax_kernel<<<grid_for(e, (long long)(c->n_vars + 1) * nbw), 256, 0, e->stream>>>(pos, neg, pos2, neg2, (long long)(c->n_vars + 1) * nbw);
        TSS_CHECK_LAUNCH(e);
        dim3 kg, kb;
        clause_geometry(e, c->n_clauses, nbw, kg, kb, CHECK_WPT);
        cnf_check_kernel<<<kg, kb, 0, e->stream>>>(c->lits, c->offsets, c->n_clauses, nbw, n, pos2, neg2, cnt, first, cnf_chunked());
        TSS_CHECK_LAUNCH(e);
        e->stats.kernel_launches += 2;
    }
    TSS_CUDA(e, cudaEventRecord(e->ev1, e->stream));
    uint8_t* a = (uint8_t*)e->scratch[0].ptr;   // (pos / neg: whichever half of the double buffer holds the fixpoint)
    cnf_unpack_kernel<<<grid_for(e, (long long)(c->n_vars + 1) * nbw), 256, 0, e->stream>>>(pos, neg, n, c->n_vars, nbw, a);
    TSS_CHECK_LAUNCH(e);
    e->stats.kernel_launches++;
    TSS_CUDA(e, cudaMemcpyAsync(assignments, a, (size_t)n * (c->n_vars + 1), cudaMemcpyDeviceToHost, e->stream));
    std::vector<int> cnt_host((size_t)n);
    if (out_conflict) {
        TSS_CUDA(e, cudaMemcpyAsync(out_conflict, first, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, e->stream));
        TSS_CUDA(e, cudaMemcpyAsync(cnt_host.data(), cnt, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, e->stream));
    }
    TSS_CUDA(e, cudaStreamSynchronize(e->stream));
    float ms = 0;
    cudaEventElapsedTime(&ms, e->ev0, e->ev1);
    e->stats.device_ms = ms;
    if (out_conflict)
        for (int64_t i = 0; i < n; i++)
            if (cnt_host[(size_t)i] == 0) out_conflict[i] = -1;
    return TSS_OK;
}

}  // extern "C"

