// Kernel (b), one chain per THREAD, for grids of at most 16 rows and 26 columns (rect 16x16, test/ex1-3).
//
// ncu on the warp-per-chain kernels (profiles/r1_sls_kernel.md, r1_sls_h16_kernel.md) shows the ALU pipe as the limiter
// with ~276 warp instructions per chain step, most of them cross-lane plumbing (indexed shuffles to fetch window rows,
// ballots, min/max butterflies) and per-chain "uniform" work that all lanes of a (half-)warp repeat.  Here a chain is
// one thread: its bitboards live in shared memory as [row][thread] words (bank = thread, conflict free for any row
// index), every candidate layout is scored by that thread alone (7 row loads, shift/mask/pack, 2 AND+POPC), there are
// no shuffles, ballots or reductions in the step loop, and a warp instruction advances 32 layouts.
//
// The step rule, RNG and tie-breaks are EXACTLY those of sls_spec.hpp: the parity tests replay this kernel against the
// same CPU model as sls.cu / sls_h16.cu (bit-identical trajectories).  Implementation choices that differ:
//   * columns are stored shifted left by 3 and the reach windows re-anchored at x-3 when the table is loaded, so a
//     window row is always `(row >> x) & 0x7f` (no anchor clamp); rows are addressed modulo 16 (rows outside the grid
//     meet zero window bits);
//   * the site list lives in global memory as [index][chain] words (coalesced across the warp, L1 resident) and carries
//     the step at which each support was added — for a current support that IS its last flip, so the "young support"
//     test of the spec needs no per-site stamp array;
//   * "recently removed" (the tabu test of add candidates, which are never current supports) is answered from a ring
//     of the sites removed in the last 32 steps (one slot per step), folded once per step into a 7x7 window mask
//     around the chosen uncovered tile.  Equivalent to the spec's 16-bit stamps for epochs of at most 32768 steps —
//     the engine never launches longer ones (tss_search_run splits them).
//   * add candidates: the 13x13 neighbourhood of the uncovered tile is fetched once into registers; the 25 diamond
//     cells are then scored with static shifts in a fully unrolled loop.
#include "engine.hpp"
#include "sls_spec.hpp"

namespace tss {
namespace slst {

using namespace tss::sls;

constexpr int NT = 128;          // chains per CTA
constexpr int MAXW = 26;         // 3 + w + 3 <= 32 bits
constexpr uint32_t NONE = 0xffffu;

struct Board {
    uint32_t w[16][NT];
};
struct Smem {
    uint2 tab[512];              // reach windows of the 16x32 sites, anchored at x-3
    uint32_t C[16];              // terrain rows, columns shifted by 3
    Board U, O, c0, c1, c2, c3, c4;
    uint16_t ring[32][NT];       // site removed at step s in slot s & 31 (NONE if none)
};

__device__ __forceinline__ int pick_rotated(uint32_t bits, uint32_t o) {
    uint32_t rot = __funnelshift_r(bits, bits, o);
    return (int)((__ffs(rot) - 1 + o) & 31u);
}

// popcount(B & R(site (x, y))): every thread scores its own site on its own board
__device__ __forceinline__ int score(const Board& B, int tid, int x, int y, uint2 win) {
    uint32_t lo = 0, hi = 0;
#pragma unroll
    for (int j = 6; j >= 4; j--) hi = hi * 128u + ((B.w[(y - 3 + j) & 15][tid] >> x) & 0x7fu);
#pragma unroll
    for (int j = 3; j >= 0; j--) lo = lo * 128u + ((B.w[(y - 3 + j) & 15][tid] >> x) & 0x7fu);
    return __popc(lo & win.x) + __popc(hi & win.y);
}

// adds (ADD) or removes the cover of site v on the five count planes and refreshes U, O and the non-empty-row mask
template <bool ADD>
__device__ __forceinline__ void flip(Smem& sm, int tid, int v, uint32_t& rowmask) {
    const int x = v & 31, y = v >> 5;
    const uint2 win = sm.tab[v];
#pragma unroll
    for (int j = 0; j < 7; j++) {
        const int r = (y - 3 + j) & 15;
        uint32_t m = ((j < 4 ? win.x >> (7 * j) : win.y >> (7 * (j - 4))) & 0x7fu) << x;
        uint32_t a0 = sm.c0.w[r][tid], a1 = sm.c1.w[r][tid], a2 = sm.c2.w[r][tid], t;
        if (ADD) {
            t = a0 & m; a0 ^= m; m = t;
            t = a1 & m; a1 ^= m; m = t;
            t = a2 & m; a2 ^= m; m = t;
        } else {
            t = ~a0 & m; a0 ^= m; m = t;
            t = ~a1 & m; a1 ^= m; m = t;
            t = ~a2 & m; a2 ^= m; m = t;
        }
        sm.c0.w[r][tid] = a0; sm.c1.w[r][tid] = a1; sm.c2.w[r][tid] = a2;
        uint32_t a3 = sm.c3.w[r][tid], a4 = sm.c4.w[r][tid];
        if (m) {  // a count crossing 7 <-> 8: rare
            if (ADD) { t = a3 & m; a3 ^= m; a4 ^= t; } else { t = ~a3 & m; a3 ^= m; a4 ^= t; }
            sm.c3.w[r][tid] = a3; sm.c4.w[r][tid] = a4;
        }
        const uint32_t hi = a1 | a2 | a3 | a4, C = sm.C[r];
        const uint32_t Un = C & ~(a0 | hi);
        sm.U.w[r][tid] = Un;
        sm.O.w[r][tid] = a0 & ~hi & C;
        rowmask = Un ? (rowmask | (1u << r)) : (rowmask & ~(1u << r));
    }
}

__device__ __forceinline__ constexpr int cell_dy(int i) { return (i >= 1) + (i >= 4) + (i >= 9) + (i >= 16) + (i >= 21) + (i >= 24) - 3; }
__device__ __forceinline__ constexpr int cell_dx(int i) {
    const int r = cell_dy(i) + 3;
    const int start = r <= 4 ? r * r : (r == 5 ? 21 : 24);
    const int ady = r >= 3 ? r - 3 : 3 - r;
    return (i - start) - (3 - ady);
}
__device__ __forceinline__ uint32_t shift_static(uint32_t v, int s) { return s >= 0 ? v << s : v >> (-s); }

// rows of the support bitboard of a site list (rare path: recording a best layout, writing the state back)
__device__ __noinline__ void list_to_rows(const uint32_t* sl, size_t stride, int k, uint32_t* out32) {
    uint32_t rows[16];
    for (int r = 0; r < 16; r++) rows[r] = 0;
    for (int i = 0; i < k; i++) {
        const uint32_t v = sl[(size_t)i * stride] & 0x1ffu;
        rows[v >> 5] |= 1u << (v & 31u);
    }
    for (int r = 0; r < 16; r++) out32[r] = rows[r];
}

__global__ void __launch_bounds__(NT, 3) sls_t16_kernel(const uint32_t* __restrict__ terrain_rows, const uint2* __restrict__ rtabs,
                                                       ChainState* __restrict__ states, uint32_t* __restrict__ site_lists, size_t stride,
                                                       int n_chains, int chains_per_terrain, uint32_t chain_offset, uint64_t seed,
                                                       long long steps, const int* __restrict__ bounds, int target, int noise_pct,
                                                       const volatile int* interrupt, unsigned long long* __restrict__ totals) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
    const int tid = threadIdx.x;
    const int chain = blockIdx.x * NT + tid;
    const int terrain = chains_per_terrain > 0 ? (blockIdx.x * NT) / chains_per_terrain : 0;
    for (int i = tid; i < 512; i += NT) {
        const uint2 s = rtabs[(size_t)terrain * 1024 + i];
        const int x = i & 31, sh = x < 3 ? 3 - x : 0;  // table windows are anchored at max(x-3, 0): re-anchor at x-3
        sm.tab[i] = make_uint2(s.x << sh, s.y << sh);
    }
    if (tid < 16) sm.C[tid] = terrain_rows[(size_t)terrain * 32 + tid] << 3;
#pragma unroll
    for (int r = 0; r < 16; r++) {
        sm.U.w[r][tid] = 0; sm.O.w[r][tid] = 0;
        sm.c0.w[r][tid] = 0; sm.c1.w[r][tid] = 0; sm.c2.w[r][tid] = 0; sm.c3.w[r][tid] = 0; sm.c4.w[r][tid] = 0;
    }
#pragma unroll
    for (int r = 0; r < 32; r++) sm.ring[r][tid] = (uint16_t)NONE;
    __syncthreads();

    const bool exists = chain < n_chains;
    ChainState& st = states[exists ? chain : 0];
    int done = exists ? st.done : 1;
    const bool run = !done;
    unsigned long long scored = 0;
    long long my_steps = 0;

    if (run) {
        const int epoch_bound = bounds[chains_per_terrain > 0 ? terrain : 0];
        const uint32_t base = chain_base(seed, chain_offset + (uint32_t)chain);
        const uint32_t nq7 = noise_q7(noise_pct);
        const int tenure = tenure_of(chain_offset + (uint32_t)chain);
        uint32_t* sl = site_lists + chain;
        int best = st.best, k = 0;
        uint32_t step = st.step, rowmask = 0;

        {   // site list in row-major order (spec), stamps "half a period ago"; cover planes from the list
            const uint32_t reset = (uint32_t)stamp_reset(step) << 16;
            for (int y = 0; y < 16; y++)
                for (uint32_t bits = st.S[y]; bits; bits &= bits - 1) sl[(size_t)(k++) * stride] = (uint32_t)(y * 32 + __ffs(bits) - 1) | reset;
            for (int r = 0; r < 16; r++) { const uint32_t c = sm.C[r]; sm.U.w[r][tid] = c; rowmask |= c ? 1u << r : 0u; }
            for (int i = 0; i < k; i++) flip<true>(sm, tid, (int)(sl[(size_t)i * stride] & 0x1ffu), rowmask);
        }

        for (long long it = 0; it < steps; it++) {
            if ((it & 1023) == 1023 && *interrupt) break;
            const int limit = min(epoch_bound, best);
            const uint32_t hs = step_hash(base, step);
            const int ten = effective_tenure(tenure, k);
            const bool drop = k >= limit;
            if (drop && k == 0) { done = 1; break; }
            sm.ring[step & 31u][tid] = (uint16_t)NONE;  // every consumed step owns its slot (stale entries are 32 steps old)
            if (!drop && rowmask == 0) {  // complete layout with k < limit supports
                best = k;
                list_to_rows(sl, stride, k, st.bestS);
                step++; my_steps++;
                if (k <= target || k == 0) { done = 1; break; }
                continue;
            }
            if (drop || (k == limit - 1 && k > 0)) {
                // ---- removal: min-loss support, random ties; supports younger than the tenure only as a last resort (not when dropping)
                uint32_t best_key = 0xffffffffu, best_e = 0;
                int best_i = 0;
#pragma unroll 2
                for (int i = 0; i < k; i++) {
                    const uint32_t e = sl[(size_t)i * stride];
                    const int v = (int)(e & 0x1ffu);
                    const int loss = score(sm.O, tid, v & 31, v >> 5, sm.tab[v]);
                    const uint32_t tie = tie_remove(lane_hash(hs, (uint32_t)(i & 31)), (uint32_t)(i >> 5));
                    const uint32_t young = (!drop && is_tabu(step, (uint16_t)(e >> 16), ten)) ? TABU_BIT : 0u;
                    const uint32_t key = young | ((uint32_t)loss << 16) | tie;
                    if (key < best_key) { best_key = key; best_i = i; best_e = e; }
                }
                const int u = (int)(best_e & 0x1ffu);
                const uint32_t last = sl[(size_t)(k - 1) * stride];
                sl[(size_t)best_i * stride] = last;
                sm.ring[step & 31u][tid] = (uint16_t)u;
                flip<false>(sm, tid, u, rowmask);
                scored += (unsigned)k;
                k--;
            }
            if (!drop) {
                // ---- addition at a random uncovered tile t: best-gain site of R(t)
                const int y = pick_rotated(rowmask, hs & 31u);
                const int x = pick_rotated(sm.U.w[y][tid] >> 3, (hs >> 5) & 31u);
                const uint2 wt = sm.tab[y * 32 + x];
                uint32_t R[13];  // rows y-6 .. y+6 of U; bit b of R[j] = tile column b + x - 6
#pragma unroll
                for (int j = 0; j < 13; j++) R[j] = (sm.U.w[(y - 6 + j) & 15][tid] << 3) >> x;
                unsigned long long tw = 0;  // sites removed fewer than `ten` steps ago, as a window around t
                for (int j = 0; j < ten; j++) {
                    const uint32_t e = sm.ring[(step - (uint32_t)j) & 31u][tid];
                    const int dx3 = (int)(e & 31u) - x + 3, dy3 = (int)(e >> 5) - y + 3;
                    if ((unsigned)dx3 < 7u && (unsigned)dy3 < 7u) tw |= 1ull << (7 * dy3 + dx3);
                }
                const uint32_t tw_lo = (uint32_t)tw & 0x0fffffffu, tw_hi = (uint32_t)(tw >> 28);
                const bool noise = ((hs >> 10) & 127u) < nq7;
                uint32_t mx = 0;
                int v = 0, nc = 0;
#pragma unroll
                for (int i = 0; i < 25; i++) {
                    const int dx = cell_dx(i), dy = cell_dy(i), bit = 7 * (dy + 3) + dx + 3;
                    const bool valid = ((bit < 28 ? wt.x >> bit : wt.y >> (bit - 28)) & 1u) != 0;
                    const int cv = (y + dy) * 32 + x + dx;
                    const uint2 win = sm.tab[valid ? cv : 0];
                    uint32_t lo = 0, hi = 0;
#pragma unroll
                    for (int j = 0; j < 4; j++) lo |= shift_static(R[dy + 3 + j], 7 * j - (dx + 3)) & (0x7fu << (7 * j));
#pragma unroll
                    for (int j = 0; j < 3; j++) hi |= shift_static(R[dy + 7 + j], 7 * j - (dx + 3)) & (0x7fu << (7 * j));
                    const int g = __popc(lo & win.x) + __popc(hi & win.y);
                    const bool tabu = ((bit < 28 ? tw_lo >> bit : tw_hi >> (bit - 28)) & 1u) != 0;
                    const uint32_t tie = tie_add(lane_hash(hs, (uint32_t)i));
                    uint32_t key = noise ? (0x10000u | tie) : ((tabu ? 0u : TABU_BIT) | ((uint32_t)(g + 1) << 16) | tie);
                    key = valid ? key : 0u;
                    nc += valid ? 1 : 0;
                    if (key > mx) { mx = key; v = cv; }
                }
                flip<true>(sm, tid, v, rowmask);
                sl[(size_t)k * stride] = (uint32_t)v | (step << 16);
                if (!noise) scored += (unsigned)nc;
                k++;
            }
            step++; my_steps++;
        }

        list_to_rows(sl, stride, k, st.S);
        st.k = k; st.best = best; st.step = step; st.done = done;
        const unsigned long long tot = ((unsigned long long)st.scored_hi << 32 | st.scored_lo) + scored;
        st.scored_lo = (uint32_t)tot; st.scored_hi = (uint32_t)(tot >> 32);
        st.steps_done += (uint32_t)my_steps;
    }
    // one pair of atomics per warp
    unsigned long long a = scored, b = (unsigned long long)my_steps;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
    if ((tid & 31) == 0 && (a | b)) { atomicAdd(&totals[0], a); atomicAdd(&totals[1], b); }
}

}  // namespace slst

bool sls_t16_fits(int w, int h) { return h <= 16 && w <= slst::MAXW; }
int sls_t16_cta_chains() { return slst::NT; }
size_t sls_t16_list_words(int n_chains) { return (size_t)16 * slst::MAXW * (((size_t)n_chains + 31) & ~(size_t)31); }

int sls_run_t16(tss_engine* e, const uint32_t* rows_dev, const uint2* tabs_dev, sls::ChainState* states, uint32_t* site_lists, int n_chains,
                int chains_per_terrain, uint32_t chain_offset, uint64_t seed, long long steps, const int* bounds_dev, int target,
                int noise_pct, unsigned long long* totals_dev) {
    static bool attr_set[64] = {false};
    if (e->device < 64 && !attr_set[e->device]) {
        TSS_CUDA(e, cudaFuncSetAttribute(slst::sls_t16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(slst::Smem)));
        attr_set[e->device] = true;
    }
    const size_t stride = ((size_t)n_chains + 31) & ~(size_t)31;
    const int blocks = (n_chains + slst::NT - 1) / slst::NT;
    slst::sls_t16_kernel<<<blocks, slst::NT, sizeof(slst::Smem), e->stream>>>(rows_dev, tabs_dev, states, site_lists, stride, n_chains,
                                                                             chains_per_terrain, chain_offset, seed, steps, bounds_dev, target,
                                                                             noise_pct, e->interrupt_dev, totals_dev);
    TSS_CHECK_LAUNCH(e);
    e->stats.kernel_launches++;
    return TSS_OK;
}

}  // namespace tss
