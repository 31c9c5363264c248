// Kernel (a), small-grid path (grids up to 32x32): ONE LAYOUT PER THREAD.
//
// A layout in the compact row format is at most 32 u32 words (a 16x16 layout is 8 words = 32 B, two grid rows per
// word).  A thread loads its layout with 16-byte loads (a warp reads 32 consecutive layouts = one contiguous 1 KB run
// for 16x16), keeps it in registers and runs validate()'s three ceiling-masked dilations (platform_layout.rs:127-141)
// in place: horizontal neighbours are masked shifts, vertical neighbours are the adjacent register (32-bit rows) or a
// funnel shift across two registers (8/16-bit rows).  No shuffles, no shared memory, no reduction: the two counts are
// thread-local POPC sums, written as one coalesced int2 per thread.  The first r1 version (one word per lane + warp
// shuffles) spent twice the issue slots per layout (profiles/r1_eval_kernel.md).
#include "engine.hpp"

namespace tss {

template <int ROWBITS>
struct RowMasksT;
template <> struct RowMasksT<8> { static constexpr uint32_t L = 0xFEFEFEFEu, R = 0x7F7F7F7Fu; };
template <> struct RowMasksT<16> { static constexpr uint32_t L = 0xFFFEFFFEu, R = 0x7FFF7FFFu; };
template <> struct RowMasksT<32> { static constexpr uint32_t L = 0xFFFFFFFFu, R = 0xFFFFFFFFu; };

template <int LUT>
__device__ __forceinline__ uint32_t lop3(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, %4;" : "=r"(d) : "r"(a), "r"(b), "r"(c), "n"(LUT));
    return d;
}

template <int WORDS>
__device__ __forceinline__ void load_words(const uint32_t* __restrict__ p, int wpl, bool vec, uint32_t (&v)[WORDS]) {
    if (WORDS >= 4 && vec) {  // wpl == WORDS and 16-byte aligned bases: layouts are WORDS*4 bytes apart
        const uint4* q = reinterpret_cast<const uint4*>(p);
#pragma unroll
        for (int i = 0; i < WORDS / 4; i++) {
            uint4 t = __ldg(q + i);
            v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
        }
    } else {
#pragma unroll
        for (int i = 0; i < WORDS; i++) v[i] = i < wpl ? __ldg(p + i) : 0u;
    }
}

template <int ROWBITS, int WORDS, bool PER_TERRAIN>
__global__ void __launch_bounds__(128) eval_thread_kernel(const uint32_t* __restrict__ grid, const uint32_t* __restrict__ layouts, long long n,
                                                         int wpl, bool vec, int2* __restrict__ out) {
    uint32_t C[WORDS];
    if (!PER_TERRAIN) load_words<WORDS>(grid, wpl, vec, C);
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long L = (long long)blockIdx.x * blockDim.x + threadIdx.x; L < n; L += stride) {
        uint32_t X[WORDS];
        load_words<WORDS>(layouts + L * wpl, wpl, vec, X);
        if (PER_TERRAIN) load_words<WORDS>(grid + L * wpl, wpl, vec, C);
        int count = 0;
#pragma unroll
        for (int i = 0; i < WORDS; i++) { count += __popc(X[i]); X[i] &= C[i]; }  // only terrain can be supported
#pragma unroll
        for (int round = 0; round < kTerrainSupportDistance - 1; round++) {
            uint32_t prev_old = 0;
#pragma unroll
            for (int i = 0; i < WORDS; i++) {
                const uint32_t x = X[i], next = i + 1 < WORDS ? X[i + 1] : 0u;
                const uint32_t up = ROWBITS == 32 ? prev_old : __funnelshift_l(prev_old, x, ROWBITS);
                const uint32_t down = ROWBITS == 32 ? next : __funnelshift_r(x, next, ROWBITS);
                // four 3-input logic ops per word: the row-boundary masks ride along as the third operand
                uint32_t t = lop3<0xF8>(x, x << 1, RowMasksT<ROWBITS>::L);   // x | ((x << 1) & L)
                t = lop3<0xF8>(t, x >> 1, RowMasksT<ROWBITS>::R);            // t | ((x >> 1) & R)
                t = lop3<0xFE>(t, up, down);                                 // t | up | down
                X[i] = t & C[i];
                prev_old = x;
            }
        }
        int unc = 0;
#pragma unroll
        for (int i = 0; i < WORDS; i++) unc += __popc(C[i] & ~X[i]);
        out[L] = make_int2(unc, count);
    }
}

template <int ROWBITS, int WORDS>
static int launch_thread(tss_engine* e, const void* grid, const void* layouts, int64_t n, int wpl, bool per_terrain, int2* out) {
    const int threads = 128;
    long long blocks = (n + threads - 1) / threads;
    const long long max_blocks = (long long)e->prop.multiProcessorCount * 16;  // whole waves; grid-stride beyond that
    if (blocks > max_blocks) blocks = max_blocks;
    const bool vec = wpl == WORDS && WORDS >= 4 && ((uintptr_t)grid & 15) == 0 && ((uintptr_t)layouts & 15) == 0;
    if (per_terrain)
        eval_thread_kernel<ROWBITS, WORDS, true><<<(unsigned)blocks, threads, 0, e->stream>>>((const uint32_t*)grid, (const uint32_t*)layouts, n, wpl, vec, out);
    else
        eval_thread_kernel<ROWBITS, WORDS, false><<<(unsigned)blocks, threads, 0, e->stream>>>((const uint32_t*)grid, (const uint32_t*)layouts, n, wpl, vec, out);
    TSS_CHECK_LAUNCH(e);
    e->stats.kernel_launches++;
    return TSS_OK;
}

template <int ROWBITS>
static int dispatch_words(tss_engine* e, const void* grid, const void* layouts, int64_t n, int wpl, bool pt, int2* out) {
    if (wpl <= 1) return launch_thread<ROWBITS, 1>(e, grid, layouts, n, wpl, pt, out);
    if (wpl <= 2) return launch_thread<ROWBITS, 2>(e, grid, layouts, n, wpl, pt, out);
    if (wpl <= 4) return launch_thread<ROWBITS, 4>(e, grid, layouts, n, wpl, pt, out);
    if (wpl <= 8) return launch_thread<ROWBITS, 8>(e, grid, layouts, n, wpl, pt, out);
    if (wpl <= 16) return launch_thread<ROWBITS, 16>(e, grid, layouts, n, wpl, pt, out);
    return launch_thread<ROWBITS, 32>(e, grid, layouts, n, wpl, pt, out);
}

// grids up to 32x32 in the compact row format (tss.h): wpl = layout words
int launch_eval_small(tss_engine* e, const void* grid_dev, int w, int h, const void* layouts_dev, int64_t n, bool per_layout_terrain,
                      int32_t* out_dev) {
    const int wpl = (int)(tss_layout_bytes(w, h) / 4);
    switch (tss_row_bits(w)) {
        case 8: return dispatch_words<8>(e, grid_dev, layouts_dev, n, wpl, per_layout_terrain, (int2*)out_dev);
        case 16: return dispatch_words<16>(e, grid_dev, layouts_dev, n, wpl, per_layout_terrain, (int2*)out_dev);
        default: return dispatch_words<32>(e, grid_dev, layouts_dev, n, wpl, per_layout_terrain, (int2*)out_dev);
    }
}

}  // namespace tss
