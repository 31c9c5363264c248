// GPU LOWER bound for the feasibility-and-bound loop (SURVEY.md §8(f) "idea"; not in the reference): a PACKING of ceiling
// tiles no two of which any single platform of the set can support.  Every complete layout needs a distinct platform per
// packed tile, so |packing| <= optimum; when the SLS upper bound meets it the loop (crates/repl/src/main.rs:280-366) is
// finished without a single exact-solver call — with the REPL's default-8 set on test/ex1.toml / ex3.toml the optimum is 1
// and any ceiling tile is such a packing.
//
// Which tiles CAN one platform support together?  validate() (src/encoder/platform_layout.rs:104-141) supports the ceiling
// under a footprint plus three ceiling-masked 4-neighbour dilations, so placement p supports tile a iff foot(p) meets
// B3(a), the geodesic ball of radius 3 around a.  The co-coverable set of a is therefore
//     N(a) = dilate^3( C  &  U_{keys (w,h)}  box_{w,h}( inb_{w,h} & boxT_{w,h}( B3(a) ) ) )
// where boxT maps tiles to the anchors whose w x h footprint contains them, inb keeps anchors whose footprint is inside the
// grid (out-of-bounds placements are forbidden, src/encoder.rs:601-609) and box maps anchors back to footprint tiles.  With
// 1x1 supports only this is the geodesic ball of radius 6.  All of it is shifts, ORs and row shuffles on bitboards.
//
// Kernel: one warp per randomized restart, lane r = grid row r (grids up to 32x32).  Greedy: sample a few available tiles,
// take the one that blocks the fewest available tiles, remove its co-coverable set from the available tiles, repeat.  The
// best packing of all restarts wins (64-bit atomicMax); a second one-warp kernel re-derives N(a) for every tile of the
// winning packing and checks that it meets the packing in a alone, so a reported bound never rests on the greedy's
// bookkeeping.
#include "engine.hpp"
#include "sls_spec.hpp"

namespace tss {
namespace lb {

constexpr uint32_t FULL = 0xffffffffu;
constexpr int MAX_KEYS = 16;
constexpr int SAMPLES = 3;

struct Keys {
    int n;
    int w[MAX_KEYS], h[MAX_KEYS];
};

__device__ __forceinline__ uint32_t dilate(uint32_t X, uint32_t C, int lane) {
    uint32_t up = __shfl_up_sync(FULL, X, 1), down = __shfl_down_sync(FULL, X, 1);
    if (lane == 0) up = 0;
    if (lane == 31) down = 0;
    return (X | (X << 1) | (X >> 1) | up | down) & C;
}

// rows of N(a) for the tile a = (ax, ay); lane = row
__device__ __forceinline__ uint32_t cocoverable(uint32_t C, int lane, int ax, int ay, int W, int H, const Keys& keys) {
    uint32_t B = (lane == ay ? 1u << ax : 0u) & C;
#pragma unroll
    for (int r = 0; r < kTerrainSupportDistance - 1; r++) B = dilate(B, C, lane);
    uint32_t F = 0;
    for (int k = 0; k < keys.n; k++) {
        const int w = keys.w[k], h = keys.h[k];
        if (w > W || h > H) continue;                          // never fits the grid
        uint32_t Ah = B;                                        // anchors (x - dx, y) for dx < w
        for (int dx = 1; dx < w; dx++) Ah |= B >> dx;
        uint32_t A = Ah;                                        // ... and (., y - dy) for dy < h: anchor row r takes tile row r + dy
        for (int dy = 1; dy < h; dy++) { uint32_t t = __shfl_down_sync(FULL, Ah, dy); A |= lane + dy < 32 ? t : 0u; }
        const int nx = W - w + 1;                               // anchors with the footprint inside the grid
        A &= nx >= 32 ? FULL : ((1u << nx) - 1u);
        if (lane > H - h) A = 0;
        uint32_t Fh = A;                                        // footprint tiles of those anchors
        for (int dx = 1; dx < w; dx++) Fh |= A << dx;
        uint32_t Fk = Fh;
        for (int dy = 1; dy < h; dy++) { uint32_t t = __shfl_up_sync(FULL, Fh, dy); Fk |= lane >= dy ? t : 0u; }
        F |= Fk;
    }
    F &= C;                                                     // the ceiling under a footprint is directly supported
#pragma unroll
    for (int r = 0; r < kTerrainSupportDistance - 1; r++) F = dilate(F, C, lane);
    return F;
}

__device__ __forceinline__ int pick_rotated(uint32_t bits, uint32_t o) {
    uint32_t rot = __funnelshift_r(bits, bits, o);
    return (int)((__ffs(rot) - 1 + o) & 31u);
}

// out_rows[restart][32]: the packing of every restart; best = max over restarts of (size << 32 | ~restart)
__global__ void __launch_bounds__(128) lb_pack_kernel(const uint32_t* __restrict__ rows, int W, int H, Keys keys, uint64_t seed, int n_restarts,
                                                      uint32_t* __restrict__ out_rows, unsigned long long* __restrict__ best) {
    const int lane = threadIdx.x & 31, restart = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (restart >= n_restarts) return;
    const uint32_t C = rows[lane];
    const uint32_t base = sls::chain_base(seed, (uint32_t)restart);
    uint32_t avail = C, P = 0;
    int count = 0;
    for (uint32_t step = 0; step < 1024u; step++) {
        const uint32_t rowmask = __ballot_sync(FULL, avail != 0);
        if (!rowmask) break;
        uint32_t bestN = 0;
        int best_deg = 1 << 30, bx = 0, by = 0;
        // the first restart samples one tile per pick (plain random greedy), later ones prefer low-degree tiles
        const int samples = restart == 0 ? 1 : SAMPLES;
        for (int j = 0; j < samples; j++) {
            const uint32_t hs = sls::step_hash(base, step * 8u + (uint32_t)j);
            const int y = pick_rotated(rowmask, hs & 31u);
            const uint32_t arow = __shfl_sync(FULL, avail, y);
            const int x = pick_rotated(arow, (hs >> 5) & 31u);
            const uint32_t N = cocoverable(C, lane, x, y, W, H, keys);
            const int deg = __reduce_add_sync(FULL, __popc(N & avail));
            if (deg < best_deg) { best_deg = deg; bestN = N; bx = x; by = y; }
        }
        if (lane == by) P |= 1u << bx;
        avail &= ~bestN;
        if (lane == by) avail &= ~(1u << bx);   // (a is in N(a); kept explicit so the loop always makes progress)
        count++;
    }
    out_rows[(size_t)restart * 32 + lane] = P;
    if (lane == 0) atomicMax(best, ((unsigned long long)(uint32_t)count << 32) | (uint32_t)(~(uint32_t)restart));
}

// result[0] = tiles of the packing, result[1] = number of packed tiles whose co-coverable set meets another packed tile
// (must be 0), result[2] = packed tiles that are not ceiling (must be 0)
__global__ void lb_verify_kernel(const uint32_t* __restrict__ rows, int W, int H, Keys keys, const uint32_t* __restrict__ out_rows,
                                 const unsigned long long* __restrict__ best, uint32_t* __restrict__ winner_rows, int* __restrict__ result) {
    const int lane = threadIdx.x;
    const uint32_t restart = ~(uint32_t)(*best & 0xffffffffu);
    const uint32_t C = rows[lane], P = out_rows[(size_t)restart * 32 + lane];
    winner_rows[lane] = P;
    int bad = 0;
    for (int y = 0; y < 32; y++) {
        uint32_t prow = __shfl_sync(FULL, P, y);
        while (prow) {
            const int x = __ffs(prow) - 1;
            prow &= prow - 1;
            const uint32_t N = cocoverable(C, lane, x, y, W, H, keys);
            const uint32_t others = N & P & ~(lane == y ? 1u << x : 0u);
            bad += __any_sync(FULL, others != 0) ? 1 : 0;
        }
    }
    const int n = __reduce_add_sync(FULL, __popc(P)), off = __reduce_add_sync(FULL, __popc(P & ~C));
    if (lane == 0) { result[0] = n; result[1] = bad; result[2] = off; }
}

// ------------------------------------------------------------------------------------------------------------------------
// Grids larger than 32x32 (e.g. the 256x256 portfolio), 1x1 supports: the same packing bound — tiles pairwise more than 6
// steps apart through the ceiling — built by Luby-style parallel rounds on the whole bitboard.  One CTA per restart, all
// state in shared memory.  A round: (1) every available tile counts the available tiles in its radius-6 diamond (its
// degree); (2) a tile is selected if its key (few neighbours first, then a hash) beats every available tile in that
// diamond — selected tiles are then more than 6 apart even in Manhattan distance, hence geodesically; (3) the geodesic
// radius-6 balls of the selected tiles (six ceiling-masked dilations of the whole board) leave the available set.
constexpr int BIG_THREADS = 1024;
constexpr int BIG_MAX_WORDS = 2048;      // 256 x 256

struct BigBoards {
    uint32_t *C, *avail, *P, *A, *B;     // [nw] each
    uint8_t* deg;                        // [nw * 32]
};

// `len` (<= 32) bits of row y of board X starting at column x0 (zeros outside the grid)
__device__ __forceinline__ uint32_t row_bits(const uint32_t* X, int wpr, int h, int y, int x0, int len) {
    if (y < 0 || y >= h) return 0u;
    const int wq = x0 >> 5, sh = x0 & 31;          // arithmetic shift: floor, also for negative x0
    const uint32_t lo = (wq >= 0 && wq < wpr) ? X[y * wpr + wq] : 0u;
    const uint32_t hi = (wq + 1 >= 0 && wq + 1 < wpr) ? X[y * wpr + wq + 1] : 0u;
    const uint32_t v = __funnelshift_r(lo, hi, sh);
    return len >= 32 ? v : (v & ((1u << len) - 1u));
}

__device__ __forceinline__ void dilate_board(const uint32_t* src, uint32_t* dst, const uint32_t* C, int nw, int wpr) {
    for (int i = threadIdx.x; i < nw; i += blockDim.x) {
        const int col = i % wpr;
        const uint32_t x = src[i];
        uint32_t l = x << 1, r = x >> 1;
        if (col > 0) l |= src[i - 1] >> 31;
        if (col + 1 < wpr) r |= src[i + 1] << 31;
        const uint32_t up = i >= wpr ? src[i - wpr] : 0u, down = i + wpr < nw ? src[i + wpr] : 0u;
        dst[i] = (x | l | r | up | down) & C[i];
    }
    __syncthreads();
}

__device__ __forceinline__ uint32_t big_key(uint32_t deg, uint32_t seed, uint32_t round, uint32_t tile) {
    return ((255u - (deg > 255u ? 255u : deg)) << 24) | (sls::fmix32(seed ^ (round * sls::K1) ^ (tile * sls::K2)) & 0xffffffu);
}

// out_rows[restart][nw]: packing of every restart; best = max over restarts of (size << 32 | ~restart)
__global__ void __launch_bounds__(BIG_THREADS) lb_pack_big_kernel(const uint32_t* __restrict__ rows, int W, int H, uint64_t seed, int n_restarts,
                                                                  uint32_t* __restrict__ out_rows, unsigned long long* __restrict__ best) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int wpr = (W + 31) / 32, nw = H * wpr, tid = threadIdx.x, restart = blockIdx.x;
    uint32_t* C = reinterpret_cast<uint32_t*>(smem_raw);
    uint32_t *avail = C + nw, *P = avail + nw, *A = P + nw, *B = A + nw;
    uint8_t* deg = reinterpret_cast<uint8_t*>(B + nw);
    __shared__ int any_avail, count;
    const uint32_t base = sls::chain_base(seed, (uint32_t)restart);
    for (int i = tid; i < nw; i += blockDim.x) { const uint32_t c = rows[i]; C[i] = c; avail[i] = c; P[i] = 0u; }
    if (tid == 0) count = 0;
    __syncthreads();
    for (uint32_t round = 0; round < 64u; round++) {
        if (tid == 0) any_avail = 0;
        __syncthreads();
        // (1) degrees
        for (int i = tid; i < nw; i += blockDim.x) {
            const int y = i / wpr, xb = (i % wpr) * 32;
            for (uint32_t bits = avail[i]; bits; bits &= bits - 1) {
                const int x = xb + __ffs(bits) - 1;
                int d = 0;
#pragma unroll
                for (int dy = -6; dy <= 6; dy++) { const int r = 6 - (dy < 0 ? -dy : dy); d += __popc(row_bits(avail, wpr, H, y + dy, x - r, 2 * r + 1)); }
                deg[i * 32 + (x - xb)] = (uint8_t)(d > 255 ? 255 : d);
                any_avail = 1;
            }
        }
        __syncthreads();
        if (!any_avail) break;
        // (2) selection: strict maximum of (key, lower tile index) over the available tiles of the diamond
        for (int i = tid; i < nw; i += blockDim.x) {
            const int y = i / wpr, xb = (i % wpr) * 32;
            uint32_t sel = 0;
            for (uint32_t bits = avail[i]; bits; bits &= bits - 1) {
                const int b = __ffs(bits) - 1, x = xb + b, tile = y * W + x;
                const uint32_t key = big_key(deg[i * 32 + b], base, round, (uint32_t)tile);
                bool win = true;
                for (int dy = -6; dy <= 6 && win; dy++) {
                    const int r = 6 - (dy < 0 ? -dy : dy), yy = y + dy;
                    for (uint32_t nb = row_bits(avail, wpr, H, yy, x - r, 2 * r + 1); nb && win; nb &= nb - 1) {
                        const int xx = x - r + __ffs(nb) - 1;
                        if (xx == x && dy == 0) continue;
                        const int ot = yy * W + xx;
                        const uint32_t ok = big_key(deg[(yy * wpr + (xx >> 5)) * 32 + (xx & 31)], base, round, (uint32_t)ot);
                        win = key > ok || (key == ok && tile < ot);
                    }
                }
                if (win) sel |= 1u << b;
            }
            A[i] = sel;
            if (sel) atomicAdd(&count, __popc(sel));
        }
        __syncthreads();
        // (3) the selected tiles join the packing; their geodesic radius-6 balls leave the available set
        for (int i = tid; i < nw; i += blockDim.x) P[i] |= A[i];
        __syncthreads();
        for (int r = 0; r < 3; r++) { dilate_board(A, B, C, nw, wpr); dilate_board(B, A, C, nw, wpr); }
        for (int i = tid; i < nw; i += blockDim.x) avail[i] &= ~A[i];
        __syncthreads();
    }
    for (int i = tid; i < nw; i += blockDim.x) out_rows[(size_t)restart * nw + i] = P[i];
    if (tid == 0) atomicMax(best, ((unsigned long long)(uint32_t)count << 32) | (uint32_t)(~(uint32_t)restart));
}

// Exact check of the winning packing: the radius-3 geodesic balls of the packed tiles are pairwise disjoint (a site in two of
// them would support both tiles).  Labels (packed tile index + 1) spread for three synchronous rounds through the ceiling; a tile
// that is offered two different labels is a violation.  One CTA; label / next arrays in global memory.
// result[0] = packed tiles, result[1] = violations, result[2] = packed tiles off the ceiling.
__global__ void __launch_bounds__(BIG_THREADS) lb_verify_big_kernel(const uint32_t* __restrict__ rows, int W, int H, const uint32_t* __restrict__ out_rows,
                                                                    const unsigned long long* __restrict__ best, uint32_t* __restrict__ winner_rows,
                                                                    uint32_t* __restrict__ label, uint32_t* __restrict__ next, int* __restrict__ result) {
    const int wpr = (W + 31) / 32, nw = H * wpr, tid = threadIdx.x, tiles = W * H;
    const uint32_t restart = ~(uint32_t)(*best & 0xffffffffu);
    const uint32_t* P = out_rows + (size_t)restart * nw;
    __shared__ int n_packed, n_bad, n_off;
    if (tid == 0) { n_packed = 0; n_bad = 0; n_off = 0; }
    __syncthreads();
    for (int i = tid; i < nw; i += blockDim.x) {
        winner_rows[i] = P[i];
        if (P[i]) atomicAdd(&n_packed, __popc(P[i]));
        if (P[i] & ~rows[i]) atomicAdd(&n_off, __popc(P[i] & ~rows[i]));
    }
    for (int t = tid; t < tiles; t += blockDim.x) {
        const int x = t % W, y = t / W;
        label[t] = ((P[y * wpr + (x >> 5)] >> (x & 31)) & 1u) ? (uint32_t)t + 1u : 0u;
    }
    __syncthreads();
    for (int round = 0; round < kTerrainSupportDistance - 1; round++) {
        for (int t = tid; t < tiles; t += blockDim.x) {
            const int x = t % W, y = t / W;
            uint32_t mine = label[t];
            if ((rows[y * wpr + (x >> 5)] >> (x & 31)) & 1u) {
                const int nx[4] = {x + 1, x, x - 1, x}, ny[4] = {y, y + 1, y, y - 1};
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    if (nx[k] < 0 || nx[k] >= W || ny[k] < 0 || ny[k] >= H) continue;
                    const uint32_t other = label[ny[k] * W + nx[k]];
                    if (!other) continue;
                    if (mine && mine != other) atomicAdd(&n_bad, 1);
                    if (!mine) mine = other;
                }
            }
            next[t] = mine;
        }
        __syncthreads();
        for (int t = tid; t < tiles; t += blockDim.x) label[t] = next[t];
        __syncthreads();
    }
    if (tid == 0) { result[0] = n_packed; result[1] = n_bad; result[2] = n_off; }
}

}  // namespace lb

// rows32_host: terrain rows; key_dims: effective (w, h) per dims key.  out_rows32 (host, 32 words) = the packing.
int lb_run(tss_engine* e, const uint32_t* rows32_host, int W, int H, const std::vector<int2>& key_dims, uint64_t seed, int restarts,
           uint32_t* out_rows32, int* out_count) {
    if ((int)key_dims.size() > lb::MAX_KEYS) return e->fail(TSS_E_UNSUPPORTED, "tss_lower_bound: more than %d dims keys", lb::MAX_KEYS);
    lb::Keys keys;
    keys.n = (int)key_dims.size();
    for (int i = 0; i < keys.n; i++) { keys.w[i] = key_dims[(size_t)i].x; keys.h[i] = key_dims[(size_t)i].y; }
    if (restarts <= 0) restarts = e->prop.multiProcessorCount * 16;
    // scratch slot 7: terrain rows [32] | winner rows [32] | result [4] | best key [2] | packings [restarts][32]
    uint32_t* buf = (uint32_t*)e->dev(7, sizeof(uint32_t) * (72 + (size_t)restarts * 32));
    uint32_t* host = (uint32_t*)e->pin(2, sizeof(uint32_t) * 72);
    if (!buf || !host) return TSS_E_CUDA;
    uint32_t *rows_dev = buf, *winner = buf + 32, *packs = buf + 72;
    int* result = (int*)(buf + 64);
    unsigned long long* best = (unsigned long long*)(buf + 68);
    for (int i = 0; i < 32; i++) host[i] = rows32_host[i];
    TSS_CUDA(e, cudaMemcpyAsync(rows_dev, host, sizeof(uint32_t) * 32, cudaMemcpyHostToDevice, e->stream));
    TSS_CUDA(e, cudaMemsetAsync(buf + 64, 0, sizeof(uint32_t) * 8, e->stream));
    TSS_CUDA(e, cudaEventRecord(e->ev0, e->stream));
    lb::lb_pack_kernel<<<(restarts + 3) / 4, 128, 0, e->stream>>>(rows_dev, W, H, keys, seed, restarts, packs, best);
    TSS_CHECK_LAUNCH(e);
    lb::lb_verify_kernel<<<1, 32, 0, e->stream>>>(rows_dev, W, H, keys, packs, best, winner, result);
    TSS_CHECK_LAUNCH(e);
    TSS_CUDA(e, cudaEventRecord(e->ev1, e->stream));
    e->stats.kernel_launches += 2;
    TSS_CUDA(e, cudaMemcpyAsync(host + 32, winner, sizeof(uint32_t) * 36, cudaMemcpyDeviceToHost, e->stream));
    TSS_CUDA(e, cudaStreamSynchronize(e->stream));
    float ms = 0;
    if (cudaEventElapsedTime(&ms, e->ev0, e->ev1) != cudaSuccess) cudaGetLastError();
    e->stats.device_ms = ms;
    const int* res = (const int*)(host + 64);
    if (res[1] != 0 || res[2] != 0)
        return e->fail(TSS_E_CUDA, "internal error: lower-bound packing failed verification (%d tiles share a platform, %d off the ceiling)", res[1], res[2]);
    for (int i = 0; i < 32; i++) out_rows32[i] = host[32 + i];
    *out_count = res[0];
    return TSS_OK;
}

// grids larger than 32x32, 1x1 supports.  rows_host: bit-packed rows [H * wpr]; out_rows (host) the same shape.
int lb_run_big(tss_engine* e, const uint32_t* rows_host, int W, int H, uint64_t seed, int restarts, uint32_t* out_rows, int* out_count) {
    const int wpr = (W + 31) / 32, nw = H * wpr, tiles = W * H;
    if (nw > lb::BIG_MAX_WORDS) return e->fail(TSS_E_UNSUPPORTED, "tss_lower_bound: grid %dx%d exceeds the packing kernel's shared-memory boards", W, H);
    if (restarts <= 0) restarts = e->prop.multiProcessorCount;
    const size_t smem = sizeof(uint32_t) * 5 * (size_t)nw + (size_t)nw * 32;
    TSS_CUDA(e, cudaFuncSetAttribute(lb::lb_pack_big_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // scratch slot 7: rows [nw] | winner [nw] | result [4] | best [2] | label [tiles] | next [tiles] | packings [restarts][nw]
    const size_t words = (size_t)2 * nw + 8 + (size_t)2 * tiles + (size_t)restarts * nw;
    uint32_t* buf = (uint32_t*)e->dev(7, sizeof(uint32_t) * words);
    if (!buf) return TSS_E_CUDA;
    uint32_t *rows_dev = buf, *winner = buf + nw;
    int* result = (int*)(buf + 2 * nw);
    unsigned long long* best = (unsigned long long*)(buf + 2 * nw + 4);
    uint32_t *label = buf + 2 * nw + 8, *next = label + tiles, *packs = next + tiles;
    TSS_CUDA(e, cudaMemcpyAsync(rows_dev, rows_host, sizeof(uint32_t) * (size_t)nw, cudaMemcpyHostToDevice, e->stream));
    TSS_CUDA(e, cudaMemsetAsync(buf + 2 * nw, 0, sizeof(uint32_t) * 8, e->stream));
    TSS_CUDA(e, cudaEventRecord(e->ev0, e->stream));
    lb::lb_pack_big_kernel<<<restarts, lb::BIG_THREADS, smem, e->stream>>>(rows_dev, W, H, seed, restarts, packs, best);
    TSS_CHECK_LAUNCH(e);
    lb::lb_verify_big_kernel<<<1, lb::BIG_THREADS, 0, e->stream>>>(rows_dev, W, H, packs, best, winner, label, next, result);
    TSS_CHECK_LAUNCH(e);
    TSS_CUDA(e, cudaEventRecord(e->ev1, e->stream));
    e->stats.kernel_launches += 2;
    int res[4];
    TSS_CUDA(e, cudaMemcpyAsync(res, result, sizeof res, cudaMemcpyDeviceToHost, e->stream));
    TSS_CUDA(e, cudaMemcpyAsync(out_rows, winner, sizeof(uint32_t) * (size_t)nw, cudaMemcpyDeviceToHost, e->stream));
    TSS_CUDA(e, cudaStreamSynchronize(e->stream));
    float ms = 0;
    if (cudaEventElapsedTime(&ms, e->ev0, e->ev1) != cudaSuccess) cudaGetLastError();
    e->stats.device_ms = ms;
    if (res[1] != 0 || res[2] != 0)
        return e->fail(TSS_E_CUDA, "internal error: lower-bound packing failed verification (%d shared sites, %d off the ceiling)", res[1], res[2]);
    *out_count = res[0];
    return TSS_OK;
}

}  // namespace tss
