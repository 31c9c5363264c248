// Greedy placement cover for grids larger than 32x32 with platform sets beyond {1x1}.
//
// The placement SEARCH (sls_multi.cu) handles grids up to 32x32; larger grids are searched with 1x1 supports (lns.cu)
// and the supports are merged into larger platforms afterwards (engine.cu merge_supports).  This file adds a second
// constructor for those grids, a parallel greedy set cover over placements: every round scores, for every anchor, the
// placement (largest platforms first) that newly supports the most ceiling tiles — exactly validate()'s rule: the ceiling
// under the footprint plus three ceiling-masked 4-neighbour dilations (src/encoder/platform_layout.rs:104-141), evaluated
// in the placement's own (w+6) x (h+6) window — keeps the placements that are the best of their neighbourhood (so their
// windows are pairwise disjoint and the gains stay exact), and applies them.  Footprints stay inside the grid and
// pairwise disjoint (src/encoder.rs:546-609).  The engine returns whichever of the two layouts has fewer platforms,
// after prune_redundant and the usual re-validation by kernel (a).
#include <algorithm>

#include "engine.hpp"

namespace tss {
namespace greedy {

constexpr int WIN = 12;   // window rows / columns: footprint (<= 6) + 3 on each side
constexpr int MAX_KEYS = 16;

// 12 columns starting at column x0 (may be negative) of grid row y (zeros outside the grid)
__device__ __forceinline__ uint32_t window_bits(const uint32_t* __restrict__ rows, int wpr, int h, int y, int x0) {
    if (y < 0 || y >= h) return 0u;
    uint32_t out = 0;
    const int w0 = x0 >> 5;   // floor(x0 / 32), also for negative x0
#pragma unroll
    for (int k = 0; k < 2; k++) {
        const int wi = w0 + k;
        if (wi < 0 || wi >= wpr) continue;
        const uint32_t word = rows[(size_t)y * wpr + wi];
        const int shift = wi * 32 - x0;   // window position of the word's bit 0
        out |= shift >= 0 ? (shift < 32 ? word << shift : 0u) : word >> (-shift);
    }
    return out & 0xfffu;
}

// reach of placement (x, y, d) inside its window (column 0 = grid column x-3, row 0 = grid row y-3); false if the
// footprint overlaps an occupied tile
__device__ __forceinline__ bool placement_window(const uint32_t* __restrict__ C, const uint32_t* __restrict__ Occ, int wpr, int h, int x, int y, int2 d,
                                                 uint32_t (&X)[WIN]) {
    const uint32_t fmask = ((1u << d.x) - 1u) << 3;
    uint32_t Cw[WIN], hit = 0;
#pragma unroll
    for (int j = 0; j < WIN; j++) {
        const bool used = j < d.y + 6;
        Cw[j] = used ? window_bits(C, wpr, h, y - 3 + j, x - 3) : 0u;
        const bool frow = j >= 3 && j < 3 + d.y;
        X[j] = frow ? (Cw[j] & fmask) : 0u;
        if (frow) hit |= window_bits(Occ, wpr, h, y - 3 + j, x - 3) & fmask;
    }
    if (hit) return false;
#pragma unroll
    for (int round = 0; round < kTerrainSupportDistance - 1; round++) {
        uint32_t N[WIN];
#pragma unroll
        for (int j = 0; j < WIN; j++) {
            uint32_t v = X[j] | (X[j] << 1) | (X[j] >> 1);
            if (j > 0) v |= X[j - 1];
            if (j < WIN - 1) v |= X[j + 1];
            N[j] = v & Cw[j];
        }
#pragma unroll
        for (int j = 0; j < WIN; j++) X[j] = N[j];
    }
    return true;
}

// per anchor: the best placement (most newly supported tiles; larger platforms win ties: keys are sorted by area)
__global__ void gain_kernel(const uint32_t* __restrict__ C, const uint32_t* __restrict__ U, const uint32_t* __restrict__ Occ, int w, int h, int wpr,
                            const int2* __restrict__ keys, int n_keys, uint32_t* __restrict__ best) {
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= w * h) return;
    const int x = a % w, y = a / w;
    uint32_t best_gain = 0, best_key = 0;
    for (int k = 0; k < n_keys; k++) {
        const int2 d = keys[k];
        if (x + d.x > w || y + d.y > h) continue;
        uint32_t X[WIN];
        if (!placement_window(C, Occ, wpr, h, x, y, d, X)) continue;
        uint32_t g = 0;
#pragma unroll
        for (int j = 0; j < WIN; j++) g += __popc(window_bits(U, wpr, h, y - 3 + j, x - 3) & X[j]);
        if (g > best_gain) { best_gain = g; best_key = (uint32_t)k; }
    }
    best[a] = (best_gain << 8) | best_key;   // gain <= 144
}

// a placement is applied this round iff no anchor within reach-interaction distance holds a better one (gain, then lower index)
__global__ void select_kernel(const uint32_t* __restrict__ best, int w, int h, int radius, uint8_t* __restrict__ chosen) {
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= w * h) return;
    const uint32_t mine = best[a] >> 8;
    bool ok = mine > 0;
    const int x = a % w, y = a / w;
    for (int dy = -radius; dy <= radius && ok; dy++) {
        const int yy = y + dy;
        if (yy < 0 || yy >= h) continue;
        for (int dx = -radius; dx <= radius; dx++) {
            const int xx = x + dx;
            if (xx < 0 || xx >= w || (dx == 0 && dy == 0)) continue;
            const int b = yy * w + xx;
            const uint32_t other = best[b] >> 8;
            if (other > mine || (other == mine && b < a)) { ok = false; break; }
        }
    }
    chosen[a] = ok ? 1 : 0;
}

__device__ __forceinline__ void scatter_bits(uint32_t* rows, int wpr, int h, int y, int x0, uint32_t bits12, bool set) {
    if (y < 0 || y >= h || !bits12) return;
    const int w0 = x0 >> 5;
#pragma unroll
    for (int k = 0; k < 2; k++) {
        const int wi = w0 + k;
        if (wi < 0 || wi >= wpr) continue;
        const int shift = wi * 32 - x0;
        const uint32_t m = shift >= 0 ? (shift < 32 ? bits12 >> shift : 0u) : bits12 << (-shift);
        if (!m) continue;
        if (set) atomicOr(&rows[(size_t)y * wpr + wi], m); else atomicAnd(&rows[(size_t)y * wpr + wi], ~m);
    }
}

__global__ void apply_kernel(const uint32_t* __restrict__ C, uint32_t* __restrict__ U, uint32_t* __restrict__ Occ, int w, int h, int wpr,
                             const int2* __restrict__ keys, const uint32_t* __restrict__ best, const uint8_t* __restrict__ chosen,
                             int4* __restrict__ out, int* __restrict__ n_out, int cap) {
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= w * h || !chosen[a]) return;
    const int x = a % w, y = a / w, k = (int)(best[a] & 0xffu);
    const int2 d = keys[k];
    uint32_t X[WIN];
    // (chosen placements have pairwise disjoint windows, so Occ / U read here are not touched by the others)
    if (!placement_window(C, Occ, wpr, h, x, y, d, X)) return;
    const uint32_t fmask = ((1u << d.x) - 1u) << 3;
#pragma unroll
    for (int j = 0; j < WIN; j++) {
        scatter_bits(U, wpr, h, y - 3 + j, x - 3, X[j], false);
        if (j >= 3 && j < 3 + d.y) scatter_bits(Occ, wpr, h, y - 3 + j, x - 3, fmask, true);
    }
    const int slot = atomicAdd(n_out, 1);
    if (slot < cap) out[slot] = make_int4(x, y, k, 0);
}

__global__ void count_bits_kernel(const uint32_t* __restrict__ rows, int n_words, int* __restrict__ out) {
    int c = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_words; i += gridDim.x * blockDim.x) c += __popc(rows[i]);
    c = __reduce_add_sync(0xffffffffu, c);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, c);
}

}  // namespace greedy

// C_rows_host: h * wpr packed ceiling rows.  keys: effective (w, h) per dims key, any order (sorted by area here);
// out: (x, y, key index into `keys`) per placement.  Synchronises the engine stream.
int greedy_cover(tss_engine* e, const uint32_t* C_rows_host, int w, int h, const std::vector<int2>& keys, std::vector<int4>& out) {
    out.clear();
    const int wpr = (w + 31) / 32, nw = h * wpr, tiles = w * h;
    if (keys.empty() || (int)keys.size() > greedy::MAX_KEYS) return e->fail(TSS_E_UNSUPPORTED, "greedy_cover: %zu dims keys", keys.size());
    std::vector<int> order(keys.size());
    for (size_t i = 0; i < order.size(); i++) order[i] = (int)i;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return keys[a].x * keys[a].y > keys[b].x * keys[b].y; });
    std::vector<int2> sorted(keys.size());
    int maxdim = 1;
    for (size_t i = 0; i < order.size(); i++) {
        sorted[i] = keys[order[i]];
        if (sorted[i].x > 6 || sorted[i].y > 6) return e->fail(TSS_E_UNSUPPORTED, "greedy_cover: platforms larger than 6x6");
        maxdim = std::max(maxdim, std::max(sorted[i].x, sorted[i].y));
    }
    uint32_t *C = nullptr, *U = nullptr, *Occ = nullptr, *best = nullptr;
    uint8_t* chosen = nullptr;
    int2* keys_dev = nullptr;
    int4* out_dev = nullptr;
    int* counters = nullptr;   // [0] placements, [1] uncovered tiles
    cudaError_t err = cudaMalloc(&C, sizeof(uint32_t) * nw);
    auto A = [&](void** p, size_t bytes) { if (err == cudaSuccess) err = cudaMalloc(p, bytes); };
    A((void**)&U, sizeof(uint32_t) * nw); A((void**)&Occ, sizeof(uint32_t) * nw); A((void**)&best, sizeof(uint32_t) * tiles);
    A((void**)&chosen, tiles); A((void**)&keys_dev, sizeof(int2) * sorted.size()); A((void**)&out_dev, sizeof(int4) * tiles); A((void**)&counters, sizeof(int) * 2);
    int rc = TSS_OK;
    int host_counters[2] = {0, 0};
    if (err == cudaSuccess) err = cudaMemcpyAsync(C, C_rows_host, sizeof(uint32_t) * nw, cudaMemcpyHostToDevice, e->stream);
    if (err == cudaSuccess) err = cudaMemcpyAsync(U, C_rows_host, sizeof(uint32_t) * nw, cudaMemcpyHostToDevice, e->stream);
    if (err == cudaSuccess) err = cudaMemsetAsync(Occ, 0, sizeof(uint32_t) * nw, e->stream);
    if (err == cudaSuccess) err = cudaMemcpyAsync(keys_dev, sorted.data(), sizeof(int2) * sorted.size(), cudaMemcpyHostToDevice, e->stream);
    if (err == cudaSuccess) err = cudaMemsetAsync(counters, 0, sizeof(int) * 2, e->stream);
    const int tb = 128, gb = (tiles + tb - 1) / tb;
    for (int round = 0; err == cudaSuccess && round < tiles; round++) {
        greedy::gain_kernel<<<gb, tb, 0, e->stream>>>(C, U, Occ, w, h, wpr, keys_dev, (int)sorted.size(), best);
        greedy::select_kernel<<<gb, tb, 0, e->stream>>>(best, w, h, maxdim + 5, chosen);
        greedy::apply_kernel<<<gb, tb, 0, e->stream>>>(C, U, Occ, w, h, wpr, keys_dev, best, chosen, out_dev, counters, tiles);
        err = cudaMemsetAsync(counters + 1, 0, sizeof(int), e->stream);
        greedy::count_bits_kernel<<<64, 256, 0, e->stream>>>(U, nw, counters + 1);
        e->stats.kernel_launches += 4;
        if (err == cudaSuccess) err = cudaMemcpyAsync(host_counters, counters, sizeof(int) * 2, cudaMemcpyDeviceToHost, e->stream);
        if (err == cudaSuccess) err = cudaStreamSynchronize(e->stream);
        if (err == cudaSuccess && host_counters[1] == 0) break;
    }
    if (err == cudaSuccess && host_counters[1] != 0) rc = e->fail(TSS_E_CUDA, "greedy_cover: %d ceiling tiles left unsupported", host_counters[1]);
    if (err == cudaSuccess && rc == TSS_OK) {
        out.resize((size_t)std::min(host_counters[0], tiles));
        if (!out.empty()) err = cudaMemcpy(out.data(), out_dev, sizeof(int4) * out.size(), cudaMemcpyDeviceToHost);
        for (auto& p : out) p.z = order[(size_t)p.z];   // back to the caller's key numbering
    }
    cudaFree(C); cudaFree(U); cudaFree(Occ); cudaFree(best); cudaFree(chosen); cudaFree(keys_dev); cudaFree(out_dev); cudaFree(counters);
    if (err != cudaSuccess) return e->fail(TSS_E_CUDA, "greedy_cover: %s", cudaGetErrorString(err));
    return rc;
}

}  // namespace tss
