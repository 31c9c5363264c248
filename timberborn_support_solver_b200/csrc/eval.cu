// Kernel (a): bit-packed coverage evaluator == PlatformLayout::validate (src/encoder/platform_layout.rs:85-149).
//
//   direct   = footprints & ceiling                              (:104-124; "only terrain can be supported")
//   3 x      { supported |= N4(supported) & ceiling }            (:127-141; TERRAIN_SUPPORT_DISTANCE-1 rounds, src/lib.rs:12)
//   unsupported = ceiling & ~supported                           (:143-146)
//
// Two shapes:
//   eval_thread  (eval_thread.cu) grids up to 32x32: a layout is <= 32 u32 words in the COMPACT row format (row stride
//                1/2/4 bytes, so a 16x16 layout is 32 B); one layout per thread, all in registers, streaming layouts
//                from HBM with a grid-stride loop (grid = multiple of the SM count).
//   eval_tiled   any grid that fits shared memory (256x256 = 8 KB per plane): one CTA per layout, planes in
//                shared memory.  Also evaluates general platform layouts (footprint stamping, overlap and
//                out-of-bounds detection) and can export the four support layers (the terrain-layer variables
//                T3..T0 of the encoder, src/encoder.rs:520-544) for witness construction.
#include "engine.hpp"

namespace tss {

int launch_eval_small(tss_engine* e, const void* grid_dev, int w, int h, const void* layouts_dev, int64_t n, bool per_layout_terrain,
                      int32_t* out_dev);  // eval_thread.cu: grids up to 32x32, one layout per thread

// ------------------------------------------------------------------------------------------------ eval_tiled
// Shared-memory planes of nw = h*wpr words.  One masked dilation round: dst = (src | shifts) & C.
__device__ __forceinline__ void tiled_round(const uint32_t* __restrict__ src, uint32_t* __restrict__ dst,
                                            const uint32_t* __restrict__ C, int nw, int wpr) {
    for (int i = threadIdx.x; i < nw; i += blockDim.x) {
        int col = i % wpr;
        uint32_t x = src[i];
        uint32_t l = x << 1, r = x >> 1;
        if (col > 0) l |= src[i - 1] >> 31;
        if (col + 1 < wpr) r |= src[i + 1] << 31;
        uint32_t up = i >= wpr ? src[i - wpr] : 0u, down = i + wpr < nw ? src[i + wpr] : 0u;
        dst[i] = (x | l | r | up | down) & C[i];
    }
}

__device__ __forceinline__ int block_sum(int v, int* smem_acc) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(smem_acc, v);
    return v;
}

// sites layouts (wide grids): layouts are packed rows [n][nw]
template <bool PER_TERRAIN>
__global__ void eval_tiled_sites_kernel(const uint32_t* __restrict__ grid, const uint32_t* __restrict__ layouts, long long n,
                                        int h, int wpr, int2* __restrict__ out) {
    extern __shared__ uint32_t sm[];
    const int nw = h * wpr;
    uint32_t *C = sm, *A = sm + nw, *B = sm + 2 * nw;
    __shared__ int acc[2];
    if (!PER_TERRAIN)
        for (int i = threadIdx.x; i < nw; i += blockDim.x) C[i] = grid[i];
    for (long long L = blockIdx.x; L < n; L += gridDim.x) {
        if (threadIdx.x < 2) acc[threadIdx.x] = 0;
        int cnt = 0;
        for (int i = threadIdx.x; i < nw; i += blockDim.x) {
            if (PER_TERRAIN) C[i] = grid[L * nw + i];
            uint32_t s = layouts[L * nw + i];
            cnt += __popc(s);
            A[i] = s & C[i];
        }
        __syncthreads();
        tiled_round(A, B, C, nw, wpr); __syncthreads();
        tiled_round(B, A, C, nw, wpr); __syncthreads();
        tiled_round(A, B, C, nw, wpr); __syncthreads();
        int unc = 0;
        for (int i = threadIdx.x; i < nw; i += blockDim.x) unc += __popc(C[i] & ~B[i]);
        block_sum(unc, &acc[0]);
        block_sum(cnt, &acc[1]);
        __syncthreads();
        if (threadIdx.x == 0) out[L] = make_int2(acc[0], acc[1]);
        __syncthreads();
    }
}

// General platform layouts.  plats: (x, y, effective w, effective h).  out[L] = (unsupported, platforms, overlapping, oob).
__global__ void eval_platforms_kernel(const uint32_t* __restrict__ grid, int w, int h, int wpr, const int4* __restrict__ plats,
                                      const uint32_t* __restrict__ offsets, long long n, int4* __restrict__ out,
                                      uint32_t* __restrict__ unsupported_rows, uint8_t* __restrict__ flags,
                                      uint32_t* __restrict__ layers) {
    extern __shared__ uint32_t sm[];
    const int nw = h * wpr;
    uint32_t *C = sm, *occ1 = sm + nw, *occ2 = sm + 2 * nw, *A = sm + 3 * nw, *B = sm + 4 * nw;
    __shared__ int acc[3];
    for (int i = threadIdx.x; i < nw; i += blockDim.x) C[i] = grid[i];
    for (long long L = blockIdx.x; L < n; L += gridDim.x) {
        const uint32_t p0 = offsets[L], p1 = offsets[L + 1];
        if (threadIdx.x < 3) acc[threadIdx.x] = 0;
        for (int i = threadIdx.x; i < nw; i += blockDim.x) { occ1[i] = 0; occ2[i] = 0; }
        __syncthreads();
        // stamp footprints (platform_layout.rs:104-124); tiles outside the grid are skipped and flag the platform
        for (uint32_t p = p0 + threadIdx.x; p < p1; p += blockDim.x) {
            int4 pl = plats[p];
            int x0 = max(pl.x, 0), y0 = max(pl.y, 0), x1 = min(pl.x + pl.z, w), y1 = min(pl.y + pl.w, h);
            for (int y = y0; y < y1; y++)
                for (int wd = x0 >> 5; wd <= (x1 - 1) >> 5 && x1 > x0; wd++) {
                    int lo = max(x0, wd * 32) - wd * 32, hi = min(x1, wd * 32 + 32) - wd * 32;  // [lo, hi) within the word
                    uint32_t m = (hi - lo == 32) ? 0xffffffffu : (((1u << (hi - lo)) - 1u) << lo);
                    uint32_t old = atomicOr(&occ1[y * wpr + wd], m);
                    if (old & m) atomicOr(&occ2[y * wpr + wd], old & m);
                }
        }
        __syncthreads();
        int n_ov = 0, n_oob = 0;
        for (uint32_t p = p0 + threadIdx.x; p < p1; p += blockDim.x) {
            int4 pl = plats[p];
            bool oob = pl.x < 0 || pl.y < 0 || pl.x + pl.z > w || pl.y + pl.w > h;
            int x0 = max(pl.x, 0), y0 = max(pl.y, 0), x1 = min(pl.x + pl.z, w), y1 = min(pl.y + pl.w, h);
            bool ov = false;
            for (int y = y0; y < y1; y++)
                for (int wd = x0 >> 5; wd <= (x1 - 1) >> 5 && x1 > x0; wd++) {
                    int lo = max(x0, wd * 32) - wd * 32, hi = min(x1, wd * 32 + 32) - wd * 32;
                    uint32_t m = (hi - lo == 32) ? 0xffffffffu : (((1u << (hi - lo)) - 1u) << lo);
                    ov = ov || (occ2[y * wpr + wd] & m);
                }
            n_ov += ov;
            n_oob += oob;
            if (flags) flags[p] = (ov ? 1 : 0) | (oob ? 2 : 0);
        }
        block_sum(n_ov, &acc[1]);
        block_sum(n_oob, &acc[2]);
        for (int i = threadIdx.x; i < nw; i += blockDim.x) {
            A[i] = occ1[i] & C[i];
            if (layers) layers[(L * 4 + 0) * nw + i] = A[i];
        }
        __syncthreads();
        uint32_t *src = A, *dst = B;
        for (int round = 1; round < kTerrainSupportDistance; round++) {
            tiled_round(src, dst, C, nw, wpr);
            __syncthreads();
            if (layers)
                for (int i = threadIdx.x; i < nw; i += blockDim.x) layers[(L * 4 + round) * nw + i] = dst[i];
            uint32_t* t = src; src = dst; dst = t;
        }
        int unc = 0;
        for (int i = threadIdx.x; i < nw; i += blockDim.x) {
            uint32_t u = C[i] & ~src[i];
            unc += __popc(u);
            if (unsupported_rows) unsupported_rows[L * nw + i] = u;
        }
        block_sum(unc, &acc[0]);
        __syncthreads();
        if (threadIdx.x == 0) out[L] = make_int4(acc[0], (int)(p1 - p0), acc[1], acc[2]);
        __syncthreads();
    }
}

// u8 masks [n][w*h] -> compact rows.  One thread per output u32 word of the compact layout.
__global__ void pack_bytes_kernel(const uint8_t* __restrict__ bytes, int w, int h, long long n, int row_bytes, int wpl,
                                  uint32_t* __restrict__ out) {
    long long total = n * wpl;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        long long L = i / wpl;
        int word = (int)(i % wpl);
        const uint8_t* src = bytes + L * (long long)w * h;
        uint32_t v = 0;
        if (row_bytes >= 4) {  // wide rows: word index = y*wpr + wx
            int wpr = row_bytes / 4, y = word / wpr, wx = word % wpr;
            for (int b = 0; b < 32; b++) { int x = wx * 32 + b; if (x < w && src[y * w + x]) v |= 1u << b; }
        } else {
            int rows_per_word = 4 / row_bytes, bits = row_bytes * 8;
            for (int s = 0; s < rows_per_word; s++) {
                int y = word * rows_per_word + s;
                if (y >= h) break;
                for (int x = 0; x < w; x++) if (src[y * w + x]) v |= 1u << (s * bits + x);
            }
        }
        out[i] = v;
    }
}

// ------------------------------------------------------------------------------------------------ launchers
int launch_eval_compact(tss_engine* e, const void* grid_dev, int w, int h, const void* layouts_dev, int64_t n,
                        bool per_layout_terrain, int32_t* out_dev) {
    if (n <= 0) return TSS_OK;
    e->stats.layouts_evaluated += (uint64_t)n;
    if (tss_is_small(w, h)) return launch_eval_small(e, grid_dev, w, h, layouts_dev, n, per_layout_terrain, out_dev);
    int wpr = (w + 31) / 32, nw = h * wpr;
    size_t smem = (size_t)3 * nw * 4;
    if (smem > 200 * 1024) return e->fail(TSS_E_UNSUPPORTED, "grid %dx%d does not fit the shared-memory evaluator", w, h);
    long long blocks = n < (long long)e->prop.multiProcessorCount * 4 ? n : (long long)e->prop.multiProcessorCount * 4;
    if (per_layout_terrain) {
        TSS_CUDA(e, cudaFuncSetAttribute(eval_tiled_sites_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        eval_tiled_sites_kernel<true><<<(unsigned)blocks, 256, smem, e->stream>>>((const uint32_t*)grid_dev, (const uint32_t*)layouts_dev, n, h, wpr, (int2*)out_dev);
    } else {
        TSS_CUDA(e, cudaFuncSetAttribute(eval_tiled_sites_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        eval_tiled_sites_kernel<false><<<(unsigned)blocks, 256, smem, e->stream>>>((const uint32_t*)grid_dev, (const uint32_t*)layouts_dev, n, h, wpr, (int2*)out_dev);
    }
    TSS_CHECK_LAUNCH(e);
    e->stats.kernel_launches++;
    return TSS_OK;
}

int launch_eval_platforms(tss_engine* e, const uint32_t* grid_rows_dev, int w, int h, const int4* plats_dev,
                          const uint32_t* offsets_dev, int64_t n, int32_t* out_dev, uint32_t* unsupported_rows_dev,
                          uint8_t* flags_dev, uint32_t* layers_dev) {
    if (n <= 0) return TSS_OK;
    int wpr = (w + 31) / 32, nw = h * wpr;
    size_t smem = (size_t)5 * nw * 4;
    if (smem > 200 * 1024) return e->fail(TSS_E_UNSUPPORTED, "grid %dx%d does not fit the shared-memory evaluator", w, h);
    TSS_CUDA(e, cudaFuncSetAttribute(eval_platforms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    long long blocks = n < (long long)e->prop.multiProcessorCount * 4 ? n : (long long)e->prop.multiProcessorCount * 4;
    int threads = nw >= 1024 ? 256 : 128;
    eval_platforms_kernel<<<(unsigned)blocks, threads, smem, e->stream>>>(grid_rows_dev, w, h, wpr, plats_dev, offsets_dev, n, (int4*)out_dev,
                                                                        unsupported_rows_dev, flags_dev, layers_dev);
    TSS_CHECK_LAUNCH(e);
    e->stats.kernel_launches++;
    e->stats.layouts_evaluated += (uint64_t)n;
    return TSS_OK;
}

int launch_pack_bytes(tss_engine* e, const uint8_t* bytes_dev, int w, int h, int64_t n, void* compact_dev) {
    if (n <= 0) return TSS_OK;
    int wpl = (int)(tss_layout_bytes(w, h) / 4);
    long long total = n * wpl;
    long long blocks = (total + 255) / 256;
    long long cap = (long long)e->prop.multiProcessorCount * 16;
    if (blocks > cap) blocks = cap;
    pack_bytes_kernel<<<(unsigned)blocks, 256, 0, e->stream>>>(bytes_dev, w, h, n, (int)tss_row_bytes(w, h), wpl, (uint32_t*)compact_dev);
    TSS_CHECK_LAUNCH(e);
    e->stats.kernel_launches++;
    return TSS_OK;
}

void pack_compact_host(const uint8_t* grid, int w, int h, uint8_t* out) {
    size_t rb = tss_row_bytes(w, h), lb = tss_layout_bytes(w, h);
    for (size_t i = 0; i < lb; i++) out[i] = 0;
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++)
            if (grid[(size_t)y * w + x]) out[(size_t)y * rb + (x >> 3)] |= (uint8_t)(1u << (x & 7));
}

void rows_to_compact_host(const uint32_t* rows, int w, int h, uint8_t* out) {
    size_t rb = tss_row_bytes(w, h), lb = tss_layout_bytes(w, h);
    int wpr = (w + 31) / 32;
    for (size_t i = 0; i < lb; i++) out[i] = 0;
    for (int y = 0; y < h; y++)
        for (size_t b = 0; b < rb; b++) out[(size_t)y * rb + b] = (uint8_t)(rows[(size_t)y * wpr + (b >> 2)] >> (8 * (b & 3)));
}

}  // namespace tss
