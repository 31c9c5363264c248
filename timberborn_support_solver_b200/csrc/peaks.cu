// Measured integer-issue and shared-memory peaks: the roofline denominators of kernels (a) and (b)
// are LOP3/SHF/POPC issue rate and warp-shuffle / shared-memory throughput, none of which MEASURED_PEAKS.json holds.
// Each micro-benchmark runs dependent-free chains at full occupancy (8 x 256 threads per SM) and reports thread-ops/s.
#include "engine.hpp"

namespace tss {

constexpr int PEAK_ITERS = 4096;
constexpr int CHAINS = 8;

__global__ void __launch_bounds__(256) peak_lop3_kernel(uint32_t* out, uint32_t seed) {
    uint32_t a[CHAINS], b = seed ^ threadIdx.x, c = seed * 0x9E3779B9u + blockIdx.x;
#pragma unroll
    for (int i = 0; i < CHAINS; i++) a[i] = seed + i * 0x85EBCA6Bu + threadIdx.x;
    for (int it = 0; it < PEAK_ITERS; it++) {
#pragma unroll
        for (int i = 0; i < CHAINS; i++) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b), "r"(c));
    }
    uint32_t r = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; i++) r ^= a[i];
    if (r == 0x12345678u) out[0] = r;
}

__global__ void __launch_bounds__(256) peak_popc_kernel(uint32_t* out, uint32_t seed) {
    uint32_t a[CHAINS];
#pragma unroll
    for (int i = 0; i < CHAINS; i++) a[i] = seed + i * 0x85EBCA6Bu + threadIdx.x * 2654435761u;
    for (int it = 0; it < PEAK_ITERS; it++) {
#pragma unroll
        for (int i = 0; i < CHAINS; i++) asm volatile("popc.b32 %0, %0;" : "+r"(a[i]));
    }
    uint32_t r = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; i++) r += a[i];
    if (r == 0x12345678u) out[0] = r;
}

__global__ void __launch_bounds__(256) peak_shfl_kernel(uint32_t* out, uint32_t seed) {
    uint32_t a[CHAINS];
    const int src = (threadIdx.x * 7 + 3) & 31;
#pragma unroll
    for (int i = 0; i < CHAINS; i++) a[i] = seed + i + threadIdx.x;
    for (int it = 0; it < PEAK_ITERS; it++) {
#pragma unroll
        for (int i = 0; i < CHAINS; i++) a[i] = __shfl_sync(0xffffffffu, a[i], src);
    }
    uint32_t r = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; i++) r += a[i];
    if (r == 0x12345678u) out[0] = r;
}

__global__ void __launch_bounds__(256) peak_smem_kernel(uint32_t* out, uint32_t seed) {
    __shared__ uint4 buf[1024];  // 16 KB
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) buf[i] = make_uint4(seed + i, i, seed, 1);
    __syncthreads();
    uint4 acc = make_uint4(0, 0, 0, 0);
    int idx = threadIdx.x;
    for (int it = 0; it < PEAK_ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 4; i++) {
            uint4 v = buf[(idx + i * 256) & 1023];  // conflict-free 16-byte loads
            acc.x ^= v.x; acc.y ^= v.y; acc.z ^= v.z; acc.w ^= v.w;
        }
        idx = (idx + 1) & 1023;
    }
    if ((acc.x ^ acc.y ^ acc.z ^ acc.w) == 0x12345678u) out[0] = acc.x;
}

__global__ void peak_clock_kernel(long long* out, int spin) {
    long long t0 = clock64();
    uint32_t x = threadIdx.x;
    for (int i = 0; i < spin; i++) asm volatile("lop3.b32 %0, %0, %0, %0, 0x96;" : "+r"(x));
    long long t1 = clock64();
    if (threadIdx.x == 0) { out[0] = t1 - t0; out[1] = x; }
}

template <class F>
static int timed(tss_engine* e, F launch, double* ms_out) {
    launch();  // warm-up
    TSS_CHECK_LAUNCH(e);
    TSS_CUDA(e, cudaStreamSynchronize(e->stream));
    double best = 1e30;
    for (int rep = 0; rep < 3; rep++) {
        TSS_CUDA(e, cudaEventRecord(e->ev0, e->stream));
        launch();
        TSS_CUDA(e, cudaEventRecord(e->ev1, e->stream));
        TSS_CUDA(e, cudaStreamSynchronize(e->stream));
        float ms = 0;
        TSS_CUDA(e, cudaEventElapsedTime(&ms, e->ev0, e->ev1));
        best = ms < best ? ms : best;
        e->stats.kernel_launches++;
    }
    *ms_out = best;
    return TSS_OK;
}

int run_peaks(tss_engine* e, double* out, int n_out) {
    uint32_t* sink = (uint32_t*)e->dev(7, 64);
    if (!sink) return TSS_E_CUDA;
    const int blocks = e->prop.multiProcessorCount * 8, threads = 256;
    const double thread_iters = (double)blocks * threads * PEAK_ITERS;
    double ms = 0.0;
    int rc;
    cudaStream_t st = e->stream;
    if ((rc = timed(e, [&] { peak_lop3_kernel<<<blocks, threads, 0, st>>>(sink, 17u); }, &ms))) return rc;
    out[0] = thread_iters * CHAINS / (ms * 1e-3) / 1e9;
    if ((rc = timed(e, [&] { peak_popc_kernel<<<blocks, threads, 0, st>>>(sink, 17u); }, &ms))) return rc;
    out[1] = thread_iters * CHAINS / (ms * 1e-3) / 1e9;
    if ((rc = timed(e, [&] { peak_shfl_kernel<<<blocks, threads, 0, st>>>(sink, 17u); }, &ms))) return rc;
    out[2] = thread_iters * CHAINS / (ms * 1e-3) / 1e9;
    if ((rc = timed(e, [&] { peak_smem_kernel<<<blocks, threads, 0, st>>>(sink, 17u); }, &ms))) return rc;
    out[3] = thread_iters * 4 * 16 / (ms * 1e-3) / 1e9;  // GB/s
    // SM clock under load: cycles of a fixed spin / its event time
    long long* cyc = (long long*)((char*)sink + 16);
    const int spin = 1 << 17;
    if ((rc = timed(e, [&] { peak_clock_kernel<<<blocks, threads, 0, st>>>(cyc, spin); }, &ms))) return rc;
    long long host_cyc[2];
    TSS_CUDA(e, cudaMemcpy(host_cyc, cyc, sizeof host_cyc, cudaMemcpyDeviceToHost));
    out[4] = (double)host_cyc[0] / (ms * 1e-3) / 1e6;
    for (int i = 5; i < n_out; i++) out[i] = 0;
    return TSS_OK;
}

}  // namespace tss
