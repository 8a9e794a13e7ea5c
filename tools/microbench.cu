// tools/microbench.cu — roofline denominators that MEASURED_PEAKS.json does not carry:
// FP32 FFMA and FP64 DFMA issue rates, L2 read bandwidth (working set << L2) and HBM read
// bandwidth (working set >> L2). Prints one JSON object. Build: see tools/build_tools.sh.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <vector>

template <typename T>
__global__ void k_fma(T* out, int iters)
{
    T a[8];
    for (int i = 0; i < 8; i++) a[i] = (T)(threadIdx.x * 1e-3 + i);
    const T b = (T)1.000001, c = (T)1e-7;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) a[i] = a[i] * b + c;
    }
    T s = 0;
    for (int i = 0; i < 8; i++) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_read(const uint4* __restrict__ p, size_t n, int reps, uint4* out)
{
    uint4 acc = make_uint4(0, 0, 0, 0);
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (int r = 0; r < reps; r++)
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
            uint4 v = p[i];
            acc.x ^= v.x; acc.y ^= v.y; acc.z ^= v.z; acc.w ^= v.w;
        }
    if (acc.x == 0x12345678u) out[0] = acc;
}

// dependent random 16-byte loads (pointer chase through a permutation): latency-bound access like a KD walk
__global__ void k_chase(const uint4* __restrict__ p, uint32_t mask, int steps, uint32_t* out)
{
    uint32_t idx = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u & mask;
    for (int s = 0; s < steps; s++) idx = p[idx].x & mask;
    if (idx == 0xffffffffu) out[0] = idx;
}

// independent random CHUNK-byte reads (hashed addresses, 4 in flight per lane): the cost of a random DRAM access as a
// function of its size -- what a KD walk pays per tree block / triangle record that misses L2
template <int CHUNK>
__global__ void k_gather(const uint4* __restrict__ p, uint32_t chunkMask, int steps, uint32_t* out)
{
    uint32_t h = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
    uint32_t acc = 0;
    for (int s = 0; s < steps; s++) {
        h = h * 1664525u + 1013904223u;
        const uint4* q = p + (size_t)((h >> 4) & chunkMask) * (CHUNK / 16);
#pragma unroll
        for (int k = 0; k < CHUNK / 16; k++) acc ^= __ldg(q + k).x;
    }
    if (acc == 0x12345u) out[0] = acc;
}

template <int CHUNK>
static double gather_rate(const uint4* buf, size_t bytes, int sms, uint32_t* o32, cudaEvent_t a, cudaEvent_t b)
{
    const int steps = 64, threads = sms * 2048;
    const uint32_t mask = (uint32_t)(bytes / CHUNK) - 1;
    float best = 1e30f;
    for (int r = 0; r < 3; r++) {
        cudaEventRecord(a); k_gather<CHUNK><<<threads / 256, 256>>>(buf, mask, steps, o32); cudaEventRecord(b);
        float ms; cudaEventSynchronize(b); cudaEventElapsedTime(&ms, a, b);
        if (r && ms < best) best = ms;
    }
    return (double)threads * steps / (best * 1e-3) / 1e9;
}

static float timeit(cudaEvent_t a, cudaEvent_t b) { float ms; cudaEventSynchronize(b); cudaEventElapsedTime(&ms, a, b); return ms; }

int main()
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    const int blocks = p.multiProcessorCount * 8, threads = 256, iters = 20000;
    float* of; double* od;
    cudaMalloc(&of, blocks * threads * 4); cudaMalloc(&od, blocks * threads * 8);
    float f32 = 1e30f, f64 = 1e30f;
    for (int r = 0; r < 4; r++) {
        cudaEventRecord(a); k_fma<float><<<blocks, threads>>>(of, iters); cudaEventRecord(b);
        float ms = timeit(a, b); if (r && ms < f32) f32 = ms;
        cudaEventRecord(a); k_fma<double><<<blocks, threads>>>(od, iters); cudaEventRecord(b);
        ms = timeit(a, b); if (r && ms < f64) f64 = ms;
    }
    const double nfma = (double)blocks * threads * iters * 8;
    // bandwidth
    const size_t big = (size_t)4 << 30, small = (size_t)32 << 20;
    uint4* buf; cudaMalloc(&buf, big); cudaMemset(buf, 1, big);
    uint4* out; cudaMalloc(&out, 64);
    float l2 = 1e30f, hbm = 1e30f;
    for (int r = 0; r < 4; r++) {
        cudaEventRecord(a); k_read<<<p.multiProcessorCount * 8, 512>>>(buf, small / 16, 64, out); cudaEventRecord(b);
        float ms = timeit(a, b); if (r && ms < l2) l2 = ms;
        cudaEventRecord(a); k_read<<<p.multiProcessorCount * 8, 512>>>(buf, big / 16, 1, out); cudaEventRecord(b);
        ms = timeit(a, b); if (r && ms < hbm) hbm = ms;
    }
    // dependent random loads: 16 MB (L2-resident) and 2 GB (HBM) tables
    float chL2 = 1e30f, chHbm = 1e30f;
    const int steps = 256, cthreads = p.multiProcessorCount * 2048;
    {
        std::vector<uint4> h((size_t)1 << 27 >> 0);  // 2 GB / 16 B = 128 M entries
        uint32_t x = 12345;
        for (size_t i = 0; i < h.size(); i++) { x = x * 1664525u + 1013904223u; h[i] = make_uint4(x >> 3, 0, 0, 0); }
        cudaMemcpy(buf, h.data(), h.size() * 16, cudaMemcpyHostToDevice);
    }
    uint32_t* o32; cudaMalloc(&o32, 64);
    for (int r = 0; r < 3; r++) {
        cudaEventRecord(a); k_chase<<<cthreads / 256, 256>>>(buf, (1u << 20) - 1, steps, o32); cudaEventRecord(b);
        float ms = timeit(a, b); if (r && ms < chL2) chL2 = ms;
        cudaEventRecord(a); k_chase<<<cthreads / 256, 256>>>(buf, (1u << 27) - 1, steps, o32); cudaEventRecord(b);
        ms = timeit(a, b); if (r && ms < chHbm) chHbm = ms;
    }
    // random access cost by size and by the L2 fetch granularity limit (2 GB table)
    {
        const int grans[3] = {32, 64, 128};
        printf("{\"random_gather_G_per_s\": {");
        for (int g = 0; g < 3; g++) {
            cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, grans[g]);
            size_t got = 0; cudaDeviceGetLimit(&got, cudaLimitMaxL2FetchGranularity);
            const size_t B = (size_t)2 << 30;
            printf("%s\"l2gran%d(got %zu)\": {\"16B\": %.2f, \"32B\": %.2f, \"64B\": %.2f, \"128B\": %.2f, \"256B\": %.2f, \"chase16B\": ", g ? ", " : "", grans[g], got,
                   gather_rate<16>(buf, B, p.multiProcessorCount, o32, a, b), gather_rate<32>(buf, B, p.multiProcessorCount, o32, a, b),
                   gather_rate<64>(buf, B, p.multiProcessorCount, o32, a, b), gather_rate<128>(buf, B, p.multiProcessorCount, o32, a, b),
                   gather_rate<256>(buf, B, p.multiProcessorCount, o32, a, b));
            float ch = 1e30f;
            for (int r = 0; r < 3; r++) {
                cudaEventRecord(a); k_chase<<<cthreads / 256, 256>>>(buf, (1u << 27) - 1, steps, o32); cudaEventRecord(b);
                float ms = timeit(a, b); if (r && ms < ch) ch = ms;
            }
            printf("%.2f}", (double)cthreads * steps / (ch * 1e-3) / 1e9);
        }
        printf("}}\n");
        cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, 64);
    }
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"fp32_tfma_per_s\": %.3f, \"fp64_tfma_per_s\": %.3f, \"l2_read_gbs\": %.1f, "
           "\"hbm_read_gbs\": %.1f, \"random16B_loads_L2_G_per_s\": %.3f, \"random16B_loads_HBM_G_per_s\": %.3f, \"cuda_err\": \"%s\"}\n",
           p.name, p.multiProcessorCount, nfma / (f32 * 1e-3) / 1e12, nfma / (f64 * 1e-3) / 1e12, (double)small * 64 / (l2 * 1e-3) / 1e9,
           (double)big / (hbm * 1e-3) / 1e9, (double)cthreads * steps / (chL2 * 1e-3) / 1e9, (double)cthreads * steps / (chHbm * 1e-3) / 1e9,
           cudaGetErrorString(cudaGetLastError()));
    return 0;
}
