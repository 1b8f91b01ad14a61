// Micro-benchmark (B200): throughput of shared-memory accumulation for the message phase of the engine.
//   mode 0  red.shared.add.u32 to 32 random distinct cells per warp instruction (what the q4 stream does)
//   mode 1  LDS + IADD + STS into a per-warp private array (no atomics)
//   mode 2  red.shared.add.u32, all lanes of a warp in distinct BANKS (conflict-free pattern)
//   mode 3  red.shared.add.u64 (two adjacent cells per lane)
// Prints warp-instructions (record rows) per cycle per SM. Build: nvcc -arch=sm_100a -O3 -o atoms_probe atoms_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256, 4) probe(const uint32_t *recs, uint32_t *out, int iters, int cells)
{
    extern __shared__ uint32_t acc[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < cells * (MODE == 1 ? 8 : 1); i += 256) acc[i] = 0;
    __syncthreads();
    const uint32_t base = static_cast<uint32_t>(__cvta_generic_to_shared(acc)) + (MODE == 1 ? warp * cells * 4 : 0);
    uint4 r = reinterpret_cast<const uint4 *>(recs)[(blockIdx.x * 256 + threadIdx.x) & 65535];
    const uint32_t mask = (cells - 1) * 4;
    for (int it = 0; it < iters; ++it)
    {
        const uint32_t q[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
        for (int u = 0; u < 4; ++u)
        {
            uint32_t a = q[u] & mask;
            if (MODE == 2) a = (a & ~124u) | (lane << 2);
            if (MODE == 0 || MODE == 2) asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(base + a), "r"(q[u] >> 14) : "memory");
            else if (MODE == 3) asm volatile("red.shared.add.u64 [%0], %1;" ::"r"(base + (a & ~7u)), "l"((unsigned long long)(q[u] >> 14)) : "memory");
            else
            {
                uint32_t v;
                asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(base + a) : "memory");
                v += q[u] >> 14;
                asm volatile("st.shared.u32 [%0], %1;" ::"r"(base + a), "r"(v) : "memory");
            }
        }
        // next pseudo-random records (cheap LCG per component)
        r.x = r.x * 1664525u + 1013904223u;
        r.y = r.y * 22695477u + 1u;
        r.z = r.z * 1103515245u + 12345u;
        r.w = r.w * 134775813u + 1u;
    }
    __syncthreads();
    uint32_t s = 0;
    for (int i = threadIdx.x; i < cells; i += 256) s += acc[i];
    if (s == 0xdeadbeef) out[blockIdx.x] = s;
}

template <int MODE> void run(const char *name, const uint32_t *recs, uint32_t *out, int cells)
{
    const int iters = 4000, grid = 148 * 4;
    const size_t smem = (MODE == 1 ? 8 : 1) * cells * 4;
    cudaFuncSetAttribute(probe<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    probe<MODE><<<grid, 256, smem>>>(recs, out, 100, cells);
    cudaEventRecord(a);
    probe<MODE><<<grid, 256, smem>>>(recs, out, iters, cells);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    int clk = 0;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const double rows = double(grid) * 8 * iters * 4; // warp-level accumulate instructions
    const double cycles = ms * 1e-3 * clk * 1e3;
    printf("%-28s cells=%5d  %.3f ms  %.2f cycles per warp-row per SM  (%.1f G records/s chip-wide) err=%s\n", name, cells, ms,
            cycles / (rows / 148.0), rows * 32 / (ms * 1e-3) / 1e9, cudaGetErrorString(cudaGetLastError()));
}

int main()
{
    uint32_t *recs, *out;
    cudaMalloc(&recs, 65536 * 16);
    cudaMalloc(&out, 4096);
    uint32_t *h = new uint32_t[65536 * 4];
    uint32_t x = 12345;
    for (int i = 0; i < 65536 * 4; ++i) { x = x * 1664525u + 1013904223u; h[i] = x ^ (x >> 13); }
    cudaMemcpy(recs, h, 65536 * 16, cudaMemcpyHostToDevice);
    for (int cells : {1024, 4096})
    {
        run<0>("red.shared.u32 random", recs, out, cells);
        run<2>("red.shared.u32 conflict-free", recs, out, cells);
        run<1>("LDS+IADD+STS private/warp", recs, out, cells);
        run<3>("red.shared.u64 random", recs, out, cells);
    }
    return 0;
}
