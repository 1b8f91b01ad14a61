// p2p_probe.cu — cost of the pieces of a peer-memory exchange between two GPUs of one box
// (development aid for the raster exchange; not part of the library).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/p2p_probe tools/p2p_probe.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { std::printf("%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

__global__ void push(const uint4 *src, uint4 *dst, uint32_t *flag, int n_vec, int mode, uint32_t value)
{
    for (int i = threadIdx.x + blockIdx.x * blockDim.x; i < n_vec; i += blockDim.x * gridDim.x) dst[i] = src[i];
    if (mode >= 1)
    {
        __syncthreads();
        if (threadIdx.x == 0)
        {
            if (mode >= 2) __threadfence_system();
            if (mode >= 3) asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(flag), "r"(value) : "memory");
        }
    }
}
__global__ void fence_only(int sys)
{
    if (sys) __threadfence_system();
    else __threadfence();
}
__global__ void wait_flag(const uint32_t *flag, uint32_t want, uint32_t *out)
{
    uint32_t v;
    long long t0 = clock64();
    do
    {
        asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
    } while (v < want && clock64() - t0 < 2000000000ll);
    *out = v;
}

int main()
{
    int n = 0;
    CK(cudaGetDeviceCount(&n));
    if (n < 2) { std::printf("need 2 GPUs\n"); return 0; }
    int can = 0;
    CK(cudaDeviceCanAccessPeer(&can, 0, 1));
    std::printf("peer access 0->1: %d\n", can);
    CK(cudaSetDevice(1));
    uint4 *remote; uint32_t *rflag, *rout;
    CK(cudaMalloc(&remote, 1 << 20)); CK(cudaMalloc(&rflag, 256)); CK(cudaMalloc(&rout, 256));
    CK(cudaMemset(rflag, 0, 256));
    CK(cudaDeviceEnablePeerAccess(0, 0));
    CK(cudaSetDevice(0));
    CK(cudaDeviceEnablePeerAccess(1, 0));
    uint4 *local, *local_dst; uint32_t *lflag;
    CK(cudaMalloc(&local, 1 << 20)); CK(cudaMalloc(&local_dst, 1 << 20)); CK(cudaMalloc(&lflag, 256));
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    const int reps = 200;
    auto time = [&](const char *what, auto &&launch) {
        for (int i = 0; i < 10; ++i) launch(i);
        cudaEventRecord(a);
        for (int i = 0; i < reps; ++i) launch(i);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms = 0;
        cudaEventElapsedTime(&ms, a, b);
        std::printf("%-58s %7.2f us per launch\n", what, 1e3 * ms / reps);
    };
    time("empty-ish kernel (fence.gpu, 1 thread)", [&](int) { fence_only<<<1, 1>>>(0); });
    time("fence.sys, 1 thread, nothing outstanding", [&](int) { fence_only<<<1, 1>>>(1); });
    for (int kb : {16, 64})
    {
        const int n_vec = kb * 1024 / 16;
        char buf[128];
        std::snprintf(buf, sizeof buf, "%d KB local copy, 256 thr", kb);
        time(buf, [&](int i) { push<<<1, 256>>>(local, local_dst, lflag, n_vec, 0, i); });
        std::snprintf(buf, sizeof buf, "%d KB peer stores only, 256 thr", kb);
        time(buf, [&](int i) { push<<<1, 256>>>(local, remote, rflag, n_vec, 0, i); });
        std::snprintf(buf, sizeof buf, "%d KB peer stores + barrier", kb);
        time(buf, [&](int i) { push<<<1, 256>>>(local, remote, rflag, n_vec, 1, i); });
        std::snprintf(buf, sizeof buf, "%d KB peer stores + barrier + fence.sys", kb);
        time(buf, [&](int i) { push<<<1, 256>>>(local, remote, rflag, n_vec, 2, i); });
        std::snprintf(buf, sizeof buf, "%d KB peer stores + barrier + fence.sys + flag", kb);
        time(buf, [&](int i) { push<<<1, 256>>>(local, remote, rflag, n_vec, 3, i + 1); });
        std::snprintf(buf, sizeof buf, "%d KB peer stores, 8 CTAs + fence.sys + flag", kb);
        time(buf, [&](int i) { push<<<8, 256>>>(local, remote, rflag, n_vec, 3, i + 1); });
    }
    // one-way latency: GPU0 pushes 16 KB + flag, GPU1 waits for the flag; both timed on GPU1's clock
    CK(cudaSetDevice(1));
    cudaStream_t s1; CK(cudaStreamCreate(&s1));
    cudaEvent_t c, d; CK(cudaEventCreate(&c)); CK(cudaEventCreate(&d));
    CK(cudaMemset(rflag, 0, 256));
    CK(cudaDeviceSynchronize());
    float total = 0;
    for (int i = 0; i < 50; ++i)
    {
        CK(cudaSetDevice(1));
        cudaEventRecord(c, s1);
        wait_flag<<<1, 1, 0, s1>>>(rflag, 1000 + i, rout);
        cudaEventRecord(d, s1);
        CK(cudaSetDevice(0));
        push<<<1, 256>>>(local, remote, rflag, 1024, 3, 1000 + i);
        CK(cudaSetDevice(1));
        CK(cudaEventSynchronize(d));
        float ms = 0; cudaEventElapsedTime(&ms, c, d);
        if (i >= 10) total += ms;
        CK(cudaSetDevice(0)); CK(cudaDeviceSynchronize());
    }
    std::printf("wait kernel on GPU1 incl. host launch of the pusher:          %7.2f us\n", 1e3 * total / 40);
    return 0;
}
