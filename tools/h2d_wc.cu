// Host-to-device copy rate of a 288 MB pinned buffer: default page-locked memory against write-combined
// (cudaHostAllocWriteCombined), one GPU.   nvcc -o h2d_wc h2d_wc.cu && ./h2d_wc
#include <cstdio>
#include <cstring>
#include <cuda_runtime.h>
int main() {
    const size_t n = 288000000;
    void *dev;
    cudaMalloc(&dev, n);
    const unsigned flags[2] = {cudaHostAllocDefault, cudaHostAllocWriteCombined};
    const char *names[2] = {"default", "write-combined"};
    for (int f = 0; f < 2; ++f) {
        void *h;
        if (cudaHostAlloc(&h, n, flags[f]) != cudaSuccess) {
            printf("%s: alloc failed\n", names[f]);
            continue;
        }
        memset(h, 7, n);
        cudaEvent_t a, b;
        cudaEventCreate(&a);
        cudaEventCreate(&b);
        for (int i = 0; i < 3; ++i)
            cudaMemcpyAsync(dev, h, n, cudaMemcpyHostToDevice, 0);
        cudaEventRecord(a, 0);
        for (int i = 0; i < 10; ++i)
            cudaMemcpyAsync(dev, h, n, cudaMemcpyHostToDevice, 0);
        cudaEventRecord(b, 0);
        cudaEventSynchronize(b);
        float ms;
        cudaEventElapsedTime(&ms, a, b);
        printf("%s: %.3f ms per 288 MB = %.1f GB/s\n", names[f], ms / 10, n / (ms / 10 * 1e-3) / 1e9);
        // 8 chunks of 36 MB each, as the pipeline issues them
        cudaEventRecord(a, 0);
        for (int i = 0; i < 10; ++i)
            for (int c = 0; c < 8; ++c)
                cudaMemcpyAsync((char *) dev + c * (n / 8), (char *) h + c * (n / 8), n / 8, cudaMemcpyHostToDevice, 0);
        cudaEventRecord(b, 0);
        cudaEventSynchronize(b);
        cudaEventElapsedTime(&ms, a, b);
        printf("%s, 8 chunks: %.3f ms per 288 MB = %.1f GB/s\n", names[f], ms / 10, n / (ms / 10 * 1e-3) / 1e9);
        cudaFreeHost(h);
    }
    return 0;
}
