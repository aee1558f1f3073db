// ubench2.cu -- throughput of the packed FP32 instructions (FFMA2 / FADD2, sm_100a) and of the int -> float
// conversions, next to their scalar forms; same method as ubench.cu (warp-instructions per clock per SM).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define ITERS 512
#define CHAINS 8

#define KBEGIN(name, T)                                                                                  \
    __global__ void __launch_bounds__(1024, 1) name(uint32_t *out, long long *cyc, uint32_t seed) {     \
        T x[CHAINS];                                                                                      \
        const uint32_t y32 = seed * 2654435761u + threadIdx.x * 40503u + 1u;                              \
        T y, z;                                                                                           \
        init(y, y32);                                                                                     \
        init(z, y32 ^ 0x9e3779b9u);                                                                       \
        _Pragma("unroll") for (int c = 0; c < CHAINS; ++c) init(x[c], y32 + c * 77u);                     \
        __syncthreads();                                                                                  \
        long long t0 = clock64();                                                                         \
        for (int it = 0; it < ITERS; ++it) {                                                              \
            _Pragma("unroll") for (int c = 0; c < CHAINS; ++c) {

#define KEND                                                                                              \
            }                                                                                             \
        }                                                                                                 \
        long long t1 = clock64();                                                                         \
        uint32_t acc = 0;                                                                                 \
        _Pragma("unroll") for (int c = 0; c < CHAINS; ++c) acc ^= fold(x[c]);                             \
        out[blockIdx.x * blockDim.x + threadIdx.x] = acc;                                                 \
        __syncthreads();                                                                                  \
        if (threadIdx.x == 0)                                                                             \
            cyc[blockIdx.x] = t1 - t0;                                                                    \
    }

__device__ __forceinline__ void init(unsigned long long &v, uint32_t s) { v = ((unsigned long long) __float_as_uint(1.0f + (s & 1023) * 1e-6f) << 32) | __float_as_uint(1.0f + (s >> 10 & 1023) * 1e-6f); }
__device__ __forceinline__ void init(float &v, uint32_t s) { v = 1.0f + (s & 1023) * 1e-6f; }
__device__ __forceinline__ void init(uint32_t &v, uint32_t s) { v = s; }
__device__ __forceinline__ uint32_t fold(unsigned long long v) { return (uint32_t) v ^ (uint32_t) (v >> 32); }
__device__ __forceinline__ uint32_t fold(float v) { return __float_as_uint(v); }
__device__ __forceinline__ uint32_t fold(uint32_t v) { return v; }

KBEGIN(k_ffma2, unsigned long long) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x[c]) : "l"(y), "l"(z)); KEND
KBEGIN(k_fadd2, unsigned long long) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(x[c]) : "l"(y)); KEND
KBEGIN(k_fmul2, unsigned long long) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(x[c]) : "l"(y)); KEND
KBEGIN(k_ffma, float) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x[c]) : "f"(y), "f"(z)); KEND
KBEGIN(k_fadd, float) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(x[c]) : "f"(y)); KEND
// conversions: u32 -> f32 and back (each step one I2F + one F2I, dependent)
KBEGIN(k_i2f_f2i, uint32_t) { float f; asm volatile("cvt.rn.f32.u32 %0, %1;" : "=f"(f) : "r"(x[c])); asm volatile("cvt.rzi.u32.f32 %0, %1;" : "=r"(x[c]) : "f"(f)); } KEND
// u32 -> f32 only (+ a bit cast back: the next conversion reads the float's bits as an integer)
KBEGIN(k_i2f, uint32_t) { float f; asm volatile("cvt.rn.f32.u32 %0, %1;" : "=f"(f) : "r"(x[c] & 0xffffu)); x[c] = __float_as_uint(f); } KEND
// u16 -> f32 (the LDS.U16 result is a zero-extended 16-bit value)
KBEGIN(k_i2f_u16, uint32_t) { float f; asm volatile("{ .reg .u16 h; cvt.u16.u32 h, %1; cvt.rn.f32.u16 %0, h; }" : "=f"(f) : "r"(x[c])); x[c] = __float_as_uint(f); } KEND
// mixes
KBEGIN(k_mix_ffma2_shf, unsigned long long) { asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x[c]) : "l"(y), "l"(z)); uint32_t lo = (uint32_t) x[c], hi = (uint32_t) (x[c] >> 32); asm volatile("shf.l.wrap.b32 %0, %0, %1, 1;" : "+r"(lo) : "r"(hi)); x[c] = ((unsigned long long) hi << 32) | lo; } KEND
KBEGIN(k_mix_ffma2_iadd_imad, unsigned long long) { asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x[c]) : "l"(y), "l"(z)); uint32_t lo = (uint32_t) x[c], hi = (uint32_t) (x[c] >> 32); asm volatile("add.u32 %0, %0, %1;" : "+r"(lo) : "r"(hi)); asm volatile("mad.lo.u32 %0, %0, 3, %1;" : "+r"(hi) : "r"(lo)); x[c] = ((unsigned long long) hi << 32) | lo; } KEND

typedef void (*kern_t)(uint32_t *, long long *, uint32_t);
static void run(const char *name, kern_t k, int nops, int sms, uint32_t *d_out, long long *d_cyc) {
    for (int rep = 0; rep < 2; ++rep)
        k<<<sms, 1024>>>(d_out, d_cyc, 12345u + rep);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        printf("%-24s FAILED: %s\n", name, cudaGetErrorString(e));
        return;
    }
    long long *h = (long long *) malloc(sms * sizeof(long long));
    cudaMemcpy(h, d_cyc, sms * sizeof(long long), cudaMemcpyDeviceToHost);
    double sum = 0;
    for (int i = 0; i < sms; ++i)
        sum += (double) h[i];
    const double cyc = sum / sms;
    printf("%-24s %8.3f source-ops/clk/SM  (%d op(s) per step, %.0f cycles; see the SASS for what each step became)\n", name,
           32.0 * ITERS * CHAINS * nops / cyc, nops, cyc);
    free(h);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    uint32_t *d_out;
    long long *d_cyc;
    cudaMalloc(&d_out, (size_t) sms * 1024 * 4);
    cudaMalloc(&d_cyc, sms * sizeof(long long));
#define RUN(k, n) run(#k, k, n, sms, d_out, d_cyc)
    RUN(k_ffma2, 1); RUN(k_fadd2, 1); RUN(k_fmul2, 1); RUN(k_ffma, 1); RUN(k_fadd, 1);
    RUN(k_i2f_f2i, 2); RUN(k_i2f, 1); RUN(k_i2f_u16, 1); RUN(k_mix_ffma2_shf, 2); RUN(k_mix_ffma2_iadd_imad, 3);
    return 0;
}
