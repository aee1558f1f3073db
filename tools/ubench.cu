// ubench.cu -- instruction-throughput probe for the integer / packed / shared-memory instructions the scan and
// slice kernels are built from (sm_100a).  Prints warp-instructions per clock per SM (4 = one per scheduler
// and clock) for each op, measured with clock64() inside one 1024-thread CTA per SM.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench tools/ubench.cu && ./ubench
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define ITERS 512
#define CHAINS 8

struct Res {
    long long cycles;
};

#define KERNEL_BEGIN(name)                                                                              \
    __global__ void __launch_bounds__(1024, 1) name(uint32_t *out, long long *cyc, uint32_t seed) {    \
        uint32_t x[CHAINS];                                                                             \
        uint32_t y = seed * 2654435761u + threadIdx.x * 40503u + 1u, z = (seed ^ 0x9e3779b9u) | 1u;    \
        _Pragma("unroll") for (int c = 0; c < CHAINS; ++c) x[c] = y + c * 77u;                          \
        __syncthreads();                                                                                \
        long long t0 = clock64();                                                                       \
        for (int it = 0; it < ITERS; ++it) {                                                            \
            _Pragma("unroll") for (int c = 0; c < CHAINS; ++c) {

#define KERNEL_END(nops)                                                                                \
            }                                                                                           \
        }                                                                                               \
        long long t1 = clock64();                                                                       \
        uint32_t acc = 0;                                                                               \
        _Pragma("unroll") for (int c = 0; c < CHAINS; ++c) acc ^= x[c];                                 \
        out[blockIdx.x * blockDim.x + threadIdx.x] = acc + y + z;                                       \
        __syncthreads();                                                                                \
        if (threadIdx.x == 0)                                                                           \
            cyc[blockIdx.x] = t1 - t0;                                                                  \
    }

KERNEL_BEGIN(k_iadd) asm volatile("add.u32 %0, %0, %1;" : "+r"(x[c]) : "r"(y)); KERNEL_END(1)
KERNEL_BEGIN(k_iadd3) asm volatile("{ .reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2; }" : "+r"(x[c]) : "r"(y), "r"(z)); KERNEL_END(1)
KERNEL_BEGIN(k_lop3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[c]) : "r"(y), "r"(z)); KERNEL_END(1)
KERNEL_BEGIN(k_shf) asm volatile("shf.l.wrap.b32 %0, %0, %1, 1;" : "+r"(x[c]) : "r"(y)); KERNEL_END(1)
KERNEL_BEGIN(k_prmt) asm volatile("prmt.b32 %0, %0, %1, 0x5410;" : "+r"(x[c]) : "r"(y)); KERNEL_END(1)
KERNEL_BEGIN(k_imad) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[c]) : "r"(y), "r"(z)); KERNEL_END(1)
KERNEL_BEGIN(k_imad_imm) asm volatile("mad.lo.u32 %0, %0, 32, %1;" : "+r"(x[c]) : "r"(y)); KERNEL_END(1)
KERNEL_BEGIN(k_ffma) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(*(float *) &x[c]) : "f"(__uint_as_float(y)), "f"(__uint_as_float(z))); KERNEL_END(1)
KERNEL_BEGIN(k_fadd) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(*(float *) &x[c]) : "f"(__uint_as_float(y))); KERNEL_END(1)
KERNEL_BEGIN(k_viadd16x2) asm volatile("add.u16x2 %0, %0, %1;" : "+r"(x[c]) : "r"(y)); KERNEL_END(1)
KERNEL_BEGIN(k_vimnmx16x2) asm volatile("max.u16x2 %0, %0, %1;" : "+r"(x[c]) : "r"(y)); KERNEL_END(1)
KERNEL_BEGIN(k_dp2a) asm volatile("dp2a.lo.u32.s32 %0, %1, %2, %0;" : "+r"(x[c]) : "r"(y), "r"(z)); KERNEL_END(1)
KERNEL_BEGIN(k_dp4a) asm volatile("dp4a.u32.s32 %0, %1, %2, %0;" : "+r"(x[c]) : "r"(y), "r"(z)); KERNEL_END(1)
KERNEL_BEGIN(k_hfma2) asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(x[c]) : "r"(y), "r"(z)); KERNEL_END(1)
KERNEL_BEGIN(k_hadd2) asm volatile("add.rn.f16x2 %0, %0, %1;" : "+r"(x[c]) : "r"(y)); KERNEL_END(1)
KERNEL_BEGIN(k_popc) asm volatile("{ .reg .u32 t; popc.b32 t, %0; add.u32 %0, %0, t; }" : "+r"(x[c])); KERNEL_END(2)
KERNEL_BEGIN(k_shfl) x[c] = __shfl_xor_sync(0xffffffffu, x[c], 1); KERNEL_END(1)
KERNEL_BEGIN(k_ballot) x[c] += __ballot_sync(0xffffffffu, x[c] & 1u); KERNEL_END(2)
KERNEL_BEGIN(k_viaddmax) x[c] = __viaddmax_s16x2(x[c], y, z); KERNEL_END(1)
KERNEL_BEGIN(k_vibmax) { bool ph, pl; x[c] = __vibmax_u16x2(x[c], y, &ph, &pl); if (ph) z ^= 1u; if (pl) y ^= 2u; } KERNEL_END(3)
// mixes: one ALU-pipe op + one FMA-pipe op per chain step
KERNEL_BEGIN(k_mix_iadd_imad) asm volatile("add.u32 %0, %0, %1;" : "+r"(x[c]) : "r"(y)); asm volatile("mad.lo.u32 %0, %0, 3, %1;" : "+r"(x[c]) : "r"(z)); KERNEL_END(2)
KERNEL_BEGIN(k_mix_lop3_ffma) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[c]) : "r"(y), "r"(z)); asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(*(float *) &x[c]) : "f"(__uint_as_float(y)), "f"(__uint_as_float(z))); KERNEL_END(2)
KERNEL_BEGIN(k_mix_viadd_imad) asm volatile("add.u16x2 %0, %0, %1;" : "+r"(x[c]) : "r"(y)); asm volatile("mad.lo.u32 %0, %0, 3, %1;" : "+r"(x[c]) : "r"(z)); KERNEL_END(2)
KERNEL_BEGIN(k_mix_iadd_shf) asm volatile("add.u32 %0, %0, %1;" : "+r"(x[c]) : "r"(y)); asm volatile("shf.l.wrap.b32 %0, %0, %1, 1;" : "+r"(x[c]) : "r"(z)); KERNEL_END(2)
KERNEL_BEGIN(k_mix_iadd_ffma_ffma) asm volatile("add.u32 %0, %0, %1;" : "+r"(x[c]) : "r"(y)); asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(*(float *) &x[c]) : "f"(__uint_as_float(y)), "f"(__uint_as_float(z))); asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(*(float *) &x[c]) : "f"(__uint_as_float(z)), "f"(__uint_as_float(y))); KERNEL_END(3)
KERNEL_BEGIN(k_mix_dp2a_iadd) asm volatile("dp2a.lo.u32.s32 %0, %1, %2, %0;" : "+r"(x[c]) : "r"(y), "r"(z)); asm volatile("add.u32 %0, %0, %1;" : "+r"(x[c]) : "r"(y)); KERNEL_END(2)
KERNEL_BEGIN(k_mix_dp2a_imad) asm volatile("dp2a.lo.u32.s32 %0, %1, %2, %0;" : "+r"(x[c]) : "r"(y), "r"(z)); asm volatile("mad.lo.u32 %0, %0, 3, %1;" : "+r"(x[c]) : "r"(z)); KERNEL_END(2)

// 64-bit multiply-accumulate (the block power sums)
__global__ void __launch_bounds__(1024, 1) k_imad_wide(uint32_t *out, long long *cyc, uint32_t seed) {
    unsigned long long x[CHAINS];
    uint32_t y = seed * 2654435761u + threadIdx.x * 40503u + 1u;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c)
        x[c] = y + c;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int c = 0; c < CHAINS; ++c)
            x[c] = (unsigned long long) (uint32_t) x[c] * y + x[c];
    }
    long long t1 = clock64();
    unsigned long long acc = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c)
        acc ^= x[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = (uint32_t) acc ^ (uint32_t) (acc >> 32);
    __syncthreads();
    if (threadIdx.x == 0)
        cyc[blockIdx.x] = t1 - t0;
}

// shared-memory loads: mode 0 = u16 conflict-free (lane-consecutive words), 1 = u16 random addresses over 128 KiB,
// 2 = u16 random over a 16x16-code noise neighbourhood of a 256x256 table, 3 = 32-bit conflict-free, 4 = 128-bit rows
// of 80 bytes (the scan ring), 5 = u16 from a per-lane replicated 256-entry table (word = e * 32 + lane)
template <int MODE>
__global__ void __launch_bounds__(1024, 1) k_lds(uint32_t *out, long long *cyc, uint32_t seed) {
    extern __shared__ __align__(16) uint32_t sm[];
    const int n_words = 128 * 1024 / 4;
    for (int i = threadIdx.x; i < n_words; i += blockDim.x)
        sm[i] = i * 2654435761u + seed;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    uint32_t x[CHAINS];
    uint32_t r = seed * 2654435761u + threadIdx.x * 40503u + 1u;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) {
        r = r * 1664525u + 1013904223u;
        x[c] = r >> 8;
    }
    uint32_t acc = 0;
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) {
            uint32_t v;
            if (MODE == 0) {
                const uint32_t a = ((x[c] & 0xffu) * 32u + lane) * 4u;
                asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a));
            } else if (MODE == 1) {
                const uint32_t a = (x[c] & 0xffffu) * 2u;
                asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a));
            } else if (MODE == 2) {
                // I in 120..135, Q in 120..135, swizzled as the scan kernel does
                uint32_t idx = (120u + (x[c] & 15u)) | ((120u + ((x[c] >> 4) & 15u)) << 8);
                idx ^= ((idx >> 8) & 31u) << 1;
                const uint32_t a = idx * 2u;
                asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a));
            } else if (MODE == 3) {
                const uint32_t a = ((x[c] & 0xffu) * 32u + lane) * 4u;
                asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
            } else if (MODE == 4) {
                const uint32_t a = (((x[c] & 0x3fu) + lane) & 63u) * 80u;
                uint32_t v1, v2, v3;
                asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v), "=r"(v1), "=r"(v2), "=r"(v3) : "r"(a));
                v ^= v1 ^ v2 ^ v3;
            } else {
                const uint32_t a = (x[c] & 0xffu) * 128u + lane * 4u;
                asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a));
            }
            acc += v;
            x[c] = x[c] * 1664525u + 1013904223u; // 1 IMAD per load: next address independent of the load
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    __syncthreads();
    if (threadIdx.x == 0)
        cyc[blockIdx.x] = t1 - t0;
}

typedef void (*kern_t)(uint32_t *, long long *, uint32_t);

static void run(const char *name, kern_t k, int nops, size_t smem, int sms, uint32_t *d_out, long long *d_cyc) {
    if (smem)
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
    for (int rep = 0; rep < 2; ++rep)
        k<<<sms, 1024, smem>>>(d_out, d_cyc, 12345u + rep);
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
    const double warp_instr = 32.0 * ITERS * CHAINS * nops; // per SM: 32 warps
    printf("%-24s %8.3f warp-instr/clk/SM  (%d op(s) per step, %.0f cycles)\n", name, warp_instr / cyc, nops, cyc);
    free(h);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    printf("device %s, %d SMs, clock %d kHz\n", p.name, sms, p.clockRate);
    uint32_t *d_out;
    long long *d_cyc;
    cudaMalloc(&d_out, (size_t) sms * 1024 * 4);
    cudaMalloc(&d_cyc, sms * sizeof(long long));
#define RUN(k, n) run(#k, k, n, 0, sms, d_out, d_cyc)
    RUN(k_iadd, 1); RUN(k_iadd3, 1); RUN(k_lop3, 1); RUN(k_shf, 1); RUN(k_prmt, 1); RUN(k_imad, 1); RUN(k_imad_imm, 1);
    RUN(k_imad_wide, 1); RUN(k_ffma, 1); RUN(k_fadd, 1); RUN(k_viadd16x2, 1); RUN(k_vimnmx16x2, 1); RUN(k_viaddmax, 1);
    RUN(k_vibmax, 3); RUN(k_dp2a, 1); RUN(k_dp4a, 1); RUN(k_hfma2, 1); RUN(k_hadd2, 1); RUN(k_popc, 2); RUN(k_shfl, 1); RUN(k_ballot, 2);
    RUN(k_mix_iadd_imad, 2); RUN(k_mix_lop3_ffma, 2); RUN(k_mix_viadd_imad, 2); RUN(k_mix_iadd_shf, 2); RUN(k_mix_iadd_ffma_ffma, 3);
    RUN(k_mix_dp2a_iadd, 2); RUN(k_mix_dp2a_imad, 2);
    // LDS probes count the load only (each step also issues one IMAD for the next address)
    run("lds_u16_conflict_free", k_lds<0>, 1, 128 * 1024, sms, d_out, d_cyc);
    run("lds_u16_random_128K", k_lds<1>, 1, 128 * 1024, sms, d_out, d_cyc);
    run("lds_u16_noise16x16_swz", k_lds<2>, 1, 128 * 1024, sms, d_out, d_cyc);
    run("lds_u32_conflict_free", k_lds<3>, 1, 128 * 1024, sms, d_out, d_cyc);
    run("lds_v4_rows80", k_lds<4>, 1, 128 * 1024, sms, d_out, d_cyc);
    run("lds_u16_lane_replica", k_lds<5>, 1, 128 * 1024, sms, d_out, d_cyc);
    return 0;
}
