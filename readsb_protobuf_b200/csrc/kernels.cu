// kernels.cu -- sm_100a kernels of the Mode S demodulator.
//
//   K1  scan_kernel      IQ -> magnitude (never leaves the SM) -> preamble scan -> PPM slice ->
//                        CRC-24 syndrome + error-table lookup -> class records
//                        replaces convert.c:63-111/215-253/332-370, demod_2400.c:98-229,257-335,
//                        crc.c:67-82,389-412 and the filter-independent half of mode_s.c:311-409
//   K2  classify_kernel  address-set test, ordered dead/live lists, re-slice + signal power of the
//                        survivors (demod_2400.c:387-399)
//   convert_kernel       IQ -> u16 magnitudes in global memory (the iq_convert_fn boundary)
//   crc_batch_kernel     CRC + diagnose for a batch of frames (the crc.h boundary)
//
// Everything here is integer/byte work bounded by HBM bandwidth or instruction issue; there is
// no dense contraction, so no tensor-core path.

#include "kernels.cuh"

#include <cuda_runtime.h>

namespace b200 {

// ------------------------------------------------------------------------------------------
// constants
// ------------------------------------------------------------------------------------------

// crc.c:59-64: syndrome of a single flipped bit, indexed from the start of a 112-bit frame
__constant__ uint32_t c_bit_syndrome[112];

cudaError_t upload_constants(const uint32_t *bit_syndromes112) {
    return cudaMemcpyToSymbol(c_bit_syndrome, bit_syndromes112, 112 * sizeof(uint32_t));
}

// demod_2400.c:73-93: the five correlators, taps for m[0..3]
__device__ __constant__ int c_slice_coef[5][4] = {
    {18, -15, -3, 0}, {14, -5, -9, 0}, {16, 5, -20, 0}, {7, 11, -18, 0}, {4, 15, -20, 1},
};

// ------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------

__device__ __forceinline__ uint4 ldg_stream(const uint4 *p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}

// float path of convert.c:231-243 / 348-360, rounding step by step like the scalar C code
// (the reference is built without FMA contraction: Makefile:12-13)
__device__ __forceinline__ uint32_t mag_from_float(float fI, float fQ, float &magsq_out, float &mag_out) {
    float magsq = __fadd_rn(__fmul_rn(fI, fI), __fmul_rn(fQ, fQ));
    if (magsq > 1.0f)
        magsq = 1.0f;
    float mag = __fsqrt_rn(magsq);
    magsq_out = magsq;
    mag_out = mag;
    return __float2uint_rz(__fadd_rn(__fmul_rn(mag, 65535.0f), 0.5f));
}

__device__ __forceinline__ uint32_t mag_sc16_word(uint32_t w, float inv_scale, float &magsq, float &mag) {
    // little-endian int16 pair: I in the low half (convert.c:231-232)
    float fI = __fmul_rn((float) (int16_t) (w & 0xffff), inv_scale); // division by 2^k == exact scaling
    float fQ = __fmul_rn((float) (int16_t) (w >> 16), inv_scale);
    return mag_from_float(fI, fQ, magsq, mag);
}

__device__ __forceinline__ unsigned long long warp_sum_u64(unsigned long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
        v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ double warp_sum_f64(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
        v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// exclusive prefix sum of one int per thread over the whole CTA; returns the total through `total`.
// s_warp must hold blockDim.x/32 + 1 ints.  Contains two __syncthreads().
__device__ __forceinline__ int block_exclusive_scan(int v, int *s_warp, int &total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o)
            inc += t;
    }
    __syncthreads(); // previous users of s_warp are done
    if (lane == 31)
        s_warp[warp] = inc;
    __syncthreads();
    int wsum = (lane < nwarps) ? s_warp[lane] : 0;
    int winc = wsum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, winc, o);
        if (lane >= o)
            winc += t;
    }
    total = __shfl_sync(0xffffffffu, winc, nwarps - 1);
    int wexc = __shfl_sync(0xffffffffu, winc - wsum, warp);
    return wexc + inc - v;
}

// crc.c:389-412 on the device: binary search of the sorted syndrome table (signed compare like
// syndrome_compare, crc.c:92-96; syndromes are 24-bit so the sign never matters)
__device__ __forceinline__ int find_syndrome(const ErrorInfo *__restrict__ tab, int n, uint32_t syndrome) {
    int lo = 0, hi = n - 1;
    while (lo <= hi) {
        int mid = (lo + hi) >> 1;
        uint32_t s = __ldg(&tab[mid].syndrome);
        if (s == syndrome)
            return mid;
        if (s < syndrome)
            lo = mid + 1;
        else
            hi = mid - 1;
    }
    return -1;
}

// mode_s.c:266-281: flip the address bits a repair touches
__device__ __forceinline__ uint32_t correct_aa(uint32_t addr, int errors, int b0, int b1) {
    if (errors >= 1 && b0 >= 8 && b0 <= 31)
        addr ^= 1u << (31 - b0);
    if (errors >= 2 && b1 >= 8 && b1 <= 31)
        addr ^= 1u << (31 - b1);
    return addr;
}

// The filter-independent half of scoreModesMessage (mode_s.c:311-409) for a sliced frame given as
// four ballot words (frame bit b is bit b%32 of w[b/32]) and its CRC syndrome.
// Returns the class; key = the address the filter will be asked about.
struct FrameClass {
    uint32_t kind, errors, key;
    int bit0, bit1;
};

__device__ __forceinline__ FrameClass classify_frame(uint32_t df, uint32_t aa, uint32_t syn, bool all_zero,
                                                     const ErrorInfo *__restrict__ tab_short, int n_short,
                                                     const ErrorInfo *__restrict__ tab_long, int n_long) {
    FrameClass fc;
    fc.kind = kKindBad;
    fc.errors = 0;
    fc.key = syn;
    fc.bit0 = fc.bit1 = -1;
    if (all_zero) // mode_s.c:325-326
        return fc;
    switch (df) {
        case 0: case 4: case 5: case 16: case 24: // mode_s.c:331-343 (DF25-31 never get here: demod_2400.c:193-205)
            fc.kind = kKindAP;
            break;
        case 20: case 21: // mode_s.c:391-403
            fc.kind = kKindAPCommB;
            break;
        case 11: { // mode_s.c:345-374
            uint32_t c2 = syn & 0xffff80u;
            if (c2 != 0) {
                int idx = find_syndrome(tab_short, n_short, c2);
                if (idx < 0)
                    return fc;
                int errors = tab_short[idx].errors;
                if (errors > 1)
                    return fc;
                fc.errors = (uint32_t) errors;
                fc.bit0 = tab_short[idx].bit[0];
                fc.bit1 = tab_short[idx].bit[1];
            }
            fc.kind = kKindDF11;
            fc.key = correct_aa(aa, (int) fc.errors, fc.bit0, fc.bit1);
            break;
        }
        case 17: case 18: { // mode_s.c:376-389
            if (syn != 0) {
                int idx = find_syndrome(tab_long, n_long, syn);
                if (idx < 0)
                    return fc;
                fc.errors = (uint32_t) tab_long[idx].errors;
                fc.bit0 = tab_long[idx].bit[0];
                fc.bit1 = tab_long[idx].bit[1];
            }
            fc.kind = kKindES;
            fc.key = correct_aa(aa, (int) fc.errors, fc.bit0, fc.bit1);
            break;
        }
        default:
            break;
    }
    return fc;
}

// demod_2400.c:193-205: frame length in bytes from the DF of the first sliced byte, 0 = give up
__device__ __forceinline__ int frame_bytes_for_df(uint32_t df) {
    // DF 0,4,5,11 -> 7 ; DF 16,17,18,20,21,24 -> 14
    const uint32_t short_set = (1u << 0) | (1u << 4) | (1u << 5) | (1u << 11);
    const uint32_t long_set = (1u << 16) | (1u << 17) | (1u << 18) | (1u << 20) | (1u << 21) | (1u << 24);
    if ((short_set >> df) & 1u)
        return 7;
    if ((long_set >> df) & 1u)
        return 14;
    return 0;
}

// One PPM bit decision (demod_2400.c:73-177 in closed form): frame bit b of a candidate whose
// preamble window starts at m[0], tried at phase try_phase, sits t = try_phase + 12*b fifths of a
// sample after m[19]; correlator t%5 over the four samples from m[19 + t/5].
__device__ __forceinline__ bool slice_bit(const uint16_t *m, int try_phase, int b, const int (*coef)[4]) {
    int t = try_phase + 12 * b;
    int s = t / 5;
    int r = t - 5 * s;
    const uint16_t *p = m + 19 + s;
    int v = coef[r][0] * (int) p[0] + coef[r][1] * (int) p[1] + coef[r][2] * (int) p[2] + coef[r][3] * (int) p[3];
    return v > 0;
}

// Warp-cooperative slice of a whole frame: lane l decides bits l, l+32, l+64, l+96; the ballots
// are the packed message (bit b of the frame = bit b%32 of w[b/32]).  Also returns the CRC
// syndrome (crc.c:67-82, by linearity the XOR of the single-bit syndromes of the set bits).
__device__ __forceinline__ void warp_slice_frame(const uint16_t *m, int try_phase, int nbits, const int (*coef)[4],
                                                 const uint32_t *s_syn, uint32_t w[4], uint32_t &syndrome) {
    const int lane = threadIdx.x & 31;
    const int off = 112 - nbits; // crc.c:143: short frames use the tail of the 112-bit syndrome list
    uint32_t x = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        int b = lane + 32 * k;
        bool bit = false;
        if (b < nbits)
            bit = slice_bit(m, try_phase, b, coef);
        w[k] = __ballot_sync(0xffffffffu, bit);
        if (bit)
            x ^= s_syn[b + off];
    }
    syndrome = __reduce_xor_sync(0xffffffffu, x);
}

// ------------------------------------------------------------------------------------------
// K1: scan kernel
// ------------------------------------------------------------------------------------------

struct ScanSmem {
    // layout computed by scan_smem_layout()
    uint16_t *lut;    // 65536 (uc8 only)
    uint16_t *mag;    // kTileSamples (+8 pad)
    uint32_t *cand;   // kMaxCand
    uint16_t *itemoff; // kMaxCand
    uint32_t *items;  // kMaxItems
    uint2 *res;       // kMaxItems
    uint32_t *syn;    // 112
    int (*coef)[4];   // 5
    int *warp;        // 32
    unsigned long long *red; // 64
};

constexpr size_t kSmemMagBytes = (kTileSamples + 8) * sizeof(uint16_t);
constexpr size_t kSmemCommon = kSmemMagBytes + kMaxCand * 4 + kMaxCand * 2 + kMaxItems * 4 + kMaxItems * 8 +
                               112 * 4 + 5 * 4 * 4 + 32 * 4 + 64 * 8 + 64;

size_t scan_smem_bytes(uint32_t format) {
    return kSmemCommon + (format == 0 ? 65536 * sizeof(uint16_t) : 0);
}

__device__ __forceinline__ ScanSmem scan_smem_layout(unsigned char *base, bool with_lut) {
    ScanSmem s;
    size_t o = 0;
    s.lut = reinterpret_cast<uint16_t *>(base);
    if (with_lut)
        o += 65536 * sizeof(uint16_t);
    s.mag = reinterpret_cast<uint16_t *>(base + o);
    o += kSmemMagBytes;
    s.res = reinterpret_cast<uint2 *>(base + o);
    o += kMaxItems * 8;
    s.red = reinterpret_cast<unsigned long long *>(base + o);
    o += 64 * 8;
    s.cand = reinterpret_cast<uint32_t *>(base + o);
    o += kMaxCand * 4;
    s.items = reinterpret_cast<uint32_t *>(base + o);
    o += kMaxItems * 4;
    s.syn = reinterpret_cast<uint32_t *>(base + o);
    o += 112 * 4;
    s.coef = reinterpret_cast<int(*)[4]>(base + o);
    o += 5 * 4 * 4;
    s.warp = reinterpret_cast<int *>(base + o);
    o += 32 * 4;
    s.itemoff = reinterpret_cast<uint16_t *>(base + o);
    return s;
}

template <int FORMAT>
struct Fmt;
template <>
struct Fmt<0> { // uc8: 2 bytes per sample, 8 samples per 16-byte chunk
    static constexpr int kBytes = 2, kChunkSamples = 8;
};
template <>
struct Fmt<1> { // sc16
    static constexpr int kBytes = 4, kChunkSamples = 4;
};
template <>
struct Fmt<2> { // sc16q11
    static constexpr int kBytes = 4, kChunkSamples = 4;
};

// Load the 16-byte chunk holding samples [s, s + kChunkSamples) of the span (s relative to the first
// new sample; negative = carried head).  lo/hi = the valid sample range inside the chunk.
template <int FORMAT>
__device__ __forceinline__ uint4 load_chunk(const ScanArgs &a, long long s, int &lo, int &hi) {
    constexpr int CS = Fmt<FORMAT>::kChunkSamples, BPS = Fmt<FORMAT>::kBytes;
    const long long n = (long long) a.nsamples;
    long long first_valid = -(long long) a.head_valid;
    long long l = first_valid - s, h = n - s;
    lo = l < 0 ? 0 : (l > CS ? CS : (int) l);
    hi = h > CS ? CS : (h < 0 ? 0 : (int) h);
    uint4 v = make_uint4(0, 0, 0, 0);
    if (hi <= lo)
        return v;
    if (s < 0) { // carried head: always fully addressable
        return ldg_stream(reinterpret_cast<const uint4 *>(a.head + (s + kHead) * BPS));
    }
    if (hi == CS)
        return ldg_stream(reinterpret_cast<const uint4 *>(a.iq + s * BPS));
    // ragged end of the span: never read past the caller's buffer
    uint32_t w[4] = {0, 0, 0, 0};
    const uint8_t *p = a.iq + s * BPS;
    for (int k = 0; k < hi * BPS; ++k)
        w[k >> 2] |= (uint32_t) p[k] << (8 * (k & 3));
    return make_uint4(w[0], w[1], w[2], w[3]);
}

template <int FORMAT, bool SLICE>
__global__ void __launch_bounds__(kScanThreads, 1) scan_kernel(const ScanArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const ScanSmem sm = scan_smem_layout(smem_raw, FORMAT == 0);
    constexpr int CS = Fmt<FORMAT>::kChunkSamples;
    constexpr int kChunks = kTileSamples / CS;
    constexpr int kChunksPerThread = (kChunks + kScanThreads - 1) / kScanThreads;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float inv_scale = (FORMAT == 1) ? (1.0f / 32768.0f) : (1.0f / 2048.0f);

    // one-time staging of the tables
    if (FORMAT == 0) {
        const uint4 *src = reinterpret_cast<const uint4 *>(a.lut);
        uint4 *dst = reinterpret_cast<uint4 *>(sm.lut);
        for (int i = tid; i < 65536 * 2 / 16; i += kScanThreads)
            dst[i] = __ldg(src + i);
    }
    if (tid < 112)
        sm.syn[tid] = c_bit_syndrome[tid];
    if (tid < 20)
        (&sm.coef[0][0])[tid] = (&c_slice_coef[0][0])[tid];
    __syncthreads();

    const long long n = (long long) a.nsamples;
    uint4 pre[kChunksPerThread];
    int pre_lo[kChunksPerThread], pre_hi[kChunksPerThread];

    auto prefetch = [&](uint32_t tile) {
        const long long s0 = (long long) tile * kTile - kHead;
#pragma unroll
        for (int k = 0; k < kChunksPerThread; ++k) {
            int c = tid + k * kScanThreads;
            pre_lo[k] = pre_hi[k] = 0;
            pre[k] = make_uint4(0, 0, 0, 0);
            if (c < kChunks)
                pre[k] = load_chunk<FORMAT>(a, s0 + (long long) c * CS, pre_lo[k], pre_hi[k]);
        }
    };

    uint32_t tile = blockIdx.x;
    if (tile < a.ntiles)
        prefetch(tile);

    for (; tile < a.ntiles; tile += gridDim.x) {
        const long long p0 = (long long) tile * kTile;

        // ---------------- phase A: IQ -> magnitudes of the tile, block sums ----------------
        unsigned long long sum_level = 0, sum_power = 0;
        double fsum_level = 0, fsum_power = 0;
        const uint32_t kb0 = (uint32_t) (p0 / a.block_samples);
        const bool one_block = ((long long) (kb0 + 1) * a.block_samples >= p0 + kTile);
#pragma unroll
        for (int k = 0; k < kChunksPerThread; ++k) {
            int c = tid + k * kScanThreads;
            if (c >= kChunks)
                continue;
            const uint4 raw = pre[k];
            const int lo = pre_lo[k], hi = pre_hi[k];
            uint32_t words[4] = {raw.x, raw.y, raw.z, raw.w};
            uint32_t m[CS];
            unsigned long long cl = 0, cp = 0;
            double fl = 0, fp = 0;
            if (FORMAT == 0) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    m[2 * j] = sm.lut[words[j] & 0xffffu];
                    m[2 * j + 1] = sm.lut[words[j] >> 16];
                }
#pragma unroll
                for (int j = 0; j < CS; ++j) {
                    if (j < lo || j >= hi)
                        m[j] = 0;
                    cl += m[j];
                    cp += (unsigned long long) m[j] * m[j];
                }
            } else {
#pragma unroll
                for (int j = 0; j < CS; ++j) {
                    float magsq, mag;
                    m[j] = mag_sc16_word(words[j], inv_scale, magsq, mag);
                    if (j < lo || j >= hi) {
                        m[j] = 0;
                    } else {
                        fl += (double) mag;
                        fp += (double) magsq;
                    }
                }
            }
            if (CS == 8) {
                uint4 packed = make_uint4(m[0] | (m[1] << 16), m[2] | (m[3] << 16), m[4 % CS] | (m[5 % CS] << 16),
                                          m[6 % CS] | (m[7 % CS] << 16));
                *reinterpret_cast<uint4 *>(sm.mag + c * CS) = packed;
            } else {
                uint2 packed = make_uint2(m[0] | (m[1] << 16), m[2] | (m[3] << 16));
                *reinterpret_cast<uint2 *>(sm.mag + c * CS) = packed;
            }
            // this tile owns the sums of samples [p0, p0 + kTile)
            const int q = c * CS;
            if (q >= kHead && q < kHead + kTile) {
                if (one_block) {
                    sum_level += cl;
                    sum_power += cp;
                    fsum_level += fl;
                    fsum_power += fp;
                } else if (hi > lo) {
                    const uint32_t kb = (uint32_t) ((p0 + q - kHead) / a.block_samples);
                    if (FORMAT == 0) {
                        atomicAdd(&a.block_sums_u64[2 * kb], cl);
                        atomicAdd(&a.block_sums_u64[2 * kb + 1], cp);
                    } else {
                        atomicAdd(&a.block_sums_f64[2 * kb], fl);
                        atomicAdd(&a.block_sums_f64[2 * kb + 1], fp);
                    }
                }
            }
        }
        if (one_block) {
            if (FORMAT == 0) {
                sum_level = warp_sum_u64(sum_level);
                sum_power = warp_sum_u64(sum_power);
                if (lane == 0) {
                    sm.red[2 * warp] = sum_level;
                    sm.red[2 * warp + 1] = sum_power;
                }
            } else {
                fsum_level = warp_sum_f64(fsum_level);
                fsum_power = warp_sum_f64(fsum_power);
                if (lane == 0) {
                    reinterpret_cast<double *>(sm.red)[2 * warp] = fsum_level;
                    reinterpret_cast<double *>(sm.red)[2 * warp + 1] = fsum_power;
                }
            }
        }

        // next tile's IQ is requested now and consumed after this tile's scan/slice work
        if (tile + gridDim.x < a.ntiles)
            prefetch(tile + gridDim.x);

        __syncthreads();
        if (one_block && warp == 0) {
            constexpr int NW = kScanThreads / 32;
            if (FORMAT == 0) {
                unsigned long long l = (lane < NW) ? sm.red[2 * lane] : 0, p = (lane < NW) ? sm.red[2 * lane + 1] : 0;
                l = warp_sum_u64(l);
                p = warp_sum_u64(p);
                if (lane == 0 && (l | p)) {
                    atomicAdd(&a.block_sums_u64[2 * kb0], l);
                    atomicAdd(&a.block_sums_u64[2 * kb0 + 1], p);
                }
            } else {
                const double *red = reinterpret_cast<const double *>(sm.red);
                double l = (lane < NW) ? red[2 * lane] : 0, p = (lane < NW) ? red[2 * lane + 1] : 0;
                l = warp_sum_f64(l);
                p = warp_sum_f64(p);
                if (lane == 0) {
                    atomicAdd(&a.block_sums_f64[2 * kb0], l);
                    atomicAdd(&a.block_sums_f64[2 * kb0 + 1], p);
                }
            }
        }

        // ---------------- phase B: preamble scan, kPosPerThread positions per thread ----------------
        // position i of the tile has its window at mag[i + 2 ...] (mag[0] is sample p0 - kHead,
        // the window of position p starts kOverlap samples before sample p)
        uint32_t masks[kPosPerThread]; // 5-bit try masks
        int ncand_mine = 0;
        {
            const int i0 = tid * kPosPerThread;
            uint32_t w[kPosPerThread + 18];
#pragma unroll
            for (int x = 0; x < kPosPerThread + 18; ++x)
                w[x] = sm.mag[i0 + 2 + x];
            const int thr = a.threshold;
#pragma unroll
            for (int i = 0; i < kPosPerThread; ++i) {
                const uint32_t *pa = &w[i];
                uint32_t mask = 0;
                // demod_2400.c:276
                if (pa[1] > pa[7] && pa[12] > pa[14] && pa[12] > pa[15]) {
                    // demod_2400.c:281-292
                    int base_noise = (int) (pa[5] + pa[8] + pa[16] + pa[17] + pa[18]);
                    int ref_level = (base_noise * thr) >> 5;
                    // demod_2400.c:298-301
                    int diff_2_3 = (int) pa[2] - (int) pa[3];
                    int sum_1_4 = (int) pa[1] + (int) pa[4];
                    int diff_10_11 = (int) pa[10] - (int) pa[11];
                    int common3456 = sum_1_4 - diff_2_3 + (int) pa[9] + (int) pa[12];
                    if (common3456 - diff_10_11 >= ref_level) // :306-312 -> phases 4, 5
                        mask |= 0x03;
                    if (common3456 + diff_10_11 >= ref_level) // :316-322 -> phases 6, 7
                        mask |= 0x0c;
                    if (sum_1_4 + 2 * diff_2_3 + diff_10_11 + (int) pa[12] >= ref_level) // :327-330 -> phase 8
                        mask |= 0x10;
                }
                if (p0 + i0 + i >= n)
                    mask = 0;
                masks[i] = mask;
                ncand_mine += (mask != 0);
                if (a.dbg_masks && p0 + i0 + i < n)
                    a.dbg_masks[p0 + i0 + i] = (uint8_t) mask;
            }
        }

        int ncand_tile;
        int my_off = block_exclusive_scan(ncand_mine, sm.warp, ncand_tile);

        if (!SLICE) {
            if (tid == 0 && ncand_tile)
                atomicAdd(&a.counters->n_cand, (unsigned long long) ncand_tile);
            __syncthreads();
            continue;
        }

        // ---------------- phases C/D in rounds that cannot overflow the shared lists ----------------
        const int nrounds = (ncand_tile <= kMaxCand) ? 1 : kSlowRounds;
        if (tid == 0 && nrounds > 1)
            atomicAdd(&a.counters->slow_tiles, 1u);
        uint32_t tile_cand_off = 0, tile_rec_off = 0, tile_ncand = 0, tile_nrec = 0;
        bool tile_ovf = false;

        for (int round = 0; round < nrounds; ++round) {
            int ncand = ncand_tile, off = my_off;
            bool mine = true;
            if (nrounds > 1) {
                // round r takes the kMaxCand positions [r*kMaxCand, (r+1)*kMaxCand)
                mine = (tid * kPosPerThread) / kMaxCand == round;
                off = block_exclusive_scan(mine ? ncand_mine : 0, sm.warp, ncand);
            }
            if (mine) {
#pragma unroll
                for (int i = 0; i < kPosPerThread; ++i)
                    if (masks[i])
                        sm.cand[off++] = (uint32_t) (tid * kPosPerThread + i) | (masks[i] << 13);
            }
            __syncthreads();

            // ---- C1: first byte of every tried phase -> DF -> frame length (demod_2400.c:188-205) ----
            int nitems = 0;
            for (int cb = 0; cb < ncand; cb += kScanThreads) { // uniform trip count
                const int c = cb + tid;
                uint32_t my_items[5];
                int my_n = 0;
                if (c < ncand) {
                    const uint32_t e = sm.cand[c];
                    const uint16_t *m = sm.mag + (e & 0x1fffu) + 2;
                    const uint32_t tm = (e >> 13) & 31u;
#pragma unroll
                    for (int ph = 0; ph < 5; ++ph) {
                        if (!((tm >> ph) & 1u))
                            continue;
                        uint32_t byte0 = 0;
#pragma unroll
                        for (int b = 0; b < 8; ++b)
                            byte0 = (byte0 << 1) | (slice_bit(m, ph + 4, b, sm.coef) ? 1u : 0u);
                        int nb = frame_bytes_for_df(byte0 >> 3);
                        if (nb)
                            my_items[my_n++] = (uint32_t) c | ((uint32_t) ph << 10) | ((nb == 14) ? (1u << 13) : 0u);
                    }
                }
                int tot;
                int ioff = block_exclusive_scan(my_n, sm.warp, tot);
                if (c < ncand)
                    sm.itemoff[c] = (uint16_t) (nitems + ioff);
                for (int k = 0; k < my_n; ++k)
                    sm.items[nitems + ioff + k] = my_items[k];
                nitems += tot;
            }
            __syncthreads();

            // ---- C2: one warp per (candidate, phase): slice, CRC, class ----
            for (int it = warp; it < nitems; it += kScanThreads / 32) {
                const uint32_t item = sm.items[it];
                const uint32_t e = sm.cand[item & 1023u];
                const uint16_t *m = sm.mag + (e & 0x1fffu) + 2;
                const int ph = (int) ((item >> 10) & 7u) + 4;
                const int nbits = (item & (1u << 13)) ? 112 : 56;
                uint32_t w[4], syn;
                warp_slice_frame(m, ph, nbits, sm.coef, sm.syn, w, syn);
                const uint32_t head32 = __brev(w[0]); // frame bits 0..31, MSB first
                const FrameClass fc = classify_frame(head32 >> 27, head32 & 0xffffffu, syn, (w[0] | w[1] | w[2] | w[3]) == 0,
                                                     a.tab_short, a.n_short, a.tab_long, a.n_long);
                if (lane == 0) {
                    sm.res[it] = make_uint2(syn | (fc.kind << 24) | (fc.errors << 28), fc.key | ((uint32_t) ph << 24));
                    // mode_s.c:717-726: only a clean DF17, or a clean DF11 with IID 0, can ever be
                    // added to the ICAO filter; remember every such address of the stream
                    const uint32_t df = head32 >> 27;
                    if (fc.kind != kKindBad && syn == 0 && (df == 17 || df == 11)) {
                        const uint32_t aa = head32 & 0xffffffu;
                        atomicOr(&a.addr_bitmap[aa >> 5], 1u << (aa & 31u));
                    }
                }
            }
            __syncthreads();

            // ---- D: ordered write-out of candidate entries and class records ----
            int nrec = 0;
            {
                // count the records first (items are ordered by candidate, then phase)
                int mine_n = 0;
                for (int it = tid; it < nitems; it += kScanThreads)
                    mine_n += ((sm.res[it].x >> 24) & 7u) != kKindBad;
                block_exclusive_scan(mine_n, sm.warp, nrec);
                // The tile's entries and records must be contiguous for K2.  A one-round tile reserves
                // exactly what it has; a slow tile reserves its worst case (5 records per candidate)
                // once, in round 0, and fills it round after round.
                if (round == 0) {
                    if (tid == 0) {
                        const unsigned long long want_c = (unsigned long long) ncand_tile;
                        const unsigned long long want_r = (nrounds == 1) ? (unsigned long long) nrec : 5ull * want_c;
                        unsigned long long co = atomicAdd(&a.counters->n_cand, want_c);
                        unsigned long long ro = atomicAdd(&a.counters->n_rec, want_r);
                        unsigned int ovf = 0;
                        if (co + want_c > a.cand_cap)
                            ovf |= 1u;
                        if (ro + want_r > a.rec_cap)
                            ovf |= 2u;
                        if (ovf)
                            atomicOr(&a.counters->overflow, ovf);
                        sm.warp[20] = (int) (uint32_t) co;
                        sm.warp[21] = (int) (uint32_t) ro;
                        sm.warp[22] = (int) ovf;
                    }
                    __syncthreads();
                    tile_cand_off = (uint32_t) sm.warp[20];
                    tile_rec_off = (uint32_t) sm.warp[21];
                    tile_ovf = sm.warp[22] != 0;
                }
                const uint32_t cand_base = tile_cand_off + tile_ncand;
                const uint32_t rec_base = tile_rec_off + tile_nrec;
                tile_ncand += (uint32_t) ncand;
                tile_nrec += (uint32_t) nrec;

                if (!tile_ovf) {
                    // records: ordered compaction over the item list, kScanThreads items at a time
                    int done = 0;
                    for (int ib = 0; ib < nitems; ib += kScanThreads) {
                        const int it = ib + tid;
                        uint2 r = make_uint2(0, 0);
                        bool keep = false;
                        if (it < nitems) {
                            r = sm.res[it];
                            keep = ((r.x >> 24) & 7u) != kKindBad;
                        }
                        int tot;
                        int o = block_exclusive_scan(keep ? 1 : 0, sm.warp, tot);
                        if (keep) {
                            const uint32_t c = sm.items[it] & 1023u;
                            PhaseRec pr;
                            pr.pos = (uint32_t) (p0 + (sm.cand[c] & 0x1fffu));
                            pr.w0 = r.x;
                            pr.w1 = r.y;
                            pr.cand = cand_base + c;
                            a.recs[rec_base + done + o] = pr;
                        }
                        done += tot;
                    }
                    // candidate entries, with the count of records each owns
                    for (int c = tid; c < ncand; c += kScanThreads) {
                        const int i0 = sm.itemoff[c];
                        const int i1 = (c + 1 < ncand) ? sm.itemoff[c + 1] : nitems;
                        uint32_t nonbad = 0;
                        for (int it = i0; it < i1; ++it)
                            nonbad += ((sm.res[it].x >> 24) & 7u) != kKindBad;
                        a.cand[cand_base + c] = sm.cand[c] | (nonbad << 18);
                    }
                }
            }
            __syncthreads();
        }

        if (tid == 0) {
            TileDesc td;
            td.cand_off = tile_cand_off;
            td.ncand = tile_ncand;
            td.rec_off = tile_rec_off;
            td.nrec = tile_nrec;
            a.tiles[tile] = td;
        }
    }
}

cudaError_t scan_configure() {
    cudaError_t e;
#define CFG(F, S)                                                                                                  \
    e = cudaFuncSetAttribute(scan_kernel<F, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) scan_smem_bytes(F)); \
    if (e != cudaSuccess)                                                                                          \
        return e;
    CFG(0, true) CFG(0, false) CFG(1, true) CFG(1, false) CFG(2, true) CFG(2, false)
#undef CFG
    return cudaSuccess;
}

cudaError_t launch_scan(const ScanArgs &a, int mode, int grid, cudaStream_t stream) {
    if (a.ntiles == 0)
        return cudaSuccess;
    if (grid > (int) a.ntiles)
        grid = (int) a.ntiles;
    const size_t smem = scan_smem_bytes(a.format);
#define LAUNCH(F)                                                              \
    if (mode)                                                                  \
        scan_kernel<F, true><<<grid, kScanThreads, smem, stream>>>(a);         \
    else                                                                       \
        scan_kernel<F, false><<<grid, kScanThreads, smem, stream>>>(a);
    switch (a.format) {
        case 0: LAUNCH(0) break;
        case 1: LAUNCH(1) break;
        case 2: LAUNCH(2) break;
        default: return cudaErrorInvalidValue;
    }
#undef LAUNCH
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// K2: classify kernel -- one CTA per tile
// ------------------------------------------------------------------------------------------

constexpr int kClassifyThreads = 256;
constexpr int kFrameSamples = 296; // samples a frame's slice + power can touch: m[0..290]

__device__ __forceinline__ bool bitmap_test(const uint32_t *__restrict__ bm, uint32_t addr) {
    return (__ldg(&bm[(addr & 0xffffffu) >> 5]) >> (addr & 31u)) & 1u;
}

// can this class record ever score >= 0?  (mode_s.c:343,364-374,386-389,393)
__device__ __forceinline__ bool record_is_live(uint32_t w0, uint32_t w1, const uint32_t *__restrict__ bm) {
    const uint32_t kind = (w0 >> 24) & 7u;
    if (kind == kKindES)
        return true;
    if (kind == kKindDF11 && (w0 & 0x7fu) == 0)
        return true; // IID 0 scores 750/375 even for an unknown address
    return bitmap_test(bm, w1 & 0xffffffu);
}

// magnitude of span sample s (relative to the first new sample), 0 outside the stream
__device__ __forceinline__ uint32_t sample_mag(const ClassifyArgs &a, long long s) {
    if (s < -(long long) a.head_valid || s >= (long long) a.nsamples)
        return 0;
    const uint8_t *base = (s < 0) ? a.head + (s + kHead) * (a.format == 0 ? 2 : 4) : a.iq + s * (a.format == 0 ? 2 : 4);
    if (a.format == 0) {
        uint32_t idx = (uint32_t) base[0] | ((uint32_t) base[1] << 8);
        return __ldg(&a.lut[idx]);
    }
    uint32_t w = *reinterpret_cast<const uint32_t *>(base);
    float magsq, mag;
    return mag_sc16_word(w, (a.format == 1) ? (1.0f / 32768.0f) : (1.0f / 2048.0f), magsq, mag);
}

__global__ void __launch_bounds__(kClassifyThreads) classify_kernel(const ClassifyArgs a) {
    __shared__ __align__(4) uint8_t s_flags[kTile]; // per candidate of the tile: bit0 live, bit1 has a -1 phase
    __shared__ uint16_t s_slot[kTile];               // first live-record slot of a live candidate
    __shared__ int s_warp[40];
    __shared__ uint32_t s_syn[112];
    __shared__ int s_coef[5][4];
    __shared__ __align__(16) uint16_t s_frame[kClassifyThreads / 32][kFrameSamples];
    __shared__ uint32_t s_bd[8];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t tile = blockIdx.x;
    if (a.counters->overflow & 3u)
        return; // K1 ran out of room: the host grows the buffers and runs the span again
    const TileDesc td = a.tiles[tile];
    const long long p0 = (long long) tile * kTile;

    if (tid < 112)
        s_syn[tid] = c_bit_syndrome[tid];
    if (tid < 20)
        (&s_coef[0][0])[tid] = (&c_slice_coef[0][0])[tid];
    if (tid < 8)
        s_bd[tid] = 0;
    for (uint32_t c = tid; c < td.ncand; c += kClassifyThreads)
        s_flags[c] = 0;
    __syncthreads();

    // ---- pass 1: which candidates have a phase that can still score >= 0, which have a -1 phase ----
    for (uint32_t r = tid; r < td.nrec; r += kClassifyThreads) {
        const PhaseRec pr = a.recs[td.rec_off + r];
        const uint32_t kind = (pr.w0 >> 24) & 7u;
        const bool live = record_is_live(pr.w0, pr.w1, a.addr_bitmap);
        uint32_t f = live ? 1u : 0u;
        // static score of a phase whose address can never be in the filter:
        // AP -> -1, DF11 with IID != 0 -> -1, Comm-B -> -2 (mode_s.c:343,373,403)
        if (!live && (kind == kKindAP || kind == kKindDF11))
            f |= 2u;
        if (f) {
            const uint32_t c = pr.cand - td.cand_off;
            // byte-wide atomic OR through the containing word
            atomicOr(reinterpret_cast<unsigned int *>(s_flags) + (c >> 2), f << (8 * (c & 3)));
        }
    }
    __syncthreads();

    // ---- pass 2: ordered dead list / live position list ----
    // count first so that one thread can reserve the tile's output ranges
    int n_dead_mine = 0, n_live_mine = 0, n_liverec_mine = 0;
    for (uint32_t c = tid; c < td.ncand; c += kClassifyThreads) {
        const uint32_t e = a.cand[td.cand_off + c];
        if (s_flags[c] & 1u) {
            ++n_live_mine;
            n_liverec_mine += (int) ((e >> 18) & 7u);
        } else {
            ++n_dead_mine;
        }
    }
    int n_dead, n_live, n_liverec;
    block_exclusive_scan(n_dead_mine, s_warp, n_dead);
    block_exclusive_scan(n_live_mine, s_warp, n_live);
    block_exclusive_scan(n_liverec_mine, s_warp, n_liverec);
    if (tid == 0) {
        unsigned long long d_off = atomicAdd(&a.counters->n_dead, (unsigned long long) n_dead);
        unsigned long long l_off = atomicAdd(&a.counters->n_live, (unsigned long long) n_live);
        unsigned long long r_off = atomicAdd(&a.counters->n_liverec, (unsigned long long) n_liverec);
        unsigned int ovf = 0;
        if (d_off + (unsigned long long) n_dead > a.dead_cap)
            ovf |= 4u;
        if (l_off + (unsigned long long) n_live > a.live_cap)
            ovf |= 8u;
        if (r_off + (unsigned long long) n_liverec > a.liverec_cap)
            ovf |= 16u;
        if (ovf)
            atomicOr(&a.counters->overflow, ovf);
        TileOut to;
        to.dead_off = (uint32_t) d_off;
        to.ndead = (uint32_t) n_dead;
        to.live_off = (uint32_t) l_off;
        to.nlive = (uint32_t) n_live;
        to.liverec_off = (uint32_t) r_off;
        to.nliverec = (uint32_t) n_liverec;
        a.tiles_out[tile] = to;
        s_warp[33] = (int) to.dead_off;
        s_warp[34] = (int) to.live_off;
        s_warp[35] = (int) to.liverec_off;
        s_warp[36] = (int) ovf;
    }
    __syncthreads();
    const uint32_t dead_off = (uint32_t) s_warp[33], live_off = (uint32_t) s_warp[34], liverec_off = (uint32_t) s_warp[35];
    if (s_warp[36])
        return; // the host grows the buffers and runs the span again

    const uint32_t kb0 = (uint32_t) (p0 / a.block_samples);
    const bool one_block = ((long long) (kb0 + 1) * a.block_samples >= p0 + kTile);
    uint32_t bd_local[8] = {0, 0, 0, 0, 0, 0, 0, 0};

    int dead_done = 0, live_done = 0, liverec_done = 0;
    for (uint32_t cb = 0; cb < td.ncand; cb += kClassifyThreads) { // uniform trip count
        const uint32_t c = cb + tid;
        bool is_dead = false, is_live = false;
        uint32_t e = 0, nonbad = 0;
        if (c < td.ncand) {
            e = a.cand[td.cand_off + c];
            nonbad = (e >> 18) & 7u;
            is_live = (s_flags[c] & 1u) != 0;
            is_dead = !is_live;
        }
        int tot_d, tot_l, tot_r;
        const int od = block_exclusive_scan(is_dead ? 1 : 0, s_warp, tot_d);
        const int ol = block_exclusive_scan(is_live ? 1 : 0, s_warp, tot_l);
        const int orr = block_exclusive_scan(is_live ? (int) nonbad : 0, s_warp, tot_r);
        if (is_dead) {
            const uint32_t unknown = (s_flags[c] >> 1) & 1u;
            a.dead[dead_off + dead_done + od] = (e & 0x3ffffu) | (unknown << 18);
            // what demodulate2400 counts for a position whose best score is negative
            // (demod_2400.c:184,339-347), provided no accepted frame skips over it
            const uint32_t tm = (e >> 13) & 31u;
            uint32_t bd[8] = {1u, unknown ? 0u : 1u, unknown, tm & 1u, (tm >> 1) & 1u, (tm >> 2) & 1u, (tm >> 3) & 1u, (tm >> 4) & 1u};
            if (one_block) {
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    bd_local[k] += bd[k];
            } else {
                const uint32_t kb = (uint32_t) ((p0 + (e & 0x1fffu)) / a.block_samples);
                uint32_t *dst = reinterpret_cast<uint32_t *>(&a.block_dead[kb]);
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    if (bd[k])
                        atomicAdd(&dst[k], bd[k]);
            }
        }
        if (is_live) {
            LivePos lp;
            lp.pos = (uint32_t) (p0 + (e & 0x1fffu));
            lp.info = ((e >> 13) & 31u) | (nonbad << 8) | ((uint32_t) (liverec_done + orr) << 16);
            a.live[live_off + live_done + ol] = lp;
            s_slot[c] = (uint16_t) (liverec_done + orr);
        }
        dead_done += tot_d;
        live_done += tot_l;
        liverec_done += tot_r;
    }
    if (one_block) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            uint32_t v = bd_local[k];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1)
                v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0 && v)
                atomicAdd(&s_bd[k], v);
        }
        __syncthreads();
        if (tid < 8 && s_bd[tid])
            atomicAdd(reinterpret_cast<uint32_t *>(&a.block_dead[kb0]) + tid, s_bd[tid]);
    }
    if (n_liverec == 0)
        return;

    // ---- pass 3: records of live positions, in (position, phase) order: re-slice + signal power ----
    // the tile's records are ordered by candidate then phase; a live position owns consecutive slots
    __syncthreads();

    for (uint32_t r = warp; r < td.nrec; r += kClassifyThreads / 32) {
        const PhaseRec pr = a.recs[td.rec_off + r];
        const uint32_t c = pr.cand - td.cand_off;
        if (!(s_flags[c] & 1u))
            continue;
        // slot: first record slot of the position + rank of this record among the position's records
        uint32_t rank = 0;
        for (uint32_t q = r; q > 0 && a.recs[td.rec_off + q - 1].cand == pr.cand; --q)
            ++rank;
        const uint32_t slot = liverec_off + s_slot[c] + rank;

        // magnitudes the frame touches: window position pos -> samples pos - kOverlap ...
        uint16_t *fm = s_frame[warp];
        const long long s_first = (long long) pr.pos - kOverlap;
        for (int x = lane; x < kFrameSamples; x += 32)
            fm[x] = (uint16_t) sample_mag(a, s_first + x);
        __syncwarp();

        const int ph = (int) ((pr.w1 >> 24) & 15u);
        uint32_t w[4], syn;
        // DF from the first five bits decides the length (demod_2400.c:193-205)
        uint32_t df = 0;
        for (int b = 0; b < 5; ++b)
            df = (df << 1) | (slice_bit(fm, ph, b, s_coef) ? 1u : 0u);
        const int nbits = (df & 0x10u) ? 112 : 56;
        warp_slice_frame(fm, ph, nbits, s_coef, s_syn, w, syn);

        // demod_2400.c:387-396: sum of m^2 over msglen*12/5 samples from m[19]
        const int signal_len = nbits * 12 / 5;
        unsigned long long power = 0;
        for (int k = lane; k < signal_len; k += 32) {
            unsigned long long v = fm[19 + k];
            power += v * v;
        }
        power = warp_sum_u64(power);

        const FrameClass fc = classify_frame(df, __brev(w[0]) & 0xffffffu, syn, (w[0] | w[1] | w[2] | w[3]) == 0,
                                             a.tab_short, a.n_short, a.tab_long, a.n_long);
        if (lane == 0) {
            LiveRec lr;
            lr.pos = pr.pos;
            lr.w0 = syn | (fc.kind << 24) | (fc.errors << 28);
            lr.w1 = fc.key | ((uint32_t) ph << 24);
            lr.errbits = (uint32_t) (uint8_t) fc.bit0 | ((uint32_t) (uint8_t) fc.bit1 << 8);
            lr.power = power;
#pragma unroll
            for (int k = 0; k < 14; ++k)
                lr.msg[k] = (uint8_t) ((__brev(w[k >> 2]) >> (24 - 8 * (k & 3))) & 0xffu);
            lr.pad[0] = lr.pad[1] = 0;
            a.liverecs[slot] = lr;
        }
        __syncwarp();
    }
}

cudaError_t launch_classify(const ClassifyArgs &a, cudaStream_t stream) {
    if (a.ntiles == 0)
        return cudaSuccess;
    classify_kernel<<<a.ntiles, kClassifyThreads, 0, stream>>>(a);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// convert_kernel: the iq_convert_fn boundary (convert.h:33-38), magnitudes materialised
// ------------------------------------------------------------------------------------------

__global__ void __launch_bounds__(256) convert_kernel(const uint8_t *__restrict__ iq, uint32_t format, uint32_t n,
                                                       const uint16_t *__restrict__ lut, uint16_t *__restrict__ mag,
                                                       unsigned long long *sums_u64, double *sums_f64) {
    unsigned long long sl = 0, sp = 0;
    double fl = 0, fp = 0;
    const float inv_scale = (format == 1) ? (1.0f / 32768.0f) : (1.0f / 2048.0f);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        uint32_t m;
        if (format == 0) {
            uint32_t idx = reinterpret_cast<const uint16_t *>(iq)[i];
            m = __ldg(&lut[idx]);
            sl += m;
            sp += (unsigned long long) m * m;
        } else {
            float magsq, fm;
            m = mag_sc16_word(reinterpret_cast<const uint32_t *>(iq)[i], inv_scale, magsq, fm);
            fl += (double) fm;
            fp += (double) magsq;
        }
        mag[i] = (uint16_t) m;
    }
    if (format == 0) {
        sl = warp_sum_u64(sl);
        sp = warp_sum_u64(sp);
        if ((threadIdx.x & 31) == 0 && (sl | sp)) {
            atomicAdd(&sums_u64[0], sl);
            atomicAdd(&sums_u64[1], sp);
        }
    } else {
        fl = warp_sum_f64(fl);
        fp = warp_sum_f64(fp);
        if ((threadIdx.x & 31) == 0) {
            atomicAdd(&sums_f64[0], fl);
            atomicAdd(&sums_f64[1], fp);
        }
    }
}

cudaError_t launch_convert(const uint8_t *iq, uint32_t format, uint32_t nsamples, const uint16_t *lut, uint16_t *mag,
                           unsigned long long *sums_u64, double *sums_f64, cudaStream_t stream) {
    if (nsamples == 0)
        return cudaSuccess;
    int grid = (int) ((nsamples + 255) / 256);
    if (grid > 148 * 8)
        grid = 148 * 8;
    convert_kernel<<<grid, 256, 0, stream>>>(iq, format, nsamples, lut, mag, sums_u64, sums_f64);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// crc_batch_kernel: the crc.h boundary (modesChecksum + modesChecksumDiagnose) for n frames
// ------------------------------------------------------------------------------------------

__global__ void __launch_bounds__(256) crc_batch_kernel(const uint8_t *__restrict__ frames, uint32_t n,
                                                         const ErrorInfo *__restrict__ tab_short, int n_short,
                                                         const ErrorInfo *__restrict__ tab_long, int n_long,
                                                         uint32_t *syndromes, int8_t *errors, int8_t *bits2) {
    const int lane = threadIdx.x & 31;
    const uint32_t wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t f = wid; f < n; f += nw) {
        const uint8_t *msg = frames + (size_t) f * 14;
        const int nbits = (msg[0] & 0x80) ? 112 : 56; // mode_s.c:81-83
        const int off = 112 - nbits;
        uint32_t x = 0;
        for (int b = lane; b < nbits; b += 32)
            if ((msg[b >> 3] >> (7 - (b & 7))) & 1)
                x ^= c_bit_syndrome[b + off];
        const uint32_t syn = __reduce_xor_sync(0xffffffffu, x);
        if (lane == 0) {
            syndromes[f] = syn;
            int e = 0, b0 = -1, b1 = -1;
            if (syn != 0) {
                const ErrorInfo *tab = (nbits == 56) ? tab_short : tab_long;
                const int idx = find_syndrome(tab, (nbits == 56) ? n_short : n_long, syn);
                if (idx < 0) {
                    e = -1;
                } else {
                    e = tab[idx].errors;
                    b0 = tab[idx].bit[0];
                    b1 = tab[idx].bit[1];
                }
            }
            errors[f] = (int8_t) e;
            bits2[2 * f] = (int8_t) b0;
            bits2[2 * f + 1] = (int8_t) b1;
        }
    }
}

cudaError_t launch_crc_batch(const uint8_t *frames14, uint32_t n, const ErrorInfo *tab_short, int n_short,
                             const ErrorInfo *tab_long, int n_long, uint32_t *syndromes, int8_t *errors, int8_t *bits2,
                             cudaStream_t stream) {
    if (n == 0)
        return cudaSuccess;
    int grid = (int) ((n + 7) / 8);
    if (grid > 148 * 8)
        grid = 148 * 8;
    crc_batch_kernel<<<grid, 256, 0, stream>>>(frames14, n, tab_short, n_short, tab_long, n_long, syndromes, errors, bits2);
    return cudaGetLastError();
}

} // namespace b200
