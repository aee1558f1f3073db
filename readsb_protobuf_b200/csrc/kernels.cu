// kernels.cu -- sm_100a kernels of the Mode S demodulator.
//
//   K1a scan2_kernel (uc8,     IQ -> magnitude (uc8 table, sc16 / sc16q11 float path, or format 4: the sc16q11 table
//       scan2.inl) /
//       scan_kernel (others)
//                              of a -DSC16Q11_TABLE_BITS build, convert.c:264-328) -> per-block sums, preamble pre-check + three correlators for
//                              every scan position -> position-ordered candidate list per tile; the u16
//                              magnitudes go to HBM for K1b / K2 / Mode A/C
//                              replaces convert.c:63-111/215-253/332-370 and demod_2400.c:257-335
//   K1b slice_kernel           PPM slice of every (candidate, phase), CRC-24 syndrome, error-table lookup ->
//                              class records; replaces demod_2400.c:98-229, crc.c:67-82,389-412 and the
//                              filter-independent half of mode_s.c:311-409
//   K2  classify_warp_kernel   address-set test, ordered dead / live lists, live records (K1b's record, frame
//       (classify_kernel)      bytes included) + signal power of the survivors (demod_2400.c:387-399); the
//                              CTA-per-tile variant serves the exact-slab retry of very dense chunks
//   live_offsets / live_gather K2's per-tile live lists packed into stream order, straight into pinned host memory,
//                              with the dead positions a frame accepted at each live position would hide
//   modeac_kernel              demodulate2400AC's framing-pulse search (demod_2400.c:522-683), --modeac only
//   float_block_sums_kernel    sc16 / sc16q11 mean_level / mean_power in the reference's summation order
//   dc_prepare / dc_chain /    --dcfilter: convert_*_generic (convert.c:113-213, 374-423), the one-pole DC block
//   dc_magnitude_kernel        walked as the reference's own chain of float operations -> a u16 magnitude stream
//                              that K1a reads as format 3
//   convert_kernel             IQ -> u16 magnitudes in global memory (the iq_convert_fn boundary)
//   crc_batch_kernel           CRC + diagnose for a batch of frames (the crc.h boundary)
//
// Everything here is integer/byte work bounded by instruction issue (and, behind that, HBM bandwidth);
// there is no dense contraction, so no tensor-core path.

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
    // little-endian int16 pair: I in the low half (convert.c:231-232).  (float) v / 2^k without the quarter-rate
    // I2F: 0x4b40_0000 + v is the float 1.5 * 2^23 + v (|v| < 2^22), and (1.5 * 2^23 + v) * 2^-k - 1.5 * 2^(23-k)
    // is exact in one FMA (a power-of-two product, a representable difference).
    int vI, vQ; // the sign-extended halves: prmt's selector bit 3 replicates the byte's sign (__byte_perm masks it off)
    asm("prmt.b32 %0, %1, 0, 0x9910;" : "=r"(vI) : "r"(w));
    asm("prmt.b32 %0, %1, 0, 0xbb32;" : "=r"(vQ) : "r"(w));
    const float bias = __fmul_rn(-12582912.0f, inv_scale);
    float fI = __fmaf_rn(__int_as_float(0x4b400000 + vI), inv_scale, bias);
    float fQ = __fmaf_rn(__int_as_float(0x4b400000 + vQ), inv_scale, bias);
    return mag_from_float(fI, fQ, magsq, mag);
}

// convert_sc16q11_table's table index of one sample (convert.c:312-314): I in the low half of the word
__device__ __forceinline__ uint32_t sc16q11_table_index(uint32_t w, int bits) {
    const uint32_t I = (uint32_t) abs((int) (int16_t) (w & 0xffffu)) & 2047u;
    const uint32_t Q = (uint32_t) abs((int) (int16_t) (w >> 16)) & 2047u;
    return ((I >> (11 - bits)) << bits) | (Q >> (11 - bits));
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

// bloom: optional 2^14-bit filter over the syndromes of both tables (bit hash(s) set for every
// entry): a miss proves the syndrome is in neither table without walking it.
constexpr int kBloomBits = 1 << 14;
__device__ __forceinline__ uint32_t bloom_hash(uint32_t syn) {
    return (syn ^ (syn >> 10)) & (kBloomBits - 1);
}
__device__ __forceinline__ bool bloom_miss(const uint32_t *bloom, uint32_t syn) {
    const uint32_t h = bloom_hash(syn);
    return bloom && !((bloom[h >> 5] >> (h & 31u)) & 1u);
}

__device__ __forceinline__ FrameClass classify_frame(uint32_t df, uint32_t aa, uint32_t syn, bool all_zero,
                                                     const ErrorInfo *__restrict__ tab_short, int n_short,
                                                     const ErrorInfo *__restrict__ tab_long, int n_long,
                                                     const uint32_t *bloom = nullptr) {
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
                if (bloom_miss(bloom, c2))
                    return fc;
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
                if (bloom_miss(bloom, syn))
                    return fc;
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

// ------------------------------------------------------------------------------------------
// K1: scan kernel -- warp-autonomous streaming
//
// A warp owns a tile (kTile scan positions) at a time, taken from a global work queue, and
// streams through it in steps of kStep samples with no block-wide synchronisation at all:
//   convert   16 samples per lane (uint4 loads prefetched one step ahead, 64 K-entry magnitude
//             table in shared memory for uc8, IEEE float path for sc16) -> warp-private buffer,
//             block sums of the samples the tile owns
//   scan      one step behind: 16 positions per lane, 36-sample register window, pre-check +
//             the three preamble correlators -> three 16-bit maps per lane
//   slice     candidates -> (position, phase) items; 8 lanes per item, 4 items per pass, one bit
//             of every byte per lane, ballot = 4 message bytes; CRC syndrome by XOR of single-bit
//             syndromes; syndrome-table lookup; class record into the tile's slab
// ------------------------------------------------------------------------------------------

template <int FORMAT>
struct Fmt;
template <>
struct Fmt<0> { // uc8: 2 bytes per sample, 8 samples per 16-byte unit
    static constexpr int kBytes = 2, kUnitSamples = 8;
};
template <>
struct Fmt<1> { // sc16
    static constexpr int kBytes = 4, kUnitSamples = 4;
};
template <>
struct Fmt<2> { // sc16q11
    static constexpr int kBytes = 4, kUnitSamples = 4;
};

template <>
struct Fmt<4> { // sc16q11 through the magnitude table of a -DSC16Q11_TABLE_BITS build (convert.c:264-328)
    static constexpr int kBytes = 4, kUnitSamples = 4;
};
template <>
struct Fmt<3> { // u16 magnitudes already converted (the --dcfilter front end, dc_* kernels below)
    static constexpr int kBytes = 2, kUnitSamples = 8;
};

constexpr size_t kSmemLut = 65536 * sizeof(uint16_t);
constexpr size_t kSmemWarp = kWarpBuf * sizeof(uint32_t);
constexpr size_t kSmemTail = 0;

size_t scan_smem_bytes(uint32_t format) {
    return ((format == 0 || format == 4) ? kSmemLut : 0) + kScanWarps * kSmemWarp + kSmemTail;
}

// Load the 16-byte unit holding samples [s, s + kUnitSamples) of the span (s relative to the first
// new sample; negative = carried head).  lo/hi = the valid sample range inside the unit.
template <int FORMAT>
__device__ __forceinline__ uint4 load_unit(const ScanArgs &a, long long s, int &lo, int &hi) {
    constexpr int US = Fmt<FORMAT>::kUnitSamples, BPS = Fmt<FORMAT>::kBytes;
    const long long n = (long long) a.nsamples;
    const long long l = -(long long) a.head_valid - s, h = n - s;
    lo = l < 0 ? 0 : (l > US ? US : (int) l);
    hi = h > US ? US : (h < 0 ? 0 : (int) h);
    if (hi <= lo)
        return make_uint4(0, 0, 0, 0);
    if (s < 0) // carried head: always fully addressable
        return ldg_stream(reinterpret_cast<const uint4 *>(a.head + (s + kHead) * BPS));
    if (hi == US)
        return ldg_stream(reinterpret_cast<const uint4 *>(a.iq + s * BPS));
    // ragged end of the span: never read past the caller's buffer
    uint32_t w[4] = {0, 0, 0, 0};
    const uint8_t *p = a.iq + s * BPS;
    for (int k = 0; k < hi * BPS; ++k)
        w[k >> 2] |= (uint32_t) p[k] << (8 * (k & 3));
    return make_uint4(w[0], w[1], w[2], w[3]);
}

struct WarpCtx { // a tile's candidate output cursor
    uint32_t *cand_out;
    uint32_t cand_cap, ncand;
    unsigned long long ncand_total; // over the warp's tiles
};

__device__ __noinline__ void flush_sums_u64(unsigned long long *dst, unsigned long long level, unsigned long long power) {
    const unsigned long long l = warp_sum_u64(level), p = warp_sum_u64(power);
    if ((threadIdx.x & 31) == 0 && (l | p)) {
        atomicAdd(&dst[0], l);
        atomicAdd(&dst[1], p);
    }
}

__device__ __noinline__ void flush_sums_f64(double *dst, double level, double power) {
    if (!dst)
        return; // the float formats' block sums come from float_block_sums_kernel
    const double l = warp_sum_f64(level), p = warp_sum_f64(power);
    if ((threadIdx.x & 31) == 0 && (l != 0 || p != 0)) {
        atomicAdd(&dst[0], l);
        atomicAdd(&dst[1], p);
    }
}

// a mag_buf boundary inside a chunk: every 16-byte unit lies on one side of it and adds its sums itself
__device__ __noinline__ void unit_sums_u64(unsigned long long *block_sums, long long kb, uint4 lo, uint4 hi) {
    const uint32_t m[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
    unsigned long long cl = 0, cp = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        cl += m[j];
        cp += (unsigned long long) m[j] * m[j];
    }
    if (cl | cp) {
        atomicAdd(&block_sums[2 * kb], cl);
        atomicAdd(&block_sums[2 * kb + 1], cp);
    }
}

__device__ __noinline__ void unit_sums_f64(double *block_sums, long long kb, float4 mag, float4 magsq) {
    if (!block_sums)
        return;
    const double fl = (double) mag.x + (double) mag.y + (double) mag.z + (double) mag.w;
    const double fp = (double) magsq.x + (double) magsq.y + (double) magsq.z + (double) magsq.w;
    atomicAdd(&block_sums[2 * kb], fl);
    atomicAdd(&block_sums[2 * kb + 1], fp);
}

// debug tap (b200_debug_scan): the 5-bit try mask of every position of a lane's step
__device__ __noinline__ void store_dbg_masks(uint8_t *dst, uint32_t b45, uint32_t b67, uint32_t b8, uint32_t vmask) {
    for (int i = 0; i < kLanePos; ++i)
        if ((vmask >> i) & 1u)
            dst[i] = (uint8_t) ((((b45 >> i) & 1u) * 3u) | (((b67 >> i) & 1u) * 12u) | (((b8 >> i) & 1u) * 16u));
}

// uc8 table index with the shared-memory bank swizzle applied to both samples of a word:
// entry i lives at i ^ (((i >> 8) & 31) << 1), so that samples with equal I but different Q
// (receiver noise sits in a few codes around 127) do not collide on one bank
__device__ __forceinline__ uint32_t swizzle_pair(uint32_t w) {
    return w ^ ((w >> 7) & 0x003e003eu);
}

// scan2_kernel's layout of the same table: entry i at i ^ ((i >> 5) & 0x38) (scan2.inl)
__device__ __forceinline__ uint32_t swizzle_pair2(uint32_t w) {
    return w ^ ((w >> 5) & 0x00380038u);
}

// SWZ: which shared-memory layout the uc8 table is staged in (1: scan_kernel's, 2: scan2_kernel's)
template <int FORMAT, bool SLICE, bool EDGE, int SWZ = 1>
__device__ __forceinline__ void process_tile(const ScanArgs &a, WarpCtx &cx, const uint32_t tile, const uint16_t *s_lut, uint32_t *s_buf) {
    constexpr int US = Fmt<FORMAT>::kUnitSamples, BPS = Fmt<FORMAT>::kBytes;
    constexpr int UNITS = kLanePos / US; // 16-byte units per lane per step
    const int lane = threadIdx.x & 31;
    constexpr bool TABLE = (FORMAT == 0 || FORMAT == 4); // table converters: integer block sums (convert.c:104-110, 321-327)
    const float inv_scale = (FORMAT == 1) ? (1.0f / 32768.0f) : (1.0f / 2048.0f);
    const long long n = (long long) a.nsamples;
    const long long B = (long long) a.block_samples;
    const int thr = a.threshold;

    const long long c0 = (long long) tile * kTile - kHead; // first window-start sample of the tile
    // the tile's slab of the candidate array (and, for K1b, of the record array)
    uint32_t cand_off, rec_off;
    if (a.tile_off) {
        cand_off = a.tile_off[2 * tile];
        rec_off = a.tile_off[2 * tile + 1];
        cx.cand_cap = a.tile_off[2 * tile + 2] - cand_off;
    } else {
        cand_off = tile * a.cand_slab;
        rec_off = tile * a.rec_slab;
        cx.cand_cap = a.cand_slab;
    }
    cx.cand_out = a.cand + cand_off;
    cx.ncand = 0;
    uint32_t ncand_lane = 0; // scan-only mode: candidates seen by this lane

    // block sums of the samples this tile owns: [c0, c0 + kTile) within [0, n)
    unsigned long long sum_level = 0, sum_power = 0;
    double fsum_level = 0, fsum_power = 0;
    long long blk = (c0 > 0 ? c0 : 0) / B, next_bound = (blk + 1) * B;
    auto flush_sums = [&]() { // rare (once per tile and per mag_buf boundary): kept out of the hot code
        if (FORMAT == 3)
            return;
        if (TABLE)
            flush_sums_u64(a.block_sums_u64 + 2 * blk, sum_level, sum_power);
        else
            flush_sums_f64(a.block_sums_f64 ? a.block_sums_f64 + 2 * blk : nullptr, fsum_level, fsum_power);
        sum_level = sum_power = 0;
        fsum_level = fsum_power = 0;
    };

    uint4 pre[UNITS];
    // interior tiles: every sample of the tile's 17 chunks is a new sample of the span
    const uint4 *gp = reinterpret_cast<const uint4 *>(a.iq + (c0 + lane * kLanePos) * BPS);
    auto prefetch = [&](int k) {
        if (EDGE) {
            const long long ls = c0 + (long long) k * kStep + lane * kLanePos;
#pragma unroll
            for (int u = 0; u < UNITS; ++u) {
                int lo, hi;
                pre[u] = load_unit<FORMAT>(a, ls + u * US, lo, hi);
            }
        } else {
            const uint4 *p = gp + (size_t) k * (kStep * BPS / 16);
#pragma unroll
            for (int u = 0; u < UNITS; ++u)
                pre[u] = ldg_stream(p + u);
        }
    };
    prefetch(0);

    for (int k = 0; k <= kScanSteps; ++k) {
        const long long cs = c0 + (long long) k * kStep; // first sample of chunk k
        const long long ls = cs + lane * kLanePos;        // this lane's first sample
        // ---------------- convert chunk k ----------------
        uint32_t m[kLanePos];
        float fmag[kLanePos], fmagsq[kLanePos];
        if (FORMAT == 0) {
#pragma unroll
            for (int u = 0; u < UNITS; ++u) {
                const uint32_t words[4] = {SWZ == 2 ? swizzle_pair2(pre[u].x) : swizzle_pair(pre[u].x), SWZ == 2 ? swizzle_pair2(pre[u].y) : swizzle_pair(pre[u].y),
                                           SWZ == 2 ? swizzle_pair2(pre[u].z) : swizzle_pair(pre[u].z), SWZ == 2 ? swizzle_pair2(pre[u].w) : swizzle_pair(pre[u].w)};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    m[u * US + 2 * j] = s_lut[words[j] & 0xffffu];
                    m[u * US + 2 * j + 1] = s_lut[words[j] >> 16];
                }
            }
        } else if (FORMAT == 4) {
            // convert_sc16q11_table (convert.c:312-316): the top table_bits bits of |I| & 2047 and |Q| & 2047
            // index the table (staged with the uc8 table's bank swizzle)
            const int bits = (int) a.table_bits, lose = 11 - bits;
#pragma unroll
            for (int u = 0; u < UNITS; ++u) {
                const uint32_t words[4] = {pre[u].x, pre[u].y, pre[u].z, pre[u].w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const uint32_t I = (uint32_t) abs((int) (int16_t) (words[j] & 0xffffu)) & 2047u;
                    const uint32_t Q = (uint32_t) abs((int) (int16_t) (words[j] >> 16)) & 2047u;
                    const uint32_t idx = ((I >> lose) << bits) | (Q >> lose);
                    // up to 8 bits the table is in shared memory (the uc8 table's slot and swizzle); the larger ones
                    // (9..11 bits: 512 KiB .. 8 MiB) are read through L2
                    m[(u * US + j) % kLanePos] = bits <= 8 ? (uint32_t) s_lut[idx ^ (((idx >> 8) & 31u) << 1)] : (uint32_t) __ldg(&a.lut[idx]);
                }
            }
        } else if (FORMAT == 3) {
            // the stream already holds magnitudes (DC-filtered front end): two per word
#pragma unroll
            for (int u = 0; u < UNITS; ++u) {
                const uint32_t words[4] = {pre[u].x, pre[u].y, pre[u].z, pre[u].w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    m[u * US + (2 * j) % US] = words[j] & 0xffffu;
                    m[u * US + (2 * j + 1) % US] = words[j] >> 16;
                }
            }
        } else {
#pragma unroll
            for (int u = 0; u < UNITS; ++u) {
                const uint32_t words[4] = {pre[u].x, pre[u].y, pre[u].z, pre[u].w};
#pragma unroll
                for (int j = 0; j < US; ++j)
                    m[u * US + j] = mag_sc16_word(words[j % 4], inv_scale, fmagsq[u * US + j], fmag[u * US + j]);
            }
        }
        if (EDGE) {
            // samples outside the stream have magnitude 0 (fifo.c:47) and are not summed
            const long long l = -(long long) a.head_valid - ls, h = n - ls;
            if (l > 0 || h < kLanePos) {
#pragma unroll
                for (int j = 0; j < kLanePos; ++j)
                    if (j < l || j >= h) {
                        m[j] = 0;
                        if (FORMAT == 1 || FORMAT == 2)
                            fmag[j] = fmagsq[j] = 0;
                    }
            }
        }
        // request the next chunk now; it is consumed one iteration later
        if (k < kScanSteps)
            prefetch(k + 1);

        // store: chunk k occupies rows (k&1)*32 + lane of the warp's ring
        {
            const int row = (k & 1) * 32 + lane;
            uint4 *dst = reinterpret_cast<uint4 *>(s_buf + row * kRowWords);
#pragma unroll
            for (int q = 0; q < kLanePos / 4; ++q)
                dst[q] = make_uint4(m[4 * q], m[4 * q + 1], m[4 * q + 2], m[4 * q + 3]);
        }
        // the tile's own 16 chunks also go to the u16 magnitude array K1b and K2 slice from
        // (index = sample + kHead; samples outside the stream are 0)
        if (SLICE && k < kScanSteps) {
            uint4 *g = reinterpret_cast<uint4 *>(a.mag + (size_t) tile * kTile + (size_t) k * kStep + lane * kLanePos);
            g[0] = make_uint4(m[0] | (m[1] << 16), m[2] | (m[3] << 16), m[4] | (m[5] << 16), m[6] | (m[7] << 16));
            g[1] = make_uint4(m[8] | (m[9] << 16), m[10] | (m[11] << 16), m[12] | (m[13] << 16), m[14] | (m[15] << 16));
        }

        // block sums: chunks 0..kScanSteps-1 are owned by this tile (format 3: the DC front end made them)
        if (FORMAT != 3 && k < kScanSteps) {
            const long long own_lo = cs > 0 ? cs : 0, own_hi = (cs + kStep < n) ? cs + kStep : n;
            if (!EDGE || own_hi > own_lo) {
                if (own_lo >= next_bound) {
                    flush_sums();
                    blk = own_lo / B;
                    next_bound = (blk + 1) * B;
                }
                if (own_hi <= next_bound) {
                    if (!EDGE || ls >= 0) { // head samples are not this span's (a lane never straddles 0: kHead % 16 == 8 is
                                            // handled below for the one lane that does)
                        if (TABLE) {
                            uint32_t s32 = 0;
                            unsigned long long p64 = 0;
#pragma unroll
                            for (int j = 0; j < kLanePos; ++j) {
                                s32 += m[j];
                                p64 += (unsigned long long) m[j] * m[j];
                            }
                            sum_level += s32;
                            sum_power += p64;
                        } else {
                            float fl = 0, fp = 0;
#pragma unroll
                            for (int j = 0; j < kLanePos; ++j) {
                                fl += fmag[j];
                                fp += fmagsq[j];
                            }
                            fsum_level += (double) fl;
                            fsum_power += (double) fp;
                        }
                    } else if (ls + kLanePos > 0) {
                        // the lane whose 16 samples straddle the start of the span: only samples >= 0
#pragma unroll
                        for (int j = 0; j < kLanePos; ++j)
                            if (ls + j >= 0) {
                                if (TABLE) {
                                    sum_level += m[j];
                                    sum_power += (unsigned long long) m[j] * m[j];
                                } else {
                                    fsum_level += (double) fmag[j];
                                    fsum_power += (double) fmagsq[j];
                                }
                            }
                    }
                } else {
                    // a mag_buf boundary inside the chunk: every 16-byte unit lies on one side of it
                    flush_sums();
#pragma unroll
                    for (int u = 0; u < UNITS; ++u) {
                        const long long us = ls + u * US;
                        if (us >= 0 && us < n) {
                            const long long kb = us / B;
                            if (TABLE)
                                unit_sums_u64(a.block_sums_u64, kb, make_uint4(m[u * US], m[u * US + 1], m[u * US + 2], m[u * US + 3]),
                                              US == 8 ? make_uint4(m[u * US + 4 % US], m[u * US + 5 % US], m[u * US + 6 % US], m[u * US + 7 % US])
                                                      : make_uint4(0, 0, 0, 0)); // a 16-byte unit of format 4 holds four samples
                            else
                                unit_sums_f64(a.block_sums_f64, kb, make_float4(fmag[u * US], fmag[u * US + 1], fmag[u * US + 2], fmag[u * US + 3]),
                                              make_float4(fmagsq[u * US], fmagsq[u * US + 1], fmagsq[u * US + 2], fmagsq[u * US + 3]));
                        }
                    }
                    blk = own_hi / B;
                    next_bound = (blk + 1) * B;
                }
            }
        }
        __syncwarp();

        // ---------------- scan chunk k-1 (its look-ahead, chunk k, is now in the buffer) ----------------
        if (k >= 1) {
            const int j = k - 1;
            const int row0 = (j & 1) * 32; // first row of chunk j
            const long long pos0 = c0 + (long long) j * kStep + kOverlap; // its scan position
            const long long lp0 = pos0 + lane * kLanePos;
            uint32_t vmask = 0xffffu;
            if (EDGE) {
                // positions of this lane that exist: 0 <= p < n
                const int vlo = (lp0 < 0) ? (int) (-lp0 > kLanePos ? kLanePos : -lp0) : 0;
                const long long vh = n - lp0;
                const int vhi = vh > kLanePos ? kLanePos : (vh < 0 ? 0 : (int) vh);
                vmask = (vhi > vlo) ? (((1u << vhi) - 1u) & ~((1u << vlo) - 1u)) : 0u;
            }

            uint32_t b45 = 0, b67 = 0, b8 = 0;
            if (!EDGE || __any_sync(0xffffffffu, vmask != 0)) {
                uint32_t w[kLanePos + 20];
                {
                    // 36 consecutive magnitudes: this lane's row, the next row, 4 of the one after
                    const uint4 *ra = reinterpret_cast<const uint4 *>(s_buf + (row0 + lane) * kRowWords);
                    const uint4 *rb = reinterpret_cast<const uint4 *>(s_buf + ((row0 + lane + 1) & (kRows - 1)) * kRowWords);
                    const uint4 *rc = reinterpret_cast<const uint4 *>(s_buf + ((row0 + lane + 2) & (kRows - 1)) * kRowWords);
#pragma unroll
                    for (int q = 0; q < 9; ++q) {
                        const uint4 v = (q < 4) ? ra[q] : (q < 8 ? rb[q - 4] : rc[0]);
                        w[4 * q] = v.x;
                        w[4 * q + 1] = v.y;
                        w[4 * q + 2] = v.z;
                        w[4 * q + 3] = v.w;
                    }
                }
                // The three correlators (demod_2400.c:298-330) as sign tests.  With
                //   Q[x] = m[x] + m[x+3], D[x] = m[x] - m[x+1], T = m[16] + m[17] + m[18]
                // base_noise = Q[5] + T and common3456 = Q[1] + Q[9] - D[2]; since
                // X >= (N >> 5)  <=>  32 X + 31 - N >= 0 for integers (N >= 0), every test is the sign of
                //   E0  = 32 (Q[1] + Q[9] - D[2]) + 31 - thr * base_noise
                //   E45 = E0 - 32 D[10]      E67 = E0 + 32 D[10]      E8 = E67 + 96 D[2] - 32 m[9]
                // (|32 X| < 2^24 and N < 2^28: no overflow).  The pre-check of demod_2400.c:276 is the sign
                // of (m[7]-m[1]) & (m[14]-m[12]) & (m[15]-m[12]).  Signs are shifted into per-lane masks
                // with one funnel shift each; positions run downwards so that bit i is position i.
                int Q[kLanePos + 9], D[kLanePos + 11];
#pragma unroll
                for (int x2 = 1; x2 < kLanePos + 9; ++x2)
                    Q[x2] = (int) (w[x2] + w[x2 + 3]);
#pragma unroll
                for (int x2 = 2; x2 < kLanePos + 11; ++x2)
                    D[x2] = (int) w[x2] - (int) w[x2 + 1];
                uint32_t s45 = 0, s67 = 0, s8 = 0, pm = 0;
                const int nthr = -thr;
#pragma unroll
                for (int i = kLanePos - 1; i >= 0; --i) {
                    const int T = (int) (w[i + 16] + w[i + 17] + w[i + 18]);
                    const int c = Q[i + 1] + Q[i + 9] - D[i + 2];
                    const int bn = Q[i + 5] + T;
                    const int E0 = nthr * bn + (c * 32 + 31);
                    const int E45 = D[i + 10] * -32 + E0;
                    const int E67 = D[i + 10] * 32 + E0;
                    const int E8 = (int) w[i + 9] * -32 + (D[i + 2] * 96 + E67);
                    const int g = ((int) w[i + 7] - (int) w[i + 1]) & ((int) w[i + 14] - (int) w[i + 12]) & ((int) w[i + 15] - (int) w[i + 12]);
                    s45 = __funnelshift_l((uint32_t) E45, s45, 1);
                    s67 = __funnelshift_l((uint32_t) E67, s67, 1);
                    s8 = __funnelshift_l((uint32_t) E8, s8, 1);
                    pm = __funnelshift_l((uint32_t) g, pm, 1);
                }
                b45 = pm & ~s45;
                b67 = pm & ~s67;
                b8 = pm & ~s8;
                if (EDGE) {
                    b45 &= vmask;
                    b67 &= vmask;
                    b8 &= vmask;
                }
            }

            if (a.dbg_masks)
                store_dbg_masks(a.dbg_masks + lp0, b45, b67, b8, vmask);

            const uint32_t any = b45 | b67 | b8;
            if (SLICE && lane == 0) // where step j's candidates start in the tile's list (K1b cuts the list into units)
                a.step_off[tile * kScanSteps + j] = (uint16_t) cx.ncand;
            if (!SLICE) {
                ncand_lane += __popc(any);
            } else {
                const uint32_t lanes = __ballot_sync(0xffffffffu, any != 0);
                if (lanes) {
                    // candidate entries of the step in position order.  Almost always a lane holds at most one
                    // (6 candidates per 512 positions on noise): its slot then follows from the ballot alone;
                    // only a lane with several sends the warp through a prefix sum.
                    const uint32_t mine = (uint32_t) __popc(any);
                    const uint32_t multi = __ballot_sync(0xffffffffu, mine > 1);
                    uint32_t ci, total;
                    if (!multi) {
                        ci = cx.ncand + __popc(lanes & ((1u << lane) - 1u));
                        total = (uint32_t) __popc(lanes);
                    } else {
                        uint32_t inc = mine;
#pragma unroll
                        for (int o = 1; o < 32; o <<= 1) {
                            const uint32_t up = __shfl_up_sync(0xffffffffu, inc, o);
                            if (lane >= o)
                                inc += up;
                        }
                        ci = cx.ncand + inc - mine;
                        total = __shfl_sync(0xffffffffu, inc, 31);
                    }
                    cx.ncand += total;
                    uint32_t u = any;
                    while (u) {
                        const int i = __ffs(u) - 1;
                        u &= u - 1;
                        const uint32_t tm = (((b45 >> i) & 1u) * 3u) | (((b67 >> i) & 1u) * 12u) | (((b8 >> i) & 1u) * 16u);
                        if (ci < cx.cand_cap)
                            cx.cand_out[ci] = (uint32_t) (j * kStep + lane * kLanePos + i) | (tm << 13);
                        ++ci;
                    }
                }
            }
            __syncwarp();
        }
    }
    flush_sums();

    // ---- tile descriptor ----
    if (!SLICE) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
            ncand_lane += __shfl_xor_sync(0xffffffffu, ncand_lane, o);
        cx.ncand = ncand_lane;
    }
    if (lane == 0) {
        if (SLICE) {
            TileDesc td;
            td.cand_off = cand_off;
            td.ncand = cx.ncand;
            td.rec_off = rec_off;
            td.nrec = 0; // K1b
            a.tiles[tile] = td;
            if (cx.ncand > cx.cand_cap)
                atomicOr(&a.counters->overflow, 1u);
        }
    }
    cx.ncand_total += cx.ncand;
}

template <int FORMAT, bool SLICE>
__global__ void __launch_bounds__(kScanThreads, 1) scan_kernel(const ScanArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    // shared memory carve-up
    unsigned char *sp = smem_raw;
    const uint16_t *s_lut = reinterpret_cast<const uint16_t *>(sp);
    if (FORMAT == 0 || FORMAT == 4)
        sp += kSmemLut;
    uint32_t *s_buf = reinterpret_cast<uint32_t *>(sp + (size_t) warp * kWarpBuf * sizeof(uint32_t));
    sp += (size_t) kScanWarps * kWarpBuf * sizeof(uint32_t);

    // one-time staging of the tables (the only block-wide barrier of the kernel).  The magnitude
    // table arrives pre-swizzled (32-bit word j of row Q at word j ^ (Q & 31)): all of a thread's
    // 16-byte loads are in flight at once, one round trip to L2 per CTA.
    if (FORMAT == 0 || (FORMAT == 4 && a.table_bits <= 8)) {
        const uint4 *src = reinterpret_cast<const uint4 *>(a.lut_swz);
        uint4 *dst = reinterpret_cast<uint4 *>(smem_raw);
        constexpr int kUnits = (int) (kSmemLut / 16);                      // 8192 sixteen-byte units
        constexpr int kPer = (kUnits + kScanThreads - 1) / kScanThreads;  // per thread, all in flight at once
        uint4 v[kPer];
#pragma unroll
        for (int q = 0; q < kPer; ++q)
            if (q * kScanThreads + tid < kUnits)
                v[q] = __ldg(src + q * kScanThreads + tid);
#pragma unroll
        for (int q = 0; q < kPer; ++q)
            if (q * kScanThreads + tid < kUnits)
                dst[q * kScanThreads + tid] = v[q];
    }
    __syncthreads();

    WarpCtx cx;
    cx.ncand_total = 0;

    const long long n = (long long) a.nsamples;
    for (;;) {
        // ---- next tile from the work queue ----
        uint32_t tile = 0;
        if (lane == 0)
            tile = atomicAdd(&a.counters->next_tile, 1u);
        tile = __shfl_sync(0xffffffffu, tile, 0);
        if (tile >= a.fast_lo) // tiles [fast_lo, fast_hi) belong to scan2_kernel
            tile += a.fast_hi - a.fast_lo;
        if (tile >= a.ntiles)
            break;
        // interior tile: all 17 chunks are new samples of the span and all positions exist
        const long long c0 = (long long) tile * kTile - kHead;
        const bool interior = c0 >= 0 && c0 + (long long) (kScanSteps + 1) * kStep <= n;
        if (interior)
            process_tile<FORMAT, SLICE, false>(a, cx, tile, s_lut, s_buf);
        else
            process_tile<FORMAT, SLICE, true>(a, cx, tile, s_lut, s_buf);
    }
    if (lane == 0 && cx.ncand_total) // one same-address atomic per warp, not per tile
        atomicAdd(&a.counters->n_cand, cx.ncand_total);
}

cudaError_t scan_configure() {
    cudaError_t e;
#define CFG(F, S)                                                                                                  \
    e = cudaFuncSetAttribute(scan_kernel<F, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) scan_smem_bytes(F)); \
    if (e != cudaSuccess)                                                                                          \
        return e;
    CFG(0, true) CFG(0, false) CFG(1, true) CFG(1, false) CFG(2, true) CFG(2, false) CFG(3, true) CFG(3, false) CFG(4, true) CFG(4, false)
#undef CFG
    return cudaSuccess;
}

cudaError_t launch_scan(const ScanArgs &a, int mode, int grid, cudaStream_t stream) {
    if (a.ntiles == 0)
        return cudaSuccess;
    const uint32_t mine = a.ntiles - (a.fast_hi - a.fast_lo);
    if (mine == 0)
        return cudaSuccess;
    const int max_useful = (int) ((mine + kScanWarps - 1) / kScanWarps);
    if (grid > max_useful)
        grid = max_useful;
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
        case 3: LAUNCH(3) break;
        case 4: LAUNCH(4) break;
        default: return cudaErrorInvalidValue;
    }
#undef LAUNCH
    return cudaGetLastError();
}

#include "scan2.inl"
#include "scan3.inl"

// ------------------------------------------------------------------------------------------
// K1b: slice kernel -- PPM slice + CRC class of every (candidate position, phase)
//
// Warp-autonomous like K1a: a warp takes a unit (1024 scan positions = two K1a steps of a tile) from
// a queue, stages the unit's magnitudes (+ the 297 a frame can reach past the last window) into its
// own shared-memory pair array, expands the unit's candidates into (position, phase) items, and then
//   1. one lane per item slices group 0 = the DF field -> frame length (demod_2400.c:188-205)
//   2. one lane per (frame, 5-bit group) slices five bits, ORs them into the frame's message words
//      and XORs the group's CRC contribution (crc.c:59-64: the syndrome is linear in the bits) into
//      its syndrome, both in shared memory
//   3. one lane per frame classifies it and appends the class record to the tile's slab
// Work is pooled over the unit, so the rounds of step 2 run with nearly all lanes busy, and there is
// no block-wide barrier after the table set-up: 32 independent warps per SM hide each other's latency.
// ------------------------------------------------------------------------------------------

constexpr int kSliceWarps = 32;
constexpr int kSliceThreads = kSliceWarps * 32;
constexpr int kUnit = 2 * kStep;              // scan positions per unit
constexpr int kUnitsPerTile = kTile / kUnit;  // 8
constexpr int kUnitMag = kUnit + 304;         // staged magnitudes (a frame reaches 297 past its window start)
constexpr int kItemMax = 5 * 32;              // (position, phase) items of a batch of 32 candidates
constexpr int kFrameBytes = 24;               // one byte per 5-bit group of a frame (23 used)
constexpr int kGroupsLong = 23, kGroupsShort = 12; // 5-bit groups of a 112 / 56 bit frame

// dynamic shared memory: per warp the pair array, item list, frame list and 32 frames' group bytes
constexpr size_t kSliceWarpSmem = kUnitMag * sizeof(uint32_t) + 2 * kItemMax * sizeof(uint16_t) + 32 * kFrameBytes;
constexpr size_t kSliceSmem = kSliceWarps * kSliceWarpSmem + (kGroupsLong + kGroupsShort) * 32 * sizeof(uint32_t) + 32 * sizeof(int2) + kBloomBits / 8;
static_assert(kSliceWarpSmem % 16 == 0, "per-warp slice buffers must stay 16-byte aligned");

__device__ __forceinline__ int dp2a_lo(uint32_t pair, int taps, int acc) {
    int d;
    asm("dp2a.lo.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(pair), "r"(taps), "r"(acc));
    return d;
}
__device__ __forceinline__ int dp2a_hi(uint32_t pair, int taps, int acc) {
    int d;
    asm("dp2a.hi.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(pair), "r"(taps), "r"(acc));
    return d;
}

// Five consecutive PPM bit decisions (demod_2400.c:73-177 in closed form).  Frame bit b of a
// candidate tried at phase p sits t = p + 12 b fifths of a sample after m[19]; five bits later the
// pattern repeats 12 samples on, so group k of a frame (bits 5k .. 5k+4) reads the 15 samples from
// m[19 + 12k] with offsets and correlators that depend on the phase only.
// P = the pair array at the frame's window start: P[x] = m[x] | m[x+1] << 16, so the four samples
// of a bit are two aligned 32-bit loads whatever x is, and the correlator is two 16x8-bit dot
// products (taps.x = the four taps as signed bytes, taps.y = first sample of the bit).
// phi = phase - 4.  Returns the five decisions, bit c = frame bit 5k + c.
// The pair array is stored swizzled: pair x lives at word x ^ ((x >> 5) & 3).  The groups of one
// frame are 12 samples apart, which without the swizzle puts lanes k, k + 8 and k + 16 of a round on
// the same bank (a 3-way conflict on every load); the swizzle moves them to different banks.
__device__ __forceinline__ int pair_slot(int x) {
    return x ^ ((x >> 5) & 3);
}

// x0 = the frame's window start in the unit
__device__ __forceinline__ uint32_t slice_group(const uint32_t *P, int x0, int phi, int k, const int2 *taps) {
    const int g = x0 + 19 + 12 * k;
    const int2 *tp = taps + phi * 5;
    uint32_t v5 = 0;
#pragma unroll
    for (int c = 0; c < 5; ++c) {
        const int2 t = tp[c];
        const int x = g + t.y;
        const int v = dp2a_hi(P[pair_slot(x + 2)], t.x, dp2a_lo(P[pair_slot(x)], t.x, 0));
        v5 |= (v > 0) ? (1u << c) : 0u;
    }
    return v5;
}

__global__ void __launch_bounds__(kSliceThreads, 1) slice_kernel(const SliceArgs a) {
    extern __shared__ __align__(16) unsigned char slice_smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    unsigned char *wp = slice_smem + (size_t) warp * kSliceWarpSmem;
    uint32_t *s_pair = reinterpret_cast<uint32_t *>(wp);
    uint8_t *s_bits = reinterpret_cast<uint8_t *>(s_pair + kUnitMag); // [32 frames][kFrameBytes]: five sliced bits per group
    uint16_t *s_items = reinterpret_cast<uint16_t *>(s_bits + 32 * kFrameBytes); // pos_in_unit[9:0] | (phase - 4)[12:10]
    uint16_t *s_frames = s_items + kItemMax;                                  // long frames from the front, short from the back
    uint32_t *s_gsyn = reinterpret_cast<uint32_t *>(slice_smem + (size_t) kSliceWarps * kSliceWarpSmem);
    int2 *s_taps = reinterpret_cast<int2 *>(s_gsyn + (kGroupsLong + kGroupsShort) * 32);
    uint32_t *s_bloom = reinterpret_cast<uint32_t *>(s_taps + 32);

    for (int i = tid; i < kBloomBits / 32; i += kSliceThreads)
        s_bloom[i] = 0;
    __syncthreads();
    for (int i = tid; i < a.n_short + a.n_long; i += kSliceThreads) {
        const uint32_t h = bloom_hash(i < a.n_short ? a.tab_short[i].syndrome : a.tab_long[i - a.n_short].syndrome);
        atomicOr(&s_bloom[h >> 5], 1u << (h & 31u));
    }
    // group syndromes: row g < 23 = bits 5g .. 5g+4 of a long frame, row 23 + g = of a short frame
    // (crc.c:143: a short frame uses the tail of the 112-entry single-bit syndrome list)
    for (int i = tid; i < (kGroupsLong + kGroupsShort) * 32; i += kSliceThreads) {
        const int g = i >> 5, v = i & 31;
        const bool is_long = g < kGroupsLong;
        const int b0 = 5 * (is_long ? g : g - kGroupsLong), nbits = is_long ? 112 : 56;
        uint32_t x = 0;
        for (int c = 0; c < 5; ++c)
            if (((v >> c) & 1) && b0 + c < nbits)
                x ^= c_bit_syndrome[b0 + c + (112 - nbits)];
        s_gsyn[i] = x;
    }
    if (tid < 25) { // bit c of a group at phase p: t = p + 12 c fifths, sample t / 5, correlator t % 5
        const int t = (tid / 5 + 4) + 12 * (tid % 5);
        const int r = t % 5;
        const uint32_t packed = (uint32_t) (uint8_t) c_slice_coef[r][0] | ((uint32_t) (uint8_t) c_slice_coef[r][1] << 8) |
                                ((uint32_t) (uint8_t) c_slice_coef[r][2] << 16) | ((uint32_t) (uint8_t) c_slice_coef[r][3] << 24);
        s_taps[tid] = make_int2((int) packed, t / 5);
    }
    __syncthreads(); // the only block-wide barrier

    const uint32_t below = (1u << lane) - 1u;
    const uint32_t nunits = a.ntiles * kUnitsPerTile;
    // Units are dealt round-robin to the warps of the grid (a shared atomic queue would serialise
    // 140 K same-address atomics per 144 M samples); the next unit's descriptors are requested one
    // iteration before they are needed, so nothing here waits on L2.
    auto load_desc = [&](uint32_t u, uint4 &td, uint32_t &so) {
        if (u < nunits) {
            const uint32_t t = u / kUnitsPerTile, sb = u % kUnitsPerTile;
            td = *reinterpret_cast<const uint4 *>(&a.tiles[t]); // cand_off, ncand, rec_off, (nrec: being updated, unused)
            // candidates in front of the unit's two steps [15:0] and in front of the next unit [31:16]
            so = (uint32_t) a.step_off[t * kScanSteps + 2 * sb] |
                 ((sb + 1 < kUnitsPerTile ? (uint32_t) a.step_off[t * kScanSteps + 2 * sb + 2] : 0xffffu) << 16);
        }
    };
    unsigned long long nrec_warp = 0;
    const uint32_t unit_stride = gridDim.x * kSliceWarps;
    uint32_t unit = blockIdx.x * kSliceWarps + warp;
    uint4 desc = make_uint4(0, 0, 0, 0);
    uint32_t desc_so = 0;
    load_desc(unit, desc, desc_so);
    for (; unit < nunits;) {
        const uint32_t tile = unit / kUnitsPerTile, sub = unit % kUnitsPerTile;
        const uint4 td = desc;
        const uint32_t so = desc_so;
        unit += unit_stride;
        load_desc(unit, desc, desc_so);

        const uint32_t t_ncand = td.y, t_cand_off = td.x, t_rec_off = td.z;
        uint32_t cand_cap, rec_cap;
        if (a.tile_off) {
            cand_cap = a.tile_off[2 * tile + 2] - a.tile_off[2 * tile];
            rec_cap = a.tile_off[2 * tile + 3] - a.tile_off[2 * tile + 1];
        } else {
            cand_cap = a.cand_slab;
            rec_cap = a.rec_slab;
        }
        const uint32_t ncand = t_ncand < cand_cap ? t_ncand : cand_cap; // a truncated tile is run again by the host
        // the unit's slice of the tile's (position-ordered) candidate list: K1a noted the count at every step
        uint32_t c_lo = so & 0xffffu;
        uint32_t c_hi = (sub + 1 < kUnitsPerTile) ? (so >> 16) : t_ncand;
        c_lo = c_lo < ncand ? c_lo : ncand;
        c_hi = c_hi < ncand ? c_hi : ncand;
        if (c_hi <= c_lo)
            continue;
        const uint32_t *cand = a.cand + t_cand_off;
        // first batch of candidates: requested now, used after the staging
        uint32_t e_first = 0;
        if (c_lo + lane < c_hi)
            e_first = __ldg(&cand[c_lo + lane]);

        // ---- stage the unit's magnitudes as the pair array (loads in two waves of three) ----
        {
            const uint32_t *src = reinterpret_cast<const uint32_t *>(a.mag + (size_t) tile * kTile + (size_t) sub * kUnit);
            constexpr int kUnits16 = kUnitMag / 8; // 166 sixteen-byte units
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                uint4 w[3];
                uint32_t nx[3];
#pragma unroll
                for (int q = 0; q < 3; ++q) {
                    const int i = (half * 3 + q) * 32 + lane;
                    if (i < kUnits16) {
                        w[q] = __ldg(reinterpret_cast<const uint4 *>(src) + i);
                        nx[q] = __ldg(src + 4 * i + 4); // first pair of the next unit (kMagSlack keeps it addressable)
                    }
                }
#pragma unroll
                for (int q = 0; q < 3; ++q) {
                    const int i = (half * 3 + q) * 32 + lane;
                    if (i < kUnits16) {
                        // pairs 8i .. 8i+7 lie in one block of 32, whose swizzle (i >> 2) & 3 = (lane >> 2) & 3 permutes
                        // each aligned group of four: word j of the group goes to j ^ f
                        uint4 lo = make_uint4(w[q].x, __funnelshift_r(w[q].x, w[q].y, 16), w[q].y, __funnelshift_r(w[q].y, w[q].z, 16));
                        uint4 hi = make_uint4(w[q].z, __funnelshift_r(w[q].z, w[q].w, 16), w[q].w, __funnelshift_r(w[q].w, nx[q], 16));
                        if (lane & 4) {
                            lo = make_uint4(lo.y, lo.x, lo.w, lo.z);
                            hi = make_uint4(hi.y, hi.x, hi.w, hi.z);
                        }
                        if (lane & 8) {
                            lo = make_uint4(lo.z, lo.w, lo.x, lo.y);
                            hi = make_uint4(hi.z, hi.w, hi.x, hi.y);
                        }
                        uint4 *dst = reinterpret_cast<uint4 *>(s_pair + 8 * i);
                        dst[0] = lo;
                        dst[1] = hi;
                    }
                }
            }
        }
        PhaseRec *recs = a.recs + t_rec_off;
        const long long p0 = (long long) tile * kTile + (long long) sub * kUnit - kPosShift; // scan position of unit-local index 0

        for (uint32_t cb = c_lo; cb < c_hi; cb += 32) { // uniform
            // ---- candidates -> (position, phase) items ----
            uint32_t e = e_first;
            if (cb != c_lo)
                e = (cb + lane < c_hi) ? __ldg(&cand[cb + lane]) : 0u;
            const uint32_t tm = (e >> 13) & 31u;
            const uint32_t ul = (e & 0x1fffu) - sub * kUnit;
            int inc = __popc(tm);
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int up = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o)
                    inc += up;
            }
            const int nitems = __shfl_sync(0xffffffffu, inc, 31);
            int off = inc - __popc(tm);
#pragma unroll
            for (int ph = 0; ph < 5; ++ph)
                if ((tm >> ph) & 1u)
                    s_items[off++] = (uint16_t) (ul | ((uint32_t) ph << 10));
            __syncwarp(); // items and magnitudes are in place

            // ---- 1. DF field -> frame length ----
            int nl = 0, ns = 0;
            for (int it0 = 0; it0 < nitems; it0 += 32) { // uniform
                const bool active = it0 + lane < nitems;
                const uint32_t item = s_items[active ? it0 + lane : 0];
                const uint32_t v5 = slice_group(s_pair, (int) (item & 1023u), (int) (item >> 10), 0, s_taps);
                const int nb = active ? frame_bytes_for_df(__brev(v5) >> 27) : 0;
                const uint32_t lm = __ballot_sync(0xffffffffu, nb == 14), sm = __ballot_sync(0xffffffffu, nb == 7);
                if (nb == 14)
                    s_frames[nl + __popc(lm & below)] = (uint16_t) item;
                else if (nb == 7)
                    s_frames[kItemMax - 1 - (ns + __popc(sm & below))] = (uint16_t) item;
                nl += __popc(lm);
                ns += __popc(sm);
            }
            const int nframes = nl + ns;
            for (int q0 = 0; q0 < nframes; q0 += 32) { // uniform
                const int nb = min(32, nframes - q0);
                const int nlb = max(0, min(nb, nl - q0)); // long frames come first
                const int ntasks = kGroupsLong * nlb + kGroupsShort * (nb - nlb);
                __syncwarp(); // s_frames complete, the previous batch is done with s_bits
                // ---- 2. one lane per (frame, group): five bits into the frame's group byte ----
                for (int t = lane; t < ntasks; t += 32) {
                    int bi, k;
                    const int ts = t - kGroupsLong * nlb;
                    if (ts < 0) {
                        bi = (t * 2850) >> 16; // t / 23 for t < 23 * 32
                        k = t - kGroupsLong * bi;
                    } else {
                        const int sb = (ts * 5462) >> 16; // ts / 12 for ts < 12 * 32
                        bi = nlb + sb;
                        k = ts - kGroupsShort * sb;
                    }
                    const int q = q0 + bi;
                    const uint32_t item = (q < nl) ? s_frames[q] : s_frames[kItemMax - 1 - (q - nl)];
                    uint32_t v5 = slice_group(s_pair, (int) (item & 1023u), (int) (item >> 10), k, s_taps);
                    const int left = (ts < 0 ? 112 : 56) - 5 * k; // the last group of a frame is partial
                    if (left < 5)
                        v5 &= (1u << left) - 1u;
                    s_bits[bi * kFrameBytes + k] = (uint8_t) v5;
                }
                __syncwarp();
                // ---- 3. one lane per frame: message words (frame bit b -> bit b % 32 of word b / 32), CRC
                // syndrome (crc.c:59-64: linear in the bits, so the XOR of the groups' syndromes), class record
                FrameClass fc;
                fc.kind = kKindBad;
                uint32_t syn = 0, item = 0, df = 0, aa = 0;
                uint32_t w[5] = {0, 0, 0, 0, 0}; // the frame's bits (bit b of the frame = bit b % 32 of w[b / 32])
                if (lane < nb) {
                    const bool is_long = lane < nlb;
                    const int ng = is_long ? kGroupsLong : kGroupsShort;
                    const uint32_t *gs = s_gsyn + (is_long ? 0 : kGroupsLong * 32);
                    const uint32_t *bw = reinterpret_cast<const uint32_t *>(s_bits + lane * kFrameBytes);
#pragma unroll
                    for (int kw = 0; kw < 6; ++kw) {
                        const uint32_t four = bw[kw];
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk) {
                            const int k = 4 * kw + kk;
                            if (k < kGroupsLong && k < ng) {
                                const uint32_t v5 = (four >> (8 * kk)) & 31u;
                                syn ^= gs[k * 32 + v5];
                                w[(5 * k) >> 5] |= v5 << ((5 * k) & 31);
                                if (((5 * k) & 31) > 27)
                                    w[((5 * k) >> 5) + 1] |= v5 >> (32 - ((5 * k) & 31));
                            }
                        }
                    }
                    const int q = q0 + lane;
                    item = (q < nl) ? s_frames[q] : s_frames[kItemMax - 1 - (q - nl)];
                    const uint32_t head32 = __brev(w[0]); // frame bits 0..31, MSB first
                    df = head32 >> 27;
                    aa = head32 & 0xffffffu;
                    fc = classify_frame(df, aa, syn, (w[0] | w[1] | w[2] | w[3]) == 0, a.tab_short, a.n_short, a.tab_long, a.n_long, s_bloom);
                }
                const bool has = fc.kind != kKindBad;
                const uint32_t mask = __ballot_sync(0xffffffffu, has);
                if (mask) {
                    const uint32_t cnt = (uint32_t) __popc(mask);
                    uint32_t base = 0;
                    if (lane == 0) {
                        base = atomicAdd(&a.tiles[tile].nrec, cnt); // the tile's units share its slab
                        if (base + cnt > rec_cap)
                            atomicOr(&a.counters->overflow, 2u);
                    }
                    base = __shfl_sync(0xffffffffu, base, 0);
                    nrec_warp += cnt;
                    if (has) {
                        const uint32_t slot = base + __popc(mask & below);
                        if (slot < rec_cap) {
                            // the frame's bytes, MSB first (frame bit b = bit b % 32 of w[b / 32])
                            const uint32_t m0 = __byte_perm(__brev(w[0]), 0, 0x0123), m1 = __byte_perm(__brev(w[1]), 0, 0x0123),
                                           m2 = __byte_perm(__brev(w[2]), 0, 0x0123), m3 = __byte_perm(__brev(w[3]), 0, 0x0123) & 0xffffu;
                            uint4 *dst = reinterpret_cast<uint4 *>(&recs[slot]);
                            dst[0] = make_uint4((uint32_t) (p0 + (item & 1023u)), syn | (fc.kind << 24) | (fc.errors << 28),
                                                fc.key | (((item >> 10) + 4u) << 24), (uint32_t) (uint8_t) fc.bit0 | ((uint32_t) (uint8_t) fc.bit1 << 8));
                            dst[1] = make_uint4(m0, m1, m2, m3);
                        }
                        // mode_s.c:717-726: only a clean DF17, or a clean DF11 with IID 0, can ever be added to
                        // the ICAO filter; remember every such address of the stream
                        if (syn == 0 && (df == 17 || df == 11))
                            atomicOr(&a.addr_bitmap[aa >> 5], 1u << (aa & 31u));
                    }
                }
            }
        }
    }
    if (lane == 0 && nrec_warp) // one same-address atomic per warp, not per unit
        atomicAdd(&a.counters->n_rec, (unsigned long long) nrec_warp);
}

cudaError_t slice_configure() {
    return cudaFuncSetAttribute(slice_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) kSliceSmem);
}

cudaError_t launch_slice(const SliceArgs &a, int grid, cudaStream_t stream) {
    if (a.ntiles == 0)
        return cudaSuccess;
    const uint32_t useful = (a.ntiles * kUnitsPerTile + kSliceWarps - 1) / kSliceWarps;
    if ((uint32_t) grid > useful)
        grid = (int) useful;
    slice_kernel<<<grid, kSliceThreads, kSliceSmem, stream>>>(a);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// K2: classify kernel -- one CTA per tile, position-indexed so that K1's output order is free
// ------------------------------------------------------------------------------------------

constexpr int kClassifyThreads = 256;

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

// index of the candidate entry of tile-local position pl (the entries are in position order)
__device__ __forceinline__ uint32_t find_cand(const uint32_t *__restrict__ cand, uint32_t ncand, uint32_t pl) {
    uint32_t lo = 0, hi = ncand;
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if ((__ldg(&cand[mid]) & 0x1fffu) < pl)
            lo = mid + 1;
        else
            hi = mid;
    }
    return lo;
}

__global__ void __launch_bounds__(kClassifyThreads) classify_kernel(const ClassifyArgs a) {
    // per candidate of the tile (K1 emits them in position order)
    __shared__ __align__(16) uint8_t s_flags[kTile]; // live[0] | has a -1 phase[1]
    __shared__ __align__(16) uint8_t s_nb[kTile];    // which phases own a class record
    __shared__ uint16_t s_slot[kTile];               // first live-record slot of a live candidate
    __shared__ int s_warp[40];
    __shared__ uint32_t s_bd[8];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t tile = blockIdx.x;
    if (a.counters->overflow & 3u)
        return; // K1 ran out of room: the host places the slabs exactly and runs the span again
    const TileDesc td = a.tiles[tile];
    const long long p0 = (long long) tile * kTile - kPosShift; // position of tile-local index 0
    const uint32_t *cand = a.cand + td.cand_off;
    if (td.ncand == 0) {
        if (tid == 0) {
            TileOut to;
            to.dead_off = to.ndead = to.live_off = to.nlive = to.liverec_off = to.nliverec = 0;
            a.tiles_out[tile] = to;
        }
        return;
    }

    if (tid < 8)
        s_bd[tid] = 0;
    for (uint32_t i = tid; i < (td.ncand + 3) / 4; i += kClassifyThreads) {
        reinterpret_cast<uint32_t *>(s_flags)[i] = 0;
        reinterpret_cast<uint32_t *>(s_nb)[i] = 0;
    }
    __syncthreads();

    // ---- pass 1: class records -> can the position still be accepted, does it hold a -1 phase ----
    for (uint32_t r = tid; r < td.nrec; r += kClassifyThreads) {
        const PhaseRec pr = a.recs[td.rec_off + r];
        const uint32_t kind = (pr.w0 >> 24) & 7u;
        const uint32_t c = find_cand(cand, td.ncand, (uint32_t) ((long long) pr.pos - p0));
        const uint32_t ph = (pr.w1 >> 24) & 15u;
        const bool live = record_is_live(pr.w0, pr.w1, a.addr_bitmap);
        uint32_t f = live ? 1u : 0u;
        // static score of a phase whose address can never be in the filter:
        // AP -> -1, DF11 with IID != 0 -> -1, Comm-B -> -2 (mode_s.c:343,373,403)
        if (!live && (kind == kKindAP || kind == kKindDF11))
            f |= 2u;
        if (f)
            atomicOr(reinterpret_cast<unsigned int *>(s_flags) + (c >> 2), f << (8 * (c & 3)));
        atomicOr(reinterpret_cast<unsigned int *>(s_nb) + (c >> 2), (1u << (ph - 4)) << (8 * (c & 3)));
    }
    __syncthreads();

    // ---- pass 2: count, reserve the tile's output ranges ----
    int n_dead = 0, n_live = 0, n_liverec = 0;
    {
        int mine = 0; // dead[9:0] | live[19:10] | live records[31:20], per batch of kClassifyThreads candidates
        for (uint32_t c = tid; c < td.ncand; c += kClassifyThreads) {
            if (s_flags[c] & 1u)
                mine += (1 << 10) + ((int) __popc((uint32_t) s_nb[c]) << 20);
            else
                mine += 1;
        }
        // counts of one thread stay small (<= 32 candidates per thread), sum them per field
        int d = mine & 1023, l = (mine >> 10) & 1023, r = (mine >> 20) & 4095;
        block_exclusive_scan(d, s_warp, n_dead);
        block_exclusive_scan(l, s_warp, n_live);
        block_exclusive_scan(r, s_warp, n_liverec);
    }
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

    // ---- pass 3: ordered dead list / live position list, kClassifyThreads candidates at a time ----
    {
        const long long B = (long long) a.block_samples;
        const long long first_pos = p0 < 0 ? 0 : p0;
        const uint32_t kb0 = (uint32_t) (first_pos / B);
        const bool one_block = ((long long) (kb0 + 1) * B >= p0 + kTile);
        uint32_t bd_local[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        int d_done = 0, l_done = 0, r_done = 0;
        for (uint32_t cb = 0; cb < td.ncand; cb += kClassifyThreads) { // uniform trip count
            const uint32_t c = cb + tid;
            uint32_t e = 0, nrec = 0, fl = 0;
            bool is_dead = false, is_live = false;
            if (c < td.ncand) {
                e = __ldg(&cand[c]);
                fl = s_flags[c];
                is_live = (fl & 1u) != 0;
                is_dead = !is_live;
                nrec = (uint32_t) __popc((uint32_t) s_nb[c]);
            }
            // one scan for the three counts: dead[9:0] | live[19:10] | live records[31:20]
            int tot;
            const int packed = (is_dead ? 1 : 0) | (is_live ? (1 << 10) | ((int) nrec << 20) : 0);
            const int off = block_exclusive_scan(packed, s_warp, tot);
            const int od = off & 1023, ol = (off >> 10) & 1023, orr = (off >> 20) & 4095;
            const uint32_t pl = e & 0x1fffu, tm = (e >> 13) & 31u;
            if (is_live) {
                LivePos lp;
                lp.pos = (uint32_t) (p0 + pl);
                lp.info = tm | (nrec << 8) | ((uint32_t) (r_done + orr) << 16);
                lp.dead_rank = (uint32_t) (d_done + od);
                lp.pad = 0;
                a.live[live_off + l_done + ol] = lp;
                s_slot[c] = (uint16_t) (r_done + orr);
            } else if (is_dead) {
                const uint32_t unknown = (fl >> 1) & 1u;
                a.dead[dead_off + d_done + od] = pl | (tm << 13) | (unknown << 18);
                // what demodulate2400 counts for a position whose best score is negative
                // (demod_2400.c:184,339-347), provided no accepted frame skips over it
                const uint32_t bd[8] = {1u, unknown ? 0u : 1u, unknown, tm & 1u, (tm >> 1) & 1u, (tm >> 2) & 1u, (tm >> 3) & 1u, (tm >> 4) & 1u};
                if (one_block) {
#pragma unroll
                    for (int q = 0; q < 8; ++q)
                        bd_local[q] += bd[q];
                } else {
                    const uint32_t kb = (uint32_t) ((p0 + pl) / B);
                    uint32_t *dst = reinterpret_cast<uint32_t *>(&a.block_dead[kb]);
#pragma unroll
                    for (int q = 0; q < 8; ++q)
                        if (bd[q])
                            atomicAdd(&dst[q], bd[q]);
                }
            }
            d_done += tot & 1023;
            l_done += (tot >> 10) & 1023;
            r_done += (tot >> 20) & 4095;
        }
        if (one_block) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                uint32_t v = bd_local[q];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1)
                    v += __shfl_xor_sync(0xffffffffu, v, o);
                if (lane == 0 && v)
                    atomicAdd(&s_bd[q], v);
            }
            __syncthreads();
            if (tid < 8 && s_bd[tid])
                atomicAdd(reinterpret_cast<uint32_t *>(&a.block_dead[kb0]) + tid, s_bd[tid]);
        }
    }
    if (n_liverec == 0)
        return;
    __syncthreads();

    // ---- pass 4: class records of live positions -> live records: the frame as K1b sliced it + its signal power ----
    // a live position owns consecutive output slots, one per recorded phase in phase order
    for (uint32_t r = warp; r < td.nrec; r += kClassifyThreads / 32) {
        const uint4 *src = reinterpret_cast<const uint4 *>(&a.recs[td.rec_off + r]);
        const uint4 ra = src[0]; // pos, w0, w1, errbits
        const uint32_t c = find_cand(cand, td.ncand, (uint32_t) ((long long) ra.x - p0));
        if (!(s_flags[c] & 1u))
            continue;
        const uint4 rm = src[1]; // the frame's bytes
        const int ph = (int) ((ra.z >> 24) & 15u);
        const uint32_t rank = (uint32_t) __popc((uint32_t) s_nb[c] & ((1u << (ph - 4)) - 1u));
        const uint32_t slot = liverec_off + s_slot[c] + rank;

        // demod_2400.c:387-396: sum of m^2 over msglen*12/5 samples from m[19]; K1a's magnitudes: window position
        // pos starts at magnitude index pos + kPosShift
        const int signal_len = (rm.x & 0x80u) ? 268 : 134;
        const uint16_t *fm = a.mag + (size_t) ra.x + kPosShift + 19;
        unsigned long long power = 0;
        for (int k = lane; k < signal_len; k += 32) {
            const unsigned long long v = fm[k];
            power += v * v;
        }
        power = warp_sum_u64(power);
        if (lane == 0) {
            const uint32_t w0 = ra.y | (bitmap_test(a.addr_bitmap, ra.z & 0xffffffu) ? 0x80000000u : 0u); // bit 31: key in S
            uint2 *dst = reinterpret_cast<uint2 *>(&a.liverecs[slot]);
            dst[0] = make_uint2(ra.x, w0);
            dst[1] = make_uint2(ra.z, ra.w);
            dst[2] = make_uint2((uint32_t) power, (uint32_t) (power >> 32));
            dst[3] = make_uint2(rm.x, rm.y);
            dst[4] = make_uint2(rm.z, rm.w & 0xffffu);
        }
    }
}

// ------------------------------------------------------------------------------------------
// K2, warp-per-tile variant: the same passes as classify_kernel with no block-wide barrier, for the
// normal case of fixed per-tile slabs (at most kCwCap candidates per tile).  A warp keeps the tile's
// candidate entries and per-candidate flags in its own shared memory; a survivor's record is K1b's
// (frame bytes included), its signal power comes straight from K1a's magnitude array.
// ------------------------------------------------------------------------------------------

constexpr int kCwWarps = 8;
constexpr int kCwCap = 640; // candidates per tile this kernel can hold (== the default candidate slab)

__device__ __forceinline__ uint32_t find_cand_smem(const uint32_t *cand, uint32_t ncand, uint32_t pl) {
    uint32_t lo = 0, hi = ncand;
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if ((cand[mid] & 0x1fffu) < pl)
            lo = mid + 1;
        else
            hi = mid;
    }
    return lo;
}

__global__ void __launch_bounds__(kCwWarps * 32) classify_warp_kernel(const ClassifyArgs a) {
    __shared__ uint32_t s_cand[kCwWarps][kCwCap];
    __shared__ uint32_t s_flagw[kCwWarps][kCwCap / 4]; // per candidate byte: live[0] | has a -1 phase[1]
    __shared__ uint32_t s_nbw[kCwWarps][kCwCap / 4];   // per candidate byte: which phases own a class record
    __shared__ uint16_t s_slot[kCwWarps][kCwCap];      // first live-record slot of a live candidate

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (a.counters->overflow & 3u)
        return; // K1 ran out of room: the host places the slabs exactly and runs the span again

    uint32_t *cand = s_cand[warp];
    uint32_t *flagw = s_flagw[warp], *nbw = s_nbw[warp];
    const uint8_t *flags = reinterpret_cast<const uint8_t *>(flagw), *nbs = reinterpret_cast<const uint8_t *>(nbw);
    uint16_t *slot = s_slot[warp];
    const uint32_t below = (1u << lane) - 1u;
    const long long B = (long long) a.block_samples;

    for (uint32_t tile = blockIdx.x * kCwWarps + warp; tile < a.ntiles; tile += gridDim.x * kCwWarps) {
        const TileDesc td = a.tiles[tile];
        const long long p0 = (long long) tile * kTile - kPosShift; // position of tile-local index 0
        __syncwarp();
        if (td.ncand == 0) {
            if (lane == 0) {
                TileOut to;
                to.dead_off = to.ndead = to.live_off = to.nlive = to.liverec_off = to.nliverec = 0;
                a.tiles_out[tile] = to;
            }
            continue;
        }
        const uint32_t ncand = td.ncand; // <= kCwCap: K1a's slab
        for (uint32_t i = lane; i < ncand; i += 32)
            cand[i] = __ldg(&a.cand[td.cand_off + i]);
        for (uint32_t i = lane; i < (ncand + 3) / 4; i += 32) {
            flagw[i] = 0;
            nbw[i] = 0;
        }
        __syncwarp();

        // ---- pass 1: class records -> can the position still be accepted, does it hold a -1 phase ----
        for (uint32_t r = lane; r < td.nrec; r += 32) {
            const PhaseRec pr = a.recs[td.rec_off + r];
            const uint32_t kind = (pr.w0 >> 24) & 7u;
            const uint32_t c = find_cand_smem(cand, ncand, (uint32_t) ((long long) pr.pos - p0));
            const uint32_t ph = (pr.w1 >> 24) & 15u;
            const bool live = record_is_live(pr.w0, pr.w1, a.addr_bitmap);
            uint32_t f = live ? 1u : 0u;
            // static score of a phase whose address can never be in the filter:
            // AP -> -1, DF11 with IID != 0 -> -1, Comm-B -> -2 (mode_s.c:343,373,403)
            if (!live && (kind == kKindAP || kind == kKindDF11))
                f |= 2u;
            if (f)
                atomicOr(&flagw[c >> 2], f << (8 * (c & 3)));
            atomicOr(&nbw[c >> 2], (1u << (ph - 4)) << (8 * (c & 3)));
        }
        __syncwarp();

        // ---- pass 2: count, reserve the tile's output ranges ----
        uint32_t n_dead = 0, n_live = 0, n_liverec = 0;
        for (uint32_t cb = 0; cb < ncand; cb += 32) {
            const uint32_t c = cb + lane;
            const bool valid = c < ncand;
            const bool is_live = valid && (flags[valid ? c : 0] & 1u);
            const uint32_t lm = __ballot_sync(0xffffffffu, is_live);
            n_live += __popc(lm);
            n_dead += __popc(__ballot_sync(0xffffffffu, valid)) - __popc(lm);
            n_liverec += __reduce_add_sync(0xffffffffu, is_live ? (uint32_t) __popc((uint32_t) nbs[c]) : 0u);
        }
        uint32_t dead_off = 0, live_off = 0, liverec_off = 0, ovf = 0;
        if (lane == 0) {
            const unsigned long long d_off = atomicAdd(&a.counters->n_dead, (unsigned long long) n_dead);
            const unsigned long long l_off = atomicAdd(&a.counters->n_live, (unsigned long long) n_live);
            const unsigned long long r_off = atomicAdd(&a.counters->n_liverec, (unsigned long long) n_liverec);
            if (d_off + n_dead > a.dead_cap)
                ovf |= 4u;
            if (l_off + n_live > a.live_cap)
                ovf |= 8u;
            if (r_off + n_liverec > a.liverec_cap)
                ovf |= 16u;
            if (ovf)
                atomicOr(&a.counters->overflow, ovf);
            TileOut to;
            to.dead_off = (uint32_t) d_off;
            to.ndead = n_dead;
            to.live_off = (uint32_t) l_off;
            to.nlive = n_live;
            to.liverec_off = (uint32_t) r_off;
            to.nliverec = n_liverec;
            a.tiles_out[tile] = to;
            dead_off = to.dead_off;
            live_off = to.live_off;
            liverec_off = to.liverec_off;
        }
        ovf = __shfl_sync(0xffffffffu, ovf, 0);
        if (ovf)
            continue; // the host grows the buffers and runs the span again
        dead_off = __shfl_sync(0xffffffffu, dead_off, 0);
        live_off = __shfl_sync(0xffffffffu, live_off, 0);
        liverec_off = __shfl_sync(0xffffffffu, liverec_off, 0);

        // ---- pass 3: ordered dead list / live position list ----
        {
            const long long first_pos = p0 < 0 ? 0 : p0;
            const uint32_t kb0 = (uint32_t) (first_pos / B);
            const bool one_block = ((long long) (kb0 + 1) * B >= p0 + kTile);
            uint32_t bd_local[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            uint32_t d_done = 0, l_done = 0, r_done = 0;
            for (uint32_t cb = 0; cb < ncand; cb += 32) {
                const uint32_t c = cb + lane;
                const bool valid = c < ncand;
                const uint32_t e = valid ? cand[c] : 0u, fl = valid ? flags[c] : 0u;
                const bool is_live = valid && (fl & 1u), is_dead = valid && !is_live;
                const uint32_t nrec = is_live ? (uint32_t) __popc((uint32_t) nbs[c]) : 0u;
                const uint32_t dm = __ballot_sync(0xffffffffu, is_dead), lm = __ballot_sync(0xffffffffu, is_live);
                const uint32_t pl = e & 0x1fffu, tm = (e >> 13) & 31u;
                if (lm) {
                    uint32_t inc = nrec; // live records in front of this candidate (rare path)
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const uint32_t up = __shfl_up_sync(0xffffffffu, inc, o);
                        if (lane >= o)
                            inc += up;
                    }
                    if (is_live) {
                        const uint32_t rs = r_done + inc - nrec;
                        LivePos lp;
                        lp.pos = (uint32_t) (p0 + pl);
                        lp.info = tm | (nrec << 8) | (rs << 16);
                        lp.dead_rank = d_done + __popc(dm & below);
                        lp.pad = 0;
                        a.live[live_off + l_done + __popc(lm & below)] = lp;
                        slot[c] = (uint16_t) rs;
                    }
                    r_done += __shfl_sync(0xffffffffu, inc, 31);
                }
                if (is_dead) {
                    const uint32_t unknown = (fl >> 1) & 1u;
                    a.dead[dead_off + d_done + __popc(dm & below)] = pl | (tm << 13) | (unknown << 18);
                    // what demodulate2400 counts for a position whose best score is negative
                    // (demod_2400.c:184,339-347), provided no accepted frame skips over it
                    const uint32_t bd[8] = {1u, unknown ? 0u : 1u, unknown, tm & 1u, (tm >> 1) & 1u, (tm >> 2) & 1u, (tm >> 3) & 1u, (tm >> 4) & 1u};
                    if (one_block) {
#pragma unroll
                        for (int q = 0; q < 8; ++q)
                            bd_local[q] += bd[q];
                    } else {
                        const uint32_t kb = (uint32_t) ((p0 + pl) / B);
                        uint32_t *dst = reinterpret_cast<uint32_t *>(&a.block_dead[kb]);
#pragma unroll
                        for (int q = 0; q < 8; ++q)
                            if (bd[q])
                                atomicAdd(&dst[q], bd[q]);
                    }
                }
                d_done += __popc(dm);
                l_done += __popc(lm);
            }
            if (one_block) {
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const uint32_t v = __reduce_add_sync(0xffffffffu, bd_local[q]);
                    if (lane == q && v)
                        atomicAdd(reinterpret_cast<uint32_t *>(&a.block_dead[kb0]) + q, v);
                }
            }
        }
        if (n_liverec == 0)
            continue;
        __syncwarp();

        // ---- pass 4: class records of live positions -> live records: the frame as K1b sliced it + its signal power ----
        // a live position owns consecutive output slots, one per recorded phase in phase order.  Four records at a
        // time, eight lanes each: demod_2400.c:387-396, the sum of m^2 over msglen * 12 / 5 samples from m[19].
        for (uint32_t rb = 0; rb < td.nrec; rb += 32) {
            const uint32_t r = rb + lane;
            uint4 ra = make_uint4(0, 0, 0, 0), rm = make_uint4(0, 0, 0, 0);
            uint32_t out = 0;
            bool mine = false;
            if (r < td.nrec) {
                const uint4 *src = reinterpret_cast<const uint4 *>(&a.recs[td.rec_off + r]);
                ra = src[0]; // pos, w0, w1, errbits
                const uint32_t c = find_cand_smem(cand, ncand, (uint32_t) ((long long) ra.x - p0));
                mine = flags[c] & 1u;
                if (mine) {
                    rm = src[1]; // the frame's bytes
                    const uint32_t ph = (ra.z >> 24) & 15u;
                    out = liverec_off + slot[c] + (uint32_t) __popc((uint32_t) nbs[c] & ((1u << (ph - 4)) - 1u));
                }
            }
            uint32_t todo = __ballot_sync(0xffffffffu, mine);
            unsigned long long my_power = 0;
            while (todo) {
                // group g = lane / 8 takes the g-th record still to do (if there is one)
                const int g = lane >> 3, sub = lane & 7;
                uint32_t t = todo;
                int src = -1;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int s0 = t ? __ffs(t) - 1 : -1;
                    if (k == g)
                        src = s0;
                    t &= t - 1;
                }
                todo = t;
                const int from = src < 0 ? 0 : src;
                const uint32_t pos = __shfl_sync(0xffffffffu, ra.x, from);
                const uint32_t first = __shfl_sync(0xffffffffu, rm.x, from); // msg[0..3]: the DF bit 0x80 of msg[0] decides the length
                unsigned long long power = 0;
                if (src >= 0) {
                    const int signal_len = (first & 0x80u) ? 268 : 134;
                    // K1a's magnitudes: window position pos starts at magnitude index pos + kPosShift
                    const uint16_t *fm = a.mag + (size_t) pos + kPosShift + 19;
                    for (int k = sub; k < signal_len; k += 8) {
                        const unsigned long long v = fm[k];
                        power += v * v;
                    }
                }
                power += __shfl_xor_sync(0xffffffffu, power, 4);
                power += __shfl_xor_sync(0xffffffffu, power, 2);
                power += __shfl_xor_sync(0xffffffffu, power, 1);
                // hand the sums back to the lanes that own the records
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int owner = __shfl_sync(0xffffffffu, src, 8 * k);
                    const unsigned long long pk = __shfl_sync(0xffffffffu, power, 8 * k);
                    if (owner == lane)
                        my_power = pk;
                }
            }
            if (mine) {
                // LiveRec: pos, w0 (+ bit 31: the key is in S), w1, errbits, power, msg[14]
                const uint32_t w0 = ra.y | (bitmap_test(a.addr_bitmap, ra.z & 0xffffffu) ? 0x80000000u : 0u);
                uint2 *dst = reinterpret_cast<uint2 *>(&a.liverecs[out]);
                dst[0] = make_uint2(ra.x, w0);
                dst[1] = make_uint2(ra.z, ra.w);
                dst[2] = make_uint2((uint32_t) my_power, (uint32_t) (my_power >> 32));
                dst[3] = make_uint2(rm.x, rm.y);
                dst[4] = make_uint2(rm.z, rm.w & 0xffffu);
            }
        }
    }
}

cudaError_t launch_classify(const ClassifyArgs &a, cudaStream_t stream) {
    if (a.ntiles == 0)
        return cudaSuccess;
    if (a.mag && a.max_cand_per_tile <= (uint32_t) kCwCap) {
        int grid = (int) ((a.ntiles + kCwWarps - 1) / kCwWarps);
        if (grid > 148 * 5)
            grid = 148 * 5;
        classify_warp_kernel<<<grid, kCwWarps * 32, 0, stream>>>(a);
    } else {
        classify_kernel<<<a.ntiles, kClassifyThreads, 0, stream>>>(a);
    }
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// order_live: K2's per-tile live lists -> flat arrays in stream order, packed in device memory (cabi.cu DMAs them)
//
// K2 reserves a tile's slice of the live-position and live-record lists with atomics, so the slices lie in
// the order the warps got there.  The host walks the positions in stream order; handing it one
// position-ordered array (and the records in the same order) turns that walk into two sequential reads:
// live_offsets_kernel (one CTA) scans the tiles' counts, live_gather_kernel (a warp per tile) copies each
// tile's slice to its place and rewrites a position's record index from tile-relative to absolute.
// ------------------------------------------------------------------------------------------

__global__ void __launch_bounds__(1024) live_offsets_kernel(const TileOut *__restrict__ tiles_out, uint32_t ntiles,
                                                             const ScanCounters *__restrict__ counters, uint2 *__restrict__ base) {
    __shared__ uint32_t s_l[32], s_r[32];
    __shared__ uint32_t s_carry[2];
    if (counters->overflow) // the host re-runs the chunk with larger buffers
        return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0)
        s_carry[0] = s_carry[1] = 0;
    __syncthreads();
    for (uint32_t t0 = 0; t0 < ntiles; t0 += 1024) {
        const uint32_t t = t0 + tid;
        uint32_t l = 0, r = 0;
        if (t < ntiles) {
            l = tiles_out[t].nlive;
            r = tiles_out[t].nliverec;
        }
        uint32_t il = l, ir = r; // inclusive warp scans
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t ul = __shfl_up_sync(0xffffffffu, il, o), ur = __shfl_up_sync(0xffffffffu, ir, o);
            if (lane >= o) {
                il += ul;
                ir += ur;
            }
        }
        if (lane == 31) {
            s_l[warp] = il;
            s_r[warp] = ir;
        }
        __syncthreads();
        if (warp == 0) {
            uint32_t wl = s_l[lane], wr = s_r[lane];
            uint32_t xl = wl, xr = wr;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t ul = __shfl_up_sync(0xffffffffu, xl, o), ur = __shfl_up_sync(0xffffffffu, xr, o);
                if (lane >= o) {
                    xl += ul;
                    xr += ur;
                }
            }
            s_l[lane] = xl - wl; // exclusive prefix of the warps' totals
            s_r[lane] = xr - wr;
        }
        __syncthreads();
        const uint32_t cl = s_carry[0], cr = s_carry[1];
        if (t < ntiles)
            base[t] = make_uint2(cl + s_l[warp] + il - l, cr + s_r[warp] + ir - r);
        __syncthreads();
        if (tid == 1023) {
            s_carry[0] = cl + s_l[31] + il;
            s_carry[1] = cr + s_r[31] + ir;
        }
        __syncthreads();
    }
}

// What a frame accepted at live position `pos` hides from the per-block dead totals (LiveHidden): the dead
// entries in (pos, pos + 134] and (pos, pos + 268], both cut at the last position of pos's mag_buf.  One walk
// over the tile's ordered dead list from the position's rank (and into the following tiles when the frame
// body crosses a tile boundary), a snapshot taken where the short body ends.
__device__ __forceinline__ LiveHidden hidden_counts(const TileOut *__restrict__ tiles_out, uint32_t ntiles, const uint32_t *__restrict__ dead,
                                                    uint32_t tile, uint32_t pos, uint32_t rank, uint64_t nsamples, uint32_t block_samples) {
    const unsigned long long p = pos, B = block_samples;
    unsigned long long b1 = (p / B + 1) * B;
    if (b1 > nsamples)
        b1 = nsamples;
    const unsigned long long end_s = (p + 134 < b1 - 1) ? p + 134 : b1 - 1, end_l = (p + 268 < b1 - 1) ? p + 268 : b1 - 1;
    unsigned long long lo = 0, hi = 0;
    LiveHidden h;
    h.short_lo = h.short_hi = 0;
    bool snap = false;
    if (end_l > p) {
        const uint32_t t1 = (uint32_t) ((end_l + kPosShift) / kTile);
        for (uint32_t t = tile; t <= t1 && t < ntiles; ++t) {
            const TileOut to = tiles_out[t];
            const long long base = (long long) t * kTile - kPosShift;
            const uint32_t *it = dead + to.dead_off + (t == tile ? rank : 0u), *dend = dead + to.dead_off + to.ndead;
            const long long last_l = (long long) end_l - base, last_s = (long long) end_s - base;
            for (; it != dend; ++it) {
                const uint32_t e = __ldg(it);
                const long long pl = (long long) (e & 0x1fffu);
                if (pl > last_l)
                    break;
                if (!snap && pl > last_s) {
                    h.short_lo = lo;
                    h.short_hi = hi;
                    snap = true;
                }
                const uint32_t tm = (e >> 13) & 31u, unk = (e >> 18) & 1u;
                lo += 1ull | ((unsigned long long) (unk ^ 1u) << 16) | ((unsigned long long) unk << 32) | ((unsigned long long) (tm & 1u) << 48);
                hi += (unsigned long long) ((tm >> 1) & 1u) | ((unsigned long long) ((tm >> 2) & 1u) << 16) |
                      ((unsigned long long) ((tm >> 3) & 1u) << 32) | ((unsigned long long) ((tm >> 4) & 1u) << 48);
            }
        }
    }
    if (!snap) {
        h.short_lo = lo;
        h.short_hi = hi;
    }
    h.long_lo = lo;
    h.long_hi = hi;
    return h;
}

__global__ void __launch_bounds__(256) live_gather_kernel(const TileOut *__restrict__ tiles_out, uint32_t ntiles,
                                                           const ScanCounters *__restrict__ counters, const uint2 *__restrict__ base,
                                                           const LivePos *__restrict__ live, const LiveRec *__restrict__ recs,
                                                           const uint32_t *__restrict__ dead, uint64_t nsamples, uint32_t block_samples,
                                                           uint8_t *__restrict__ packed) {
    if (counters->overflow)
        return;
    // [n_live LivePos | n_live LiveHidden | n_liverec LiveRec], back to back: one download of exactly these bytes
    const size_t n_live_total = (size_t) counters->n_live;
    LivePos *__restrict__ live_out = reinterpret_cast<LivePos *>(packed);
    LiveHidden *__restrict__ hidden_out = reinterpret_cast<LiveHidden *>(packed + n_live_total * sizeof(LivePos));
    LiveRec *__restrict__ recs_out = reinterpret_cast<LiveRec *>(packed + n_live_total * (sizeof(LivePos) + sizeof(LiveHidden)));
    const int lane = threadIdx.x & 31;
    const uint32_t warps = gridDim.x * (blockDim.x >> 5);
    for (uint32_t t = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); t < ntiles; t += warps) {
        const TileOut to = tiles_out[t];
        if (!to.nlive)
            continue;
        const uint2 b = base[t];
        for (uint32_t i = lane; i < to.nlive; i += 32) {
            LivePos lp = live[to.live_off + i];
            const LiveHidden h = hidden_counts(tiles_out, ntiles, dead, t, lp.pos, lp.dead_rank, nsamples, block_samples);
            lp.pad = b.y + (lp.info >> 16); // first record of the position, now an absolute index
            live_out[b.x + i] = lp;
            uint4 *ho = reinterpret_cast<uint4 *>(hidden_out + b.x + i);
            ho[0] = make_uint4((uint32_t) h.short_lo, (uint32_t) (h.short_lo >> 32), (uint32_t) h.short_hi, (uint32_t) (h.short_hi >> 32));
            ho[1] = make_uint4((uint32_t) h.long_lo, (uint32_t) (h.long_lo >> 32), (uint32_t) h.long_hi, (uint32_t) (h.long_hi >> 32));
        }
        // records: whole 8-byte units (the struct holds a uint64_t, so every record starts on one)
        static_assert(sizeof(LiveRec) % 8 == 0 && alignof(LiveRec) == 8, "LiveRec is copied in 8-byte units");
        constexpr uint32_t kUnits = sizeof(LiveRec) / 8;
        const uint2 *src = reinterpret_cast<const uint2 *>(recs + to.liverec_off);
        uint2 *dst = reinterpret_cast<uint2 *>(recs_out + b.y);
        for (uint32_t i = lane; i < kUnits * to.nliverec; i += 32)
            dst[i] = src[i];
    }
}

cudaError_t launch_order_live(const TileOut *tiles_out, uint32_t ntiles, const ScanCounters *counters, uint2 *base, const LivePos *live,
                              const LiveRec *recs, const uint32_t *dead, uint64_t nsamples, uint32_t block_samples, uint8_t *packed,
                              cudaStream_t stream) {
    if (ntiles == 0)
        return cudaSuccess;
    live_offsets_kernel<<<1, 1024, 0, stream>>>(tiles_out, ntiles, counters, base);
    int grid = (int) ((ntiles + 7) / 8);
    if (grid > 148 * 4)
        grid = 148 * 4;
    live_gather_kernel<<<grid, 256, 0, stream>>>(tiles_out, ntiles, counters, base, live, recs, dead, nsamples, block_samples, packed);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// float_block_sums_kernel: mean_level / mean_power of the float converters, bit-exactly
//
// convert_sc16_nodc / convert_sc16q11_nodc add mag and magsq of every sample to two float
// accumulators in stream order (convert.c:228,241-242 / 345,358-359).  Float addition is not
// associative -- but while the running sum S stays inside one binade [2^k, 2^(k+1)) every step
// S <- RN(S + x) moves it by a whole number of its (fixed) ulps u = 2^(k-23):
//     RN(S + x) = S + RN_u(x),   RN_u(x) = x rounded to the nearest multiple of u,
// unless x falls exactly half-way between two multiples (the tie then goes to the even neighbour of
// S + x, which depends on S).  RN_u(x) does not depend on S, so a batch of samples is rounded in
// parallel -- the hardware does it: bits(x + 2^k) - bits(2^k) is RN_u(x) / u when x < 2^k -- the
// batch's counts are added up as integers, and S moves by their total (an integer add on the float's
// bits).  A batch that holds a tie, or that would carry S into the next binade, or that starts below
// S = 4 (x <= 1 < 2^k must hold), is added one sample after the other instead; that happens about ten
// times per mag_buf and sum (one per binade S passes through, plus the first batch).
// One CTA of four warps per mag_buf, 512 samples per batch, two barriers per batch (the batch's
// total has to be in S before the next batch knows its binade).
// ------------------------------------------------------------------------------------------

constexpr int kSumBatch = 512;

__global__ void __launch_bounds__(128) float_block_sums_kernel(const uint8_t *__restrict__ iq, uint32_t format, uint64_t nsamples,
                                                                uint32_t block_samples, uint32_t nblocks, double *__restrict__ sums) {
    __shared__ __align__(16) float s_val[2][kSumBatch]; // [0 = mag, 1 = magsq][sample]: read only by the sequential fall-back
    __shared__ uint32_t s_part[2][4];                   // [sum][warp]: the warp's total of RN_u(x) / u
    __shared__ uint32_t s_tie[2][4];                    // [sum][warp]: a tie in the warp's share
    __shared__ uint32_t s_S[2];                         // the two running sums (float bits)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t k = blockIdx.x;
    const uint64_t b0 = (uint64_t) k * block_samples;
    const uint64_t nk = nsamples > b0 ? (nsamples - b0 < block_samples ? nsamples - b0 : block_samples) : 0;
    const float inv_scale = (format == 1) ? (1.0f / 32768.0f) : (1.0f / 2048.0f);
    const uint32_t *src = reinterpret_cast<const uint32_t *>(iq) + b0; // one 32-bit word per sample

    // 4 samples per thread (block_samples % 8 == 0 and 16-byte aligned spans: whole uint4s except in the stream's
    // ragged last batch); a sample past the end reads as zero IQ = magnitude +0, which changes no sum
    auto load = [&](uint64_t base) {
        uint4 v = make_uint4(0, 0, 0, 0);
        const uint64_t s0 = base + 4 * (uint64_t) tid;
        if (s0 + 4 <= nk) {
            v = ldg_stream(reinterpret_cast<const uint4 *>(src + s0));
        } else if (s0 < nk) {
            uint32_t w[4] = {0, 0, 0, 0};
            for (int j = 0; j < 4; ++j)
                if (s0 + j < nk)
                    w[j] = __ldg(src + s0 + j);
            v = make_uint4(w[0], w[1], w[2], w[3]);
        }
        return v;
    };

    const uint64_t nbatches = (nk + kSumBatch - 1) / kSumBatch;
    uint4 ahead[3] = {load(0), load(kSumBatch), load(2 * kSumBatch)};
    if (tid < 2)
        s_S[tid] = 0; // +0.0f
    __syncthreads();
    for (uint64_t b = 0; b < nbatches; ++b) {
        const uint4 v = ahead[0];
        ahead[0] = ahead[1];
        ahead[1] = ahead[2];
        ahead[2] = load((b + 3) * kSumBatch);
        float x[2][4];
        mag_sc16_word(v.x, inv_scale, x[1][0], x[0][0]);
        mag_sc16_word(v.y, inv_scale, x[1][1], x[0][1]);
        mag_sc16_word(v.z, inv_scale, x[1][2], x[0][2]);
        mag_sc16_word(v.w, inv_scale, x[1][3], x[0][3]);
#pragma unroll
        for (int q = 0; q < 2; ++q)
            *reinterpret_cast<float4 *>(&s_val[q][4 * tid]) = make_float4(x[q][0], x[q][1], x[q][2], x[q][3]);
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const uint32_t Sb = s_S[q];
            const uint32_t Mb = Sb & 0x7f800000u;                  // 2^k
            const float M = __uint_as_float(Mb);
            const float half_u = __uint_as_float(Mb - (24u << 23)); // 2^(k-24); garbage below S = 4, where the batch is redone anyway
            uint32_t acc = 0;
            bool tie = false;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float z = __fadd_rn(x[q][j], M);                       // M + RN_u(x)
                acc += __float_as_uint(z) - Mb;                              // RN_u(x) / u
                const float err = __fsub_rn(x[q][j], __fsub_rn(z, M));       // x - RN_u(x), exact
                tie |= fabsf(err) == half_u;
            }
            const uint32_t total = __reduce_add_sync(0xffffffffu, acc);
            const bool any_tie = __any_sync(0xffffffffu, tie);
            if (lane == 0) {
                s_part[q][warp] = total;
                s_tie[q][warp] = any_tie ? 1u : 0u;
            }
        }
        __syncthreads();
        if (lane == 0 && warp < 2) { // thread 0 owns the level sum, thread 32 the power sum
            const int q = warp;
            const uint32_t Sb = s_S[q];
            const uint32_t T = s_part[q][0] + s_part[q][1] + s_part[q][2] + s_part[q][3];
            const bool tie = (s_tie[q][0] | s_tie[q][1] | s_tie[q][2] | s_tie[q][3]) != 0;
            // x <= 1 (convert.c:236-237 clamps magsq), so below 2^21 counts per sample and 2^30 per batch once S >= 4
            if (Sb >= 0x40800000u /* 4.0f */ && !tie && (Sb & 0x007fffffu) + T < (1u << 23)) {
                s_S[q] = Sb + T;
            } else {
                float accf = __uint_as_float(Sb);
                const float4 *vals = reinterpret_cast<const float4 *>(s_val[q]);
#pragma unroll 1
                for (int h = 0; h < kSumBatch / 128; ++h) {
                    float4 r[32]; // all loads first: they do not depend on the chain
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                        r[i] = vals[32 * h + i];
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        accf = __fadd_rn(accf, r[i].x);
                        accf = __fadd_rn(accf, r[i].y);
                        accf = __fadd_rn(accf, r[i].z);
                        accf = __fadd_rn(accf, r[i].w);
                    }
                }
                s_S[q] = __float_as_uint(accf);
            }
        }
        __syncthreads();
    }
    if (tid < 2)
        sums[2 * k + tid] = (double) __uint_as_float(s_S[tid]); // exactly the float the reference divides by nsamples
}

cudaError_t launch_float_block_sums(const uint8_t *iq, uint32_t format, uint64_t nsamples, uint32_t block_samples, uint32_t nblocks,
                                    double *sums, cudaStream_t stream) {
    if (nblocks == 0 || format == 0 || format == 4)
        return cudaSuccess;
    float_block_sums_kernel<<<nblocks, 128, 0, stream>>>(iq, format, nsamples, block_samples, nblocks, sums);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// Mode A/C (demodulate2400AC, demod_2400.c:522-708): F1/F2 framing-pulse search over K1a's magnitudes
//
// Whether a reply with its first framing pulse at data index f1 of a mag_buf decodes is a pure
// function of the magnitudes and of the block's noise level, so every position is tested
// independently (one thread each; the first two comparisons reject almost all of them) and the hits
// go to an unordered list.  The only sequential part -- skipping 69 samples past an accepted reply
// (demod_2400.c:707) -- is left to the host, over the sorted hits.
// Float steps are spelled out (no FMA, round-to-nearest) exactly as the scalar C code takes them.
// ------------------------------------------------------------------------------------------

// demod_2400.c:529-530 from the block's converter sums (convert.c:104-110 / 246-252)
__global__ void modeac_noise_kernel(const ModeacArgs a) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= a.nblocks)
        return;
    const unsigned long long B = a.block_samples, b0 = (unsigned long long) k * B;
    const unsigned nk = (unsigned) (a.nsamples > b0 ? (a.nsamples - b0 < B ? a.nsamples - b0 : B) : 0);
    double mean_level, mean_power;
    if (a.format == 0 || a.format == 4) {
        mean_level = __ddiv_rn(__ddiv_rn((double) a.sums_u64[2 * k], 65536.0), (double) nk); // sic: 65536
        mean_power = __ddiv_rn(__ddiv_rn(__ddiv_rn((double) a.sums_u64[2 * k + 1], 65535.0), 65535.0), (double) nk);
    } else {
        mean_level = (double) __fdiv_rn((float) a.sums_f64[2 * k], (float) nk);
        mean_power = (double) __fdiv_rn((float) a.sums_f64[2 * k + 1], (float) nk);
    }
    const double noise_stddev = __dsqrt_rn(__dsub_rn(mean_power, __dmul_rn(mean_level, mean_level)));
    a.noise_level[k] = nk ? __double2uint_rz(__dadd_rn(__dmul_rn(__dadd_rn(mean_power, noise_stddev), 65535.0), 0.5)) : 0u;
}

// a framing pulse at m[0]: rising edge, quiet third sample, 6 dB above noise (demod_2400.c:577-588, 604-614)
__device__ __forceinline__ bool ac_framing_pulse(const uint16_t *m, uint32_t noise_level, uint32_t &level) {
    const uint32_t m0 = m[0], m1 = m[1], m2 = m[2];
    if (!(m[-1] < m0) || m2 > m0 || m2 > m1)
        return false;
    level = (m0 + m1) >> 1;
    return !(noise_level * 2u > level);
}

__global__ void __launch_bounds__(256) modeac_kernel(const ModeacArgs a) {
    const uint32_t B = a.block_samples;
    const uint32_t lane = threadIdx.x & 31;
    // q = block * B + data index of F1 (data[0] of a block is 326 samples before its first new sample);
    // the loop is warp-uniform so that the hit list can be appended to with one atomic per warp
    const unsigned long long total = a.nsamples, stride = (unsigned long long) gridDim.x * blockDim.x;
    for (unsigned long long base = (unsigned long long) blockIdx.x * blockDim.x + (threadIdx.x & ~31u); base < total; base += stride) {
        const unsigned long long q = base + lane;
        bool hit = false;
        uint32_t f1_clock = 0, modeac = 0;
        const uint32_t k = (uint32_t) (q / B), f1 = (uint32_t) (q - (unsigned long long) k * B);
        if (q < total && f1 >= 1) { // demod_2400.c:532: f1_sample runs from 1 to mlen - 1
            const uint16_t *d = a.mag + (size_t) k * B + kPosShift; // data[] of the block
            const uint32_t noise_level = a.noise_level[k];
            uint32_t f1_level, f2_level;
            if (ac_framing_pulse(d + f1, noise_level, f1_level)) {
                // clock phase from the share of power in the second sample (:593-596)
                const float fa = __uint2float_rn(d[f1]), fb = __uint2float_rn(d[f1 + 1]);
                const float pa = __fmul_rn(fa, fa), pb = __fmul_rn(fb, fb);
                const float fraction = __fdiv_rn(pb, __fadd_rn(pa, pb));
                const float pos = __fmul_rn(25.0f, __fadd_rn(__uint2float_rn(f1), __fmul_rn(fraction, fraction)));
                f1_clock = __double2uint_rz(__dadd_rn((double) pos, 0.5));
                const uint32_t f2_clock = f1_clock + 87 * 14; // :600
                if (ac_framing_pulse(d + f2_clock / 25, noise_level, f2_level)) {
                    const uint32_t top = f1_level > f2_level ? f1_level : f2_level;
                    const float midpoint = __fsqrt_rn(__uint2float_rn(noise_level * top)); // :618
                    const uint32_t signal_threshold = __double2uint_rz(__dadd_rn(__dmul_rn((double) midpoint, 1.41421356237309504880), 0.5));
                    const uint32_t noise_threshold = __double2uint_rz(__dadd_rn(__ddiv_rn((double) midpoint, 1.41421356237309504880), 0.5));
                    uint32_t bits = 0;
                    bool bad = false; // noisy or uncertain (:629-651, :664)
                    uint32_t clock = f1_clock;
                    for (int bit = 0; bit < 20; ++bit, clock += 87) {
                        const uint16_t *p = d + clock / 25;
                        const uint32_t p0 = p[0], p1 = p[1], p2 = p[2];
                        bits <<= 1;
                        if (p2 >= signal_threshold)
                            bad = true;
                        if (p0 >= signal_threshold || p1 >= signal_threshold)
                            bits |= 1u;
                        else if (p0 > noise_threshold && p1 > noise_threshold)
                            bad = true;
                    }
                    if ((bits & 0x80020u) == 0x80020u && (bits & 0x0101Bu) == 0 && !bad) {
                        // 00 A4 A2 A1  00 B4 B2 B1  SPI C4 C2 C1  00 D4 D2 D1 (:670-683)
                        modeac = ((bits & 0x40000u) ? 0x0010u : 0) | ((bits & 0x20000u) ? 0x1000u : 0) | ((bits & 0x10000u) ? 0x0020u : 0) |
                                 ((bits & 0x08000u) ? 0x2000u : 0) | ((bits & 0x04000u) ? 0x0040u : 0) | ((bits & 0x02000u) ? 0x4000u : 0) |
                                 ((bits & 0x00800u) ? 0x0100u : 0) | ((bits & 0x00400u) ? 0x0001u : 0) | ((bits & 0x00200u) ? 0x0200u : 0) |
                                 ((bits & 0x00100u) ? 0x0002u : 0) | ((bits & 0x00080u) ? 0x0400u : 0) | ((bits & 0x00040u) ? 0x0004u : 0) |
                                 ((bits & 0x00004u) ? 0x0080u : 0);
                        hit = true;
                    }
                }
            }
        }
        const uint32_t hm = __ballot_sync(0xffffffffu, hit);
        if (hm) {
            uint32_t slot = 0;
            if (lane == 0)
                slot = atomicAdd(&a.counters->n_modeac_hits, (unsigned int) __popc(hm));
            slot = __shfl_sync(0xffffffffu, slot, 0) + __popc(hm & ((1u << lane) - 1u));
            if (hit) {
                if (slot < a.hit_cap)
                    a.hits[slot] = AcHit{(uint32_t) q, f1_clock, modeac, 0u};
                else
                    atomicOr(&a.counters->overflow, 32u);
            }
        }
    }
}

cudaError_t launch_modeac(const ModeacArgs &a, cudaStream_t stream) {
    if (a.nsamples == 0 || a.nblocks == 0)
        return cudaSuccess;
    modeac_noise_kernel<<<(a.nblocks + 127) / 128, 128, 0, stream>>>(a);
    unsigned long long want = (a.nsamples + 255) / 256;
    const int grid = (int) (want < 148ull * 8 ? want : 148ull * 8);
    modeac_kernel<<<grid, 256, 0, stream>>>(a);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// DC-filter front end (--dcfilter): convert_uc8_generic / convert_sc16_generic / convert_sc16q11_generic
// (convert.c:113-213, 374-423), bit-exactly
//
// Per rail the reference runs z = f*a + z*b over the WHOLE stream (the state lives in struct
// converter_state across calls) and takes the magnitude of f - z.  The recurrence is a chain of
// dependent float operations, one multiply and one add per sample; rounding makes it non-associative,
// so the same bits need the same chain.  Everything around the chain is parallel:
//   dc_prepare_kernel    every sample: a = f * dc_a for both rails            (planar float arrays)
//   dc_chain_kernel      ONE warp: lane 0 walks the I chain, lane 1 the Q chain (z replaces a in place);
//                        all 32 lanes stream the batches through shared memory with cp.async, so the
//                        chain lanes never wait on HBM: 8 cycles per sample, 0.6 s per minute of signal
//   dc_magnitude_kernel  one CTA per mag_buf: |f - z| -> u16 magnitudes, and the converter's sequential
//                        float sums of mag and magsq (same scheme as float_block_sums_kernel)
// K1a then reads the magnitude stream as "format 3".
// ------------------------------------------------------------------------------------------

// fI / fQ of sample i exactly as the generic converters compute them (convert.c:131-134, 182-185, 392-395)
__device__ __forceinline__ void dc_rails(const uint8_t *__restrict__ iq, uint32_t format, uint64_t i, float &fI, float &fQ) {
    if (format == 0) {
        const uint32_t w = reinterpret_cast<const uint16_t *>(iq)[i];
        fI = __fdiv_rn(__fsub_rn((float) (w & 0xffu), 127.5f), 127.5f);
        fQ = __fdiv_rn(__fsub_rn((float) (w >> 8), 127.5f), 127.5f);
    } else {
        const uint32_t w = reinterpret_cast<const uint32_t *>(iq)[i];
        const float inv_scale = (format == 1) ? (1.0f / 32768.0f) : (1.0f / 2048.0f); // division by 2^k == exact scaling
        fI = __fmul_rn((float) (int16_t) (w & 0xffff), inv_scale);
        fQ = __fmul_rn((float) (int16_t) (w >> 16), inv_scale);
    }
}

__global__ void __launch_bounds__(256) dc_prepare_kernel(const uint8_t *__restrict__ iq, uint32_t format, uint64_t n, float dc_a,
                                                          float *__restrict__ aI, float *__restrict__ aQ) {
    for (uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t) gridDim.x * blockDim.x) {
        float fI, fQ;
        dc_rails(iq, format, i, fI, fQ);
        aI[i] = __fmul_rn(fI, dc_a);
        aQ[i] = __fmul_rn(fQ, dc_a);
    }
}

constexpr int kDcBatch = 1024; // samples per rail and batch of the chain kernel

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
    const uint32_t sa = (uint32_t) __cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gmem) : "memory");
}

// aI / aQ hold room for whole batches (the caller pads them); only the first n entries mean anything
__global__ void __launch_bounds__(32) dc_chain_kernel(float *__restrict__ aI, float *__restrict__ aQ, uint64_t n, float dc_b,
                                                       float *__restrict__ state) {
    __shared__ __align__(16) float s_buf[2][2][kDcBatch]; // [buffer][rail][sample]
    const int lane = threadIdx.x;
    const uint64_t nbatches = (n + kDcBatch - 1) / kDcBatch;
    auto fetch = [&](uint64_t b) { // all lanes: batch b of both rails -> buffer b & 1
        const uint64_t base = b * kDcBatch;
#pragma unroll
        for (int q = 0; q < kDcBatch / 4 / 32; ++q) {
            const int u = q * 32 + lane; // 16-byte unit of the batch
            cp_async16(&s_buf[b & 1][0][4 * u], aI + base + 4 * u);
            cp_async16(&s_buf[b & 1][1][4 * u], aQ + base + 4 * u);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    float z = (lane < 2) ? state[lane] : 0.0f;
    if (nbatches)
        fetch(0);
    for (uint64_t b = 0; b < nbatches; ++b) {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp();
        if (b + 1 < nbatches)
            fetch(b + 1); // lands while the chain lanes walk batch b
        if (lane < 2) {
            float *mine = s_buf[b & 1][lane];
            const uint64_t left = n - b * kDcBatch;
            if (left >= (uint64_t) kDcBatch) {
                for (int g = 0; g < kDcBatch; g += 32) {
                    float4 r[8]; // loads first: they do not depend on the chain
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        r[i] = reinterpret_cast<const float4 *>(mine + g)[i];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        // z1 = f * dc_a + z1 * dc_b (convert.c:137-138): product, then sum, each rounded
                        z = __fadd_rn(r[i].x, __fmul_rn(z, dc_b));
                        r[i].x = z;
                        z = __fadd_rn(r[i].y, __fmul_rn(z, dc_b));
                        r[i].y = z;
                        z = __fadd_rn(r[i].z, __fmul_rn(z, dc_b));
                        r[i].z = z;
                        z = __fadd_rn(r[i].w, __fmul_rn(z, dc_b));
                        r[i].w = z;
                    }
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        reinterpret_cast<float4 *>(mine + g)[i] = r[i];
                }
            } else {
                for (int i = 0; i < (int) left; ++i) {
                    z = __fadd_rn(mine[i], __fmul_rn(z, dc_b));
                    mine[i] = z;
                }
            }
        }
        __syncwarp();
        // all lanes: z of batch b back to the arrays
        const uint64_t base = b * kDcBatch;
#pragma unroll
        for (int q = 0; q < kDcBatch / 4 / 32; ++q) {
            const int u = q * 32 + lane;
            reinterpret_cast<float4 *>(aI + base)[u] = reinterpret_cast<const float4 *>(s_buf[b & 1][0])[u];
            reinterpret_cast<float4 *>(aQ + base)[u] = reinterpret_cast<const float4 *>(s_buf[b & 1][1])[u];
        }
        __syncwarp(); // buffer b & 1 is refilled by the fetch of batch b + 2, issued after the next wait
    }
    if (lane < 2)
        state[lane] = z;
}

__global__ void __launch_bounds__(64) dc_magnitude_kernel(const uint8_t *__restrict__ iq, uint32_t format, uint64_t nsamples,
                                                           uint32_t block_samples, const float *__restrict__ zI,
                                                           const float *__restrict__ zQ, uint16_t *__restrict__ mag_out,
                                                           double *__restrict__ sums) {
    // one CTA of two warps per mag_buf, as float_block_sums_kernel: warp 0 converts the next 128 samples
    // while lanes 0 and 1 of warp 1 add the current 128 to sum_level / sum_power in stream order
    __shared__ __align__(16) float s_val[2][2][128];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t k = blockIdx.x;
    const uint64_t b0 = (uint64_t) k * block_samples;
    const uint64_t nk = nsamples > b0 ? (nsamples - b0 < block_samples ? nsamples - b0 : block_samples) : 0;
    auto put = [&](int buf, uint64_t base) { // samples b0 + base + 4 * lane .. + 3
        float mg[4], sq[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint64_t i = base + 4 * (uint64_t) lane + j;
            mg[j] = sq[j] = 0.0f;
            if (i < nk) {
                float fI, fQ;
                dc_rails(iq, format, b0 + i, fI, fQ);
                fI = __fsub_rn(fI, __ldg(zI + b0 + i)); // convert.c:139-140
                fQ = __fsub_rn(fQ, __ldg(zQ + b0 + i));
                mag_out[b0 + i] = (uint16_t) mag_from_float(fI, fQ, sq[j], mg[j]);
            }
        }
        *reinterpret_cast<float4 *>(&s_val[buf][0][4 * lane]) = make_float4(mg[0], mg[1], mg[2], mg[3]);
        *reinterpret_cast<float4 *>(&s_val[buf][1][4 * lane]) = make_float4(sq[0], sq[1], sq[2], sq[3]);
    };
    const uint64_t nbatches = (nk + 127) / 128;
    if (warp == 0 && nbatches)
        put(0, 0);
    float acc = 0.0f; // warp 1, lane 0: sum_level, lane 1: sum_power
    __syncthreads();
    for (uint64_t b = 0; b < nbatches; ++b) {
        if (warp == 0) {
            if (b + 1 < nbatches)
                put((int) ((b + 1) & 1), (b + 1) * 128);
        } else if (lane < 2) {
            const float *mine = s_val[b & 1][lane];
            const uint64_t left = nk - b * 128;
            if (left >= 128) {
                float4 r[32];
#pragma unroll
                for (int i = 0; i < 32; ++i)
                    r[i] = reinterpret_cast<const float4 *>(mine)[i];
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    acc = __fadd_rn(acc, r[i].x);
                    acc = __fadd_rn(acc, r[i].y);
                    acc = __fadd_rn(acc, r[i].z);
                    acc = __fadd_rn(acc, r[i].w);
                }
            } else {
                for (int i = 0; i < (int) left; ++i)
                    acc = __fadd_rn(acc, mine[i]);
            }
        }
        __syncthreads();
    }
    if (warp == 1 && lane < 2)
        sums[2 * k + lane] = (double) acc;
}

cudaError_t launch_dc_front_end(const uint8_t *iq, uint32_t format, uint64_t nsamples, uint32_t block_samples, float dc_a, float dc_b,
                                float *aI, float *aQ, float *state, uint16_t *mag_out, double *sums, cudaStream_t stream) {
    if (nsamples == 0)
        return cudaSuccess;
    unsigned long long want = (nsamples + 255) / 256;
    const int grid = (int) (want < 148ull * 16 ? want : 148ull * 16);
    dc_prepare_kernel<<<grid, 256, 0, stream>>>(iq, format, nsamples, dc_a, aI, aQ);
    dc_chain_kernel<<<1, 32, 0, stream>>>(aI, aQ, nsamples, dc_b, state);
    const uint32_t nblocks = (uint32_t) ((nsamples + block_samples - 1) / block_samples);
    dc_magnitude_kernel<<<nblocks, 64, 0, stream>>>(iq, format, nsamples, block_samples, aI, aQ, mag_out, sums);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// convert_kernel: the iq_convert_fn boundary (convert.h:33-38), magnitudes materialised
// ------------------------------------------------------------------------------------------

__global__ void __launch_bounds__(256) convert_kernel(const uint8_t *__restrict__ iq, uint32_t format, uint32_t n,
                                                       const uint16_t *__restrict__ lut, int table_bits, uint16_t *__restrict__ mag,
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
        } else if (format == 4) {
            m = __ldg(&lut[sc16q11_table_index(reinterpret_cast<const uint32_t *>(iq)[i], table_bits)]);
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
    if (format == 0 || format == 4) {
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

cudaError_t launch_convert(const uint8_t *iq, uint32_t format, uint32_t nsamples, const uint16_t *lut, int table_bits, uint16_t *mag,
                           unsigned long long *sums_u64, double *sums_f64, cudaStream_t stream) {
    if (nsamples == 0)
        return cudaSuccess;
    int grid = (int) ((nsamples + 255) / 256);
    if (grid > 148 * 8)
        grid = 148 * 8;
    convert_kernel<<<grid, 256, 0, stream>>>(iq, format, nsamples, lut, table_bits, mag, sums_u64, sums_f64);
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
