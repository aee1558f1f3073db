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

constexpr size_t kSmemLut = 65536 * sizeof(uint16_t);
constexpr int kBatch = 16;     // frames sliced per batch (their message words live in shared memory)
constexpr int kMsgWords = 5;   // 4 message words + the CRC syndrome
constexpr int kGroupsLong = 23, kGroupsShort = 12; // 5-bit groups of a 112 / 56 bit frame
constexpr size_t kSmemWarp = kWarpBuf * sizeof(uint32_t) + 2 * kItemCap * sizeof(uint16_t) + kBatch * kMsgWords * sizeof(uint32_t);
// CTA-wide tables: group syndromes [23 + 12][32], slicer taps [5 phases][5 bits] (int4 taps + sample offset)
constexpr size_t kSmemTail = (kGroupsLong + kGroupsShort) * 32 * sizeof(uint32_t) + 25 * sizeof(int4) + 25 * sizeof(int);

size_t scan_smem_bytes(uint32_t format) {
    return (format == 0 ? kSmemLut : 0) + kScanWarps * kSmemWarp + kSmemTail;
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

struct WarpCtx {
    const uint32_t *buf;   // the warp's magnitude rows
    int row0;              // first row of the chunk being scanned
    const uint32_t *gsyn;  // group syndromes (shared): [23 long groups + 12 short groups][32]
    const int4 *taps;      // [phase - 4][bit of the group]: correlator taps (shared)
    const int *soff;       // [phase - 4][bit of the group]: first sample of the bit, relative to the group
    uint16_t *items, *valid;
    uint32_t *msg;         // [kBatch][kMsgWords]
    long long chunk_pos0;  // scan position of window start 0 of the chunk
    // tile output cursors
    uint32_t *cand_out;
    PhaseRec *rec_out;
    uint32_t cand_cap, rec_cap, ncand, nrec;
};

// Five consecutive PPM bit decisions (demod_2400.c:73-177 in closed form).  Frame bit b of a
// candidate tried at phase p sits t = p + 12 b fifths of a sample after m[19]; five bits later the
// pattern repeats 12 samples on, so group k of a frame (bits 5k .. 5k+4) reads the 15 samples from
// m[19 + 12k] with offsets and correlators that depend on the phase only.  x = window start of the
// frame in the chunk, phi = phase - 4.  Returns the five decisions, bit c = frame bit 5k + c.
// A row of the warp buffer holds 16 magnitudes and a copy of the next row's first 4, so the four
// samples of one bit never straddle a row: they start in the group's first row or in the next one.
__device__ __forceinline__ uint32_t slice_group(const WarpCtx &cx, int x, int phi, int k) {
    const int y = x + 19 + 12 * k;
    const int row = (cx.row0 + (y >> 4)) & (kRows - 1);
    const int col0 = y & 15;
    const uint32_t *r0 = cx.buf + row * kRowWords;
    const uint32_t *r1 = cx.buf + ((row + 1) & (kRows - 1)) * kRowWords - 16;
    const int4 *tp = cx.taps + phi * 5;
    const int *so = cx.soff + phi * 5;
    uint32_t v5 = 0;
#pragma unroll
    for (int c = 0; c < 5; ++c) {
        const int4 t = tp[c];
        const int col = col0 + so[c];
        const uint32_t *p = (col <= 16 ? r0 : r1) + col;
        const int v = t.x * (int) p[0] + t.y * (int) p[1] + t.z * (int) p[2] + t.w * (int) p[3];
        v5 |= (v > 0) ? (1u << c) : 0u;
    }
    return v5;
}

// Slice the queued (position, phase) items of the chunk.
//   1. one lane per item slices group 0 = the DF field -> frame length (demod_2400.c:188-205);
//      long frames are listed from the front of `valid`, short ones from its back
//   2. in batches of kBatch frames: one lane per (frame, group) slices five bits, ORs them into the
//      frame's message words and XORs the group's CRC contribution (crc.c:59-64: the syndrome is
//      linear in the bits) into its syndrome, both in shared memory
//   3. one lane per frame classifies it and appends the class record
__device__ __forceinline__ void process_items(const ScanArgs &a, WarpCtx &cx, int &nitems) {
    const int lane = threadIdx.x & 31;
    const uint32_t below = (1u << lane) - 1u;
    __syncwarp();
    int nl = 0, ns = 0;
    for (int it = 0; it < nitems; it += 32) {
        const bool active = it + lane < nitems;
        const uint32_t item = cx.items[active ? it + lane : it];
        const uint32_t v5 = slice_group(cx, (int) (item & 511u), (int) ((item >> 9) & 7u), 0);
        const int nb = active ? frame_bytes_for_df(__brev(v5) >> 27) : 0;
        const uint32_t lm = __ballot_sync(0xffffffffu, nb == 14), sm = __ballot_sync(0xffffffffu, nb == 7);
        if (nb == 14)
            cx.valid[nl + __popc(lm & below)] = (uint16_t) (item | (1u << 12));
        else if (nb == 7)
            cx.valid[kItemCap - 1 - (ns + __popc(sm & below))] = (uint16_t) item;
        nl += __popc(lm);
        ns += __popc(sm);
    }
    __syncwarp();
    const int nframes = nl + ns;
    for (int q0 = 0; q0 < nframes; q0 += kBatch) { // uniform
        const int nb = min(kBatch, nframes - q0);
        const int nlb = max(0, min(nb, nl - q0)); // long frames of the batch come first
        const int ntasks = kGroupsLong * nlb + kGroupsShort * (nb - nlb);
        for (int i = lane; i < nb * kMsgWords; i += 32)
            cx.msg[i] = 0;
        __syncwarp();
        for (int t0 = 0; t0 < ntasks; t0 += 32) {
            const int t = t0 + lane;
            if (t < ntasks) {
                int bi, k;
                const int ts = t - kGroupsLong * nlb;
                if (ts < 0) {
                    bi = (t * 2850) >> 16; // t / 23 for t < 23 * 16
                    k = t - kGroupsLong * bi;
                } else {
                    const int sb = (ts * 5462) >> 16; // ts / 12
                    bi = nlb + sb;
                    k = ts - kGroupsShort * sb;
                }
                const int q = q0 + bi;
                const uint32_t item = (q < nl) ? cx.valid[q] : cx.valid[kItemCap - 1 - (q - nl)];
                const bool is_long = ts < 0;
                uint32_t v5 = slice_group(cx, (int) (item & 511u), (int) ((item >> 9) & 7u), k);
                const int left = (is_long ? 112 : 56) - 5 * k; // the last group of a frame is partial
                if (left < 5)
                    v5 &= (1u << left) - 1u;
                uint32_t *mw = cx.msg + bi * kMsgWords;
                const int b0 = 5 * k, wi = b0 >> 5, sh = b0 & 31;
                if (v5) {
                    atomicOr(&mw[wi], v5 << sh); // frame bit b -> bit b % 32 of word b / 32
                    if (sh > 27 && (v5 >> (32 - sh)))
                        atomicOr(&mw[wi + 1], v5 >> (32 - sh));
                    atomicXor(&mw[4], cx.gsyn[((is_long ? 0 : kGroupsLong) + k) * 32 + v5]);
                }
            }
        }
        __syncwarp();
        // one lane per frame: class record
        const bool active = lane < nb;
        uint32_t w0 = 0, w1 = 0, w2 = 0, w3 = 0, syn = 0, item = 0;
        if (active) {
            const uint32_t *mw = cx.msg + lane * kMsgWords;
            w0 = mw[0], w1 = mw[1], w2 = mw[2], w3 = mw[3], syn = mw[4];
            const int q = q0 + lane;
            item = (q < nl) ? cx.valid[q] : cx.valid[kItemCap - 1 - (q - nl)];
        }
        const uint32_t head32 = __brev(w0); // frame bits 0..31, MSB first
        const uint32_t df = head32 >> 27, aa = head32 & 0xffffffu;
        FrameClass fc;
        fc.kind = kKindBad;
        if (active)
            fc = classify_frame(df, aa, syn, (w0 | w1 | w2 | w3) == 0, a.tab_short, a.n_short, a.tab_long, a.n_long);
        const bool has = active && fc.kind != kKindBad;
        const uint32_t mask = __ballot_sync(0xffffffffu, has);
        if (has) {
            const uint32_t slot = cx.nrec + __popc(mask & below);
            const uint32_t ph = ((item >> 9) & 7u) + 4u;
            if (slot < cx.rec_cap) {
                PhaseRec pr;
                pr.pos = (uint32_t) (cx.chunk_pos0 + (item & 511u));
                pr.w0 = syn | (fc.kind << 24) | (fc.errors << 28);
                pr.w1 = fc.key | (ph << 24);
                pr.pad = 0;
                *reinterpret_cast<uint4 *>(&cx.rec_out[slot]) = *reinterpret_cast<const uint4 *>(&pr);
            }
            // mode_s.c:717-726: only a clean DF17, or a clean DF11 with IID 0, can ever be added to the
            // ICAO filter; remember every such address of the stream
            if (syn == 0 && (df == 17 || df == 11))
                atomicOr(&a.addr_bitmap[aa >> 5], 1u << (aa & 31u));
        }
        cx.nrec += __popc(mask);
        __syncwarp();
    }
    nitems = 0;
}

__device__ __noinline__ void flush_sums_u64(unsigned long long *dst, unsigned long long level, unsigned long long power) {
    const unsigned long long l = warp_sum_u64(level), p = warp_sum_u64(power);
    if ((threadIdx.x & 31) == 0 && (l | p)) {
        atomicAdd(&dst[0], l);
        atomicAdd(&dst[1], p);
    }
}

__device__ __noinline__ void flush_sums_f64(double *dst, double level, double power) {
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

template <int FORMAT, bool SLICE, bool EDGE>
__device__ __forceinline__ void process_tile(const ScanArgs &a, WarpCtx &cx, const uint32_t tile, const uint16_t *s_lut, uint32_t *s_buf) {
    constexpr int US = Fmt<FORMAT>::kUnitSamples, BPS = Fmt<FORMAT>::kBytes;
    constexpr int UNITS = kLanePos / US; // 16-byte units per lane per step
    const int lane = threadIdx.x & 31;
    const float inv_scale = (FORMAT == 1) ? (1.0f / 32768.0f) : (1.0f / 2048.0f);
    const long long n = (long long) a.nsamples;
    const long long B = (long long) a.block_samples;
    const int thr = a.threshold;

    const long long c0 = (long long) tile * kTile - kHead; // first window-start sample of the tile
    uint32_t cand_off, rec_off;
    if (a.tile_off) {
        cand_off = a.tile_off[2 * tile];
        rec_off = a.tile_off[2 * tile + 1];
        cx.cand_cap = a.tile_off[2 * tile + 2] - cand_off;
        cx.rec_cap = a.tile_off[2 * tile + 3] - rec_off;
    } else {
        cand_off = tile * a.cand_slab;
        rec_off = tile * a.rec_slab;
        cx.cand_cap = a.cand_slab;
        cx.rec_cap = a.rec_slab;
    }
    cx.cand_out = a.cand + cand_off;
    cx.rec_out = a.recs + rec_off;
    cx.ncand = cx.nrec = 0;
    int nitems = 0;
    uint32_t ncand_lane = 0; // scan-only mode: candidates seen by this lane

    // block sums of the samples this tile owns: [c0, c0 + kTile) within [0, n)
    unsigned long long sum_level = 0, sum_power = 0;
    double fsum_level = 0, fsum_power = 0;
    long long blk = (c0 > 0 ? c0 : 0) / B, next_bound = (blk + 1) * B;
    auto flush_sums = [&]() { // rare (once per tile and per mag_buf boundary): kept out of the hot code
        if (FORMAT == 0)
            flush_sums_u64(a.block_sums_u64 + 2 * blk, sum_level, sum_power);
        else
            flush_sums_f64(a.block_sums_f64 + 2 * blk, fsum_level, fsum_power);
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
                const uint32_t words[4] = {swizzle_pair(pre[u].x), swizzle_pair(pre[u].y), swizzle_pair(pre[u].z), swizzle_pair(pre[u].w)};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    m[u * US + 2 * j] = s_lut[words[j] & 0xffffu];
                    m[u * US + 2 * j + 1] = s_lut[words[j] >> 16];
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
                        if (FORMAT != 0)
                            fmag[j] = fmagsq[j] = 0;
                    }
            }
        }
        // request the next chunk now; it is consumed one iteration later
        if (k < kScanSteps)
            prefetch(k + 1);

        // store: chunk k occupies rows (k&1)*32 + lane; the first four magnitudes are also copied into
        // the pad of the previous row
        {
            const int row = (k & 1) * 32 + lane;
            uint4 *dst = reinterpret_cast<uint4 *>(s_buf + row * kRowWords);
#pragma unroll
            for (int q = 0; q < kLanePos / 4; ++q)
                dst[q] = make_uint4(m[4 * q], m[4 * q + 1], m[4 * q + 2], m[4 * q + 3]);
            uint4 *pad = reinterpret_cast<uint4 *>(s_buf + ((row - 1) & (kRows - 1)) * kRowWords + kLanePos);
            *pad = make_uint4(m[0], m[1], m[2], m[3]);
        }

        // block sums: chunks 0..kScanSteps-1 are owned by this tile
        if (k < kScanSteps) {
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
                        if (FORMAT == 0) {
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
                                if (FORMAT == 0) {
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
                            if (FORMAT == 0)
                                unit_sums_u64(a.block_sums_u64, kb, make_uint4(m[u * US], m[u * US + 1], m[u * US + 2], m[u * US + 3]),
                                              make_uint4(m[u * US + 4 % US], m[u * US + 5 % US], m[u * US + 6 % US], m[u * US + 7 % US]));
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
            if (!SLICE) {
                ncand_lane += __popc(any);
            } else {
                uint32_t lanes = __ballot_sync(0xffffffffu, any != 0);
                cx.buf = s_buf;
                cx.row0 = row0;
                cx.chunk_pos0 = pos0;
                if (lanes) {
                    // candidates and (position, phase) items of the step, in position order: every lane
                    // places its own through a warp prefix sum (candidates low half, items high half).
                    // A lane's items (<= 80) always fit the queue; a step with more items than the queue
                    // holds is taken in rounds of as many whole lanes as fit.
                    const uint32_t mine = (uint32_t) __popc(any) | ((uint32_t) (2 * __popc(b45) + 2 * __popc(b67) + __popc(b8)) << 16);
                    uint32_t inc = mine;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const uint32_t up = __shfl_up_sync(0xffffffffu, inc, o);
                        if (lane >= o)
                            inc += up;
                    }
                    const uint32_t exc = inc - mine;
                    const uint32_t cand_base = cx.ncand;
                    cx.ncand += __shfl_sync(0xffffffffu, inc, 31) & 0xffffu;
                    int first_lane = 0;      // lanes below were queued in earlier rounds
                    uint32_t base_items = 0; // their items
                    while (first_lane < 32) { // uniform
                        const bool fits = lane >= first_lane && (inc >> 16) - base_items <= (uint32_t) kItemCap;
                        const uint32_t fm = __ballot_sync(0xffffffffu, fits) >> first_lane;
                        const int nl_round = __ffs(~fm) - 1; // consecutive lanes from first_lane that fit (>= 1)
                        const int end_lane = first_lane + (nl_round < 0 ? 32 - first_lane : nl_round);
                        if (lane >= first_lane && lane < end_lane) {
                            uint32_t ci = cand_base + (exc & 0xffffu);
                            int ii = (int) ((exc >> 16) - base_items);
                            uint32_t u = any;
                            while (u) {
                                const int i = __ffs(u) - 1;
                                u &= u - 1;
                                const uint32_t tm = (((b45 >> i) & 1u) * 3u) | (((b67 >> i) & 1u) * 12u) | (((b8 >> i) & 1u) * 16u);
                                const uint32_t pic = (uint32_t) (lane * kLanePos + i);
                                if (ci < cx.cand_cap)
                                    cx.cand_out[ci] = (uint32_t) (j * kStep + (int) pic) | (tm << 13);
                                ++ci;
#pragma unroll
                                for (int ph = 0; ph < 5; ++ph)
                                    if ((tm >> ph) & 1u)
                                        cx.items[ii++] = (uint16_t) (pic | ((uint32_t) ph << 9));
                            }
                        }
                        const uint32_t upto = __shfl_sync(0xffffffffu, inc, end_lane - 1) >> 16;
                        nitems = (int) (upto - base_items);
                        base_items = upto;
                        first_lane = end_lane;
                        if (nitems)
                            process_items(a, cx, nitems);
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
            td.nrec = cx.nrec;
            a.tiles[tile] = td;
            unsigned int ovf = (cx.ncand > cx.cand_cap ? 1u : 0u) | (cx.nrec > cx.rec_cap ? 2u : 0u);
            if (ovf)
                atomicOr(&a.counters->overflow, ovf);
            if (cx.nrec)
                atomicAdd(&a.counters->n_rec, (unsigned long long) cx.nrec);
        }
        if (cx.ncand)
            atomicAdd(&a.counters->n_cand, (unsigned long long) cx.ncand);
    }
}

template <int FORMAT, bool SLICE>
__global__ void __launch_bounds__(kScanThreads, 1) scan_kernel(const ScanArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    // shared memory carve-up
    unsigned char *sp = smem_raw;
    const uint16_t *s_lut = reinterpret_cast<const uint16_t *>(sp);
    if (FORMAT == 0)
        sp += kSmemLut;
    uint32_t *s_buf = reinterpret_cast<uint32_t *>(sp + (size_t) warp * kWarpBuf * sizeof(uint32_t));
    sp += (size_t) kScanWarps * kWarpBuf * sizeof(uint32_t);
    uint16_t *s_lists = reinterpret_cast<uint16_t *>(sp) + (size_t) warp * 2 * kItemCap;
    sp += (size_t) kScanWarps * 2 * kItemCap * sizeof(uint16_t);
    uint32_t *s_msg = reinterpret_cast<uint32_t *>(sp) + (size_t) warp * kBatch * kMsgWords;
    sp += (size_t) kScanWarps * kBatch * kMsgWords * sizeof(uint32_t);
    uint32_t *s_gsyn = reinterpret_cast<uint32_t *>(sp);
    sp += (kGroupsLong + kGroupsShort) * 32 * sizeof(uint32_t);
    int4 *s_taps = reinterpret_cast<int4 *>(sp);
    sp += 25 * sizeof(int4);
    int *s_soff = reinterpret_cast<int *>(sp);

    // one-time staging of the tables (the only block-wide barrier of the kernel).  The magnitude
    // table arrives pre-swizzled (32-bit word j of row Q at word j ^ (Q & 31)): all of a thread's
    // 16-byte loads are in flight at once, one round trip to L2 per CTA.
    if (FORMAT == 0) {
        const uint4 *src = reinterpret_cast<const uint4 *>(a.lut_swz);
        uint4 *dst = reinterpret_cast<uint4 *>(smem_raw);
        constexpr int kPer = (int) (kSmemLut / 16) / kScanThreads; // 16
        uint4 v[kPer];
#pragma unroll
        for (int q = 0; q < kPer; ++q)
            v[q] = __ldg(src + q * kScanThreads + tid);
#pragma unroll
        for (int q = 0; q < kPer; ++q)
            dst[q * kScanThreads + tid] = v[q];
    }
    // group syndromes: row g < 23 = bits 5g .. 5g+4 of a long frame, row 23 + g = of a short frame
    // (crc.c:143: a short frame uses the tail of the 112-entry single-bit syndrome list)
    for (int i = tid; i < (kGroupsLong + kGroupsShort) * 32; i += kScanThreads) {
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
        s_taps[tid] = make_int4(c_slice_coef[r][0], c_slice_coef[r][1], c_slice_coef[r][2], c_slice_coef[r][3]);
        s_soff[tid] = t / 5;
    }
    __syncthreads();

    WarpCtx cx;
    cx.gsyn = s_gsyn;
    cx.taps = s_taps;
    cx.soff = s_soff;
    cx.items = s_lists;
    cx.valid = s_lists + kItemCap;
    cx.msg = s_msg;

    const long long n = (long long) a.nsamples;
    for (;;) {
        // ---- next tile from the work queue ----
        uint32_t tile = 0;
        if (lane == 0)
            tile = atomicAdd(&a.counters->next_tile, 1u);
        tile = __shfl_sync(0xffffffffu, tile, 0);
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
    const int max_useful = (int) ((a.ntiles + kScanWarps - 1) / kScanWarps);
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
        default: return cudaErrorInvalidValue;
    }
#undef LAUNCH
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// K2: classify kernel -- one CTA per tile, position-indexed so that K1's output order is free
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
    const int bps = (a.format == 0) ? 2 : 4;
    const uint8_t *base = (s < 0) ? a.head + (s + kHead) * bps : a.iq + s * bps;
    if (a.format == 0) {
        const uint32_t idx = (uint32_t) base[0] | ((uint32_t) base[1] << 8);
        return __ldg(&a.lut[idx]);
    }
    const uint32_t w = *reinterpret_cast<const uint32_t *>(base);
    float magsq, mag;
    return mag_sc16_word(w, (a.format == 1) ? (1.0f / 32768.0f) : (1.0f / 2048.0f), magsq, mag);
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
    __shared__ uint32_t s_syn[112];
    __shared__ int s_coef[5][4];
    __shared__ __align__(16) uint16_t s_frame[kClassifyThreads / 32][kFrameSamples];
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

    if (tid < 112)
        s_syn[tid] = c_bit_syndrome[tid];
    if (tid < 20)
        (&s_coef[0][0])[tid] = (&c_slice_coef[0][0])[tid];
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

    // ---- pass 4: class records of live positions: re-slice the frame, signal power ----
    // a live position owns consecutive output slots, one per recorded phase in phase order
    for (uint32_t r = warp; r < td.nrec; r += kClassifyThreads / 32) {
        const PhaseRec pr = a.recs[td.rec_off + r];
        const uint32_t c = find_cand(cand, td.ncand, (uint32_t) ((long long) pr.pos - p0));
        if (!(s_flags[c] & 1u))
            continue;
        const int ph = (int) ((pr.w1 >> 24) & 15u);
        const uint32_t rank = (uint32_t) __popc((uint32_t) s_nb[c] & ((1u << (ph - 4)) - 1u));
        const uint32_t slot = liverec_off + s_slot[c] + rank;

        // magnitudes the frame touches: window position pos -> samples pos - kOverlap ...
        uint16_t *fm = s_frame[warp];
        const long long s_first = (long long) pr.pos - kOverlap;
        for (int x = lane; x < kFrameSamples; x += 32)
            fm[x] = (uint16_t) sample_mag(a, s_first + x);
        __syncwarp();

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
