// scan2.inl -- K1a for uc8 spans: the register-window scan kernel (included by kernels.cu, after scan_kernel).
//
// Same job and same outputs as scan_kernel<0, *> (kernels.cu): IQ -> magnitude (convert.c:63-111), per-mag_buf
// sums of mag and mag^2, the pre-check and the three preamble correlators of demod_2400.c:276-330 for every scan
// position, a position-ordered candidate list per tile, and the u16 magnitudes for K1b / K2.  What differs is the
// data flow, chosen so that nothing but the magnitude table goes through shared memory:
//
//   * a warp owns a tile of 8192 scan positions, and inside it every LANE owns a contiguous run of 256 positions
//     (the old kernel gave a lane 16 positions of every 512-sample step and fetched its neighbours' magnitudes
//     from a shared-memory ring: 9 + 4 128-bit shared-memory accesses per lane and step, 60 % of the
//     shared-memory pipe of the SM);
//   * the lane streams through its run with one 256-bit global load per 16 samples (LDG.256, a full 32-byte
//     sector per lane, two loads kept in flight) and keeps the last 32 magnitudes in a register window that is
//     addressed statically: the loop body covers 32 samples, so window slot (sample mod 32) is a compile-time
//     register;
//   * a position is tested when the last sample of its 19-sample window arrives, with the sign-test form of the
//     correlators (see scan_kernel); the four sign bits of 32 positions end up in four registers per body;
//   * the 18 samples a run needs past its end are the first 18 of the next lane's run: every lane leaves its
//     first nine magnitude pairs in a small per-warp apron in shared memory (the last lane's come from the first
//     samples of the next tile, converted once at tile start), so every sample is converted exactly once;
//   * block sums: sum of mag as an integer add, sum of mag^2 as two 16x8-bit dot products (IDP.2A) on the packed
//     pair -- m^2 = m * lo8(m) + 256 * m * hi8(m) -- into 32-bit partials that are widened once per 32 samples;
//     a mag_buf boundary inside a run is honoured at the 8-sample group where it falls (block sizes are multiples
//     of 8);
//   * the u16 magnitudes leave with one 256-bit store per 16 samples;
//   * candidates are emitted at the end of the tile, lane by lane, which is position order.
//
// The magnitude table is the only big consumer of shared memory (128 KiB); its layout XORs the three low bits of Q
// into bits 3..5 of the index, which spreads the 16 x 16 codes receiver noise lives in evenly over the 32 banks
// (measured 2.1 cycles per LDS.U16 against 4.0 for the r01 swizzle, tools/ubench.cu).
//
// Tiles that touch the start or the end of the span (carried head, ragged tail: two or three per chunk) go through
// scan_kernel's process_tile<.., EDGE> inside the same launch: warp 0 of a CTA owns the ring buffer that path needs
// and takes them before it joins the interior queue.

#include <stdlib.h>
#include <string.h>

#include <type_traits>

namespace {

constexpr int kRun = kTile / 32;                 // scan positions (and owned samples) per lane and tile
constexpr int kBodies = kRun / 32;               // loop bodies of 32 samples
constexpr int kScan2Warps = 16;                  // 128 registers per thread (20 warps at 96 registers spill and measure 7 % slower)
constexpr int kScan2Threads = kScan2Warps * 32;
constexpr int kApronWords = 12;                  // per lane: 9 magnitude pairs (18 samples), padded to 48 bytes
constexpr size_t kScan2Lut = 65536 * sizeof(uint16_t);
constexpr size_t kScan2Apron = (size_t) kScan2Warps * 32 * kApronWords * sizeof(uint32_t);
constexpr size_t kScan2Smem = kScan2Lut + kScan2Apron + kSmemWarp; // + one ring buffer for the edge tiles

__device__ __forceinline__ void ldg256(uint32_t (&w)[8], const uint8_t *p) {
    asm volatile("ld.global.nc.L1::no_allocate.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7])
                 : "l"(p));
}

__device__ __forceinline__ void stg256(uint16_t *p, const uint32_t (&w)[8]) {
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]),
                 "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7])
                 : "memory");
}

__device__ __forceinline__ uint32_t dp2a_lo_u(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    asm("dp2a.lo.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ uint32_t dp2a_hi_u(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    asm("dp2a.hi.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

// shared-memory layout of the magnitude table: entry i (= I | Q << 8) lives at i ^ ((i >> 5) & 0x38)
__device__ __forceinline__ uint32_t swizzle2_pair(uint32_t w) {
    return w ^ ((w >> 5) & 0x00380038u);
}

// the two magnitudes of a raw word (two uc8 samples): table byte offsets 2 * index through 16x8-bit dot products
__device__ __forceinline__ void convert_pair(const unsigned char *s_lut, uint32_t w, uint32_t &m0, uint32_t &m1) {
    const uint32_t ws = swizzle2_pair(w);
    const uint32_t a0 = dp2a_lo_u(ws, 0x00000002u, 0u); // 2 * low half
    const uint32_t a1 = dp2a_lo_u(ws, 0x00000200u, 0u); // 2 * high half
    m0 = *reinterpret_cast<const uint16_t *>(s_lut + a0);
    m1 = *reinterpret_cast<const uint16_t *>(s_lut + a1);
}

struct LaneSums {
    unsigned long long level, power; // of the lane's current mag_buf
    uint32_t level32, lo32, hi32;    // partials of the current body
    uint32_t blk;                    // mag_buf the lane is summing into
    uint32_t g8_next;                // 8-sample group of the run at which the next mag_buf starts (>= 32: not in this run)
};

// by value, so that the caller's sums stay in registers
__device__ __noinline__ void add_block_sums(unsigned long long *block_sums, uint32_t blk, unsigned long long level, unsigned long long power) {
    if (level | power) {
        atomicAdd(&block_sums[2 * (size_t) blk], level);
        atomicAdd(&block_sums[2 * (size_t) blk + 1], power);
    }
}

__device__ __forceinline__ void lane_flush(unsigned long long *block_sums, LaneSums &s) {
    s.level += s.level32;
    s.power += (unsigned long long) s.lo32 + ((unsigned long long) s.hi32 << 8);
    s.level32 = s.lo32 = s.hi32 = 0;
    add_block_sums(block_sums, s.blk, s.level, s.power);
    s.level = s.power = 0;
}

// One pair of samples enters the window (slots s0 = 2 * JJ and s0 + 1 of the 32-slot rings), and the two scan
// positions whose 19-sample windows end with it (window starts s0 - 18 and s0 - 17) are tested.
// demod_2400.c:276-330 as sign tests (see scan_kernel), arranged so that what neighbouring positions share is
// computed once, when the sample that completes it arrives:
//   D[x] = m[x] - m[x+1]                      used as m2 - m3 of position x - 2 and as m10 - m11 of position x - 10
//   F[x] = 32 (m[x+1] - m[x+2] + m[x+3] + m[x+4]) + 15
//                                             the pulse pattern of one preamble half: position x uses F[x] + F[x+8]
// and per position
//   bn  = m5 + m8 + m16 + m17 + m18
//   E0  = F[i] + F[i+8] + 1 - thr bn          (= 32 (m1 - m2 + m3 + m4 + m9 - m10 + m11 + m12) + 31 - thr bn)
//   E45 = E0                                   phases 4, 5: common3456 - diff_10_11 >= ref_level
//   E67 = E0 + 64 D[i+10]                      phases 6, 7: common3456 + diff_10_11
//   E8  = E67 + 96 D[i+2] - 32 m9              phase 8:     sum_1_4 + 2 diff_2_3 + diff_10_11 + m12
//   g   = (m7 - m1) & (m14 - m12) & (m15 - m12)
// a test passes when its E is non-negative, the pre-check when g is negative.  Everything is exact int32:
// |32 X| < 2^24 and thr * bn < 2^31 for thresholds up to 6000.
struct Window {
    int m[32], D[32], F[32];
};

template <int JJ>
__device__ __forceinline__ void test_pair(Window &w, int nthr, uint32_t &s45, uint32_t &s67, uint32_t &s8, uint32_t &pm) {
    constexpr int s0 = 2 * JJ + 64; // slot of the pair's first sample (before the & 31)
#define M_(x) w.m[(x) & 31]
#define D_(x) w.D[(x) & 31]
#define F_(x) w.F[(x) & 31]
    D_(s0 - 1) = M_(s0 - 1) - M_(s0);
    D_(s0) = M_(s0) - M_(s0 + 1);
    F_(s0 - 4) = (D_(s0 - 3) + M_(s0 - 1) + M_(s0)) * 32 + 15;
    F_(s0 - 3) = (D_(s0 - 2) + M_(s0) + M_(s0 + 1)) * 32 + 15;
#pragma unroll
    for (int odd = 0; odd < 2; ++odd) {
        const int i = s0 - 18 + odd; // window start slot
        const int bn = (M_(i + 5) + M_(i + 8) + M_(i + 16)) + (M_(i + 17) + M_(i + 18));
        const int E45 = (nthr * bn + F_(i)) + F_(i + 8) + 1;
        const int E67 = D_(i + 10) * 64 + E45;
        const int E8 = M_(i + 9) * -32 + (D_(i + 2) * 96 + E67);
        const int g = (M_(i + 7) - M_(i + 1)) & (M_(i + 14) - M_(i + 12)) & (M_(i + 15) - M_(i + 12));
        s45 = __funnelshift_l((uint32_t) E45, s45, 1);
        s67 = __funnelshift_l((uint32_t) E67, s67, 1);
        s8 = __funnelshift_l((uint32_t) E8, s8, 1);
        pm = __funnelshift_l((uint32_t) g, pm, 1);
    }
#undef M_
#undef D_
#undef F_
}

struct TileMasks { // bit k of word w: lane-local position 32 w + k - 18 (words 0..8; bits 0..17 of word 0 are not this lane's)
    uint32_t b45[kBodies + 1], b67[kBodies + 1], b8[kBodies + 1];
};

// ODD16: the runs start 16 bytes past a 32-byte boundary (the usual case: a tile starts kHead = 328 samples = 656
// bytes before a multiple of 16 KiB).  The 256-bit loads then fetch the aligned blocks around the run and a body
// takes the upper half of one block, a whole block and the lower half of a third -- register renaming, no shuffling.
template <bool SLICE, bool ODD16>
__device__ __forceinline__ void scan2_tile(const ScanArgs &a, WarpCtx &cx, const uint32_t tile, const unsigned char *s_lut, uint32_t *s_apron) {
    const int lane = threadIdx.x & 31;
    const int nthr = -a.threshold;
    const long long c0 = (long long) tile * kTile - kHead; // first window-start sample of the tile (>= 0: interior)
    const long long s_run = c0 + (long long) lane * kRun;  // first sample of the lane's run
    const uint8_t *gp = a.iq + s_run * 2 - (ODD16 ? 16 : 0); // 32-byte aligned
    uint16_t *gm = a.mag + (size_t) tile * kTile + (size_t) lane * kRun;

    // ---- block sums bookkeeping ----
    LaneSums sums;
    sums.level = sums.power = 0;
    sums.level32 = sums.lo32 = sums.hi32 = 0;
    {
        const unsigned long long B = a.block_samples;
        sums.blk = (uint32_t) ((unsigned long long) s_run / B);
        const unsigned long long to_next = ((unsigned long long) sums.blk + 1) * B - (unsigned long long) s_run;
        sums.g8_next = to_next < (unsigned long long) kRun ? (uint32_t) (to_next >> 3) : 0xffffffffu;
    }

    // ---- the tile's tail: the 18 samples after its last run belong to the next tile; lanes 0..8 convert one pair each ----
    __syncwarp(); // the previous tile's apron reads are done
    if (lane < 9) {
        const uint32_t w = __ldg(reinterpret_cast<const uint32_t *>(a.iq + (c0 + kTile) * 2) + lane);
        uint32_t m0, m1;
        convert_pair(s_lut, w, m0, m1);
        s_apron[31 * kApronWords + lane] = m0 | (m1 << 16);
    }

    Window win;
#pragma unroll
    for (int i = 0; i < 32; ++i)
        win.m[i] = win.D[i] = win.F[i] = 0;
    TileMasks tm;

    uint32_t nxt[16]; // the next body's two aligned 32-byte blocks, in flight
    uint32_t carry[4] = {0, 0, 0, 0}; // ODD16: the upper half of the block the previous body ended in
    if (ODD16) {
        uint32_t first[8];
        ldg256(first, gp);
        gp += 32;
#pragma unroll
        for (int i = 0; i < 4; ++i)
            carry[i] = first[4 + i];
    }
    ldg256(*reinterpret_cast<uint32_t(*)[8]>(&nxt[0]), gp);
    ldg256(*reinterpret_cast<uint32_t(*)[8]>(&nxt[8]), gp + 32);

#pragma unroll 1
    for (int body = 0; body < kBodies; ++body) {
        uint32_t cur[16];
        if (ODD16) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
                cur[i] = carry[i];
#pragma unroll
            for (int i = 0; i < 12; ++i)
                cur[4 + i] = nxt[i];
#pragma unroll
            for (int i = 0; i < 4; ++i)
                carry[i] = nxt[12 + i];
        } else {
#pragma unroll
            for (int i = 0; i < 16; ++i)
                cur[i] = nxt[i];
        }
        if (body + 1 < kBodies) {
            ldg256(*reinterpret_cast<uint32_t(*)[8]>(&nxt[0]), gp + (body + 1) * 64);
            ldg256(*reinterpret_cast<uint32_t(*)[8]>(&nxt[8]), gp + (body + 1) * 64 + 32);
        }
        uint32_t s45 = 0, s67 = 0, s8 = 0, pm = 0;
        uint32_t V[16];
        auto step = [&](auto jj_c) {
            constexpr int JJ = decltype(jj_c)::value;
            if (JJ % 4 == 0) { // an 8-sample group starts: does a mag_buf end here?
                if ((uint32_t) (body * 4 + JJ / 4) == sums.g8_next) {
                    lane_flush(a.block_sums_u64, sums);
                    ++sums.blk;
                    const uint32_t step8 = a.block_samples >> 3;
                    sums.g8_next = (step8 < 32u - sums.g8_next) ? sums.g8_next + step8 : 0xffffffffu;
                }
            }
            uint32_t m0, m1;
            convert_pair(s_lut, cur[JJ], m0, m1);
            const uint32_t v = m0 | (m1 << 16);
            V[JJ] = v;
            // sums (convert.c:95-110): sum of mag, and of mag^2 = mag * lo8(mag) + 256 * mag * hi8(mag)
            sums.level32 += m0 + m1;
            const uint32_t bytes = __byte_perm(v, 0, 0x3120); // lo8(m0), lo8(m1), hi8(m0), hi8(m1)
            sums.lo32 = dp2a_lo_u(v, bytes, sums.lo32);
            sums.hi32 = dp2a_hi_u(v, bytes, sums.hi32);
            win.m[(2 * JJ) & 31] = (int) m0;
            win.m[(2 * JJ + 1) & 31] = (int) m1;
            test_pair<JJ>(win, nthr, s45, s67, s8, pm);
        };
#define STEP_(J) step(std::integral_constant<int, J>{});
        STEP_(0) STEP_(1) STEP_(2) STEP_(3) STEP_(4) STEP_(5) STEP_(6) STEP_(7)
        if (SLICE)
            stg256(gm + body * 32, *reinterpret_cast<uint32_t(*)[8]>(&V[0]));
        STEP_(8) STEP_(9) STEP_(10) STEP_(11) STEP_(12) STEP_(13) STEP_(14) STEP_(15)
#undef STEP_
        if (SLICE)
            stg256(gm + body * 32 + 16, *reinterpret_cast<uint32_t(*)[8]>(&V[8]));
        if (body == 0 && lane > 0) {
            // the first 18 magnitudes of this run are the previous lane's look-ahead
            uint32_t *ap = s_apron + (lane - 1) * kApronWords;
            *reinterpret_cast<uint4 *>(ap) = make_uint4(V[0], V[1], V[2], V[3]);
            *reinterpret_cast<uint4 *>(ap + 4) = make_uint4(V[4], V[5], V[6], V[7]);
            ap[8] = V[8];
        }
        // widen the body's partial sums
        sums.level += sums.level32;
        sums.power += (unsigned long long) sums.lo32 + ((unsigned long long) sums.hi32 << 8);
        sums.level32 = sums.lo32 = sums.hi32 = 0;
        // masks: the first position tested sits in the top bit
        pm = __brev(pm);
        tm.b45[body] = pm & ~__brev(s45);
        tm.b67[body] = pm & ~__brev(s67);
        tm.b8[body] = pm & ~__brev(s8);
    }

    // ---- the run's last 18 positions: their look-ahead comes from the apron ----
    __syncwarp();
    {
        const uint32_t *ap = s_apron + lane * kApronWords;
        const uint4 q0 = *reinterpret_cast<const uint4 *>(ap), q1 = *reinterpret_cast<const uint4 *>(ap + 4);
        const uint32_t V[9] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, ap[8]};
        uint32_t s45 = 0, s67 = 0, s8 = 0, pm = 0;
        auto step = [&](auto jj_c) {
            constexpr int JJ = decltype(jj_c)::value;
            win.m[(2 * JJ) & 31] = (int) (V[JJ] & 0xffffu);
            win.m[(2 * JJ + 1) & 31] = (int) (V[JJ] >> 16);
            test_pair<JJ>(win, nthr, s45, s67, s8, pm);
        };
#define STEP_(J) step(std::integral_constant<int, J>{});
        STEP_(0) STEP_(1) STEP_(2) STEP_(3) STEP_(4) STEP_(5) STEP_(6) STEP_(7) STEP_(8)
#undef STEP_
        // 18 positions tested: they sit in the low 18 bits, last one in bit 0
        pm = __brev(pm << 14);
        tm.b45[kBodies] = pm & ~__brev(s45 << 14);
        tm.b67[kBodies] = pm & ~__brev(s67 << 14);
        tm.b8[kBodies] = pm & ~__brev(s8 << 14);
    }
    // bits 0..17 of word 0 are positions of the previous lane (tested there, with its own window)
    tm.b45[0] &= ~0x3ffffu;
    tm.b67[0] &= ~0x3ffffu;
    tm.b8[0] &= ~0x3ffffu;

    // ---- block sums of the run ----
    {
        const uint32_t blk0 = __shfl_sync(0xffffffffu, sums.blk, 0);
        if (__all_sync(0xffffffffu, sums.blk == blk0)) {
            const unsigned long long l = warp_sum_u64(sums.level), p = warp_sum_u64(sums.power);
            if (lane == 0 && (l | p)) {
                atomicAdd(&a.block_sums_u64[2 * (size_t) blk0], l);
                atomicAdd(&a.block_sums_u64[2 * (size_t) blk0 + 1], p);
            }
        } else {
            lane_flush(a.block_sums_u64, sums); // a mag_buf boundary inside the tile: every lane for itself
        }
    }

    // ---- candidates, in position order: lane after lane ----
    uint32_t mine = 0;
#pragma unroll
    for (int w = 0; w <= kBodies; ++w)
        mine += (uint32_t) __popc(tm.b45[w] | tm.b67[w] | tm.b8[w]);
    uint32_t inc = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t up = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o)
            inc += up;
    }
    const uint32_t total = __shfl_sync(0xffffffffu, inc, 31);
    if (a.dbg_masks) {
        uint8_t *dst = a.dbg_masks + ((long long) tile * kTile - kPosShift + (long long) lane * kRun);
#pragma unroll
        for (int w = 0; w <= kBodies; ++w)
            for (int k = 0; k < 32; ++k) {
                const int r = 32 * w + k - 18;
                if (r >= 0 && r < kRun)
                    dst[r] = (uint8_t) ((((tm.b45[w] >> k) & 1u) * 3u) | (((tm.b67[w] >> k) & 1u) * 12u) | (((tm.b8[w] >> k) & 1u) * 16u));
            }
    }
    if (SLICE) {
        uint32_t cand_off, rec_off, cand_cap;
        if (a.tile_off) {
            cand_off = a.tile_off[2 * tile];
            rec_off = a.tile_off[2 * tile + 1];
            cand_cap = a.tile_off[2 * tile + 2] - cand_off;
        } else {
            cand_off = tile * a.cand_slab;
            rec_off = tile * a.rec_slab;
            cand_cap = a.cand_slab;
        }
        uint32_t *out = a.cand + cand_off;
        uint32_t ci = inc - mine;
        // K1b cuts a tile's list into units of 1024 positions: candidates in front of every 512-position step
        if ((lane & 1) == 0)
            a.step_off[tile * kScanSteps + (lane >> 1)] = (uint16_t) ci;
        if (mine) {
#pragma unroll
            for (int w = 0; w <= kBodies; ++w) {
                uint32_t u = tm.b45[w] | tm.b67[w] | tm.b8[w];
                while (u) {
                    const int k = __ffs(u) - 1;
                    u &= u - 1;
                    const uint32_t t5 = (((tm.b45[w] >> k) & 1u) * 3u) | (((tm.b67[w] >> k) & 1u) * 12u) | (((tm.b8[w] >> k) & 1u) * 16u);
                    if (ci < cand_cap)
                        out[ci] = (uint32_t) (lane * kRun + 32 * w + k - 18) | (t5 << 13);
                    ++ci;
                }
            }
        }
        if (lane == 0) {
            TileDesc td;
            td.cand_off = cand_off;
            td.ncand = total;
            td.rec_off = rec_off;
            td.nrec = 0; // K1b
            a.tiles[tile] = td;
            if (total > cand_cap)
                atomicOr(&a.counters->overflow, 1u);
        }
    }
    cx.ncand_total += total;
}

template <bool SLICE, bool ODD16>
__global__ void __launch_bounds__(kScan2Threads, 1) scan2_kernel(const ScanArgs a) {
    extern __shared__ __align__(16) unsigned char smem2[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint32_t *s_apron = reinterpret_cast<uint32_t *>(smem2 + kScan2Lut) + (size_t) warp * 32 * kApronWords;

    // one-time staging of the magnitude table, already in its shared-memory layout (the only block-wide barrier)
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(a.lut_swz2);
        uint4 *dst = reinterpret_cast<uint4 *>(smem2);
        constexpr int kUnits = (int) (kScan2Lut / 16);
        constexpr int kPer = (kUnits + kScan2Threads - 1) / kScan2Threads;
        uint4 v[kPer];
#pragma unroll
        for (int q = 0; q < kPer; ++q)
            if (q * kScan2Threads + tid < kUnits)
                v[q] = __ldg(src + q * kScan2Threads + tid);
#pragma unroll
        for (int q = 0; q < kPer; ++q)
            if (q * kScan2Threads + tid < kUnits)
                dst[q * kScan2Threads + tid] = v[q];
    }
    __syncthreads();

    WarpCtx cx;
    cx.ncand_total = 0;
    if (warp == 0) {
        // The edge tiles of the span first (the generic path of scan_kernel, with this kernel's table layout): they
        // take two to three times as long as an interior tile, and handed out last they would be the tail of the
        // launch.  The warps that take them simply come to the interior queue later.
        uint32_t *s_ring = reinterpret_cast<uint32_t *>(smem2 + kScan2Lut + kScan2Apron);
        for (;;) {
            uint32_t tile = 0;
            if (lane == 0)
                tile = atomicAdd(&a.counters->next_tile, 1u);
            tile = __shfl_sync(0xffffffffu, tile, 0);
            if (tile >= a.fast_lo)
                tile += a.fast_hi - a.fast_lo;
            if (tile >= a.ntiles)
                break;
            process_tile<0, SLICE, true, 2>(a, cx, tile, reinterpret_cast<const uint16_t *>(smem2), s_ring);
        }
    }
    for (;;) {
        uint32_t q = 0;
        if (lane == 0)
            q = atomicAdd(&a.counters->next_tile2, 1u);
        q = __shfl_sync(0xffffffffu, q, 0);
        const uint32_t tile = a.fast_lo + q;
        if (tile >= a.fast_hi)
            break;
        scan2_tile<SLICE, ODD16>(a, cx, tile, smem2, s_apron);
    }
    if (lane == 0 && cx.ncand_total) // one same-address atomic per warp, not per tile
        atomicAdd(&a.counters->n_cand, cx.ncand_total);
}

} // namespace

cudaError_t scan2_configure() {
    cudaError_t e;
#define CFG2(S, O)                                                                                                  \
    e = cudaFuncSetAttribute(scan2_kernel<S, O>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) kScan2Smem);   \
    if (e != cudaSuccess)                                                                                           \
        return e;
    CFG2(true, true) CFG2(true, false) CFG2(false, true) CFG2(false, false)
#undef CFG2
    return cudaSuccess;
}

bool scan2_supports(const ScanArgs &a) {
    // uc8 through the table; thresholds for which 96 m + thr * 5 m stays inside 32 bits (any configured value does)
    return a.format == 0 && a.block_samples % 8 == 0 && a.block_samples >= 8 && ((uintptr_t) a.iq & 15u) == 0;
}

void scan2_tile_range(uint64_t nsamples, uint32_t &lo, uint32_t &hi) {
    // interior tiles: the tile's samples and the look-ahead of its last position lie inside the span's own samples.
    // Tile t starts at sample t * kTile - kHead (t >= 1) and the old kernel's criterion -- all 17 of its 512-sample
    // steps inside [0, n) -- is kept, so both kernels agree on which tiles are edge tiles.
    lo = hi = 0;
    const long long n = (long long) nsamples;
    const long long last = (n + kHead - (long long) (kScanSteps + 1) * kStep) / kTile; // largest t with t*kTile - kHead + 17*512 <= n
    if (n + kHead >= (long long) (kScanSteps + 1) * kStep && last >= 1) {
        lo = 1;
        hi = (uint32_t) last + 1;
    }
}

// scan3.inl: the packed-FP32 edition.  Bit-identical output (it runs the whole GPU suite when selected) and 25 % fewer
// instructions, but measured slower on the B200 (scan only: 0.177 of the HBM roofline against 0.266): its loop body
// is 23.6 KB of SASS, which the instruction caches do not hold for 12 unsynchronised warps (ncu: top stall
// no_instruction, issue slots 42 % busy), and its 168 registers leave 12 warps per SM.  Kept selectable
// (B200_K1A=scan3) as the measured alternative; scan2_kernel is the product path.
cudaError_t launch_scan3(const ScanArgs &a, int mode, int grid, cudaStream_t stream);
bool scan3_supports(const ScanArgs &a);
int scan3_warps_per_cta();

static bool use_scan3() {
    static const bool on = getenv("B200_K1A") && !strcmp(getenv("B200_K1A"), "scan3"); // development switch
    return on;
}

int k1a_warps_per_cta(uint32_t format) {
    return format == 0 ? (use_scan3() ? scan3_warps_per_cta() : kScan2Warps) : kScanWarps;
}

cudaError_t launch_scan2(const ScanArgs &a, int mode, int grid, cudaStream_t stream) {
    if (a.fast_hi <= a.fast_lo)
        return cudaSuccess;
    if (use_scan3() && scan3_supports(a))
        return launch_scan3(a, mode, grid, stream);
    const int useful = (int) ((a.fast_hi - a.fast_lo + kScan2Warps - 1) / kScan2Warps);
    if (grid > useful)
        grid = useful;
    // a run starts at iq + 2 * (tile * kTile - kHead + lane * 256): 16 or 0 bytes past a 32-byte boundary
    const bool odd16 = (((uintptr_t) a.iq - 2 * (uintptr_t) kHead) & 31u) != 0;
    if (mode) {
        if (odd16)
            scan2_kernel<true, true><<<grid, kScan2Threads, kScan2Smem, stream>>>(a);
        else
            scan2_kernel<true, false><<<grid, kScan2Threads, kScan2Smem, stream>>>(a);
    } else {
        if (odd16)
            scan2_kernel<false, true><<<grid, kScan2Threads, kScan2Smem, stream>>>(a);
        else
            scan2_kernel<false, false><<<grid, kScan2Threads, kScan2Smem, stream>>>(a);
    }
    return cudaGetLastError();
}
