// scan3.inl -- K1a for uc8 spans, packed-FP32 edition of the register-window kernel: an experiment that is exact but
// NOT faster on the B200 (see the note at use_scan3() in scan2.inl); selected with B200_K1A=scan3.  (Included by kernels.cu after
// scan2.inl, whose helpers and data flow it shares: lane-contiguous runs, 256-bit loads and stores, the apron for
// a run's look-ahead, IDP.2A block sums, candidates emitted at the end of a tile).
//
// scan2_kernel is bound by instruction issue: 30 thread-instructions per sample with every pipe near balance
// (profiles/r02_*).  Blackwell's packed FP32 instructions (FFMA2 / FADD2, two IEEE fp32 operations per lane and
// issue slot) halve the slots the correlators take, provided the two halves are independent streams -- pairing
// neighbouring positions would need every magnitude in two register pairs.  So a lane splits its run of 256
// positions into stream A (positions 0..127) and stream B (128..255), walks both in lockstep and keeps the window
// as float2 (A, B) per sample.  Per window start i (two positions, one per stream):
//   bn  = m5 + m8 + m16 + m17 + m18                  d2 = m2 - m3      d10 = m10 - m11
//   E45 = 32 (m1 + m4 + m9 + m12 - d2 - d10) + 31 - thr bn
//   E67 = E45 + 64 d10        E8 = E67 + 96 d2 - 32 m9        g = sign(m7 - m1) & sign(m14 - m12) & sign(m15 - m12)
// 19 packed operations instead of 2 x 17 integer ones; the sign bits are collected per stream as before.
//
// Exactness.  Magnitudes are integers <= 65535, so every sum of them that appears here is an integer below 2^24
// in absolute value and exact in fp32, up to the thr * bn term.  An fp32 result is inexact only if its exact value
// is >= 2^24 in absolute value, and rounding never changes a sign, so a sign can only be wrong if a value was
// rounded earlier in the chain and the chain then came back to zero.  It cannot: 32 (...) + 31 <= 32 * 6 * 65535
// + 31 < 2^24, so a rounded E45 (or E67) is below -2^24, and the terms added afterwards (64 d10 <= 4.2 M, 96 d2
// <= 6.3 M, -32 m9 <= 0) leave E67 and E8 below -6 M; the one positive intermediate that can pass 2^24
// (E67 + 96 d2 <= 18.9 M) is followed only by -32 m9 >= -2.1 M, which leaves E8 above 14 M.  The pre-check
// differences are exact, and x - x is +0, whose clear sign bit is "not greater", as the reference's > demands.

namespace {

constexpr int kScan3Warps = 12;                  // 170 registers per thread
constexpr int kScan3Threads = kScan3Warps * 32;
constexpr int kHalfRun = kRun / 2;               // positions per stream
constexpr int kBodies3 = kHalfRun / 32;          // 4 loop bodies of 32 samples per stream
constexpr size_t kScan3Apron = (size_t) kScan3Warps * 2 * 32 * kApronWords * sizeof(uint32_t);
constexpr size_t kScan3Smem = kScan2Lut + kScan3Apron + kSmemWarp; // + one ring buffer for the edge tiles

typedef unsigned long long f2; // two fp32 in a register pair: low half = stream A, high half = stream B

__device__ __forceinline__ f2 f2_pack(float a, float b) {
    f2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ f2 f2_add(f2 a, f2 b) {
    f2 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f2 f2_fma(f2 a, f2 b, f2 c) {
    f2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ uint32_t f2_lo(f2 a) { return (uint32_t) a; }
__device__ __forceinline__ uint32_t f2_hi(f2 a) { return (uint32_t) (a >> 32); }

struct Consts3 {
    f2 neg1, c32, c31, c64, c96, cn32, nthr;
};

struct Signs3 { // sign bits of the tests of one stream, newest position in bit 0
    uint32_t s45, s67, s8, pm;
};

// The two samples of both streams that just arrived sit in ring slots s0 = 2 * JJ and s0 + 1; the windows that end
// with them start at s0 - 18 and s0 - 17.
template <int JJ>
__device__ __forceinline__ void test_pair3(const f2 (&m)[32], const Consts3 &k, Signs3 &sa, Signs3 &sb) {
    constexpr int s0 = 2 * JJ + 64;
#define M_(x) m[(x) & 31]
#pragma unroll
    for (int odd = 0; odd < 2; ++odd) {
        const int i = s0 - 18 + odd;
        const f2 bn = f2_add(f2_add(f2_add(M_(i + 5), M_(i + 8)), f2_add(M_(i + 16), M_(i + 17))), M_(i + 18));
        const f2 d2 = f2_fma(M_(i + 3), k.neg1, M_(i + 2));
        const f2 d10 = f2_fma(M_(i + 11), k.neg1, M_(i + 10));
        const f2 cc = f2_add(f2_add(M_(i + 1), M_(i + 4)), f2_add(M_(i + 9), M_(i + 12)));
        const f2 c = f2_fma(d10, k.neg1, f2_fma(d2, k.neg1, cc));
        const f2 E45 = f2_fma(bn, k.nthr, f2_fma(c, k.c32, k.c31));
        const f2 E67 = f2_fma(d10, k.c64, E45);
        const f2 E8 = f2_fma(M_(i + 9), k.cn32, f2_fma(d2, k.c96, E67));
        const f2 g1 = f2_fma(M_(i + 1), k.neg1, M_(i + 7));
        const f2 g2 = f2_fma(M_(i + 12), k.neg1, M_(i + 14));
        const f2 g3 = f2_fma(M_(i + 12), k.neg1, M_(i + 15));
        sa.s45 = __funnelshift_l(f2_lo(E45), sa.s45, 1);
        sb.s45 = __funnelshift_l(f2_hi(E45), sb.s45, 1);
        sa.s67 = __funnelshift_l(f2_lo(E67), sa.s67, 1);
        sb.s67 = __funnelshift_l(f2_hi(E67), sb.s67, 1);
        sa.s8 = __funnelshift_l(f2_lo(E8), sa.s8, 1);
        sb.s8 = __funnelshift_l(f2_hi(E8), sb.s8, 1);
        sa.pm = __funnelshift_l(f2_lo(g1) & f2_lo(g2) & f2_lo(g3), sa.pm, 1);
        sb.pm = __funnelshift_l(f2_hi(g1) & f2_hi(g2) & f2_hi(g3), sb.pm, 1);
    }
#undef M_
}

struct TileMasks3 { // [stream][word]: bit k of word w = stream-local position 32 w + k - 18 (word 4: the 18 tail positions)
    uint32_t b45[2][kBodies3 + 1], b67[2][kBodies3 + 1], b8[2][kBodies3 + 1];
};

__device__ __forceinline__ void finish_masks3(const Signs3 &s, int shift, uint32_t &b45, uint32_t &b67, uint32_t &b8) {
    // the first position tested sits in the top bit (of the low 32 - shift bits): reverse
    const uint32_t pm = __brev(s.pm << shift);
    b45 = pm & ~__brev(s.s45 << shift);
    b67 = pm & ~__brev(s.s67 << shift);
    b8 = pm & ~__brev(s.s8 << shift);
}

// SPLIT: a mag_buf boundary falls into some lane's run of this tile, so the two streams of a lane may sum into
// different mag_bufs: per-stream sums with a boundary check every 8 samples.  Otherwise one set of sums, no checks.
template <bool SLICE, bool ODD16, bool SPLIT>
__device__ __forceinline__ void scan3_tile(const ScanArgs &a, WarpCtx &cx, const uint32_t tile, const unsigned char *s_lut, uint32_t *s_apron) {
    const int lane = threadIdx.x & 31;
    const long long c0 = (long long) tile * kTile - kHead; // first window-start sample of the tile (>= 0: interior)
    const long long s_run = c0 + (long long) lane * kRun;  // first sample of the lane's run
    const uint8_t *gpa = a.iq + s_run * 2 - (ODD16 ? 16 : 0); // 32-byte aligned
    const uint8_t *gpb = gpa + kHalfRun * 2;
    uint16_t *gma = a.mag + (size_t) tile * kTile + (size_t) lane * kRun, *gmb = gma + kHalfRun;
    uint32_t *apron_a = s_apron, *apron_b = s_apron + 32 * kApronWords; // [lane][kApronWords] each

    Consts3 k;
    {
        const float thr = (float) a.threshold;
        k.neg1 = f2_pack(-1.0f, -1.0f);
        k.c32 = f2_pack(32.0f, 32.0f);
        k.c31 = f2_pack(31.0f, 31.0f);
        k.c64 = f2_pack(64.0f, 64.0f);
        k.c96 = f2_pack(96.0f, 96.0f);
        k.cn32 = f2_pack(-32.0f, -32.0f);
        k.nthr = f2_pack(-thr, -thr);
    }

    // ---- block sums bookkeeping: sums[0] alone unless SPLIT ----
    LaneSums sums[2];
#pragma unroll
    for (int x = 0; x < 2; ++x) {
        sums[x].level = sums[x].power = 0;
        sums[x].level32 = sums[x].lo32 = sums[x].hi32 = 0;
        const unsigned long long B = a.block_samples, s0 = (unsigned long long) s_run + (unsigned long long) (x * kHalfRun);
        sums[x].blk = (uint32_t) (s0 / B);
        const unsigned long long to_next = ((unsigned long long) sums[x].blk + 1) * B - s0;
        sums[x].g8_next = to_next < (unsigned long long) kHalfRun ? (uint32_t) (to_next >> 3) : 0xffffffffu;
    }

    // ---- the tile's tail: the 18 samples after its last run belong to the next tile; lanes 0..8 convert one pair each ----
    __syncwarp(); // the previous tile's apron reads are done
    if (lane < 9) {
        const uint32_t w = __ldg(reinterpret_cast<const uint32_t *>(a.iq + (c0 + kTile) * 2) + lane);
        uint32_t m0, m1;
        convert_pair(s_lut, w, m0, m1);
        apron_a[31 * kApronWords + lane] = m0 | (m1 << 16);
    }

    f2 m[32];
#pragma unroll
    for (int i = 0; i < 32; ++i)
        m[i] = 0;
    TileMasks3 tm;

    // 16 samples (one 256-bit block) per stream in flight
    uint32_t nxa[8], nxb[8], cya[4] = {0, 0, 0, 0}, cyb[4] = {0, 0, 0, 0};
    if (ODD16) {
        uint32_t fa[8], fb[8];
        ldg256(fa, gpa);
        ldg256(fb, gpb);
        gpa += 32;
        gpb += 32;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            cya[i] = fa[4 + i];
            cyb[i] = fb[4 + i];
        }
    }
    ldg256(nxa, gpa);
    ldg256(nxb, gpb);

#pragma unroll 1
    for (int body = 0; body < kBodies3; ++body) {
        Signs3 sa = {0, 0, 0, 0}, sb = {0, 0, 0, 0};
        uint32_t va[16], vb[16];
        uint32_t cua[8], cub[8];
        auto refill = [&](int half) { // the next 16 samples of both streams become current; their successors are requested
            if (ODD16) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    cua[i] = cya[i];
                    cub[i] = cyb[i];
                    cua[4 + i] = nxa[i];
                    cub[4 + i] = nxb[i];
                    cya[i] = nxa[4 + i];
                    cyb[i] = nxb[4 + i];
                }
            } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    cua[i] = nxa[i];
                    cub[i] = nxb[i];
                }
            }
            // (ODD16: gpa / gpb already point one block on, so the same index fetches the block whose lower half
            // ends the next group)
            const int nexth = 2 * body + half + 1; // the next 16-sample group of the streams, 8 = past their end
            if (nexth < 2 * kBodies3) {
                ldg256(nxa, gpa + nexth * 32);
                ldg256(nxb, gpb + nexth * 32);
            }
        };
        auto step = [&](auto jj_c) {
            constexpr int JJ = decltype(jj_c)::value;
            if (SPLIT && JJ % 4 == 0) { // an 8-sample group starts: does a mag_buf end here, in either stream?
#pragma unroll
                for (int x = 0; x < 2; ++x)
                    if ((uint32_t) (body * 4 + JJ / 4) == sums[x].g8_next) {
                        lane_flush(a.block_sums_u64, sums[x]);
                        ++sums[x].blk;
                        sums[x].g8_next = 0xffffffffu; // block sizes >= 256: one boundary per stream at most
                    }
            }
            uint32_t a0, a1, b0, b1;
            convert_pair(s_lut, cua[JJ & 7], a0, a1);
            convert_pair(s_lut, cub[JJ & 7], b0, b1);
            const uint32_t wa = a0 | (a1 << 16), wb = b0 | (b1 << 16);
            va[JJ] = wa;
            vb[JJ] = wb;
            // sums (convert.c:95-110): sum of mag, and of mag^2 = mag * lo8(mag) + 256 * mag * hi8(mag)
            LaneSums &su_a = sums[0], &su_b = sums[SPLIT ? 1 : 0];
            su_a.level32 += a0 + a1;
            su_b.level32 += b0 + b1;
            const uint32_t ba = __byte_perm(wa, 0, 0x3120), bb = __byte_perm(wb, 0, 0x3120);
            su_a.lo32 = dp2a_lo_u(wa, ba, su_a.lo32);
            su_a.hi32 = dp2a_hi_u(wa, ba, su_a.hi32);
            su_b.lo32 = dp2a_lo_u(wb, bb, su_b.lo32);
            su_b.hi32 = dp2a_hi_u(wb, bb, su_b.hi32);
            m[(2 * JJ) & 31] = f2_pack((float) a0, (float) b0);
            m[(2 * JJ + 1) & 31] = f2_pack((float) a1, (float) b1);
            test_pair3<JJ>(m, k, sa, sb);
        };
#define STEP_(J) step(std::integral_constant<int, J>{});
        refill(0);
        STEP_(0) STEP_(1) STEP_(2) STEP_(3) STEP_(4) STEP_(5) STEP_(6) STEP_(7)
        if (SLICE) {
            stg256(gma + body * 32, *reinterpret_cast<uint32_t(*)[8]>(&va[0]));
            stg256(gmb + body * 32, *reinterpret_cast<uint32_t(*)[8]>(&vb[0]));
        }
        refill(1);
        STEP_(8) STEP_(9) STEP_(10) STEP_(11) STEP_(12) STEP_(13) STEP_(14) STEP_(15)
#undef STEP_
        if (SLICE) {
            stg256(gma + body * 32 + 16, *reinterpret_cast<uint32_t(*)[8]>(&va[8]));
            stg256(gmb + body * 32 + 16, *reinterpret_cast<uint32_t(*)[8]>(&vb[8]));
        }
        if (body == 0) {
            // the first 18 magnitudes of a stream are the look-ahead of the stream in front of it
            if (lane > 0) {
                uint32_t *ap = apron_a + (lane - 1) * kApronWords;
                *reinterpret_cast<uint4 *>(ap) = make_uint4(va[0], va[1], va[2], va[3]);
                *reinterpret_cast<uint4 *>(ap + 4) = make_uint4(va[4], va[5], va[6], va[7]);
                ap[8] = va[8];
            }
            uint32_t *ap = apron_b + lane * kApronWords;
            *reinterpret_cast<uint4 *>(ap) = make_uint4(vb[0], vb[1], vb[2], vb[3]);
            *reinterpret_cast<uint4 *>(ap + 4) = make_uint4(vb[4], vb[5], vb[6], vb[7]);
            ap[8] = vb[8];
        }
        // widen the body's partial sums
#pragma unroll
        for (int x = 0; x < (SPLIT ? 2 : 1); ++x) {
            sums[x].level += sums[x].level32;
            sums[x].power += (unsigned long long) sums[x].lo32 + ((unsigned long long) sums[x].hi32 << 8);
            sums[x].level32 = sums[x].lo32 = sums[x].hi32 = 0;
        }
        finish_masks3(sa, 0, tm.b45[0][body], tm.b67[0][body], tm.b8[0][body]);
        finish_masks3(sb, 0, tm.b45[1][body], tm.b67[1][body], tm.b8[1][body]);
    }

    // ---- the streams' last 18 positions: A looks ahead into B's first samples, B into the next lane's A ----
    __syncwarp();
    {
        const uint32_t *pa = apron_b + lane * kApronWords, *pb = apron_a + lane * kApronWords;
        const uint4 qa0 = *reinterpret_cast<const uint4 *>(pa), qa1 = *reinterpret_cast<const uint4 *>(pa + 4);
        const uint4 qb0 = *reinterpret_cast<const uint4 *>(pb), qb1 = *reinterpret_cast<const uint4 *>(pb + 4);
        const uint32_t wa[9] = {qa0.x, qa0.y, qa0.z, qa0.w, qa1.x, qa1.y, qa1.z, qa1.w, pa[8]};
        const uint32_t wb[9] = {qb0.x, qb0.y, qb0.z, qb0.w, qb1.x, qb1.y, qb1.z, qb1.w, pb[8]};
        Signs3 sa = {0, 0, 0, 0}, sb = {0, 0, 0, 0};
        auto step = [&](auto jj_c) {
            constexpr int JJ = decltype(jj_c)::value;
            m[(2 * JJ) & 31] = f2_pack((float) (wa[JJ] & 0xffffu), (float) (wb[JJ] & 0xffffu));
            m[(2 * JJ + 1) & 31] = f2_pack((float) (wa[JJ] >> 16), (float) (wb[JJ] >> 16));
            test_pair3<JJ>(m, k, sa, sb);
        };
#define STEP_(J) step(std::integral_constant<int, J>{});
        STEP_(0) STEP_(1) STEP_(2) STEP_(3) STEP_(4) STEP_(5) STEP_(6) STEP_(7) STEP_(8)
#undef STEP_
        finish_masks3(sa, 14, tm.b45[0][kBodies3], tm.b67[0][kBodies3], tm.b8[0][kBodies3]);
        finish_masks3(sb, 14, tm.b45[1][kBodies3], tm.b67[1][kBodies3], tm.b8[1][kBodies3]);
    }
    // bits 0..17 of a stream's word 0 are the last positions of the stream in front of it (tested there)
#pragma unroll
    for (int x = 0; x < 2; ++x) {
        tm.b45[x][0] &= ~0x3ffffu;
        tm.b67[x][0] &= ~0x3ffffu;
        tm.b8[x][0] &= ~0x3ffffu;
    }

    // ---- block sums of the run ----
    if (!SPLIT) {
        const unsigned long long l = warp_sum_u64(sums[0].level), p = warp_sum_u64(sums[0].power);
        if (lane == 0 && (l | p)) {
            atomicAdd(&a.block_sums_u64[2 * (size_t) sums[0].blk], l);
            atomicAdd(&a.block_sums_u64[2 * (size_t) sums[0].blk + 1], p);
        }
    } else {
        lane_flush(a.block_sums_u64, sums[0]); // a mag_buf boundary inside the tile: every lane and stream for itself
        lane_flush(a.block_sums_u64, sums[1]);
    }

    // ---- candidates, in position order: lane after lane, stream A before stream B ----
    uint32_t mine = 0;
#pragma unroll
    for (int x = 0; x < 2; ++x)
#pragma unroll
        for (int w = 0; w <= kBodies3; ++w)
            mine += (uint32_t) __popc(tm.b45[x][w] | tm.b67[x][w] | tm.b8[x][w]);
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
        for (int x = 0; x < 2; ++x)
#pragma unroll
            for (int w = 0; w <= kBodies3; ++w)
                for (int b = 0; b < 32; ++b) {
                    const int r = 32 * w + b - 18;
                    if (r >= 0 && r < kHalfRun)
                        dst[x * kHalfRun + r] = (uint8_t) ((((tm.b45[x][w] >> b) & 1u) * 3u) | (((tm.b67[x][w] >> b) & 1u) * 12u) |
                                                           (((tm.b8[x][w] >> b) & 1u) * 16u));
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
            for (int x = 0; x < 2; ++x)
#pragma unroll
                for (int w = 0; w <= kBodies3; ++w) {
                    uint32_t u = tm.b45[x][w] | tm.b67[x][w] | tm.b8[x][w];
                    while (u) {
                        const int b = __ffs(u) - 1;
                        u &= u - 1;
                        const uint32_t t5 = (((tm.b45[x][w] >> b) & 1u) * 3u) | (((tm.b67[x][w] >> b) & 1u) * 12u) | (((tm.b8[x][w] >> b) & 1u) * 16u);
                        if (ci < cand_cap)
                            out[ci] = (uint32_t) (lane * kRun + x * kHalfRun + 32 * w + b - 18) | (t5 << 13);
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
__global__ void __launch_bounds__(kScan3Threads, 1) scan3_kernel(const ScanArgs a) {
    extern __shared__ __align__(16) unsigned char smem3[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint32_t *s_apron = reinterpret_cast<uint32_t *>(smem3 + kScan2Lut) + (size_t) warp * 2 * 32 * kApronWords;

    // one-time staging of the magnitude table, already in its shared-memory layout (the only block-wide barrier)
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(a.lut_swz2);
        uint4 *dst = reinterpret_cast<uint4 *>(smem3);
        constexpr int kUnits = (int) (kScan2Lut / 16);
        constexpr int kPer = (kUnits + kScan3Threads - 1) / kScan3Threads;
        for (int q0 = 0; q0 < kPer; q0 += 8) { // eight 16-byte loads per thread in flight at a time
            uint4 v[8];
#pragma unroll
            for (int q = 0; q < 8; ++q)
                if ((q0 + q) * kScan3Threads + tid < kUnits)
                    v[q] = __ldg(src + (q0 + q) * kScan3Threads + tid);
#pragma unroll
            for (int q = 0; q < 8; ++q)
                if ((q0 + q) * kScan3Threads + tid < kUnits)
                    dst[(q0 + q) * kScan3Threads + tid] = v[q];
        }
    }
    __syncthreads();

    WarpCtx cx;
    cx.ncand_total = 0;
    if (warp == 0) {
        // the edge tiles of the span first (see scan2_kernel)
        uint32_t *s_ring = reinterpret_cast<uint32_t *>(smem3 + kScan2Lut + kScan3Apron);
        for (;;) {
            uint32_t tile = 0;
            if (lane == 0)
                tile = atomicAdd(&a.counters->next_tile, 1u);
            tile = __shfl_sync(0xffffffffu, tile, 0);
            if (tile >= a.fast_lo)
                tile += a.fast_hi - a.fast_lo;
            if (tile >= a.ntiles)
                break;
            process_tile<0, SLICE, true, 2>(a, cx, tile, reinterpret_cast<const uint16_t *>(smem3), s_ring);
        }
    }
    const unsigned long long B = a.block_samples;
    for (;;) {
        uint32_t q = 0;
        if (lane == 0)
            q = atomicAdd(&a.counters->next_tile2, 1u);
        q = __shfl_sync(0xffffffffu, q, 0);
        const uint32_t tile = a.fast_lo + q;
        if (tile >= a.fast_hi)
            break;
        // does a mag_buf boundary fall inside the tile's samples?
        const unsigned long long first = (unsigned long long) tile * kTile - kHead;
        if (first / B != (first + kTile - 1) / B)
            scan3_tile<SLICE, ODD16, true>(a, cx, tile, smem3, s_apron);
        else
            scan3_tile<SLICE, ODD16, false>(a, cx, tile, smem3, s_apron);
    }
    if (lane == 0 && cx.ncand_total) // one same-address atomic per warp, not per tile
        atomicAdd(&a.counters->n_cand, cx.ncand_total);
}

} // namespace

cudaError_t scan3_configure() {
    cudaError_t e;
#define CFG3(S, O)                                                                                                  \
    e = cudaFuncSetAttribute(scan3_kernel<S, O>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) kScan3Smem);   \
    if (e != cudaSuccess)                                                                                           \
        return e;
    CFG3(true, true) CFG3(true, false) CFG3(false, true) CFG3(false, false)
#undef CFG3
    return cudaSuccess;
}

bool scan3_supports(const ScanArgs &a) {
    // one mag_buf boundary per stream of 128 samples at most
    return a.block_samples >= 256;
}

int scan3_warps_per_cta() {
    return kScan3Warps;
}

cudaError_t launch_scan3(const ScanArgs &a, int mode, int grid, cudaStream_t stream) {
    if (a.fast_hi <= a.fast_lo)
        return cudaSuccess;
    const int useful = (int) ((a.fast_hi - a.fast_lo + kScan3Warps - 1) / kScan3Warps);
    if (grid > useful)
        grid = useful;
    const bool odd16 = (((uintptr_t) a.iq - 2 * (uintptr_t) kHead) & 31u) != 0;
    if (mode) {
        if (odd16)
            scan3_kernel<true, true><<<grid, kScan3Threads, kScan3Smem, stream>>>(a);
        else
            scan3_kernel<true, false><<<grid, kScan3Threads, kScan3Smem, stream>>>(a);
    } else {
        if (odd16)
            scan3_kernel<false, true><<<grid, kScan3Threads, kScan3Smem, stream>>>(a);
        else
            scan3_kernel<false, false><<<grid, kScan3Threads, kScan3Smem, stream>>>(a);
    }
    return cudaGetLastError();
}
