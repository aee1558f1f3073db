// device_types.h -- records exchanged between the kernels and the host resolver.
#pragma once

#include <stdint.h>

namespace b200 {

// ---- geometry ---------------------------------------------------------------------------
constexpr int kOverlap = 326;  // Modes.trailing_samples at 2.4 MS/s (readsb.c:198)
constexpr int kHead = 328;     // samples carried in front of a span: kOverlap rounded up to 16 bytes
constexpr int kPosShift = kHead - kOverlap; // 2

// A tile (K1 "segment") is kTile consecutive scan positions whose preamble windows start at a
// 16-byte aligned sample: tile t covers positions [t*kTile - kPosShift, (t+1)*kTile - kPosShift),
// i.e. window-start samples [t*kTile - kHead, (t+1)*kTile - kHead) relative to the span's first
// new sample.  It also owns the block sums of exactly those samples.
constexpr int kTile = 8192;
constexpr int kStep = 512;                 // samples one warp converts per step (16 per lane)
constexpr int kLanePos = kStep / 32;       // 16 scan positions per lane per step
constexpr int kScanSteps = kTile / kStep;  // 16 steps of window starts
constexpr int kScanWarps = 19;             // warps per K1a CTA (shared memory: 128 KiB table + 5 KiB ring per warp)
constexpr int kScanThreads = kScanWarps * 32;
// Warp buffer: a ring of two chunks of u32 magnitudes, one row per lane.  A row is the lane's 16
// magnitudes plus 16 bytes of padding: the 80-byte row stride makes every 128-bit access of a
// quarter warp hit 8 distinct 16-byte bank groups.
constexpr int kRowWords = kLanePos + 4;     // 20
constexpr int kRows = 64;                   // 2 chunks x 32 lanes
constexpr int kWarpBuf = kRows * kRowWords; // u32 words per warp (5120 bytes)
constexpr int kMagSlack = 1024;            // magnitudes K1b may stage past the last tile (never used by a frame)

inline uint32_t tiles_for(uint64_t nsamples) {
    // every position < nsamples and every sample < nsamples must fall into a tile
    return nsamples ? (uint32_t) ((nsamples + kHead - 1) / kTile + 1) : 0;
}

// ---- scoring classes of a sliced frame that does not score -2 outright ------------------
// (the filter-independent half of scoreModesMessage, mode_s.c:311-409)
enum : uint32_t {
    kKindBad = 0,     // -2 whatever the ICAO filter holds
    kKindAP = 1,      // DF0/4/5/16/24: filter(crc) ? 1000 : -1
    kKindAPCommB = 2, // DF20/21:       filter(crc) ? 1000 : -2
    kKindDF11 = 3,    // all-call reply, syndrome (ignoring IID) clean or 1-bit repairable
    kKindES = 4,      // DF17/18, syndrome clean or repairable
};

// K1b -> K2: one per (candidate position, phase) whose class is not kKindBad, with the frame as sliced: K2 only
// has to add the signal power of the ones that stay live.  32 bytes.
struct PhaseRec {
    uint32_t pos;     // scan position within the span
    uint32_t w0;      // crc[23:0] | kind[26:24] | errors[29:28]
    uint32_t w1;      // key[23:0] (address the score depends on) | phase[27:24]
    uint32_t errbits; // bit0[7:0] | bit1[15:8]: the frame bits the syndrome table would repair (0xff = none)
    uint8_t msg[14];  // the frame as sliced, MSB first; bytes past a short frame are 0
    uint8_t pad[2];
};

// K1 per tile: where the tile's candidate entries and class records are, and how many
struct TileDesc {
    uint32_t cand_off, ncand;
    uint32_t rec_off, nrec;
};

// candidate entry (K1): pos_in_tile[12:0] | trymask[17:13]
// dead entry (K2):      pos_in_tile[12:0] | trymask[17:13] | unknown_icao[18]

// K2 per tile
struct TileOut {
    uint32_t dead_off, ndead;
    uint32_t live_off, nlive;
    uint32_t liverec_off, nliverec;
};

// K2 -> host: a position the resolver must look at.  16 bytes.
struct LivePos {
    uint32_t pos;       // scan position within the span
    uint32_t info;      // trymask[4:0] | nrec[10:8] | first live record (relative to the tile's liverec_off) [31:16]
    uint32_t dead_rank; // dead entries of the tile in front of this position (where a skip-ahead starts counting)
    uint32_t pad;       // 0 from K2; in the ordered list the host walks: index of the position's first record
};

// order_live -> host: what an accepted frame at a live position hides from the per-block dead totals -- the dead
// positions in (pos, pos + 134] (a 56-bit frame) and in (pos, pos + 268] (a 112-bit frame), both cut at the last
// position of the position's mag_buf (demod_2400.c:416: the for loop ends with the block).  Eight 16-bit counters
// per case: lo = preambles | rejected_bad << 16 | rejected_unknown << 32 | phase[0] << 48, hi = phase[1..4].  32 bytes.
struct LiveHidden {
    uint64_t short_lo, short_hi;
    uint64_t long_lo, long_hi;
};

// K2 -> host: a sliced frame of a live position.  40 bytes.
struct LiveRec {
    uint32_t pos;
    uint32_t w0;    // as PhaseRec, plus bit 31: the key is in the device-side address set S (outside it the ICAO filter says no)
    uint32_t w1;
    uint32_t errbits; // bit0[7:0] | bit1[15:8] (0xff = none)
    uint64_t power; // sum of m^2 over the frame's 134/268 samples (demod_2400.c:393-396)
    uint8_t msg[14];
    uint8_t pad[2];
};

// Mode A/C kernel -> host: a position where a reply decodes (demod_2400.c:577-683).  16 bytes.
struct AcHit {
    uint32_t q;        // block * block_samples + data index of F1 (data[0] = 326 samples before the block's first new one)
    uint32_t f1_clock; // 60 MHz ticks from data[0] (demod_2400.c:596)
    uint32_t modeac;   // 00 A4 A2 A1  00 B4 B2 B1  SPI C4 C2 C1  00 D4 D2 D1
    uint32_t pad;
};

// per mag_buf counters of positions that can never be accepted (K2)
struct BlockDead {
    uint32_t preambles;
    uint32_t rejected_bad;
    uint32_t rejected_unknown;
    uint32_t phase[5];
};

struct ScanCounters {
    unsigned long long n_cand;
    unsigned long long n_rec;
    unsigned long long n_dead;
    unsigned long long n_live;
    unsigned long long n_liverec;
    unsigned int overflow; // bit0 cand, bit1 rec, bit2 dead, bit3 live, bit4 liverec, bit5 Mode A/C hits
    unsigned int next_tile; // K1a work queue (scan_kernel: the edge tiles, or every tile)
    unsigned int next_tile2; // K1a work queue of scan2_kernel (the interior tiles)
    unsigned int n_modeac_hits;   // Mode A/C kernel
};

struct ErrorInfo { // struct errorinfo, crc.h:32-37
    uint32_t syndrome;
    int32_t errors;
    int8_t bit[2];
    uint16_t padding;
};

} // namespace b200
