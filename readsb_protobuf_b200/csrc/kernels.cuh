// kernels.cuh -- launch interface of the sm_100a kernels (kernels.cu).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "device_types.h"

namespace b200 {

struct ScanArgs {
    // input
    const uint8_t *iq;    // new samples of the span, 16-byte aligned
    const uint8_t *head;  // kHead samples carried from the previous span, 16-byte aligned
    uint32_t head_valid;  // how many of them are real (0 at stream start: magnitude 0, fifo.c:47)
    uint32_t format;      // B200_INPUT_*, 3 = the stream already holds u16 magnitudes (--dcfilter front end),
                          // 4 = sc16q11 through the magnitude table of a -DSC16Q11_TABLE_BITS build
    uint64_t nsamples;    // new samples == scan positions of the span
    int32_t threshold;    // Modes.preambleThreshold
    uint32_t block_samples;
    uint32_t ntiles;
    // tables
    const uint16_t *lut;  // magnitude table of the format (65536 entries): uc8, or format 4's sc16q11 table
    const uint16_t *lut_swz; // the same table in K1's bank-swizzled shared-memory layout
    const uint16_t *lut_swz2; // ... and in scan2_kernel's (entry i at i ^ ((i >> 5) & 0x38))
    uint32_t fast_lo, fast_hi; // tiles [fast_lo, fast_hi) are scan2_kernel's: scan_kernel leaves them out
    int32_t table_bits;   // format 4: SC16Q11_TABLE_BITS (<= 8)
    const ErrorInfo *tab_short;
    const ErrorInfo *tab_long;
    int32_t n_short, n_long;
    uint32_t *addr_bitmap; // 2^24 bits: every address icaoFilterAdd could ever see
    // output: every tile writes into its own slab of the two arrays
    uint32_t *cand;
    PhaseRec *recs;
    TileDesc *tiles;
    uint16_t *step_off;   // [ntiles][kScanSteps]: candidates of the tile in front of each scan step
    uint16_t *mag;        // [ntiles * kTile + kMagSlack] magnitudes, index = sample + kHead (K1a writes, K1b/K2 read)
    // slab placement: tile t owns cand[t*cand_slab ...) / recs[t*rec_slab ...), or, when tile_off is
    // given (the exact-fit retry), cand[tile_off[2t] .. ) and recs[tile_off[2t+1] ..) with the caps
    // taken from the next tile's offsets
    uint32_t cand_slab, rec_slab;
    const uint32_t *tile_off; // [2 * (ntiles + 1)] or nullptr
    ScanCounters *counters;
    unsigned long long *block_sums_u64; // [nblocks][2]: sum mag, sum mag^2 (table formats)
    double *block_sums_f64;             // [nblocks][2]: sum mag, sum magsq (float formats)
    uint8_t *dbg_masks;                 // optional: try mask per scan position
};

struct SliceArgs {
    const uint16_t *mag;  // K1a's magnitudes
    const uint32_t *cand; // K1a's candidate entries
    const uint16_t *step_off;
    PhaseRec *recs;
    TileDesc *tiles;      // cand_off / ncand / rec_off from K1a; K1b fills nrec
    uint32_t ntiles;
    uint32_t cand_slab, rec_slab;
    const uint32_t *tile_off; // as ScanArgs
    const ErrorInfo *tab_short;
    const ErrorInfo *tab_long;
    int32_t n_short, n_long;
    uint32_t *addr_bitmap;
    ScanCounters *counters;
};

struct ClassifyArgs {
    const uint8_t *iq;
    const uint8_t *head;
    uint32_t head_valid;
    uint32_t format;
    uint64_t nsamples;
    uint32_t block_samples;
    uint32_t ntiles;
    const uint16_t *lut;
    int32_t table_bits;
    const ErrorInfo *tab_short;
    const ErrorInfo *tab_long;
    int32_t n_short, n_long;
    const uint32_t *addr_bitmap;
    const uint32_t *cand;
    const PhaseRec *recs;
    const TileDesc *tiles;
    const uint16_t *mag;         // K1a's magnitudes (index = sample + kHead), or nullptr
    uint32_t max_cand_per_tile;  // the candidate slab K1a ran with (bounds a tile's list)
    // output
    uint32_t *dead;
    LivePos *live;
    LiveRec *liverecs;
    TileOut *tiles_out;
    uint32_t dead_cap, live_cap, liverec_cap;
    ScanCounters *counters;
    BlockDead *block_dead; // [nblocks]
};

struct ModeacArgs {
    const uint16_t *mag;  // K1a's magnitudes (index = chunk sample + kHead)
    uint64_t nsamples;    // new samples of the chunk
    uint32_t block_samples;
    uint32_t format;
    uint32_t nblocks;     // mag_bufs of the chunk, the short final one included
    const unsigned long long *sums_u64; // K1a's block sums
    const double *sums_f64;
    uint32_t *noise_level; // [nblocks] scratch
    AcHit *hits;
    uint32_t hit_cap;
    ScanCounters *counters;
};

// dynamic shared memory the scan kernel needs for a format
size_t scan_smem_bytes(uint32_t format);
cudaError_t scan_configure();
cudaError_t slice_configure();
// mode 0: magnitude + preamble scan only (candidates counted); 1: + slice + CRC + records
cudaError_t launch_scan(const ScanArgs &a, int mode, int grid, cudaStream_t stream);
// scan2.cu: the register-window K1a for the interior tiles of a uc8 span (same outputs as scan_kernel)
cudaError_t scan2_configure();
cudaError_t scan3_configure();
bool scan2_supports(const ScanArgs &a);
void scan2_tile_range(uint64_t nsamples, uint32_t &lo, uint32_t &hi);
cudaError_t launch_scan2(const ScanArgs &a, int mode, int grid, cudaStream_t stream);
// resident K1a warps per CTA for a format (chunks are sized to whole waves of tiles)
int k1a_warps_per_cta(uint32_t format);
cudaError_t launch_slice(const SliceArgs &a, int grid, cudaStream_t stream);
cudaError_t launch_classify(const ClassifyArgs &a, cudaStream_t stream);
// K2's per-tile live lists (device memory) -> position-ordered arrays (pinned host memory), plus per live position
// what a frame accepted there hides of the dead list (which therefore never leaves the device); base = ntiles scratch
cudaError_t launch_order_live(const TileOut *tiles_out, uint32_t ntiles, const ScanCounters *counters, uint2 *base, const LivePos *live,
                              const LiveRec *recs, const uint32_t *dead, uint64_t nsamples, uint32_t block_samples, uint8_t *packed,
                              cudaStream_t stream);
// sc16 / sc16q11: per-mag_buf sum of mag and of magsq as the reference's sequential float accumulators
// leave them (sums[2k], sums[2k+1]); iq = the span's first new sample
cudaError_t launch_float_block_sums(const uint8_t *iq, uint32_t format, uint64_t nsamples, uint32_t block_samples, uint32_t nblocks,
                                    double *sums, cudaStream_t stream);
// --dcfilter front end (convert.c:113-213, 374-423): raw IQ -> DC-filtered u16 magnitudes (the stream K1a then
// reads as format 3) and the converter's per-mag_buf float sums.  aI / aQ: scratch of nsamples floats rounded
// up to whole batches of 1024 (+ 1024); state: the two filter states z1_I, z1_Q, carried from call to call
cudaError_t launch_dc_front_end(const uint8_t *iq, uint32_t format, uint64_t nsamples, uint32_t block_samples, float dc_a, float dc_b,
                                float *aI, float *aQ, float *state, uint16_t *mag_out, double *sums, cudaStream_t stream);
// Mode A/C framing-pulse search over K1a's magnitudes (after K1a of the same chunk)
cudaError_t launch_modeac(const ModeacArgs &a, cudaStream_t stream);

// IQ -> u16 magnitudes materialised in global memory (+ sums into sums_u64[2] / sums_f64[2])
cudaError_t launch_convert(const uint8_t *iq, uint32_t format, uint32_t nsamples, const uint16_t *lut, int table_bits,
                           uint16_t *mag, unsigned long long *sums_u64, double *sums_f64, cudaStream_t stream);

// frames14[n][14] -> syndrome, errors (-1 = uncorrectable), bits[n][2]
cudaError_t launch_crc_batch(const uint8_t *frames14, uint32_t n, const ErrorInfo *tab_short, int n_short,
                             const ErrorInfo *tab_long, int n_long, uint32_t *syndromes, int8_t *errors,
                             int8_t *bits2, cudaStream_t stream);

cudaError_t upload_constants(const uint32_t *bit_syndromes112);

} // namespace b200
