/*
 * readsb_b200.h -- C ABI of the B200-native Mode S demodulator (libreadsb_b200.so).
 *
 * Drop-in scope: the IQ -> magnitude -> preamble scan -> PPM slice -> CRC/score -> resolve
 * path of Mictronics/readsb-protobuf, i.e. what the reference does between
 *     ifile.converter(...)            sdr_ifile.c:214   (convert.h:33-43, convert.c)
 * and demodulate2400(struct mag_buf*) readsb.c:830      (demod_2400.h:37, demod_2400.c:236-428)
 * including the CRC/score/filter calls those make (crc.h:39-43, mode_s.c:311-555,717-726,
 * icao_filter.c:73-164).  Plain pointers and sizes only; no CUDA or torch types.
 *
 * Everything runs on the GPU except the order-dependent resolve step (skip-ahead, ICAO filter,
 * best-phase pick), which walks the few surviving candidates on the host.  There is no CPU
 * fallback: every entry point fails with B200_ERR_CUDA when no sm_100 device is usable.
 *
 * readsb-shaped entry points (init_converter, demodulate2400, ifile*) that bind this ABI to
 * the reference's own structs live in readsb_protobuf_b200/shim/ (see INTEGRATION.md).
 */
#ifndef READSB_B200_H
#define READSB_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200_ABI_VERSION 3

/* input_format_t of the reference (convert.h:25-31), same numeric values */
enum { B200_INPUT_UC8 = 0, B200_INPUT_SC16 = 1, B200_INPUT_SC16Q11 = 2 };

enum {
    B200_OK = 0,
    B200_ERR_ARG = -1,      /* bad argument / configuration */
    B200_ERR_CUDA = -2,     /* no usable device, or a CUDA call failed (see b200_last_error) */
    B200_ERR_NOMEM = -3,
    B200_ERR_CAPACITY = -4, /* span larger than max_span_samples */
    B200_ERR_STATE = -5     /* call order violated (e.g. data after the final span) */
};

#define B200_OVERLAP_SAMPLES 326        /* Modes.trailing_samples at 2.4 MS/s, readsb.c:198 */
#define B200_DEFAULT_BLOCK_SAMPLES 131072 /* MODES_MAG_BUF_SAMPLES, readsb.h:98-99 */

/* process flags */
#define B200_FLAG_FINAL 1u   /* this span ends the stream (EOF in ifileRun, sdr_ifile.c:196-209) */

typedef struct b200_demod b200_demod; /* one receiver stream */

typedef struct b200_demod_config {
    int32_t abi_version;         /* B200_ABI_VERSION */
    int32_t device;              /* CUDA device ordinal */
    int32_t input_format;        /* B200_INPUT_*            (--iformat, sdr_ifile.c:86-99) */
    int32_t nfix_crc;            /* Modes.nfix_crc 0|1|2    (readsb.c:173, 541-543) */
    int32_t preamble_threshold;  /* Modes.preambleThreshold (readsb.c:142-151, 503-505) */
    uint32_t block_samples;      /* samples per mag_buf     (readsb.h:98-99); 0 = default */
    uint64_t startup_time_ms;    /* Modes.startup_time      (readsb.c:739) */
    uint64_t max_span_samples;   /* largest span one process call may carry; 0 = 64 Mi samples */
    int32_t mode_ac;             /* Modes.mode_ac (--modeac, readsb.c:831-833): also demodulate Mode A/C replies */
    int32_t filter_dc;           /* Modes.dc_filter (--dcfilter, readsb.c:486): the convert_*_generic converters
                                    (convert.c:113-213, 374-423) with their 1 Hz DC block at 2.4 MS/s */
    int32_t sc16q11_table_bits;  /* 0, or the SC16Q11_TABLE_BITS (1..11) the reference was built with: sc16q11 input then
                                    goes through convert_sc16q11_table (convert.c:264-328; the armhf package uses 8,
                                    debian/rules:19; oneoff/convert_benchmark.c also times 9..11) instead of the float
                                    path, unless filter_dc picks the generic one */
    int32_t reserved;
} b200_demod_config;

/* One accepted message: the fields of struct modesMessage (readsb.h:340-547) that the path
 * itself sets (demod_2400.c:353-399, mode_s.c:424-562).  Layout is packed, 68 bytes. */
#pragma pack(push, 1)
typedef struct b200_message {
    uint64_t timestampMsg;     /* 12 MHz ticks, demod_2400.c:358 */
    uint64_t sysTimestampMsg;  /* ms, demod_2400.c:361 */
    double signalLevel;        /* demod_2400.c:398-399 */
    uint32_t crc;              /* mode_s.c:440 */
    uint32_t addr;             /* mode_s.c:465,544,561 */
    int32_t score;             /* demod_2400.c:368 */
    uint8_t msgbits;           /* 56 | 112 */
    uint8_t msgtype;           /* DF */
    uint8_t correctedbits;
    uint8_t bestphase;         /* 4..8, demod_2400.c:222 */
    uint8_t msg[14];           /* after CRC repair; bytes past msgbits/8 are zero */
    uint8_t verbatim[14];      /* as sliced, before repair (mode_s.c:427-430) */
} b200_message;

/* The demodulator's share of struct stats (stats.h:57-121).  Layout is packed, 136 bytes. */
typedef struct b200_demod_stats {
    uint32_t demod_preambles;
    uint32_t demod_rejected_bad;
    uint32_t demod_rejected_unknown_icao;
    uint32_t demod_accepted[3];
    uint32_t demod_preamblePhase[5];
    uint32_t demod_bestPhase[5];
    uint32_t strong_signal_count;
    uint32_t messages_total;
    uint64_t samples_processed;
    uint64_t noise_power_count;
    uint64_t signal_power_count;
    double noise_power_sum;
    double signal_power_sum;
    double peak_signal_power;
    double reserved[2];
} b200_demod_stats;

/* Per-mag_buf converter outputs (convert.h:33-38: out_mean_level / out_mean_power) */
typedef struct b200_block_info {
    double mean_level;
    double mean_power;
} b200_block_info;
#pragma pack(pop)

/* Device-side timing of the last process call, CUDA events on the library's stream (ms). */
typedef struct b200_timing {
    float h2d_ms;      /* host -> device copy of the span (0 for device-resident input) */
    float scan_ms;     /* K1a: magnitude + preamble scan (+ candidate list, magnitudes for K1b) */
    float classify_ms; /* K2: address-set test, live records and signal power of survivors; + packing the live lists (and their hidden-dead counts) into stream order */
    float d2h_ms;      /* 0 since ABI 2: the survivors' download (one DMA per chunk of exactly the live data) runs under the next chunk's kernels */
    float resolve_ms;  /* host: order-dependent resolve (wall clock) */
    float total_ms;    /* wall clock of the whole call */
    uint64_t n_candidates;   /* scan positions with a non-empty try mask */
    uint64_t n_phase_records;/* (position, phase) pairs that survived the CRC class test */
    uint64_t n_live;         /* positions handed to the host resolver */
    uint32_t scan_launches;  /* kernels launched by the call */
    uint32_t chunks;         /* pipeline chunks the span was cut into */
    uint64_t d2h_bytes;      /* bytes copied device -> host by the call (counters, lists, survivors) */
    float slice_ms;          /* K1b: PPM slice + CRC class of every (candidate, phase) */
    uint32_t reserved;
} b200_timing;

/* ---- lifetime ---- */

/* replaces: modesInit()'s set-up of the path (readsb.c:195-243: modesChecksumInit, icaoFilterInit,
 * fifo_create) and init_converter() (convert.c:446-491) */
int b200_demod_create(const b200_demod_config *cfg, b200_demod **out);
void b200_demod_destroy(b200_demod *d);
/* forget the stream (filter, overlap carry, counters); configuration is kept */
int b200_demod_reset(b200_demod *d);
const char *b200_last_error(void);

/* ---- the hot path ---- */

/* replaces: the ifileRun loop body + demodulate2400 for every mag_buf in the span
 * (sdr_ifile.c:178-231, fifo.c:180-188, readsb.c:830-836, readsb.c:331).
 * `iq` holds nsamples raw IQ pairs in host memory.  Unless B200_FLAG_FINAL is set, nsamples
 * must be a multiple of block_samples.  Messages and block infos of the span are kept in the
 * context until the next process call. */
int b200_demod_process(b200_demod *d, const void *iq, uint64_t nsamples, uint32_t flags);

/* Page-locked host memory for the sample buffers handed to b200_demod_process (replaces the malloc of ifileOpen's
 * read buffer, sdr_ifile.c:142-146): the H2D copy of a pinned buffer runs at PCIe speed and overlaps the kernels,
 * a pageable one is staged through the driver at a fraction of that.  NULL when no device is usable or the
 * allocation fails (the caller may then fall back to malloc: any host pointer is accepted by process). */
void *b200_host_alloc(size_t bytes);
void b200_host_free(void *p);

/* Same, with the span already resident in device memory (16-byte aligned), launched on
 * `cuda_stream` (a cudaStream_t passed as void*, NULL = the library's own stream). */
int b200_demod_process_device(b200_demod *d, const void *d_iq, uint64_t nsamples, uint32_t flags,
                              void *cuda_stream);

/* results of the last process call (owned by the context) */
uint64_t b200_demod_message_count(const b200_demod *d);
const b200_message *b200_demod_messages(const b200_demod *d);
uint64_t b200_demod_block_count(const b200_demod *d);
const b200_block_info *b200_demod_blocks(const b200_demod *d);
/* running totals since create/reset (Modes.stats_current of the reference) */
int b200_demod_get_stats(const b200_demod *d, b200_demod_stats *out);
int b200_demod_get_timing(const b200_demod *d, b200_timing *out);
/* Modes.stats_current.demod_modeac (demod_2400.c:708): Mode A/C replies since create/reset.  With
 * mode_ac set, a block's replies follow its Mode S messages in the message list as entries with
 * msgtype 32, msgbits 16, msg[0..1] = the Mode A code, addr = (code & 0xFF7F) | 1 << 24
 * (decodeModeAMessage, mode_ac.c:168-181), timestamped at the second framing pulse. */
uint64_t b200_demod_modeac_count(const b200_demod *d);

/* ---- output formats (host; no device needed) ---- */

/* replaces: modesSendBeastOutput (net_io.c:769-835): Beast binary frames (0x1a, type '1' | '2' | '3', 48-bit
 * big-endian 12 MHz timestamp, signal byte round(sqrt(signalLevel) * 255), payload; every 0x1a doubled)
 * of n messages, in order.  net_verbatim = Modes.net_verbatim (send the frame as sliced, not as
 * repaired).  Returns the byte count needed; only the first `cap` bytes are stored (out may be NULL). */
uint64_t b200_format_beast(const b200_message *msgs, uint64_t n, int net_verbatim, uint8_t *out, uint64_t cap);
/* replaces: modesSendRawOutput (net_io.c:870-896): "*<hex>;\n", or "@<12 hex digits of the timestamp><hex>;\n"
 * with mlat (Modes.mlat) and a non-zero timestamp.  Same return convention. */
uint64_t b200_format_raw(const b200_message *msgs, uint64_t n, int net_verbatim, int mlat, char *out, uint64_t cap);

/* ---- kernel-level entry points (measurement and unit parity; device pointers) ---- */

/* Only K1 over a device-resident span, no host work and no result download: the kernels the
 * roofline is quoted on.  mode 0 = magnitude + preamble scan only (candidates counted), 1 = K1a as
 * the pipeline runs it (+ candidate list + u16 magnitudes for K1b), 2 = K1a followed by K1b
 * (slice + CRC class of every candidate phase).
 * Returns the kernel's duration in *ms_out (CUDA events on `cuda_stream`). */
int b200_scan_device(b200_demod *d, const void *d_iq, uint64_t nsamples, int mode,
                     void *cuda_stream, float *ms_out, uint64_t *n_candidates_out);

/* replaces: iq_convert_fn (convert.h:33-38; convert.c:63-111, 215-253, 332-370), materialising
 * the u16 magnitudes: host in, host out.  mean_level / mean_power may be NULL. */
int b200_convert(b200_demod *d, const void *iq, uint32_t nsamples, uint16_t *mag,
                 double *mean_level, double *mean_power);

/* replaces: init_uc8_lookup (convert.c:35-61); copies the 65536-entry table the kernels use */
int b200_uc8_table(b200_demod *d, uint16_t *table65536);

/* Debug/parity taps of K1 over a host span (stream start, zero overlap):
 *   try_masks[nsamples]  5-bit "phases to try" per scan position (demod_2400.c:276-330)
 * Either pointer may be NULL. */
#pragma pack(push, 1)
typedef struct b200_phase_record {
    uint32_t position;   /* scan position j within the span (overlap coordinates) */
    uint32_t crc;        /* modesChecksum of the sliced frame */
    uint32_t key;        /* address the score depends on (CRC-derived or corrected AA) */
    uint8_t phase;       /* 4..8 */
    uint8_t kind;        /* B200_KIND_* */
    uint8_t errors;      /* bits the syndrome table would repair */
    uint8_t reserved;
} b200_phase_record;
#pragma pack(pop)
enum { B200_KIND_AP = 1, B200_KIND_AP_COMMB = 2, B200_KIND_DF11 = 3, B200_KIND_ES = 4 };

int b200_debug_scan(b200_demod *d, const void *iq, uint64_t nsamples, uint8_t *try_masks,
                    b200_phase_record *records, uint64_t record_cap, uint64_t *n_records);

/* replaces: modesChecksum (crc.c:67-82) + modesChecksumDiagnose (crc.c:389-412) on the device,
 * for n frames of 14 bytes each (short frames use the first 7): syndromes[n], and per frame
 * errors (-1 = not correctable) and up to two bit positions. */
int b200_crc_batch(b200_demod *d, const uint8_t *frames14, uint32_t n, uint32_t *syndromes,
                   int8_t *errors, int8_t *bits2);

/* host-side copy of the syndrome table the device uses (crc.c:184-383); returns entry count */
#pragma pack(push, 1)
typedef struct b200_errorinfo {
    uint32_t syndrome;
    int32_t errors;
    int8_t bit[2];
    uint16_t padding;
} b200_errorinfo; /* struct errorinfo, crc.h:32-37 */
#pragma pack(pop)
int b200_error_table(b200_demod *d, int bits, b200_errorinfo *out, int cap);

/* ---- host-only helpers (no device needed): the tables and the filter the resolver uses ---- */

/* sizeof of the packed ABI structs: 0 b200_message, 1 b200_demod_stats, 2 b200_block_info,
 * 3 b200_timing, 4 b200_phase_record, 5 b200_errorinfo, 6 b200_demod_config */
int b200_abi_sizeof(int which);
/* modesChecksum (crc.c:67-82) on the host */
uint32_t b200_host_checksum(const uint8_t *msg, int bits);
/* prepareErrorTable (crc.c:184-354) on the host for nfix in 0..2; returns the entry count */
int b200_host_error_table(int nfix, int bits, b200_errorinfo *out, int cap);
/* init_uc8_lookup (convert.c:35-61) on the host */
void b200_host_uc8_table(uint16_t *table65536);
/* icao_filter.c:73-164 driven step by step: ops[i] = 0 add, 1 test, 2 expire(arg = now ms);
 * results[i] = test outcome (0/1), else 0 */
int b200_host_filter_script(const uint8_t *ops, const uint64_t *args, uint32_t n, uint8_t *results);
/* The order-dependent tail of demodulate2400 (demod_2400.c:236-428: best-phase pick, decode-time rejects, ICAO
 * filter, skip-ahead, statistics) run on the host alone over recorded kernel outputs: `paths` = the files a
 * process call wrote with B200_DUMP_SPAN=<dir> set (one per pipeline chunk), in stream order.  Lets the host
 * logic be tested against the oracle without a GPU.  Returns B200_OK, or B200_ERR_ARG / B200_ERR_CAPACITY. */
int b200_host_resolve_dumps(const char *const *paths, uint32_t npaths, int nfix_crc, b200_message *msgs, uint64_t msg_cap,
                            uint64_t *n_msgs, b200_block_info *blocks, uint64_t block_cap, uint64_t *n_blocks,
                            b200_demod_stats *stats);

#ifdef __cplusplus
}
#endif

#endif
