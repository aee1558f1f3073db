/*
 * TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's
 * IQ -> magnitude -> preamble scan -> PPM slice -> CRC/score -> resolve path.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this;
 * the product (readsb_protobuf_b200/csrc) never links, imports or calls it.
 *
 * PARITY PIN: the reference's own tests hold no golden vector for this path (SURVEY.md
 * section 4 / 8c: the only known-answer frame is the comment 8D4B969699155600E87406F5B69F at
 * net_io.c:1645).  The restatement is therefore pinned against the UNMODIFIED reference
 * compiled here from /root/reference (oracle/Makefile target `ref`, harness
 * oracle/ref_harness.c): tests/test_oracle_vs_reference.py requires byte-identical result
 * files on seeded streams, and tests/golden/ holds reference-generated fixtures (made by
 * tests/golden/make_golden.py) that travel to the GPU box.
 */
#ifndef MODES_ORACLE_H
#define MODES_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { MO_UC8 = 0, MO_SC16 = 1, MO_SC16Q11 = 2 };

#define MO_OVERLAP 326          /* readsb.c:198 at 2.4 MHz */
#define MO_BLOCK_SAMPLES 131072 /* readsb.h:98-99 */

#pragma pack(push, 1)
typedef struct { /* same layout as oracle/ref_harness.c struct result_msg */
    uint64_t timestampMsg;
    uint64_t sysTimestampMsg;
    double signalLevel;
    uint32_t crc;
    uint32_t addr;
    int32_t score;
    uint8_t msgbits;
    uint8_t msgtype;
    uint8_t correctedbits;
    uint8_t reserved;
    uint8_t msg[14];
    uint8_t verbatim[14];
} mo_msg;

typedef struct { /* same layout as struct result_stats */
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
    double convert_cpu_s;
    double demod_cpu_s;
} mo_stats;

typedef struct {
    double mean_level;
    double mean_power;
} mo_block;
#pragma pack(pop)

typedef struct {
    uint32_t syndrome;
    int32_t errors;
    int8_t bit[2];
    uint16_t padding;
} mo_errorinfo; /* crc.h:32-37 */

/* ---- building blocks (each checked against the reference in tests/) ---- */

/* convert.c:35-61: the 65536-entry uc8 magnitude table, indexed by the little-endian u16 (I | Q<<8) */
void mo_uc8_table(uint16_t *table65536);

/* convert.c:63-111 / 215-253 / 332-370 (no DC filter).  Returns 0, or -1 for a bad format */
int mo_convert(int format, const void *iq, uint32_t nsamples, uint16_t *mag,
               double *mean_level, double *mean_power);

/* struct converter_state (convert.c:25-30) and the --dcfilter converters convert_*_generic
 * (convert.c:113-213, 374-423): one call = one mag_buf, the filter state runs on across calls */
typedef struct {
    float dc_a, dc_b, z1_I, z1_Q;
} mo_dc_state;
void mo_dc_init(mo_dc_state *st, double sample_rate); /* init_converter, convert.c:476-488 */
int mo_convert_dc(int format, const void *iq, uint32_t nsamples, mo_dc_state *st, uint16_t *mag,
                  double *mean_level, double *mean_power);

/* convert.c:264-328, only in builds with -DSC16Q11_TABLE_BITS=bits (the armhf package: 8, debian/rules:19):
 * sc16q11 through a 2^(2*bits)-entry magnitude table indexed by the top `bits` bits of |I| & 2047 and |Q| & 2047;
 * integer sums like the uc8 converter.  bits in 1..11. */
void mo_sc16q11_table(int bits, uint16_t *table /* 1 << (2 * bits) entries */);
int mo_convert_sc16q11_table(int bits, const void *iq, uint32_t nsamples, uint16_t *mag,
                             double *mean_level, double *mean_power);

/* crc.c:67-82 */
uint32_t mo_checksum(const uint8_t *msg, int bits);
/* crc.c:42-65: syndrome of a single flipped bit, indexed from the start of a 112-bit frame */
uint32_t mo_single_bit_syndrome(int bit);
/* crc.c:184-383: sorted error table for 56/112 bit frames; returns the entry count (<= cap) */
int mo_error_table(int nfix, int bits, mo_errorinfo *out, int cap);

/* demod_2400.c:276-330: the 5-bit "phases to try" mask of scan position j (bit p-4 set => try phase p) */
int mo_try_mask(const uint16_t *m, uint32_t j, int threshold);
/* demod_2400.c:98-209: slice nbytes message bytes for (j, try_phase) */
void mo_slice(const uint16_t *m, uint32_t j, int try_phase, int nbytes, uint8_t *msg);

/* demod_2400.c:529-530: noise level of a block from the converter's means */
unsigned mo_modeac_noise_level(double mean_level, double mean_power);
/* demod_2400.c:577-683: does a Mode A/C reply with F1 at data index f1 (>= 1) decode? */
int mo_modeac_at(const uint16_t *m, uint32_t f1, unsigned noise_level, uint32_t *f1_clock_out, uint32_t *modeac_out);

/* ---- whole-stream run: ifileRun + fifo overlap + demodulate2400 + backgroundTasks ---- */

typedef struct {
    int32_t format;        /* MO_UC8 ... */
    int32_t nfix;          /* Modes.nfix_crc: 0, 1 or 2 */
    int32_t threshold;     /* Modes.preambleThreshold */
    uint32_t block_samples; /* samples per mag_buf (MO_BLOCK_SAMPLES) */
    int32_t modeac;        /* Modes.mode_ac: also run the Mode A/C demodulator on every block */
    int32_t dcfilter;      /* Modes.dc_filter (--dcfilter): the convert_*_generic converters */
    int32_t sc16q11_table_bits; /* != 0: a reference built with -DSC16Q11_TABLE_BITS=N (sc16q11 without --dcfilter only) */
} mo_config;

typedef struct {
    mo_msg *msgs;
    uint64_t n_msgs;
    mo_block *blocks;
    uint64_t n_blocks;
    mo_stats stats;
    uint64_t n_samples;
} mo_result;

/* net_io.c:769-835 / 870-896: the bytes the Beast and raw output services write for these messages */
size_t mo_format_beast(const mo_msg *msgs, uint64_t n, int net_verbatim, uint8_t *out, size_t cap);
size_t mo_format_raw(const mo_msg *msgs, uint64_t n, int net_verbatim, int mlat, char *out, size_t cap);

/* Runs the whole stream; result arrays are malloc'd, release with mo_result_free. */
int mo_run(const mo_config *cfg, const void *iq, uint64_t nsamples, mo_result *res);
void mo_result_free(mo_result *res);

#ifdef __cplusplus
}
#endif

#endif
