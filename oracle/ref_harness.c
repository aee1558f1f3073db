/*
 * TEST INFRASTRUCTURE ONLY -- not part of the product, never linked into it.
 *
 * ref_demod: drives the UNMODIFIED reference objects (convert.o, demod_2400.o,
 * crc.o, mode_s.o, icao_filter.o, track.o ... compiled by oracle/Makefile from
 * the sources where they lie under /root/reference) over an IQ file, single
 * threaded, and writes every decoded modesMessage plus the demod statistics to
 * a small binary result file (format: oracle/result_format.md, reader:
 * readsb_protobuf_b200/results.py).
 *
 * Why a harness and not the stock binary: the reference's fifo_enqueue() never
 * advances fifo_tail (fifo.c:192-197), so an un-throttled `readsb --ifile`
 * silently drops queued buffers.  This harness restates only the plumbing around
 * the path, and calls the reference for the path itself:
 *   - block loop, timestamps:   sdr_ifile.c:164-231 (ifileRun)
 *   - overlap carry:            fifo.c:180-188      (fifo_enqueue)
 *   - per-block housekeeping:   readsb.c:830-836, readsb.c:331-332
 *   - init:                     readsb.c:138-243 (the fields the path reads)
 * One process = one stream (icaoFilterExpire keeps a function-static next_flip,
 * icao_filter.c:151, which cannot be reset in-process).
 *
 * usage: ref_demod --in FILE --out FILE [--format uc8|sc16|sc16q11] [--nfix 0|1|2]
 *                  [--threshold N] [--block N] [--dcfilter] [--mag-out FILE]
 *                  [--max-samples N] [--repeat R] [--modeac]
 */
#include "readsb.h"

#include <fcntl.h>
#include <unistd.h>

struct _Modes Modes;

extern void (*oracle_capture_hook)(struct modesMessage *mm);

/* ---- result file records (keep in sync with readsb_protobuf_b200/results.py) ---- */

#pragma pack(push, 1)
struct result_header {
    char magic[4]; /* "MDSR" */
    uint32_t version; /* 1 */
    uint64_t n_msgs;
    uint64_t n_blocks;
    uint64_t n_samples;
};

struct result_stats {
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
};

struct result_msg {
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
};

struct result_block {
    double mean_level;
    double mean_power;
};
#pragma pack(pop)

static struct result_msg *msgs;
static size_t n_msgs, cap_msgs;

static void capture(struct modesMessage *mm) {
    if (n_msgs == cap_msgs) {
        cap_msgs = cap_msgs ? cap_msgs * 2 : 4096;
        msgs = realloc(msgs, cap_msgs * sizeof (*msgs));
        if (!msgs) {
            fprintf(stderr, "ref_demod: out of memory\n");
            exit(2);
        }
    }
    struct result_msg *r = &msgs[n_msgs++];
    memset(r, 0, sizeof (*r));
    r->timestampMsg = mm->timestampMsg;
    r->sysTimestampMsg = mm->sysTimestampMsg;
    r->signalLevel = mm->signalLevel;
    r->crc = mm->crc;
    r->addr = mm->addr;
    r->score = mm->score;
    r->msgbits = (uint8_t) mm->msgbits;
    r->msgtype = (uint8_t) mm->msgtype;
    r->correctedbits = (uint8_t) mm->correctedbits;
    /* bytes past msgbits/8 are stale stack data in the reference (demod_2400.c:239) */
    memcpy(r->msg, mm->msg, mm->msgbits / 8);
    memcpy(r->verbatim, mm->verbatim, mm->msgbits / 8);
}

static double thread_cpu_s(void) {
    struct timespec ts;
    clock_gettime(CLOCK_THREAD_CPUTIME_ID, &ts);
    return ts.tv_sec + ts.tv_nsec * 1e-9;
}

int main(int argc, char **argv) {
    const char *in_path = NULL, *out_path = NULL, *mag_path = NULL;
    input_format_t format = INPUT_UC8;
    int nfix = 1, threshold = 58, dcfilter = 0, repeat = 1, modeac = 0;
    unsigned block = MODES_MAG_BUF_SAMPLES;
    uint64_t max_samples = UINT64_MAX;

    for (int i = 1; i < argc; ++i) {
        const char *a = argv[i];
        const char *v = (i + 1 < argc) ? argv[i + 1] : NULL;
        if (!strcmp(a, "--in") && v) { in_path = v; ++i; }
        else if (!strcmp(a, "--out") && v) { out_path = v; ++i; }
        else if (!strcmp(a, "--mag-out") && v) { mag_path = v; ++i; }
        else if (!strcmp(a, "--nfix") && v) { nfix = atoi(v); ++i; }
        else if (!strcmp(a, "--threshold") && v) { threshold = atoi(v); ++i; }
        else if (!strcmp(a, "--block") && v) { block = (unsigned) atol(v); ++i; }
        else if (!strcmp(a, "--max-samples") && v) { max_samples = strtoull(v, NULL, 10); ++i; }
        else if (!strcmp(a, "--repeat") && v) { repeat = atoi(v); ++i; }
        else if (!strcmp(a, "--dcfilter")) { dcfilter = 1; }
        else if (!strcmp(a, "--modeac")) { modeac = 1; }
        else if (!strcmp(a, "--format") && v) {
            if (!strcasecmp(v, "uc8")) format = INPUT_UC8;
            else if (!strcasecmp(v, "sc16")) format = INPUT_SC16;
            else if (!strcasecmp(v, "sc16q11")) format = INPUT_SC16Q11;
            else { fprintf(stderr, "ref_demod: bad format %s\n", v); return 2; }
            ++i;
        } else {
            fprintf(stderr, "ref_demod: bad argument %s\n", a);
            return 2;
        }
    }
    if (!in_path || !out_path || block == 0) {
        fprintf(stderr, "usage: ref_demod --in FILE --out FILE [options]\n");
        return 2;
    }

    /* the fields of Modes the path reads; values as modesInitConfig/modesInit set them
     * (readsb.c:138-243) with the fixed flags of BASELINE.md section 3 */
    memset(&Modes, 0, sizeof (Modes));
    Modes.sample_rate = 2400000.0; /* readsb.c:195 */
    Modes.trailing_samples = (unsigned) ((MODES_PREAMBLE_US + MODES_LONG_MSG_BITS + 16) * 1e-6 * Modes.sample_rate); /* readsb.c:198 */
    Modes.preambleThreshold = (uint32_t) threshold;
    Modes.nfix_crc = (int8_t) nfix;
    Modes.sdr_type = SDR_IFILE;
    Modes.dc_filter = (int8_t) dcfilter;
    Modes.mode_ac = (int8_t) modeac;
    Modes.quiet = 1;
    Modes.net = 1;
    Modes.net_verbatim = 1;
    Modes.maxRange = 1852 * 300.0;
    Modes.startup_time = 0;
    Modes.check_crc = 1;
    Modes.filter_persistence = 8;

    modesChecksumInit(Modes.nfix_crc); /* readsb.c:241 */
    icaoFilterInit(); /* readsb.c:242 */
    modeACInit(); /* readsb.c:243 */
    oracle_capture_hook = capture;

    struct converter_state *cstate = NULL;
    iq_convert_fn converter = init_converter(format, Modes.sample_rate, Modes.dc_filter, &cstate);
    if (!converter) {
        fprintf(stderr, "ref_demod: init_converter failed\n");
        return 2;
    }

    const unsigned bytes_per_sample = (format == INPUT_UC8) ? 2 : 4; /* sdr_ifile.c:127-139 */
    const unsigned overlap = Modes.trailing_samples;

    struct mag_buf buf;
    memset(&buf, 0, sizeof (buf));
    buf.totalLength = block + overlap;
    buf.data = calloc(buf.totalLength, sizeof (uint16_t)); /* fifo.c:57 */
    buf.overlap = overlap;
    uint16_t *overlap_buffer = calloc(overlap, sizeof (uint16_t)); /* fifo.c:47 */
    char *readbuf = malloc((size_t) block * bytes_per_sample);
    if (!buf.data || !overlap_buffer || !readbuf) {
        fprintf(stderr, "ref_demod: out of memory\n");
        return 2;
    }

    int fd = open(in_path, O_RDONLY);
    if (fd < 0) {
        perror(in_path);
        return 2;
    }
    FILE *magf = NULL;
    if (mag_path && !(magf = fopen(mag_path, "wb"))) {
        perror(mag_path);
        return 2;
    }

    struct result_block *blocks = NULL;
    size_t n_blocks = 0, cap_blocks = 0;
    double convert_cpu = 0, demod_cpu = 0;
    uint64_t sampleCounter = 0;

    /* --repeat R replays the file R times as one continuous stream (CPU-baseline timing on
     * a bounded sample without materialising a huge file) */
    for (int rep = 0; rep < repeat; ++rep) {
        if (lseek(fd, 0, SEEK_SET) < 0) {
            perror("lseek");
            return 2;
        }
        int eof = 0;
        while (!eof) {
            /* sdr_ifile.c:187-190 */
            buf.sampleTimestamp = sampleCounter * 12e6 / Modes.sample_rate;
            buf.sysTimestamp = buf.sampleTimestamp / 12000U + Modes.startup_time;

            /* sdr_ifile.c:192-211 */
            uint64_t want_samples = block;
            if (max_samples - sampleCounter < want_samples)
                want_samples = max_samples - sampleCounter;
            size_t bytes_wanted = (size_t) want_samples * bytes_per_sample;
            size_t bytes_read = 0;
            while (bytes_read < bytes_wanted) {
                ssize_t n = read(fd, readbuf + bytes_read, bytes_wanted - bytes_read);
                if (n <= 0) {
                    eof = 1;
                    break;
                }
                bytes_read += (size_t) n;
            }
            if (want_samples < block)
                eof = 1;
            unsigned samples_read = (unsigned) (bytes_read / bytes_per_sample);
            if (eof && rep + 1 < repeat && samples_read == 0)
                break; /* seamless wrap between repeats */

            /* sdr_ifile.c:214-216 */
            double t0 = thread_cpu_s();
            converter(readbuf, &buf.data[overlap], samples_read, cstate, &buf.mean_level, &buf.mean_power);
            convert_cpu += thread_cpu_s() - t0;
            buf.validLength = overlap + samples_read;
            buf.flags = 0;

            /* fifo.c:180-188 */
            memcpy(buf.data, overlap_buffer, overlap * sizeof (uint16_t));
            memcpy(overlap_buffer, &buf.data[buf.validLength - overlap], overlap * sizeof (uint16_t));

            if (magf)
                fwrite(&buf.data[overlap], sizeof (uint16_t), samples_read, magf);

            /* readsb.c:828-837 */
            t0 = thread_cpu_s();
            demodulate2400(&buf);
            if (Modes.mode_ac) /* readsb.c:831-833 */
                demodulate2400AC(&buf);
            demod_cpu += thread_cpu_s() - t0;
            Modes.stats_current.samples_processed += buf.validLength;

            /* readsb.c:331-332 (backgroundTasks) */
            icaoFilterExpire();
            trackPeriodicUpdate();

            if (n_blocks == cap_blocks) {
                cap_blocks = cap_blocks ? cap_blocks * 2 : 1024;
                blocks = realloc(blocks, cap_blocks * sizeof (*blocks));
            }
            blocks[n_blocks].mean_level = buf.mean_level;
            blocks[n_blocks].mean_power = buf.mean_power;
            ++n_blocks;

            sampleCounter += samples_read; /* sdr_ifile.c:230 */
        }
    }
    close(fd);
    if (magf)
        fclose(magf);

    struct result_header hdr;
    memcpy(hdr.magic, "MDSR", 4);
    hdr.version = 1;
    hdr.n_msgs = n_msgs;
    hdr.n_blocks = n_blocks;
    hdr.n_samples = sampleCounter;

    struct stats *st = &Modes.stats_current;
    struct result_stats rs;
    memset(&rs, 0, sizeof (rs));
    rs.demod_preambles = st->demod_preambles;
    rs.demod_rejected_bad = st->demod_rejected_bad;
    rs.demod_rejected_unknown_icao = st->demod_rejected_unknown_icao;
    for (int i = 0; i < 3; ++i)
        rs.demod_accepted[i] = st->demod_accepted[i];
    for (int i = 0; i < 5; ++i) {
        rs.demod_preamblePhase[i] = st->demod_preamblePhase[i];
        rs.demod_bestPhase[i] = st->demod_bestPhase[i];
    }
    rs.strong_signal_count = st->strong_signal_count;
    rs.messages_total = st->messages_total;
    rs.samples_processed = st->samples_processed;
    rs.noise_power_count = st->noise_power_count;
    rs.signal_power_count = st->signal_power_count;
    rs.noise_power_sum = st->noise_power_sum;
    rs.signal_power_sum = st->signal_power_sum;
    rs.peak_signal_power = st->peak_signal_power;
    rs.convert_cpu_s = convert_cpu;
    rs.demod_cpu_s = demod_cpu;

    FILE *out = fopen(out_path, "wb");
    if (!out) {
        perror(out_path);
        return 2;
    }
    fwrite(&hdr, sizeof (hdr), 1, out);
    fwrite(&rs, sizeof (rs), 1, out);
    fwrite(msgs, sizeof (*msgs), n_msgs, out);
    fwrite(blocks, sizeof (*blocks), n_blocks, out);
    fclose(out);

    fprintf(stderr, "ref_demod: %llu samples, %zu blocks, %zu messages, convert %.3f s, demod %.3f s CPU\n",
            (unsigned long long) sampleCounter, n_blocks, n_msgs, convert_cpu, demod_cpu);
    return 0;
}
