/*
 * readsb_b200_shim.c -- readsb's own symbols on top of the C ABI of include/readsb_b200.h.
 *
 * Link this file and libreadsb_b200.so INSTEAD OF readsb's convert.o, demod_2400.o and
 * sdr_ifile.o (Makefile:82-83 of the reference); keep crc.o, mode_s.o, icao_filter.o, fifo.o,
 * readsb.o.  It is compiled against readsb's own headers, so it lives outside the library.
 *
 * What replaces what:
 *   ifileOpen / ifileRun / ifileClose   sdr_ifile.c:115-255: the file is read in spans of many
 *       mag_bufs, each span goes to the GPU in one b200_demod_process() call; readsb's main loop
 *       still receives one mag_buf per 131072 samples through the FIFO (fifo.h), carrying the
 *       block's timestamps and converter means, and calls demodulate2400() for it
 *   demodulate2400                      demod_2400.c:236-428: hands the block's already resolved
 *       frames to readsb's decodeModesMessage() / useModesMessage() and adds the block's counters
 *       to Modes.stats_current
 *   init_converter / cleanup_converter  convert.c:446-499: iq_convert_fn over b200_convert() for callers that
 *       want the magnitudes of a block themselves (oneoff/convert_benchmark.c style)
 *
 * Scope: --device-type ifile only (the north star's boundary).  demodulate2400() here replays what ifileRun
 * resolved; with any other SDR front-end nothing would feed it, so init_converter() -- the first thing every
 * live front-end calls (sdr_rtlsdr.c:321-325, sdr_bladerf.c, sdr_plutosdr.c) -- refuses them loudly.
 *
 * The reference's fifo_enqueue() loses buffers when more than one is queued (it never advances
 * fifo_tail, fifo.c:192-197), so the reader hands blocks over one at a time and waits until
 * demodulate2400() has consumed the block before queueing the next.
 */
#include "readsb.h"

#include <fcntl.h>
#include <unistd.h>

#include "readsb_b200.h"

#define SPAN_BLOCKS 256 /* mag_bufs per GPU call: 14 s of samples */
#define FIRST_SPAN_BLOCKS 1 /* the first call carries one mag_buf: the main loop gets a block before its 100 ms FIFO
                               time-out runs backgroundTasks() -> icaoFilterExpire() on an empty stream (readsb.c:797-835) */

static struct {
    const char *filename;
    input_format_t input_format;
    bool throttle;
    int fd;
    unsigned bytes_per_sample;
    b200_demod *demod;
    char *span;
    bool span_pinned; /* span came from b200_host_alloc */
} ifile;

/* one mag_buf's share of a resolved span, queued from the reader to the demodulator thread */
struct block_result {
    const b200_message *msgs;
    uint64_t nmsgs;
    b200_demod_stats delta; /* added to Modes.stats_current with the block */
    bool has_delta;
};
static struct block_result pending;
static pthread_mutex_t pending_mutex = PTHREAD_MUTEX_INITIALIZER;
static pthread_cond_t pending_cond = PTHREAD_COND_INITIALIZER;
static bool pending_ready; /* set by the reader, cleared by demodulate2400 */

void ifileInitConfig(void) {
    memset(&ifile, 0, sizeof (ifile));
    ifile.input_format = INPUT_UC8;
    ifile.fd = -1;
}

bool ifileHandleOption(int argc, char *argv) {
    switch (argc) {
        case OptIfileName:
            ifile.filename = strdup(argv);
            Modes.sdr_type = SDR_IFILE;
            break;
        case OptIfileFormat:
            if (!strcasecmp(argv, "uc8")) ifile.input_format = INPUT_UC8;
            else if (!strcasecmp(argv, "sc16")) ifile.input_format = INPUT_SC16;
            else if (!strcasecmp(argv, "sc16q11")) ifile.input_format = INPUT_SC16Q11;
            else {
                fprintf(stderr, "Input format '%s' not understood (supported values: UC8, SC16, SC16Q11)\n", argv);
                return false;
            }
            break;
        case OptIfileThrottle:
            ifile.throttle = true;
            break;
    }
    return true;
}

bool ifileOpen(void) {
    if (!ifile.filename) {
        fprintf(stderr, "SDR type 'ifile' requires an --ifile argument\n");
        return false;
    }
    if (!strcmp(ifile.filename, "-")) {
        ifile.fd = STDIN_FILENO;
    } else if ((ifile.fd = open(ifile.filename, O_RDONLY)) < 0) {
        fprintf(stderr, "ifile: could not open %s: %s\n", ifile.filename, strerror(errno));
        return false;
    }
    ifile.bytes_per_sample = (ifile.input_format == INPUT_UC8) ? 2 : 4;

    b200_demod_config cfg;
    memset(&cfg, 0, sizeof (cfg));
    cfg.abi_version = B200_ABI_VERSION;
    cfg.device = 0;
    cfg.input_format = (int32_t) ifile.input_format; /* INPUT_UC8/SC16/SC16Q11 == B200_INPUT_* */
    cfg.nfix_crc = Modes.nfix_crc;
    cfg.preamble_threshold = (int32_t) Modes.preambleThreshold;
    cfg.block_samples = MODES_MAG_BUF_SAMPLES;
    cfg.startup_time_ms = Modes.startup_time;
    cfg.max_span_samples = (uint64_t) SPAN_BLOCKS * MODES_MAG_BUF_SAMPLES;
    cfg.mode_ac = Modes.mode_ac ? 1 : 0; /* --modeac: the library also runs demodulate2400AC's search */
    cfg.filter_dc = Modes.dc_filter ? 1 : 0; /* --dcfilter: init_converter(..., Modes.dc_filter, ...), sdr_ifile.c:151 */
#ifdef SC16Q11_TABLE_BITS
    cfg.sc16q11_table_bits = SC16Q11_TABLE_BITS; /* a table build (debian/rules:19): convert_sc16q11_table, convert.c:264-328 */
#endif
    if (b200_demod_create(&cfg, &ifile.demod) != B200_OK) {
        /* no CPU fallback: fail loudly, like a converter that cannot be initialised (sdr_ifile.c:155-159) */
        fprintf(stderr, "ifile: can't initialize the GPU demodulator: %s\n", b200_last_error());
        return false;
    }
    /* page-locked if possible: the span's H2D copy then runs at PCIe speed behind the kernels */
    const size_t span_bytes = (size_t) SPAN_BLOCKS * MODES_MAG_BUF_SAMPLES * ifile.bytes_per_sample;
    ifile.span = b200_host_alloc(span_bytes);
    ifile.span_pinned = ifile.span != NULL;
    if (!ifile.span)
        ifile.span = malloc(span_bytes);
    return ifile.span != NULL;
}

static size_t read_fully(int fd, char *dst, size_t want, bool *eof) {
    size_t got = 0;
    while (got < want) {
        ssize_t n = read(fd, dst + got, want - got);
        if (n <= 0) {
            *eof = true;
            break;
        }
        got += (size_t) n;
    }
    return got;
}

void ifileRun(void) {
    if (ifile.fd < 0)
        return;
    bool eof = false;
    uint64_t sampleCounter = 0;
    b200_demod_stats before, after;
    memset(&before, 0, sizeof (before));

    bool failed = false, first = true;
    while (!Modes.exit && !eof) {
        const size_t span_blocks = first ? FIRST_SPAN_BLOCKS : SPAN_BLOCKS;
        first = false;
        size_t bytes = read_fully(ifile.fd, ifile.span, span_blocks * MODES_MAG_BUF_SAMPLES * ifile.bytes_per_sample, &eof);
        uint64_t nsamples = bytes / ifile.bytes_per_sample;
        if (b200_demod_process(ifile.demod, ifile.span, nsamples, eof ? B200_FLAG_FINAL : 0) != B200_OK) {
            fprintf(stderr, "ifile: GPU demodulation failed: %s\n", b200_last_error());
            failed = true;
            break;
        }
        const b200_message *msgs = b200_demod_messages(ifile.demod);
        const uint64_t nmsgs = b200_demod_message_count(ifile.demod);
        const b200_block_info *blocks = b200_demod_blocks(ifile.demod);
        const uint64_t nblocks = b200_demod_block_count(ifile.demod);
        b200_demod_get_stats(ifile.demod, &after);

        uint64_t mi = 0;
        for (uint64_t k = 0; k < nblocks && !Modes.exit; ++k) {
            struct mag_buf *outbuf = NULL;
            while (!outbuf && !Modes.exit)
                outbuf = fifo_acquire(100 /* milliseconds */);
            if (!outbuf)
                break;
            uint64_t left = nsamples - k * (uint64_t) MODES_MAG_BUF_SAMPLES;
            unsigned n_k = left < MODES_MAG_BUF_SAMPLES ? (unsigned) left : MODES_MAG_BUF_SAMPLES;
            /* sdr_ifile.c:187-190, 215 */
            outbuf->sampleTimestamp = sampleCounter * 12e6 / Modes.sample_rate;
            outbuf->sysTimestamp = outbuf->sampleTimestamp / 12000U + Modes.startup_time;
            outbuf->validLength = outbuf->overlap + n_k;
            outbuf->flags = 0;
            outbuf->mean_level = blocks[k].mean_level;
            outbuf->mean_power = blocks[k].mean_power;

            /* the block's frames: timestampMsg = sampleTimestamp + 5 j + 768 + phase, j < n_k (demod_2400.c:358) */
            const uint64_t t_end = outbuf->sampleTimestamp + (uint64_t) n_k * 5 + 768;
            uint64_t m0 = mi;
            while (mi < nmsgs && msgs[mi].timestampMsg - msgs[mi].bestphase < t_end)
                ++mi;
            pending.msgs = msgs + m0;
            pending.nmsgs = mi - m0;
            pending.has_delta = (k + 1 == nblocks); /* the span's counters ride on its last block */
            if (pending.has_delta) {
                pending.delta = after;
                pending.delta.demod_preambles -= before.demod_preambles;
                pending.delta.demod_rejected_bad -= before.demod_rejected_bad;
                pending.delta.demod_rejected_unknown_icao -= before.demod_rejected_unknown_icao;
                for (int i = 0; i < 3; ++i) pending.delta.demod_accepted[i] -= before.demod_accepted[i];
                for (int i = 0; i < 5; ++i) {
                    pending.delta.demod_preamblePhase[i] -= before.demod_preamblePhase[i];
                    pending.delta.demod_bestPhase[i] -= before.demod_bestPhase[i];
                }
                pending.delta.strong_signal_count -= before.strong_signal_count;
                pending.delta.noise_power_sum -= before.noise_power_sum;
                pending.delta.noise_power_count -= before.noise_power_count;
                pending.delta.signal_power_sum -= before.signal_power_sum;
                pending.delta.signal_power_count -= before.signal_power_count;
            }
            pthread_mutex_lock(&pending_mutex);
            pending_ready = true;
            pthread_mutex_unlock(&pending_mutex);
            fifo_enqueue(outbuf);
            /* one block in flight: wait until demodulate2400() is done with it */
            pthread_mutex_lock(&pending_mutex);
            while (pending_ready && !Modes.exit) {
                struct timespec deadline;
                clock_gettime(CLOCK_REALTIME, &deadline);
                deadline.tv_nsec += 100 * 1000 * 1000;
                if (deadline.tv_nsec >= 1000000000) {
                    deadline.tv_sec += 1;
                    deadline.tv_nsec -= 1000000000;
                }
                pthread_cond_timedwait(&pending_cond, &pending_mutex, &deadline);
            }
            pthread_mutex_unlock(&pending_mutex);
            sampleCounter += n_k;
        }
        before = after;
    }
    fifo_drain();
    /* sdr_ifile.c:234-236; a stream cut short by a GPU failure is not a normal exit (readsb.c:867: "Abnormal exit") */
    Modes.exit = failed ? 2 : 1;
}

void ifileClose(void) {
    if (ifile.demod) {
        b200_demod_destroy(ifile.demod);
        ifile.demod = NULL;
    }
    if (ifile.span_pinned)
        b200_host_free(ifile.span);
    else
        free(ifile.span);
    ifile.span = NULL;
    if (ifile.fd >= 0 && ifile.fd != STDIN_FILENO) {
        close(ifile.fd);
        ifile.fd = -1;
    }
}

static void release_pending(void) {
    pending.nmsgs = 0;
    pthread_mutex_lock(&pending_mutex);
    pending_ready = false;
    pthread_cond_signal(&pending_cond);
    pthread_mutex_unlock(&pending_mutex);
}

/* demod_2400.h:37 -- the block's frames were resolved on the GPU path; this is the hand-over to
 * readsb's own field decoder, tracker and outputs */
void demodulate2400(struct mag_buf *mag) {
    static struct modesMessage zeroMessage;
    if (Modes.sdr_type == SDR_IFILE)
        Modes.ifile_now = mag->sysTimestamp; /* demod_2400.c:253-255 */

    for (uint64_t i = 0; i < pending.nmsgs; ++i) {
        const b200_message *m = &pending.msgs[i];
        if (m->msgtype == 32)
            continue; /* a Mode A/C reply: demodulate2400AC hands it over */
        struct modesMessage mm = zeroMessage;
        mm.timestampMsg = m->timestampMsg;
        mm.sysTimestampMsg = m->sysTimestampMsg;
        if (Modes.sdr_type == SDR_IFILE)
            Modes.ifile_now = mm.sysTimestampMsg; /* demod_2400.c:364-366 */
        mm.score = m->score;
        unsigned char raw[MODES_LONG_MSG_BYTES];
        memcpy(raw, m->verbatim, MODES_LONG_MSG_BYTES);
        /* The library's resolver decided with its own copy of the ICAO filter, fed in stream order and expired once
         * per mag_buf; that decision is authoritative.  readsb's copy (icao_filter.c) sees the same adds through
         * decodeModesMessage, but the main loop may expire it at a different block (backgroundTasks() also runs on
         * FIFO time-outs), so it can lag by one table generation: if it does not know the address yet, tell it. */
        int rc = decodeModesMessage(&mm, raw);
        if (rc < 0) {
            icaoFilterAdd(m->addr & 0xffffffu);
            memcpy(raw, m->verbatim, MODES_LONG_MSG_BYTES);
            mm = zeroMessage;
            mm.timestampMsg = m->timestampMsg;
            mm.sysTimestampMsg = m->sysTimestampMsg;
            mm.score = m->score;
            rc = decodeModesMessage(&mm, raw);
        }
        if (rc < 0) {
            fprintf(stderr, "readsb_b200_shim: decodeModesMessage disagrees with the GPU path at %012llx\n",
                    (unsigned long long) m->timestampMsg);
            continue;
        }
        mm.signalLevel = m->signalLevel;
        useModesMessage(&mm);
    }
    if (pending.has_delta) {
        struct stats *st = &Modes.stats_current;
        const b200_demod_stats *d = &pending.delta;
        st->demod_preambles += d->demod_preambles;
        st->demod_rejected_bad += d->demod_rejected_bad;
        st->demod_rejected_unknown_icao += d->demod_rejected_unknown_icao;
        for (int i = 0; i < 3; ++i) st->demod_accepted[i] += d->demod_accepted[i];
        for (int i = 0; i < 5; ++i) {
            st->demod_preamblePhase[i] += d->demod_preamblePhase[i];
            st->demod_bestPhase[i] += d->demod_bestPhase[i];
        }
        st->strong_signal_count += d->strong_signal_count;
        st->noise_power_sum += d->noise_power_sum;
        st->noise_power_count += d->noise_power_count;
        st->signal_power_sum += d->signal_power_sum;
        st->signal_power_count += d->signal_power_count;
        /* peak_signal_power is a running maximum */
        if (d->peak_signal_power > st->peak_signal_power)
            st->peak_signal_power = d->peak_signal_power;
    }
    pending.has_delta = false;
    if (!Modes.mode_ac)
        release_pending(); /* otherwise demodulate2400AC follows on the same mag_buf (readsb.c:831-833) */
}

/* demod_2400.h:38 -- the block's Mode A/C replies (library message entries with msgtype 32), handed to
 * readsb's own decodeModeAMessage / useModesMessage exactly as demod_2400.c:686-708 does */
void demodulate2400AC(struct mag_buf *mag) {
    MODES_NOTUSED(mag);
    struct modesMessage mm;
    memset(&mm, 0, sizeof (mm)); /* once per block, like the reference (demod_2400.c:527) */
    for (uint64_t i = 0; i < pending.nmsgs; ++i) {
        const b200_message *m = &pending.msgs[i];
        if (m->msgtype != 32)
            continue;
        mm.timestampMsg = m->timestampMsg;
        mm.sysTimestampMsg = m->sysTimestampMsg;
        decodeModeAMessage(&mm, (m->msg[0] << 8) | m->msg[1]);
        useModesMessage(&mm);
        Modes.stats_current.demod_modeac++;
    }
    release_pending();
}

/* ---- convert.h: a converter for callers that want the magnitudes themselves ---- */

struct converter_state {
    b200_demod *demod;
};

static void convert_via_gpu(void *iq_data, uint16_t *mag_data, unsigned nsamples, struct converter_state *state,
        double *out_mean_level, double *out_mean_power) {
    if (b200_convert(state->demod, iq_data, nsamples, mag_data, out_mean_level, out_mean_power) != B200_OK) {
        /* iq_convert_fn cannot fail (convert.h:33-38): say so and hand back silence rather than stale memory */
        fprintf(stderr, "convert: %s\n", b200_last_error());
        memset(mag_data, 0, (size_t) nsamples * sizeof (uint16_t));
        if (out_mean_level) *out_mean_level = 0;
        if (out_mean_power) *out_mean_power = 0;
    }
}

iq_convert_fn init_converter(input_format_t format, double sample_rate, int filter_dc, struct converter_state **out_state) {
    if (Modes.sdr_type != SDR_IFILE && Modes.sdr_type != SDR_NONE) {
        /* a live front-end: its mag_bufs would reach demodulate2400() above, which only replays what ifileRun resolved */
        fprintf(stderr, "readsb_b200_shim: this build demodulates --device-type ifile only; "
                "SDR type %d would decode nothing (feed live blocks through b200_demod_process, INTEGRATION.md)\n", (int) Modes.sdr_type);
        return NULL;
    }
    if (filter_dc && sample_rate != 2400000.0) {
        /* the library's DC block is the 1 Hz one at the demodulator's 2.4 MS/s (convert.c:476-480) */
        fprintf(stderr, "no suitable converter for format=%d dc=%d at %.0f samples/s\n", format, filter_dc, sample_rate);
        return NULL;
    }
    b200_demod_config cfg;
    memset(&cfg, 0, sizeof (cfg));
    cfg.abi_version = B200_ABI_VERSION;
    cfg.input_format = (int32_t) format;
    cfg.filter_dc = filter_dc ? 1 : 0; /* convert_*_generic, convert.c:113-213, 374-423 */
#ifdef SC16Q11_TABLE_BITS
    cfg.sc16q11_table_bits = SC16Q11_TABLE_BITS;
#endif
    cfg.nfix_crc = 1;
    cfg.preamble_threshold = 58;
    *out_state = malloc(sizeof (struct converter_state));
    if (!*out_state || b200_demod_create(&cfg, &(*out_state)->demod) != B200_OK) {
        fprintf(stderr, "can't allocate converter state: %s\n", b200_last_error());
        free(*out_state);
        *out_state = NULL;
        return NULL;
    }
    return convert_via_gpu;
}

void cleanup_converter(struct converter_state *state) {
    if (!state)
        return;
    b200_demod_destroy(state->demod);
    free(state);
}
