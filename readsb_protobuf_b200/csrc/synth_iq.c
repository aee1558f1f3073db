/*
 * synth_iq.c -- seeded synthetic 2.4 MS/s Mode S IQ streams (uc8 / sc16 / sc16q11).
 *
 * Workload generator for the tests and for bench.py; it is not on
 * the demodulation path.  The envelope is built on the 12 MHz tick grid readsb timestamps
 * in (readsb's own description of the waveform: demod_2400.c:31-37, 264-273): preamble
 * pulses at 0, 1.0, 3.5 and 4.5 us, each 0.5 us wide; data bit i occupies 8+i us with the
 * high half first for a 1.  One 2.4 MHz sample integrates 5 ticks, so the frame's start tick
 * modulo 5 exercises all five demodulator phases.  Mode A/C replies (df = 32 in the plan) follow
 * readsb's description at demod_2400.c:513-520, 532-557: 20 bit periods of 1.45 us (87 cycles of a
 * 60 MHz clock) with a 0.45 us (27 cycle) pulse where the bit is set, F1 at bit 0, F2 at bit 14.
 *
 * Determinism: the frame plan comes from one sequential PRNG; noise is seeded per 65536
 * sample chunk from (seed, chunk index), so any span renders identically whatever the
 * thread count or the span boundaries.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define SYNTH_CHUNK 65536u
#define FRAME_TICKS_MAX (96 + 112 * 12)

typedef struct {
    uint64_t seed;
    uint64_t nsamples;
    int32_t format; /* 0 = uc8, 1 = sc16, 2 = sc16q11 */
    int32_t n_icao; /* size of the aircraft pool */
    double frames_per_s; /* Poisson arrival rate */
    double noise_sigma; /* per rail, fraction of full scale */
    double amp_min, amp_max; /* frame amplitude, uniform, fraction of full scale */
    double frac_biterror; /* fraction of frames with one flipped bit in bits 5..n-1 */
    double frac_df17, frac_df11; /* the rest is split between DF4, DF5, DF20, DF21 */
    double modeac_per_s; /* Mode A/C replies (a second, independent Poisson process; 0 = none) */
} synth_cfg;

typedef struct {
    uint64_t start_tick; /* 12 MHz tick of the first preamble pulse */
    float amp;
    float phase0; /* carrier phase at the frame start, radians */
    float dphase; /* carrier phase advance per sample, radians */
    int16_t errbit; /* flipped bit index or -1 */
    uint8_t nbytes; /* 7 or 14 */
    uint8_t df;
    uint8_t msg[14]; /* as transmitted (error already applied) */
    uint8_t pad[2];
} synth_frame;

/* ---- PRNG: splitmix64 seeding + xoshiro256** ---- */

typedef struct { uint64_t s[4]; } rng_t;

static uint64_t splitmix64(uint64_t *x) {
    uint64_t z = (*x += 0x9e3779b97f4a7c15ULL);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}

static void rng_seed(rng_t *r, uint64_t a, uint64_t b) {
    uint64_t x = a * 0xd1342543de82ef95ULL + b + 0x2545f4914f6cdd1dULL;
    for (int i = 0; i < 4; ++i)
        r->s[i] = splitmix64(&x);
}

static inline uint64_t rotl64(uint64_t x, int k) {
    return (x << k) | (x >> (64 - k));
}

static inline uint64_t rng_next(rng_t *r) {
    uint64_t *s = r->s;
    const uint64_t result = rotl64(s[1] * 5, 7) * 9;
    const uint64_t t = s[1] << 17;
    s[2] ^= s[0];
    s[3] ^= s[1];
    s[1] ^= s[2];
    s[0] ^= s[3];
    s[2] ^= t;
    s[3] = rotl64(s[3], 45);
    return result;
}

static inline double rng_uniform(rng_t *r) {
    return (double) (rng_next(r) >> 11) * (1.0 / 9007199254740992.0);
}

/* ---- Mode S CRC-24 (generator 0xFFF409), bitwise; used to build valid frames ---- */

static uint32_t crc24(const uint8_t *msg, int nbytes) {
    uint32_t rem = 0;
    for (int i = 0; i < nbytes - 3; ++i) {
        rem ^= (uint32_t) msg[i] << 16;
        for (int b = 0; b < 8; ++b)
            rem = (rem & 0x800000) ? ((rem << 1) ^ 0xfff409u) & 0xffffff : (rem << 1) & 0xffffff;
    }
    return rem;
}

/* ---- Gaussian table: inverse normal CDF at 65536 mid-points (Acklam's rational approximation) ---- */

static float gauss_table[65536];
static int gauss_ready;

static double inv_norm_cdf(double p) {
    static const double a[] = {-3.969683028665376e+01, 2.209460984245205e+02, -2.759285104469687e+02,
                               1.383577518672690e+02, -3.066479806614716e+01, 2.506628277459239e+00};
    static const double b[] = {-5.447609879822406e+01, 1.615858368580409e+02, -1.556989798598866e+02,
                               6.680131188771972e+01, -1.328068155288572e+01};
    static const double c[] = {-7.784894002430293e-03, -3.223964580411365e-01, -2.400758277161838e+00,
                               -2.549732539343734e+00, 4.374664141464968e+00, 2.938163982698783e+00};
    static const double d[] = {7.784695709041462e-03, 3.224671290700398e-01, 2.445134137142996e+00,
                               3.754408661907416e+00};
    const double plow = 0.02425;
    if (p < plow) {
        double q = sqrt(-2 * log(p));
        return (((((c[0] * q + c[1]) * q + c[2]) * q + c[3]) * q + c[4]) * q + c[5]) /
               ((((d[0] * q + d[1]) * q + d[2]) * q + d[3]) * q + 1);
    }
    if (p > 1 - plow) {
        double q = sqrt(-2 * log(1 - p));
        return -(((((c[0] * q + c[1]) * q + c[2]) * q + c[3]) * q + c[4]) * q + c[5]) /
               ((((d[0] * q + d[1]) * q + d[2]) * q + d[3]) * q + 1);
    }
    double q = p - 0.5, r = q * q;
    return (((((a[0] * r + a[1]) * r + a[2]) * r + a[3]) * r + a[4]) * r + a[5]) * q /
           (((((b[0] * r + b[1]) * r + b[2]) * r + b[3]) * r + b[4]) * r + 1);
}

static void gauss_init(void) {
    if (gauss_ready)
        return;
    for (int i = 0; i < 65536; ++i)
        gauss_table[i] = (float) inv_norm_cdf((i + 0.5) / 65536.0);
    gauss_ready = 1;
}

/* ---- frame plan ---- */

static void put_bits(uint8_t *msg, int firstbit, int nbits, uint32_t value) {
    /* firstbit is 1-based, MSB first, like the ICAO annex numbering */
    for (int i = 0; i < nbits; ++i) {
        int bit = firstbit - 1 + i;
        if ((value >> (nbits - 1 - i)) & 1)
            msg[bit >> 3] |= (uint8_t) (1 << (7 - (bit & 7)));
    }
}

/* Fill `frames` (capacity cap) for cfg; returns the number of frames the stream holds
 * (which may exceed cap: call again with a larger buffer). */
int64_t synth_plan(const synth_cfg *cfg, synth_frame *frames, int64_t cap) {
    rng_t r;
    rng_seed(&r, cfg->seed, 0x706c616e);

    int n_icao = cfg->n_icao > 0 ? cfg->n_icao : 1;
    uint32_t *pool = malloc(sizeof (uint32_t) * (size_t) n_icao);
    for (int i = 0; i < n_icao; ++i)
        pool[i] = (uint32_t) (rng_next(&r) % 0xfffffeu) + 1;

    const double total_ticks = (double) cfg->nsamples * 5.0;
    const double mean_gap = 12e6 / (cfg->frames_per_s > 0 ? cfg->frames_per_s : 1e-9);
    double t = 0;
    int64_t n = 0;

    for (;;) {
        double u = rng_uniform(&r);
        t += -log(1.0 - u) * mean_gap;
        if (!(t < total_ticks) || cfg->frames_per_s <= 0)
            break;

        synth_frame f;
        memset(&f, 0, sizeof (f));
        f.start_tick = (uint64_t) t;
        f.amp = (float) (cfg->amp_min + (cfg->amp_max - cfg->amp_min) * rng_uniform(&r));
        f.phase0 = (float) (rng_uniform(&r) * 6.283185307179586);
        /* residual carrier offset up to +-100 kHz -> radians per 2.4 MHz sample */
        f.dphase = (float) ((rng_uniform(&r) - 0.5) * 2.0 * 6.283185307179586 * 100e3 / 2.4e6);
        f.errbit = -1;

        uint32_t icao = pool[rng_next(&r) % (uint64_t) n_icao];
        double kind = rng_uniform(&r);
        uint64_t payload = rng_next(&r);
        uint32_t parity_xor;

        if (kind < cfg->frac_df17) {
            f.df = 17;
            f.nbytes = 14;
            put_bits(f.msg, 1, 5, 17);
            put_bits(f.msg, 6, 3, 5); /* CA = airborne */
            put_bits(f.msg, 9, 24, icao);
            put_bits(f.msg, 33, 32, (uint32_t) (payload >> 24));
            put_bits(f.msg, 65, 24, (uint32_t) (payload & 0xffffff));
            parity_xor = 0;
        } else if (kind < cfg->frac_df17 + cfg->frac_df11) {
            f.df = 11;
            f.nbytes = 7;
            put_bits(f.msg, 1, 5, 11);
            put_bits(f.msg, 6, 3, 5);
            put_bits(f.msg, 9, 24, icao);
            parity_xor = 0; /* IID 0 acquisition squitter */
        } else {
            static const uint8_t dfs[4] = {4, 5, 20, 21};
            f.df = dfs[payload & 3];
            f.nbytes = (f.df >= 16) ? 14 : 7;
            put_bits(f.msg, 1, 5, f.df);
            put_bits(f.msg, 6, 27, (uint32_t) ((payload >> 2) & 0x7ffffff));
            if (f.nbytes == 14) {
                uint64_t more = rng_next(&r);
                put_bits(f.msg, 33, 32, (uint32_t) (more >> 32));
                put_bits(f.msg, 65, 24, (uint32_t) (more & 0xffffff));
            }
            parity_xor = icao; /* Address/Parity */
        }

        uint32_t pi = crc24(f.msg, f.nbytes) ^ parity_xor;
        f.msg[f.nbytes - 3] = (uint8_t) (pi >> 16);
        f.msg[f.nbytes - 2] = (uint8_t) (pi >> 8);
        f.msg[f.nbytes - 1] = (uint8_t) pi;

        if (rng_uniform(&r) < cfg->frac_biterror) {
            int nbits = f.nbytes * 8;
            int bit = 5 + (int) (rng_next(&r) % (uint64_t) (nbits - 5));
            f.msg[bit >> 3] ^= (uint8_t) (1 << (7 - (bit & 7)));
            f.errbit = (int16_t) bit;
        }

        if (n < cap)
            frames[n] = f;
        ++n;
    }

    free(pool);

    /* Mode A/C replies: their own PRNG stream, so that a plan without them is unchanged */
    if (cfg->modeac_per_s > 0) {
        rng_t ra;
        rng_seed(&ra, cfg->seed, 0x6d6f6461);
        const double gap_ac = 12e6 / cfg->modeac_per_s;
        const int64_t n_s = n < cap ? n : cap; /* Mode S frames actually stored */
        int64_t n_ac = 0, stored = n_s;
        double ta = 0;
        for (;;) {
            ta += -log(1.0 - rng_uniform(&ra)) * gap_ac;
            if (!(ta < total_ticks))
                break;
            synth_frame f;
            memset(&f, 0, sizeof (f));
            f.start_tick = (uint64_t) ta;
            f.amp = (float) (cfg->amp_min + (cfg->amp_max - cfg->amp_min) * rng_uniform(&ra));
            f.phase0 = (float) (rng_uniform(&ra) * 6.283185307179586);
            f.dphase = (float) ((rng_uniform(&ra) - 0.5) * 2.0 * 6.283185307179586 * 100e3 / 2.4e6);
            f.errbit = -1;
            f.df = 32;
            f.nbytes = 3;
            /* pulse pattern, bit 19 = F1 ... bit 0 = X5 (the order demod_2400.c:629-651 shifts them in):
             * framing pulses, twelve random data pulses, SPI in one reply of sixteen */
            uint64_t x = rng_next(&ra);
            uint32_t pat = 0x80020u;
            static const int data_bits[12] = {18, 17, 16, 15, 14, 13, 11, 10, 9, 8, 7, 6};
            for (int i = 0; i < 12; ++i)
                if ((x >> i) & 1)
                    pat |= 1u << data_bits[i];
            if (((x >> 12) & 15) == 0)
                pat |= 1u << 2;
            f.msg[0] = (uint8_t) (pat >> 16);
            f.msg[1] = (uint8_t) (pat >> 8);
            f.msg[2] = (uint8_t) pat;
            if (stored < cap)
                frames[stored++] = f;
            ++n_ac;
        }
        /* merge by start tick (insertion from the back: both runs are sorted) */
        if (stored == n_s + n_ac && n == n_s) {
            synth_frame *tmp = malloc(sizeof (synth_frame) * (size_t) (stored > 0 ? stored : 1));
            int64_t i = 0, j = n_s, k = 0;
            while (i < n_s || j < stored) {
                if (j >= stored || (i < n_s && frames[i].start_tick <= frames[j].start_tick))
                    tmp[k++] = frames[i++];
                else
                    tmp[k++] = frames[j++];
            }
            memcpy(frames, tmp, sizeof (synth_frame) * (size_t) stored);
            free(tmp);
        }
        n += n_ac;
    }
    return n;
}

/* ---- rendering ---- */

static void frame_ticks(const synth_frame *f, uint8_t *ticks, int *nticks) {
    int total = 96 + f->nbytes * 8 * 12;
    memset(ticks, 0, (size_t) total + 8);
    static const int pulses[4] = {0, 12, 42, 54};
    for (int p = 0; p < 4; ++p)
        memset(ticks + pulses[p], 1, 6);
    for (int i = 0; i < f->nbytes * 8; ++i) {
        int bit = (f->msg[i >> 3] >> (7 - (i & 7))) & 1;
        memset(ticks + 96 + 12 * i + (bit ? 0 : 6), 1, 6);
    }
    *nticks = total;
}

static void render_chunk(const synth_cfg *cfg, const synth_frame *frames, int64_t nframes,
                         uint64_t chunk_index, uint64_t lo, uint64_t hi, uint8_t *out, uint64_t out_first) {
    /* samples [lo, hi) of chunk chunk_index; out is indexed from out_first */
    const uint64_t c0 = chunk_index * (uint64_t) SYNTH_CHUNK;
    float *bi = malloc(2 * sizeof (float) * SYNTH_CHUNK);
    float *bq = bi + SYNTH_CHUNK;
    if (!bi)
        return;

    rng_t r;
    rng_seed(&r, cfg->seed, 0x6e6f6973 + chunk_index * 2654435761ULL);
    const float sigma = (float) cfg->noise_sigma;
    for (uint32_t k = 0; k < SYNTH_CHUNK; ++k) {
        uint64_t x = rng_next(&r);
        bi[k] = sigma * gauss_table[x & 0xffff];
        bq[k] = sigma * gauss_table[(x >> 16) & 0xffff];
    }

    /* frames overlapping this chunk: a frame spans at most ceil(1440/5)+1 samples */
    const uint64_t chunk_tick0 = c0 * 5, chunk_tick1 = (c0 + SYNTH_CHUNK) * 5;
    int64_t a = 0, b = nframes;
    uint64_t want = chunk_tick0 > (uint64_t) (FRAME_TICKS_MAX + 5) ? chunk_tick0 - (FRAME_TICKS_MAX + 5) : 0;
    while (a < b) { /* first frame with start_tick >= want */
        int64_t m = (a + b) / 2;
        if (frames[m].start_tick < want)
            a = m + 1;
        else
            b = m;
    }
    uint8_t ticks[FRAME_TICKS_MAX + 16];
    for (int64_t i = a; i < nframes && frames[i].start_tick < chunk_tick1; ++i) {
        const synth_frame *f = &frames[i];
        if (f->df == 32) { /* Mode A/C reply on the 60 MHz grid: a sample is 25 cycles */
            const uint32_t pat = ((uint32_t) f->msg[0] << 16) | ((uint32_t) f->msg[1] << 8) | f->msg[2];
            const int64_t cyc0 = (int64_t) f->start_tick * 5;
            const uint64_t s_first_ac = f->start_tick / 5;
            for (int bit = 0; bit < 20; ++bit) {
                if (!((pat >> (19 - bit)) & 1))
                    continue;
                const int64_t on0 = cyc0 + 87 * bit, on1 = on0 + 27;
                for (int64_t s = on0 / 25; s * 25 < on1; ++s) {
                    if (s < (int64_t) c0 || s >= (int64_t) (c0 + SYNTH_CHUNK))
                        continue;
                    const int64_t lo25 = s * 25 > on0 ? s * 25 : on0, hi25 = (s + 1) * 25 < on1 ? (s + 1) * 25 : on1;
                    if (hi25 <= lo25)
                        continue;
                    float env = f->amp * (float) (hi25 - lo25) * 0.04f;
                    float ph = f->phase0 + f->dphase * (float) (s - (int64_t) s_first_ac);
                    bi[s - c0] += env * cosf(ph);
                    bq[s - c0] += env * sinf(ph);
                }
            }
            continue;
        }
        int nticks;
        frame_ticks(f, ticks, &nticks);
        uint64_t s_first = f->start_tick / 5;
        uint64_t s_last = (f->start_tick + (uint64_t) nticks + 4) / 5;
        for (uint64_t s = s_first; s <= s_last; ++s) {
            if (s < c0 || s >= c0 + SYNTH_CHUNK)
                continue;
            int64_t rel = (int64_t) (s * 5) - (int64_t) f->start_tick;
            int high = 0;
            for (int k = 0; k < 5; ++k) {
                int64_t tk = rel + k;
                if (tk >= 0 && tk < nticks)
                    high += ticks[tk];
            }
            if (!high)
                continue;
            float env = f->amp * (float) high * 0.2f;
            float ph = f->phase0 + f->dphase * (float) (int64_t) (s - s_first);
            bi[s - c0] += env * cosf(ph);
            bq[s - c0] += env * sinf(ph);
        }
    }

    for (uint64_t s = lo; s < hi; ++s) {
        float fi = bi[s - c0], fq = bq[s - c0];
        uint64_t o = s - out_first;
        if (cfg->format == 0) {
            float vi = rintf(fi * 127.5f + 127.5f), vq = rintf(fq * 127.5f + 127.5f);
            vi = vi < 0 ? 0 : (vi > 255 ? 255 : vi);
            vq = vq < 0 ? 0 : (vq > 255 ? 255 : vq);
            out[2 * o] = (uint8_t) vi;
            out[2 * o + 1] = (uint8_t) vq;
        } else {
            float scale = (cfg->format == 1) ? 32767.0f : 2047.0f;
            float vi = rintf(fi * scale), vq = rintf(fq * scale);
            vi = vi < -scale ? -scale : (vi > scale ? scale : vi);
            vq = vq < -scale ? -scale : (vq > scale ? scale : vq);
            int16_t ii = (int16_t) vi, qq = (int16_t) vq;
            /* little-endian int16 pairs, as convert.c:231-232 reads them */
            out[4 * o] = (uint8_t) (ii & 0xff);
            out[4 * o + 1] = (uint8_t) ((ii >> 8) & 0xff);
            out[4 * o + 2] = (uint8_t) (qq & 0xff);
            out[4 * o + 3] = (uint8_t) ((qq >> 8) & 0xff);
        }
    }
    free(bi);
}

/* Render samples [first, first+count) of the stream into out (count * bytes_per_sample bytes). */
void synth_render(const synth_cfg *cfg, const synth_frame *frames, int64_t nframes,
                  uint64_t first, uint64_t count, void *out) {
    gauss_init();
    if (count == 0)
        return;
    const uint64_t end = first + count;
    const int64_t ch0 = (int64_t) (first / SYNTH_CHUNK), ch1 = (int64_t) ((end - 1) / SYNTH_CHUNK);
#pragma omp parallel for schedule(dynamic, 4)
    for (int64_t ch = ch0; ch <= ch1; ++ch) {
        uint64_t lo = (uint64_t) ch * SYNTH_CHUNK, hi = lo + SYNTH_CHUNK;
        if (lo < first)
            lo = first;
        if (hi > end)
            hi = end;
        render_chunk(cfg, frames, nframes, (uint64_t) ch, lo, hi, (uint8_t *) out, first);
    }
}

uint32_t synth_crc24(const uint8_t *msg, int nbytes) {
    return crc24(msg, nbytes);
}

int synth_frame_size(void) {
    return (int) sizeof (synth_frame);
}
