/*
 * TEST INFRASTRUCTURE ONLY -- see modes_oracle.h.  Plain-C restatement of the reference path;
 * every function cites the reference file:line it follows.  Not linked into the product.
 */
#define _GNU_SOURCE
#include "modes_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <stdio.h>
#include <string.h>
#include <time.h>

/* ======================================================================================
 * CRC-24 and syndrome tables (crc.c)
 * ==================================================================================== */

#define GENERATOR_POLY 0xfff409u /* crc.c:31 */

static uint32_t crc_table[256];
static uint32_t bit_syndrome[112];
static int crc_ready;

static uint32_t checksum_raw(const uint8_t *msg, int bits) {
    /* crc.c:67-82 */
    uint32_t rem = 0;
    int n = bits / 8;
    for (int i = 0; i < n - 3; ++i) {
        rem = (rem << 8) ^ crc_table[msg[i] ^ ((rem & 0xff0000) >> 16)];
        rem &= 0xffffff;
    }
    return rem ^ ((uint32_t) msg[n - 3] << 16) ^ ((uint32_t) msg[n - 2] << 8) ^ msg[n - 1];
}

static void crc_init(void) {
    /* crc.c:42-65 */
    if (crc_ready)
        return;
    for (int i = 0; i < 256; ++i) {
        uint32_t c = (uint32_t) i << 16;
        for (int j = 0; j < 8; ++j)
            c = (c & 0x800000) ? (c << 1) ^ GENERATOR_POLY : (c << 1);
        crc_table[i] = c & 0xffffff;
    }
    uint8_t msg[14];
    memset(msg, 0, sizeof (msg));
    for (int i = 0; i < 112; ++i) {
        msg[i / 8] ^= (uint8_t) (1 << (7 - (i & 7)));
        bit_syndrome[i] = checksum_raw(msg, 112);
        msg[i / 8] ^= (uint8_t) (1 << (7 - (i & 7)));
    }
    crc_ready = 1;
}

uint32_t mo_checksum(const uint8_t *msg, int bits) {
    crc_init();
    return checksum_raw(msg, bits);
}

uint32_t mo_single_bit_syndrome(int bit) {
    crc_init();
    return bit_syndrome[bit];
}

static int cmp_syndrome(const void *x, const void *y) {
    /* crc.c:92-96 */
    return (int) ((const mo_errorinfo *) x)->syndrome - (int) ((const mo_errorinfo *) y)->syndrome;
}

static int choose(int n, int k) {
    /* crc.c:100-115 */
    if (k == 0 || k == n)
        return 1;
    if (k > n)
        return 0;
    int r = 1;
    for (int i = 1; i <= k; ++i) {
        r = r * n / i;
        --n;
    }
    return r;
}

static int fill_subtable(mo_errorinfo *t, int n, int offset, int startbit, int endbit,
                         const mo_errorinfo *base, int error_bit, int max_errors) {
    /* crc.c:133-152: every combination of up to max_errors flipped bits, depth first */
    if (error_bit >= max_errors)
        return n;
    for (int i = startbit; i < endbit; ++i) {
        t[n] = *base;
        t[n].syndrome ^= bit_syndrome[i + offset];
        t[n].errors = error_bit + 1;
        t[n].bit[error_bit] = (int8_t) i;
        ++n;
        n = fill_subtable(t, n, offset, i + 1, endbit, &t[n - 1], error_bit + 1, max_errors);
    }
    return n;
}

static mo_errorinfo *find_syndrome(mo_errorinfo *t, int size, uint32_t syndrome) {
    int lo = 0, hi = size - 1;
    while (lo <= hi) {
        int mid = (lo + hi) / 2;
        if (t[mid].syndrome == syndrome)
            return &t[mid];
        if ((int) t[mid].syndrome < (int) syndrome)
            lo = mid + 1;
        else
            hi = mid - 1;
    }
    return NULL;
}

static int flag_collisions(mo_errorinfo *t, int size, int offset, int startbit, int endbit,
                           uint32_t base_syndrome, int error_bit, int first_error, int last_error) {
    /* crc.c:154-178: patterns of first_error..last_error bits that alias a table entry */
    if (error_bit > last_error)
        return 0;
    int count = 0;
    for (int i = startbit; i < endbit; ++i) {
        uint32_t s = base_syndrome ^ bit_syndrome[i + offset];
        if (error_bit >= first_error) {
            mo_errorinfo *hit = find_syndrome(t, size, s);
            if (hit && hit->errors != -1) {
                ++count;
                hit->errors = -1;
            }
        }
        count += flag_collisions(t, size, offset, i + 1, endbit, s, error_bit + 1, first_error, last_error);
    }
    return count;
}

static mo_errorinfo *build_error_table(int bits, int max_correct, int max_detect, int *size_out) {
    /* crc.c:184-354 */
    crc_init();
    *size_out = 0;
    if (!max_correct)
        return NULL;
    int maxsize = 0;
    for (int i = 1; i <= max_correct; ++i)
        maxsize += choose(bits - 5, i);
    mo_errorinfo *t = calloc((size_t) maxsize + 1, sizeof (*t));
    mo_errorinfo base;
    memset(&base, 0, sizeof (base));
    base.bit[0] = base.bit[1] = -1;
    /* the 5 DF bits are never corrected (crc.c:214-215) */
    int used = fill_subtable(t, 0, 112 - bits, 5, bits, &base, 0, max_correct);
    qsort(t, (size_t) used, sizeof (*t), cmp_syndrome);

    /* crc.c:247-267: drop every syndrome that more than one pattern produces */
    int j = 0;
    for (int i = 0; i < used; ++i) {
        if (i < used - 1 && t[i + 1].syndrome == t[i].syndrome) {
            while (i < used - 1 && t[i + 1].syndrome == t[i].syndrome)
                ++i;
            continue;
        }
        t[j++] = t[i];
    }
    used = j;

    /* crc.c:269-298 */
    if (max_detect > max_correct) {
        int flagged = flag_collisions(t, used, 112 - bits, 5, bits, 0, 1, max_correct + 1, max_detect);
        if (flagged > 0) {
            j = 0;
            for (int i = 0; i < used; ++i)
                if (t[i].errors != -1)
                    t[j++] = t[i];
            used = j;
        }
    }
    *size_out = used;
    return t;
}

typedef struct {
    mo_errorinfo *short_table, *long_table;
    int short_size, long_size;
} crc_tables;

static void crc_tables_init(crc_tables *ct, int nfix) {
    /* crc.c:358-383 */
    memset(ct, 0, sizeof (*ct));
    if (nfix == 1) {
        ct->short_table = build_error_table(56, 1, 1, &ct->short_size);
        ct->long_table = build_error_table(112, 1, 1, &ct->long_size);
    } else if (nfix >= 2) {
        ct->short_table = build_error_table(56, 2, 4, &ct->short_size);
        ct->long_table = build_error_table(112, 2, 4, &ct->long_size);
    }
}

static void crc_tables_free(crc_tables *ct) {
    free(ct->short_table);
    free(ct->long_table);
}

int mo_error_table(int nfix, int bits, mo_errorinfo *out, int cap) {
    crc_tables ct;
    crc_tables_init(&ct, nfix);
    mo_errorinfo *t = (bits == 56) ? ct.short_table : ct.long_table;
    int n = (bits == 56) ? ct.short_size : ct.long_size;
    for (int i = 0; i < n && i < cap; ++i)
        out[i] = t[i];
    crc_tables_free(&ct);
    return n;
}

static const mo_errorinfo NO_ERRORS = {0, 0, {0, 0}, 0}; /* crc.c:28 */

static const mo_errorinfo *diagnose(const crc_tables *ct, uint32_t syndrome, int bitlen) {
    /* crc.c:389-412 */
    if (syndrome == 0)
        return &NO_ERRORS;
    mo_errorinfo *t = (bitlen == 56) ? ct->short_table : ct->long_table;
    int n = (bitlen == 56) ? ct->short_size : ct->long_size;
    if (!t)
        return NULL;
    return find_syndrome(t, n, syndrome);
}

static void apply_fix(uint8_t *msg, const mo_errorinfo *ei) {
    /* crc.c:417-425 */
    for (int i = 0; i < ei->errors; ++i)
        msg[ei->bit[i] >> 3] ^= (uint8_t) (1 << (7 - (ei->bit[i] & 7)));
}

/* ======================================================================================
 * Recently-seen ICAO address filter (icao_filter.c)
 * ==================================================================================== */

#define FILTER_SIZE 8192      /* icao_filter.c:27 */
#define FILTER_TTL 60000      /* icao_filter.c:30 */
#define FILTER_EMPTY 0xffffffffu

typedef struct {
    uint32_t a[FILTER_SIZE], b[FILTER_SIZE];
    uint32_t *active;
    uint64_t next_flip;
} icao_filter;

static uint32_t filter_hash(uint32_t a) {
    /* icao_filter.c:44-65, Jenkins one-at-a-time over 3 bytes */
    uint32_t h = 0;
    for (int k = 0; k < 3; ++k) {
        h += (a >> (8 * k)) & 0xff;
        h += h << 10;
        h ^= h >> 6;
    }
    h += h << 3;
    h ^= h >> 11;
    h += h << 15;
    return h & (FILTER_SIZE - 1);
}

static void filter_init(icao_filter *f) {
    /* icao_filter.c:67-71 */
    memset(f->a, 0xff, sizeof (f->a));
    memset(f->b, 0xff, sizeof (f->b));
    f->active = f->a;
    f->next_flip = 0;
}

static void filter_add(icao_filter *f, uint32_t addr) {
    /* icao_filter.c:73-97 */
    uint32_t h, h0;
    h0 = h = filter_hash(addr);
    while (f->active[h] != FILTER_EMPTY && f->active[h] != addr) {
        h = (h + 1) & (FILTER_SIZE - 1);
        if (h == h0)
            return;
    }
    if (f->active[h] == FILTER_EMPTY)
        f->active[h] = addr;

    h0 = h = filter_hash(addr & 0x00ffff);
    while (f->active[h] != FILTER_EMPTY && (f->active[h] & 0x00ffff) != (addr & 0x00ffff)) {
        h = (h + 1) & (FILTER_SIZE - 1);
        if (h == h0)
            return;
    }
    if (f->active[h] == FILTER_EMPTY)
        f->active[h] = addr;
}

static int filter_probe(const uint32_t *t, uint32_t addr) {
    uint32_t h, h0;
    h0 = h = filter_hash(addr);
    while (t[h] != FILTER_EMPTY && t[h] != addr) {
        h = (h + 1) & (FILTER_SIZE - 1);
        if (h == h0)
            break;
    }
    return t[h] == addr;
}

static int filter_test(const icao_filter *f, uint32_t addr) {
    /* icao_filter.c:99-122 */
    return filter_probe(f->a, addr) || filter_probe(f->b, addr);
}

static void filter_expire(icao_filter *f, uint64_t now) {
    /* icao_filter.c:150-164; `now` is mstime() == Modes.ifile_now for ifile (util.c:61-64) */
    if (now >= f->next_flip) {
        if (f->active == f->a) {
            memset(f->b, 0xff, sizeof (f->b));
            f->active = f->b;
        } else {
            memset(f->a, 0xff, sizeof (f->a));
            f->active = f->a;
        }
        f->next_flip = now + FILTER_TTL;
    }
}

/* ======================================================================================
 * IQ -> magnitude (convert.c)
 * ==================================================================================== */

static uint16_t uc8_table[65536];
static int uc8_ready;

static void uc8_init(void) {
    /* convert.c:45-58.  The table index is the little-endian u16 the converter loads
     * (convert.c:69,80): first byte (I) in the low half. */
    if (uc8_ready)
        return;
    for (int i = 0; i <= 255; i++) {
        for (int q = 0; q <= 255; q++) {
            float fI, fQ, magsq;
            fI = (i - 127.5) / 127.5;
            fQ = (q - 127.5) / 127.5;
            magsq = fI * fI + fQ * fQ;
            if (magsq > 1)
                magsq = 1;
            float mag = sqrtf(magsq);
            uc8_table[(i * 256) + q] = (uint16_t) (mag * 65535.0f + 0.5f);
        }
    }
    uc8_ready = 1;
}

void mo_uc8_table(uint16_t *out) {
    uc8_init();
    memcpy(out, uc8_table, sizeof (uc8_table));
}

static void convert_uc8(const uint8_t *in, uint32_t n, uint16_t *mag, double *mean_level, double *mean_power) {
    /* convert.c:63-111 */
    uc8_init();
    uint64_t sum_level = 0, sum_power = 0;
    for (uint32_t i = 0; i < n; ++i) {
        uint16_t m = uc8_table[(uint32_t) in[2 * i] | ((uint32_t) in[2 * i + 1] << 8)];
        mag[i] = m;
        sum_level += m;
        sum_power += (uint32_t) m * (uint32_t) m;
    }
    if (mean_level)
        *mean_level = sum_level / 65536.0 / n; /* sic: 65536, convert.c:105 */
    if (mean_power)
        *mean_power = sum_power / 65535.0 / 65535.0 / n;
}

static void convert_sc16_scaled(const uint8_t *in, uint32_t n, float scale, uint16_t *mag,
                                double *mean_level, double *mean_power) {
    /* convert.c:215-253 (scale 32768) and convert.c:332-370 (scale 2048) */
    float sum_level = 0, sum_power = 0;
    for (uint32_t i = 0; i < n; ++i) {
        int16_t I = (int16_t) ((uint16_t) in[4 * i] | ((uint16_t) in[4 * i + 1] << 8));
        int16_t Q = (int16_t) ((uint16_t) in[4 * i + 2] | ((uint16_t) in[4 * i + 3] << 8));
        float fI = I / scale;
        float fQ = Q / scale;
        float magsq = fI * fI + fQ * fQ;
        if (magsq > 1)
            magsq = 1;
        float m = sqrtf(magsq);
        sum_power += magsq;
        sum_level += m;
        mag[i] = (uint16_t) (m * 65535.0f + 0.5f);
    }
    if (mean_level)
        *mean_level = sum_level / n;
    if (mean_power)
        *mean_power = sum_power / n;
}

/* convert.c:113-162 (uc8), 164-213 (sc16), 374-423 (sc16q11): the "generic" float converters that
 * --dcfilter selects.  A one-pole DC block z = f*a + z*b runs per rail over the whole stream (its state
 * is carried from call to call in struct converter_state), the magnitude is taken of f - z.  Every step
 * is a separate float operation in the order the C code spells (no contraction: -ffp-contract=off,
 * like the reference's x86-64 -O2 build). */
void mo_dc_init(mo_dc_state *st, double sample_rate) {
    /* init_converter, convert.c:476-488 with filter_dc != 0: float fields assigned from double expressions */
    st->z1_I = 0;
    st->z1_Q = 0;
    st->dc_b = (float) exp(-2.0 * M_PI * 1.0 / sample_rate);
    st->dc_a = (float) (1.0 - st->dc_b);
}

int mo_convert_dc(int format, const void *iq, uint32_t n, mo_dc_state *st, uint16_t *mag, double *mean_level,
                  double *mean_power) {
    if (format < MO_UC8 || format > MO_SC16Q11)
        return -1;
    const uint8_t *in = iq;
    float z1_I = st->z1_I, z1_Q = st->z1_Q;
    const float dc_a = st->dc_a, dc_b = st->dc_b;
    float sum_level = 0, sum_power = 0;
    for (uint32_t i = 0; i < n; ++i) {
        float fI, fQ;
        if (format == MO_UC8) { /* convert.c:131-134 */
            fI = (in[2 * i] - 127.5f) / 127.5f;
            fQ = (in[2 * i + 1] - 127.5f) / 127.5f;
        } else { /* convert.c:182-185 / 392-395 */
            const float scale = (format == MO_SC16) ? 32768.0f : 2048.0f;
            int16_t I = (int16_t) ((uint16_t) in[4 * i] | ((uint16_t) in[4 * i + 1] << 8));
            int16_t Q = (int16_t) ((uint16_t) in[4 * i + 2] | ((uint16_t) in[4 * i + 3] << 8));
            fI = I / scale;
            fQ = Q / scale;
        }
        /* DC block, convert.c:136-140 */
        z1_I = fI * dc_a + z1_I * dc_b;
        z1_Q = fQ * dc_a + z1_Q * dc_b;
        fI -= z1_I;
        fQ -= z1_Q;
        float magsq = fI * fI + fQ * fQ;
        if (magsq > 1)
            magsq = 1;
        float m = sqrtf(magsq);
        sum_power += magsq;
        sum_level += m;
        mag[i] = (uint16_t) (m * 65535.0f + 0.5f);
    }
    st->z1_I = z1_I;
    st->z1_Q = z1_Q;
    if (mean_level)
        *mean_level = sum_level / n;
    if (mean_power)
        *mean_power = sum_power / n;
    return 0;
}

void mo_sc16q11_table(int bits, uint16_t *table) {
    /* init_sc16q11_lookup, convert.c:270-294 (USE_BITS = bits, LOSE_BITS = 11 - bits) */
    const int lose = 11 - bits;
    for (int i = 0; i < 2048; i += (1 << lose)) {
        for (int q = 0; q < 2048; q += (1 << lose)) {
            float fI = i / 2048.0, fQ = q / 2048.0; /* double division, rounded to float */
            float magsq = fI * fI + fQ * fQ;
            if (magsq > 1)
                magsq = 1;
            float m = sqrtf(magsq);
            unsigned index = ((unsigned) (i >> lose) << bits) | (unsigned) (q >> lose);
            table[index] = (uint16_t) (m * 65535.0f + 0.5f);
        }
    }
}

int mo_convert_sc16q11_table(int bits, const void *iq, uint32_t n, uint16_t *mag, double *mean_level, double *mean_power) {
    /* convert_sc16q11_table, convert.c:296-328 */
    if (bits < 1 || bits > 11)
        return -1;
    static uint16_t *table;
    static int table_bits;
    if (!table || table_bits != bits) {
        free(table);
        table = malloc(sizeof (uint16_t) << (2 * bits));
        mo_sc16q11_table(bits, table);
        table_bits = bits;
    }
    const uint8_t *in = iq;
    const int lose = 11 - bits;
    uint64_t sum_level = 0, sum_power = 0;
    for (uint32_t i = 0; i < n; ++i) {
        int16_t sI = (int16_t) ((uint16_t) in[4 * i] | ((uint16_t) in[4 * i + 1] << 8));
        int16_t sQ = (int16_t) ((uint16_t) in[4 * i + 2] | ((uint16_t) in[4 * i + 3] << 8));
        uint16_t I = abs(sI) & 2047; /* abs() of the promoted int: -32768 -> 32768 -> 0 */
        uint16_t Q = abs(sQ) & 2047;
        uint16_t m = table[((unsigned) (I >> lose) << bits) | (unsigned) (Q >> lose)];
        mag[i] = m;
        sum_level += m;
        sum_power += (uint32_t) m * (uint32_t) m;
    }
    if (mean_level)
        *mean_level = sum_level / 65536.0 / n;
    if (mean_power)
        *mean_power = sum_power / 65535.0 / 65535.0 / n;
    return 0;
}

int mo_convert(int format, const void *iq, uint32_t n, uint16_t *mag, double *mean_level, double *mean_power) {
    /* converter choice: convert.c:425-444 with filter_dc == 0 */
    switch (format) {
        case MO_UC8:
            convert_uc8(iq, n, mag, mean_level, mean_power);
            return 0;
        case MO_SC16:
            convert_sc16_scaled(iq, n, 32768.0f, mag, mean_level, mean_power);
            return 0;
        case MO_SC16Q11:
            convert_sc16_scaled(iq, n, 2048.0f, mag, mean_level, mean_power);
            return 0;
    }
    return -1;
}

/* ======================================================================================
 * Preamble scan and PPM slicer (demod_2400.c)
 * ==================================================================================== */

int mo_try_mask(const uint16_t *m, uint32_t j, int threshold) {
    const uint16_t *pa = &m[j];
    /* demod_2400.c:276 */
    if (!(pa[1] > pa[7] && pa[12] > pa[14] && pa[12] > pa[15]))
        return 0;
    /* demod_2400.c:281-292 */
    int32_t base_noise = pa[5] + pa[8] + pa[16] + pa[17] + pa[18];
    int32_t ref_level = (base_noise * threshold) >> 5;
    /* demod_2400.c:298-301 */
    int32_t diff_2_3 = pa[2] - pa[3];
    int32_t sum_1_4 = pa[1] + pa[4];
    int32_t diff_10_11 = pa[10] - pa[11];
    int32_t common3456 = sum_1_4 - diff_2_3 + pa[9] + pa[12];
    int mask = 0;
    if (common3456 - diff_10_11 >= ref_level) /* demod_2400.c:306-312 */
        mask |= 0x03;
    if (common3456 + diff_10_11 >= ref_level) /* demod_2400.c:316-322 */
        mask |= 0x0c;
    if (sum_1_4 + 2 * diff_2_3 + diff_10_11 + pa[12] >= ref_level) /* demod_2400.c:327-330 */
        mask |= 0x10;
    return mask;
}

/* demod_2400.c:73-93: the five correlators as one coefficient table */
static const int slice_coeff[5][4] = {
    {18, -15, -3, 0}, {14, -5, -9, 0}, {16, 5, -20, 0}, {7, 11, -18, 0}, {4, 15, -20, 1},
};

void mo_slice(const uint16_t *m, uint32_t j, int try_phase, int nbytes, uint8_t *msg) {
    /* demod_2400.c:98-177,188-209.  slice_byte() walks pPtr/phase so that bit b of the frame is
     * taken at sub-sample offset t = try_phase + 12*b (in fifths of a sample) after m[j+19]:
     * sample index t/5, correlator t%5. */
    for (int k = 0; k < nbytes; ++k) {
        uint8_t byte = 0;
        for (int i = 0; i < 8; ++i) {
            int t = try_phase + 12 * (8 * k + i);
            const uint16_t *p = &m[j + 19 + t / 5];
            const int *c = slice_coeff[t % 5];
            int v = c[0] * p[0] + c[1] * p[1] + c[2] * p[2] + c[3] * p[3];
            if (v > 0)
                byte |= (uint8_t) (0x80 >> i);
        }
        msg[k] = byte;
    }
}

/* ======================================================================================
 * Scoring and the CRC-dependent part of decode (mode_s.c)
 * ==================================================================================== */

typedef struct {
    mo_config cfg;
    crc_tables crc;
    icao_filter filter;
    mo_stats stats;
    uint64_t ifile_now; /* Modes.ifile_now */
} demod_state;

static uint32_t aa_field(const uint8_t *msg) {
    /* getbits(msg, 9, 32) */
    return ((uint32_t) msg[1] << 16) | ((uint32_t) msg[2] << 8) | msg[3];
}

static void correct_aa(uint32_t *addr, const mo_errorinfo *ei) {
    /* mode_s.c:266-281 */
    for (int i = 0; i < ei->errors; ++i)
        if (ei->bit[i] >= 8 && ei->bit[i] <= 31)
            *addr ^= 1u << (31 - ei->bit[i]);
}

static int score_message(demod_state *s, const uint8_t *msg, int validbits) {
    /* mode_s.c:311-409 */
    static const uint8_t zeros[14] = {0};
    if (validbits < 56)
        return -2;
    int msgtype = msg[0] >> 3;
    int msgbits = (msgtype & 0x10) ? 112 : 56; /* mode_s.c:81-83 */
    if (validbits < msgbits)
        return -2;
    if (!memcmp(zeros, msg, (size_t) msgbits / 8))
        return -2;
    uint32_t crc = checksum_raw(msg, msgbits);
    const mo_errorinfo *ei;
    uint32_t addr;

    switch (msgtype) {
        case 0: case 4: case 5: case 16:
        case 24: case 25: case 26: case 27: case 28: case 29: case 30: case 31:
            return filter_test(&s->filter, crc) ? 1000 : -1;
        case 11: {
            uint32_t iid = crc & 0x7f;
            crc &= 0xffff80;
            addr = aa_field(msg);
            ei = diagnose(&s->crc, crc, msgbits);
            if (!ei)
                return -2;
            if (ei->errors > 1)
                return -2;
            correct_aa(&addr, ei);
            if (iid == 0)
                return (filter_test(&s->filter, addr) ? 1600 : 750) / (ei->errors + 1);
            return filter_test(&s->filter, addr) ? 1000 / (ei->errors + 1) : -1;
        }
        case 17: case 18:
            ei = diagnose(&s->crc, crc, msgbits);
            if (!ei)
                return -2;
            addr = aa_field(msg);
            correct_aa(&addr, ei);
            return (filter_test(&s->filter, addr) ? 1800 : 1400) / (ei->errors + 1);
        case 20: case 21:
            return filter_test(&s->filter, crc) ? 1000 : -2;
        default:
            return -2;
    }
}

/* The part of decodeModesMessage() that can reject a message or touch the filter
 * (mode_s.c:424-555 and 717-726).  msg is corrected in place. */
/* DF18: is the AA field something other than an ICAO address?  The extended-squitter decoder then
 * flags mm->addr with MODES_NON_ICAO_ADDRESS (1 << 24): by CF alone (mode_s.c:1379-1428), or for CF 2 / 3 /
 * 6 by the IMF bit of the ME field, whose position depends on the ME type (mode_s.c:806, 927, 966-968,
 * 1054, 1064, 1259, 1404-1406).  msg = the frame after CRC repair. */
static int me_bit(const uint8_t *me, int n) { /* 1-based, MSB first (getbit, mode_s.c) */
    return (me[(n - 1) >> 3] >> (7 - ((n - 1) & 7))) & 1;
}

static int df18_non_icao(const uint8_t *msg) {
    const uint8_t *me = msg + 4;
    const unsigned cf = msg[0] & 7, metype = me[0] >> 3, mesub3 = me[0] & 7;
    switch (cf) {
        case 0: return 0;
        case 1: case 5: return 1;
        case 3: return me_bit(me, 1);
        case 2: case 6: break; /* look for the IMF bit */
        default: return 1;     /* unknown format: assumed non-ICAO */
    }
    if (metype == 19)
        return mesub3 >= 1 && mesub3 <= 4 && me_bit(me, 9);
    if (metype >= 5 && metype <= 8)
        return me_bit(me, 21);
    if (metype == 0 || (metype >= 9 && metype <= 18) || (metype >= 20 && metype <= 22))
        return me_bit(me, 8);
    if (metype == 28)
        return mesub3 == 1 && me_bit(me, 56);
    if (metype == 29)
        return me_bit(me, 51);
    if (metype == 31)
        return me_bit(me, 56);
    return 0;
}

static int decode_crc_part(demod_state *s, mo_msg *mm, const uint8_t *raw) {
    static const uint8_t zeros[7] = {0};
    uint8_t *msg = mm->msg;
    memcpy(mm->msg, raw, 14);
    memcpy(mm->verbatim, raw, 14); /* Modes.net_verbatim (mode_s.c:427-430) */
    if (!memcmp(zeros, msg, 7))
        return -2;
    mm->msgtype = msg[0] >> 3;
    mm->msgbits = (mm->msgtype & 0x10) ? 112 : 56;
    mm->crc = checksum_raw(msg, mm->msgbits);
    mm->correctedbits = 0;
    mm->addr = 0;
    uint32_t iid = 0;

    switch (mm->msgtype) {
        case 0: case 4: case 5: case 16:
        case 24: case 25: case 26: case 27: case 28: case 29: case 30: case 31:
            if (!filter_test(&s->filter, mm->crc))
                return -1;
            mm->addr = mm->crc;
            break;
        case 11:
            iid = mm->crc & 0x7f;
            if (mm->crc & 0xffff80) {
                const mo_errorinfo *ei = diagnose(&s->crc, mm->crc & 0xffff80, mm->msgbits);
                if (!ei)
                    return -2;
                if (ei->errors > 1)
                    return -2;
                mm->correctedbits = (uint8_t) ei->errors;
                apply_fix(msg, ei);
                if (!filter_test(&s->filter, aa_field(msg)))
                    return -1;
            }
            break;
        case 17: case 18:
            if (mm->crc != 0) {
                const mo_errorinfo *ei = diagnose(&s->crc, mm->crc, mm->msgbits);
                if (!ei)
                    return -2;
                uint32_t addr1 = aa_field(msg);
                mm->correctedbits = (uint8_t) ei->errors;
                apply_fix(msg, ei);
                uint32_t addr2 = aa_field(msg);
                if (addr1 != addr2 && !filter_test(&s->filter, addr2))
                    return -1;
            }
            break;
        case 20: case 21:
            if (!filter_test(&s->filter, mm->crc))
                return -1;
            mm->addr = mm->crc;
            break;
        default:
            return -2;
    }

    if (mm->msgtype == 11 || mm->msgtype == 17 || mm->msgtype == 18)
        mm->addr = aa_field(msg); /* mode_s.c:560-562 */

    /* mode_s.c:717-726: the only place addresses enter the filter */
    if (!mm->correctedbits && (mm->msgtype == 17 || (mm->msgtype == 11 && iid == 0)))
        filter_add(&s->filter, mm->addr);
    /* decodeExtendedSquitter (mode_s.c:1373-1428) runs later in decodeModesMessage and may flag the address */
    if (mm->msgtype == 18 && df18_non_icao(mm->msg))
        mm->addr |= 1u << 24;
    return 0;
}

/* ======================================================================================
 * demodulate2400() over one mag_buf (demod_2400.c:236-428)
 * ==================================================================================== */

typedef struct {
    mo_msg *msgs;
    uint64_t n, cap;
} msg_list;

static void push_msg(msg_list *l, const mo_msg *m) {
    if (l->n == l->cap) {
        l->cap = l->cap ? l->cap * 2 : 1024;
        l->msgs = realloc(l->msgs, l->cap * sizeof (mo_msg));
    }
    l->msgs[l->n++] = *m;
}

static void demod_block(demod_state *s, const uint16_t *m, uint32_t mlen, uint64_t sampleTimestamp,
                        uint64_t sysTimestamp, double mean_power, msg_list *out) {
    uint64_t sum_scaled_signal_power = 0;
    s->ifile_now = sysTimestamp; /* demod_2400.c:253-255 */

    for (uint32_t j = 0; j < mlen; j++) {
        int mask = mo_try_mask(m, j, s->cfg.threshold);
        if (!mask)
            continue;

        uint8_t bufs[2][14], *msg = bufs[0], *bestmsg = NULL;
        int bestscore = -42, bestphase = -1;
        memset(bufs, 0, sizeof (bufs));

        for (int try_phase = 4; try_phase <= 8; ++try_phase) {
            if (!(mask & (1 << (try_phase - 4))))
                continue;
            /* score_phase(), demod_2400.c:183-229 */
            s->stats.demod_preamblePhase[try_phase - 4]++;
            mo_slice(m, j, try_phase, 1, msg);
            int bytelen;
            switch (msg[0] >> 3) {
                case 0: case 4: case 5: case 11:
                    bytelen = 7;
                    break;
                case 16: case 17: case 18: case 20: case 21: case 24:
                    bytelen = 14;
                    break;
                default:
                    bytelen = 1;
                    break;
            }
            int score = -2;
            if (bytelen > 1) {
                mo_slice(m, j, try_phase, bytelen, msg);
                score = score_message(s, msg, bytelen * 8);
            }
            if (score > bestscore) {
                bestmsg = msg;
                bestscore = score;
                bestphase = try_phase;
                msg = (msg == bufs[0]) ? bufs[1] : bufs[0];
            }
        }

        s->stats.demod_preambles++; /* demod_2400.c:339 */
        if (bestscore < 0) { /* demod_2400.c:342-348 */
            if (bestscore == -1)
                s->stats.demod_rejected_unknown_icao++;
            else
                s->stats.demod_rejected_bad++;
            continue;
        }

        int msglen = (bestmsg[0] & 0x80) ? 112 : 56; /* demod_2400.c:350 */
        mo_msg mm;
        memset(&mm, 0, sizeof (mm));
        mm.timestampMsg = sampleTimestamp + (uint64_t) j * 5 + (8 + 56) * 12 + (uint64_t) bestphase; /* :358 */
        mm.sysTimestampMsg = sysTimestamp + (mm.timestampMsg - sampleTimestamp) / 12000U; /* :361, util.c:79-81 */
        s->ifile_now = mm.sysTimestampMsg; /* :364-366 */
        mm.score = bestscore;

        int result = decode_crc_part(s, &mm, bestmsg); /* :372 */
        if (result < 0) {
            if (result == -1)
                s->stats.demod_rejected_unknown_icao++;
            else
                s->stats.demod_rejected_bad++;
            continue;
        }
        s->stats.demod_accepted[mm.correctedbits]++;
        s->stats.demod_bestPhase[bestphase - 4]++;

        /* demod_2400.c:387-408 */
        uint64_t scaled = 0;
        int signal_len = msglen * 12 / 5;
        for (int k = 0; k < signal_len; ++k) {
            uint32_t v = m[j + 19 + k];
            scaled += v * v;
        }
        double signal_power = scaled / 65535.0 / 65535.0;
        mm.signalLevel = signal_power / signal_len;
        s->stats.signal_power_sum += signal_power;
        s->stats.signal_power_count += (uint64_t) signal_len;
        sum_scaled_signal_power += scaled;
        if (mm.signalLevel > s->stats.peak_signal_power)
            s->stats.peak_signal_power = mm.signalLevel;
        if (mm.signalLevel > 0.50119)
            s->stats.strong_signal_count++;

        j += (uint32_t) (msglen * 12 / 5); /* :416 */

        /* useModesMessage(), mode_s.c:2146-2173 */
        s->stats.messages_total++;
        memset(mm.msg + mm.msgbits / 8, 0, 14 - mm.msgbits / 8);
        memset(mm.verbatim + mm.msgbits / 8, 0, 14 - mm.msgbits / 8);
        push_msg(out, &mm);
    }

    /* demod_2400.c:423-427 */
    double sum_signal_power = sum_scaled_signal_power / 65535.0 / 65535.0;
    s->stats.noise_power_sum += (mean_power * mlen - sum_signal_power);
    s->stats.noise_power_count += mlen;
}

/* ======================================================================================
 * Whole stream: ifileRun (sdr_ifile.c:164-237) + fifo_enqueue overlap (fifo.c:180-188) +
 * the main loop's per-block housekeeping (readsb.c:830-836, 331)
 * ==================================================================================== */

static double thread_cpu_s(void) {
    struct timespec ts;
    clock_gettime(CLOCK_THREAD_CPUTIME_ID, &ts);
    return ts.tv_sec + ts.tv_nsec * 1e-9;
}

/* ------------------------------------------------------------------------------------------
 * Mode A/C (demod_2400.c:522-708, mode_ac.c:168-203), only with mo_config.modeac
 *
 * A reply is 20 bit periods of 1.45 us (87 ticks of a 60 MHz clock; a sample is 25 ticks): framing
 * pulses F1 (bit 0) and F2 (bit 14), data pulses between and after, quiet zones at bits 7, 15, 16,
 * 18, 19.  Whether a reply with F1 at data index f1 decodes is a pure function of the magnitudes
 * and of the block's noise level; the only sequential part is the skip over an accepted reply.
 * ------------------------------------------------------------------------------------------ */

/* demod_2400.c:529-530 */
unsigned mo_modeac_noise_level(double mean_level, double mean_power) {
    double noise_stddev = sqrt(mean_power - mean_level * mean_level);
    return (unsigned) ((mean_power + noise_stddev) * 65535 + 0.5);
}

/* a framing pulse at sample s: rising edge, quiet third sample, 6 dB above noise (:577-588, :604-614) */
static int ac_framing_pulse(const uint16_t *m, uint32_t s, unsigned noise_level, unsigned *level) {
    if (!(m[s - 1] < m[s]))
        return 0;
    if (m[s + 2] > m[s] || m[s + 2] > m[s + 1])
        return 0;
    *level = (unsigned) ((m[s] + m[s + 1]) / 2);
    return !(noise_level * 2 > *level);
}

/* Returns 1 and the F1 clock (60 MHz ticks from data[0]) and the Mode A code if a reply with F1 at
 * data index f1 >= 1 decodes (demod_2400.c:577-683). */
int mo_modeac_at(const uint16_t *m, uint32_t f1, unsigned noise_level, uint32_t *f1_clock_out, uint32_t *modeac_out) {
    unsigned f1_level, f2_level;
    if (!ac_framing_pulse(m, f1, noise_level, &f1_level))
        return 0;

    /* clock phase from the share of power in the second sample (:593-596) */
    float pa = (float) m[f1] * m[f1];
    float pb = (float) m[f1 + 1] * m[f1 + 1];
    float fraction = pb / (pa + pb);
    unsigned f1_clock = (unsigned) (25 * (f1 + fraction * fraction) + 0.5);

    unsigned f2_clock = f1_clock + 87 * 14; /* :600 */
    if (!ac_framing_pulse(m, f2_clock / 25, noise_level, &f2_level))
        return 0;

    unsigned top = f1_level > f2_level ? f1_level : f2_level;
    float midpoint = sqrtf(noise_level * top);                          /* :618 */
    unsigned signal_threshold = (unsigned) (midpoint * M_SQRT2 + 0.5);  /* +3 dB */
    unsigned noise_threshold = (unsigned) (midpoint / M_SQRT2 + 0.5);   /* -3 dB */

    unsigned bits = 0, bad = 0; /* bad = noisy or uncertain (:629-651, :664) */
    for (unsigned bit = 0, clock = f1_clock; bit < 20; ++bit, clock += 87) {
        const uint16_t *p = m + clock / 25;
        bits <<= 1;
        if (p[2] >= signal_threshold)
            bad = 1;
        if (p[0] >= signal_threshold || p[1] >= signal_threshold)
            bits |= 1;
        else if (p[0] > noise_threshold && p[1] > noise_threshold)
            bad = 1;
    }
    if ((bits & 0x80020) != 0x80020 || (bits & 0x0101B) != 0 || bad)
        return 0;

    /* 00 A4 A2 A1  00 B4 B2 B1  SPI C4 C2 C1  00 D4 D2 D1 (:670-683) */
    static const struct { unsigned from, to; } map[13] = {
        {0x40000, 0x0010}, {0x20000, 0x1000}, {0x10000, 0x0020}, {0x08000, 0x2000}, {0x04000, 0x0040},
        {0x02000, 0x4000}, {0x00800, 0x0100}, {0x00400, 0x0001}, {0x00200, 0x0200}, {0x00100, 0x0002},
        {0x00080, 0x0400}, {0x00040, 0x0004}, {0x00004, 0x0080},
    };
    unsigned modeac = 0;
    for (int i = 0; i < 13; ++i)
        if (bits & map[i].from)
            modeac |= map[i].to;
    *f1_clock_out = f1_clock;
    *modeac_out = modeac;
    return 1;
}

static void demod_block_ac(demod_state *s, const uint16_t *m, uint32_t mlen, uint64_t sampleTimestamp, uint64_t sysTimestamp,
                           double mean_level, double mean_power, msg_list *out) {
    unsigned noise_level = mo_modeac_noise_level(mean_level, mean_power);
    for (uint32_t f1 = 1; f1 < mlen; ++f1) {
        uint32_t f1_clock, modeac;
        if (!mo_modeac_at(m, f1, noise_level, &f1_clock, &modeac))
            continue;
        mo_msg mm;
        memset(&mm, 0, sizeof (mm));
        mm.timestampMsg = sampleTimestamp + (f1_clock + 87 * 14) / 5;                   /* :697, at F2 */
        mm.sysTimestampMsg = sysTimestamp + (mm.timestampMsg - sampleTimestamp) / 12000U; /* :700 */
        mm.msgtype = 32;                                                                 /* mode_ac.c:171 */
        mm.msgbits = 16;
        mm.msg[0] = mm.verbatim[0] = (uint8_t) (modeac >> 8);
        mm.msg[1] = mm.verbatim[1] = (uint8_t) modeac;
        mm.addr = (modeac & 0x0000FF7F) | (1u << 24);                                    /* mode_ac.c:180 */
        push_msg(out, &mm);
        s->stats.messages_total++; /* useModesMessage, mode_s.c:2149 */
        f1 += (20 * 87 / 25);      /* :707 */
    }
}

int mo_run(const mo_config *cfg, const void *iq, uint64_t nsamples, mo_result *res) {
    memset(res, 0, sizeof (*res));
    if (cfg->format < MO_UC8 || cfg->format > MO_SC16Q11 || cfg->block_samples == 0)
        return -1;
    crc_init();

    demod_state *s = calloc(1, sizeof (*s));
    s->cfg = *cfg;
    crc_tables_init(&s->crc, cfg->nfix);
    filter_init(&s->filter);

    const uint32_t block = cfg->block_samples;
    const unsigned bps = (cfg->format == MO_UC8) ? 2 : 4;
    uint16_t *data = calloc((size_t) block + MO_OVERLAP, sizeof (uint16_t));
    uint16_t carry[MO_OVERLAP];
    memset(carry, 0, sizeof (carry)); /* fifo.c:47 */

    mo_dc_state dc;
    mo_dc_init(&dc, 2400000.0); /* Modes.sample_rate, readsb.c:141 */

    msg_list list = {0};
    mo_block *blocks = NULL;
    uint64_t n_blocks = 0, cap_blocks = 0;
    uint64_t sampleCounter = 0;
    int eof = 0;

    while (!eof) {
        uint64_t sampleTimestamp = (uint64_t) (sampleCounter * 12e6 / 2400000.0); /* sdr_ifile.c:187 */
        uint64_t sysTimestamp = sampleTimestamp / 12000U; /* sdr_ifile.c:190, startup_time = 0 */

        uint64_t left = nsamples - sampleCounter;
        uint32_t n = (left < block) ? (uint32_t) left : block;
        if (n < block)
            eof = 1; /* a short (possibly empty) read ends the stream, sdr_ifile.c:196-209 */

        double mean_level, mean_power;
        double t0 = thread_cpu_s();
        if (cfg->dcfilter) /* init_converter(..., Modes.dc_filter, ...), sdr_ifile.c:151 */
            mo_convert_dc(cfg->format, (const uint8_t *) iq + sampleCounter * bps, n, &dc, data + MO_OVERLAP,
                          &mean_level, &mean_power);
        else if (cfg->format == MO_SC16Q11 && cfg->sc16q11_table_bits) /* converters_table order, convert.c:432-437 */
            mo_convert_sc16q11_table(cfg->sc16q11_table_bits, (const uint8_t *) iq + sampleCounter * bps, n, data + MO_OVERLAP,
                                     &mean_level, &mean_power);
        else
            mo_convert(cfg->format, (const uint8_t *) iq + sampleCounter * bps, n, data + MO_OVERLAP,
                       &mean_level, &mean_power);
        s->stats.convert_cpu_s += thread_cpu_s() - t0;

        memcpy(data, carry, sizeof (carry)); /* fifo.c:184 */
        memcpy(carry, data + n, sizeof (carry)); /* fifo.c:188: &data[validLength - overlap] */

        t0 = thread_cpu_s();
        demod_block(s, data, n, sampleTimestamp, sysTimestamp, mean_power, &list);
        if (cfg->modeac) /* readsb.c:831-833 */
            demod_block_ac(s, data, n, sampleTimestamp, sysTimestamp, mean_level, mean_power, &list);
        s->stats.demod_cpu_s += thread_cpu_s() - t0;
        s->stats.samples_processed += MO_OVERLAP + n; /* readsb.c:835 */

        filter_expire(&s->filter, s->ifile_now); /* readsb.c:331 */

        if (n_blocks == cap_blocks) {
            cap_blocks = cap_blocks ? cap_blocks * 2 : 1024;
            blocks = realloc(blocks, cap_blocks * sizeof (*blocks));
        }
        blocks[n_blocks].mean_level = mean_level;
        blocks[n_blocks].mean_power = mean_power;
        ++n_blocks;
        sampleCounter += n;
    }

    res->msgs = list.msgs;
    res->n_msgs = list.n;
    res->blocks = blocks;
    res->n_blocks = n_blocks;
    res->stats = s->stats;
    res->n_samples = sampleCounter;

    crc_tables_free(&s->crc);
    free(s);
    free(data);
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * Output writers: what the Beast and raw TCP services put on the wire for a
 * message.  Restated from modesSendBeastOutput (net_io.c:769-835) and modesSendRawOutput
 * (net_io.c:870-896); pinned against those functions themselves through oracle/ref_netfmt.c.
 * Both return the number of bytes the messages need; bytes past `cap` are not written.
 * ------------------------------------------------------------------------------------------ */

static size_t put_escaped(uint8_t *out, size_t cap, size_t at, uint8_t ch) { /* 0x1a is doubled */
    if (at < cap)
        out[at] = ch;
    ++at;
    if (ch == 0x1a) {
        if (at < cap)
            out[at] = ch;
        ++at;
    }
    return at;
}

size_t mo_format_beast(const mo_msg *msgs, uint64_t n, int net_verbatim, uint8_t *out, size_t cap) {
    size_t at = 0;
    for (uint64_t i = 0; i < n; ++i) {
        const mo_msg *mm = &msgs[i];
        const int len = mm->msgbits / 8;
        const uint8_t *msg = net_verbatim ? mm->verbatim : mm->msg;
        char type;
        if (len == 7) type = '2';
        else if (len == 14) type = '3';
        else if (len == 2) type = '1';
        else continue; /* net_io.c:781-789 */
        if (at < cap) out[at] = 0x1a;
        ++at;
        if (at < cap) out[at] = (uint8_t) type;
        ++at;
        for (int shift = 40; shift >= 0; shift -= 8) /* 12 MHz timestamp, big-endian */
            at = put_escaped(out, cap, at, (uint8_t) (mm->timestampMsg >> shift));
        int sig = (int) round(sqrt(mm->signalLevel) * 255); /* net_io.c:817-821 */
        if (mm->signalLevel > 0 && sig < 1)
            sig = 1;
        if (sig > 255)
            sig = 255;
        at = put_escaped(out, cap, at, (uint8_t) sig);
        for (int j = 0; j < len; ++j)
            at = put_escaped(out, cap, at, msg[j]);
    }
    return at;
}

size_t mo_format_raw(const mo_msg *msgs, uint64_t n, int net_verbatim, int mlat, char *out, size_t cap) {
    static const char hex[] = "0123456789ABCDEF";
    size_t at = 0;
    for (uint64_t i = 0; i < n; ++i) {
        const mo_msg *mm = &msgs[i];
        const int len = mm->msgbits / 8;
        const uint8_t *msg = net_verbatim ? mm->verbatim : mm->msg;
        char head[16];
        int nh = 1;
        head[0] = '*';
        if (mlat && mm->timestampMsg) /* net_io.c:879-884 */
            nh = snprintf(head, sizeof (head), "@%012llX", (unsigned long long) mm->timestampMsg);
        for (int j = 0; j < nh; ++j, ++at)
            if (at < cap) out[at] = head[j];
        for (int j = 0; j < len; ++j) {
            if (at < cap) out[at] = hex[msg[j] >> 4];
            ++at;
            if (at < cap) out[at] = hex[msg[j] & 15];
            ++at;
        }
        if (at < cap) out[at] = ';';
        ++at;
        if (at < cap) out[at] = '\n';
        ++at;
    }
    return at;
}

void mo_result_free(mo_result *res) {
    free(res->msgs);
    free(res->blocks);
    memset(res, 0, sizeof (*res));
}
