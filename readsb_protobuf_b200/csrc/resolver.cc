// resolver.cc -- see resolver.h.
#include "resolver.h"

#include <algorithm>

#include <string.h>

namespace b200 {

// ------------------------------------------------------------------------------------------
// ICAO filter (icao_filter.c)
// ------------------------------------------------------------------------------------------

uint32_t IcaoFilter::hash(uint32_t a) {
    // icao_filter.c:44-65: Jenkins one-at-a-time over the three address bytes, low byte first
    uint32_t h = 0;
    for (int shift = 0; shift < 24; shift += 8) {
        h += (a >> shift) & 0xffu;
        h += h << 10;
        h ^= h >> 6;
    }
    h += h << 3;
    h ^= h >> 11;
    h += h << 15;
    return h & (kSize - 1);
}

void IcaoFilter::reset() {
    memset(a_, 0xff, sizeof(a_));
    memset(b_, 0xff, sizeof(b_));
    active_ = a_;
    next_flip_ = 0;
    if (bits_a_.empty()) {
        bits_a_.assign((1u << 24) / 64, 0);
        bits_b_.assign((1u << 24) / 64, 0);
    } else {
        for (uint32_t x : list_a_)
            bits_a_[x >> 6] = 0;
        for (uint32_t x : list_b_)
            bits_b_[x >> 6] = 0;
    }
    list_a_.clear();
    list_b_.clear();
    dropped_ = false;
}

void IcaoFilter::add(uint32_t addr) {
    // the address itself ...
    uint32_t h0 = hash(addr), h = h0;
    bool full = false;
    while (active_[h] != kEmpty && active_[h] != addr) {
        h = (h + 1) & (kSize - 1);
        if (h == h0) {
            full = true;
            break;
        }
    }
    if (full) {
        dropped_ = true;
        return; // icao_filter.c:78-81: a full table drops the address (and skips the second insert)
    }
    if (active_[h] == kEmpty) {
        active_[h] = addr;
        if (addr < (1u << 24)) {
            std::vector<uint64_t> &bits = (active_ == a_) ? bits_a_ : bits_b_;
            bits[addr >> 6] |= 1ull << (addr & 63u);
            ((active_ == a_) ? list_a_ : list_b_).push_back(addr);
        } else {
            dropped_ = true; // not a 24-bit address: let the tables answer
        }
    }
    // ... and once more on the chain of its low 16 bits (Data/Parity lookups, icao_filter.c:87-96)
    const uint32_t low = addr & 0x00ffffu;
    h0 = h = hash(low);
    while (active_[h] != kEmpty && (active_[h] & 0x00ffffu) != low) {
        h = (h + 1) & (kSize - 1);
        if (h == h0)
            return;
    }
    if (active_[h] == kEmpty)
        active_[h] = addr;
}

bool IcaoFilter::probe(const uint32_t *t, uint32_t addr) {
    uint32_t h0 = hash(addr), h = h0;
    while (t[h] != kEmpty && t[h] != addr) {
        h = (h + 1) & (kSize - 1);
        if (h == h0)
            break;
    }
    return t[h] == addr;
}

bool IcaoFilter::test(uint32_t addr) const {
    if (!dropped_ && addr < (1u << 24))
        return ((bits_a_[addr >> 6] | bits_b_[addr >> 6]) >> (addr & 63u)) & 1ull;
    return test_tables(addr);
}

bool IcaoFilter::test_tables(uint32_t addr) const {
    // icaoFilterTest probes table a, then table b, from the same hash (icao_filter.c:99-122)
    const uint32_t h0 = hash(addr);
    for (const uint32_t *t : {a_, b_}) {
        uint32_t h = h0;
        while (t[h] != kEmpty && t[h] != addr) {
            h = (h + 1) & (kSize - 1);
            if (h == h0)
                break;
        }
        if (t[h] == addr)
            return true;
    }
    return false;
}

void IcaoFilter::expire(uint64_t now_ms) {
    if (now_ms < next_flip_)
        return;
    uint32_t *other = (active_ == a_) ? b_ : a_;
    memset(other, 0xff, sizeof(a_));
    {
        std::vector<uint64_t> &bits = (other == a_) ? bits_a_ : bits_b_;
        std::vector<uint32_t> &list = (other == a_) ? list_a_ : list_b_;
        for (uint32_t x : list)
            bits[x >> 6] = 0;
        list.clear();
    }
    active_ = other;
    next_flip_ = now_ms + 60000; // MODES_ICAO_FILTER_TTL, icao_filter.c:30
}

void IcaoFilter::collect(std::vector<uint32_t> &out) const {
    for (uint32_t i = 0; i < kSize; ++i) {
        if (a_[i] != kEmpty)
            out.push_back(a_[i]);
        if (b_[i] != kEmpty)
            out.push_back(b_[i]);
    }
}

// ------------------------------------------------------------------------------------------
// scoring and the CRC-dependent part of decode
// ------------------------------------------------------------------------------------------

void Resolver::reset() {
    filter_.reset();
    memset(&stats_, 0, sizeof(stats_));
    ifile_now_ = 0;
    mismatches_ = 0;
    modeac_ = 0;
}

// scoreModesMessage (mode_s.c:311-409) for a frame K1 already classified
int Resolver::score(const LiveRec &r) const {
    const uint32_t crc = r.w0 & 0xffffffu, kind = (r.w0 >> 24) & 7u, errors = (r.w0 >> 28) & 3u;
    const uint32_t key = r.w1 & 0xffffffu;
    switch (kind) {
        case kKindAP: // mode_s.c:343
            return filter_.test(crc) ? 1000 : -1;
        case kKindAPCommB: // mode_s.c:393-403
            return filter_.test(crc) ? 1000 : -2;
        case kKindDF11: { // mode_s.c:364-374
            const bool known = filter_.test(key);
            if ((crc & 0x7fu) == 0)
                return (known ? 1600 : 750) / (int) (errors + 1);
            return known ? 1000 / (int) (errors + 1) : -1;
        }
        case kKindES: // mode_s.c:386-389
            return (filter_.test(key) ? 1800 : 1400) / (int) (errors + 1);
        default:
            return -2;
    }
}

static inline uint32_t aa_field(const uint8_t *msg) { // getbits(msg, 9, 32)
    return ((uint32_t) msg[1] << 16) | ((uint32_t) msg[2] << 8) | (uint32_t) msg[3];
}

// The part of decodeModesMessage that can reject the frame or touches the filter
// (mode_s.c:424-555, 560-562, 717-726).  CRC and repair are recomputed on the host from the
// sliced bytes; a disagreement with the kernel's values is counted, never hidden.
/* DF18: is the AA field something other than an ICAO address?  The extended-squitter decoder then
 * flags mm->addr with MODES_NON_ICAO_ADDRESS (1 << 24): by CF alone (mode_s.c:1379-1428), or for CF 2 / 3 /
 * 6 by the IMF bit of the ME field, whose position depends on the ME type (mode_s.c:806, 927, 966-968,
 * 1054, 1064, 1259, 1404-1406).  msg = the frame after CRC repair. */
static inline int me_bit(const uint8_t *me, int n) { /* 1-based, MSB first (getbit, mode_s.c) */
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

int Resolver::decode(const LiveRec &r, b200_message &mm) {
    memcpy(mm.msg, r.msg, 14);
    memcpy(mm.verbatim, r.msg, 14);
    uint8_t *msg = mm.msg;
    static const uint8_t zeros[7] = {0, 0, 0, 0, 0, 0, 0};
    if (!memcmp(msg, zeros, 7))
        return -2;
    mm.msgtype = msg[0] >> 3;
    mm.msgbits = (mm.msgtype & 0x10) ? 112 : 56;
    mm.crc = crc_->checksum(msg, mm.msgbits);
    mm.correctedbits = 0;
    mm.addr = 0;
    if (mm.crc != (r.w0 & 0xffffffu))
        ++mismatches_;
    uint32_t iid = 0;

    switch (mm.msgtype) {
        case 0: case 4: case 5: case 16:
        case 24: case 25: case 26: case 27: case 28: case 29: case 30: case 31:
            if (!filter_.test(mm.crc))
                return -1;
            mm.addr = mm.crc;
            break;
        case 11: {
            iid = mm.crc & 0x7fu;
            if (mm.crc & 0xffff80u) {
                const ErrorInfo *ei = crc_->diagnose(mm.crc & 0xffff80u, mm.msgbits);
                if (!ei || ei->errors > 1)
                    return -2;
                mm.correctedbits = (uint8_t) ei->errors;
                CrcTables::fix(msg, ei);
                if (!filter_.test(aa_field(msg)))
                    return -1;
            }
            break;
        }
        case 17: case 18: {
            if (mm.crc != 0) {
                const ErrorInfo *ei = crc_->diagnose(mm.crc, mm.msgbits);
                if (!ei)
                    return -2;
                const uint32_t addr1 = aa_field(msg);
                mm.correctedbits = (uint8_t) ei->errors;
                CrcTables::fix(msg, ei);
                const uint32_t addr2 = aa_field(msg);
                if (addr1 != addr2 && !filter_.test(addr2))
                    return -1;
            }
            break;
        }
        case 20: case 21:
            if (!filter_.test(mm.crc))
                return -1;
            mm.addr = mm.crc;
            break;
        default:
            return -2;
    }
    if (mm.msgtype == 11 || mm.msgtype == 17 || mm.msgtype == 18)
        mm.addr = aa_field(msg);
    if (((r.w0 >> 28) & 3u) != mm.correctedbits && mm.msgtype != 11)
        ++mismatches_;
    // the only place addresses enter the filter
    if (!mm.correctedbits && (mm.msgtype == 17 || (mm.msgtype == 11 && iid == 0)))
        filter_.add(mm.addr);
    // decodeExtendedSquitter (mode_s.c:1373-1428) runs later in decodeModesMessage and may flag the address
    if (mm.msgtype == 18 && df18_non_icao(mm.msg))
        mm.addr |= 1u << 24;
    return 0;
}

// ------------------------------------------------------------------------------------------
// the walk
// ------------------------------------------------------------------------------------------

namespace {

// what skip-ahead hides, as eight 16-bit counters in two words (a frame body hides at most 268
// positions): lo = preambles | bad << 16 | unknown << 32 | phase0 << 48, hi = phase1..phase4
struct DeadCount {
    uint64_t lo = 0, hi = 0;
    uint32_t preambles() const { return (uint32_t) (lo & 0xffff); }
    uint32_t bad() const { return (uint32_t) ((lo >> 16) & 0xffff); }
    uint32_t unknown() const { return (uint32_t) ((lo >> 32) & 0xffff); }
    uint32_t phase(int k) const { return (uint32_t) (k == 0 ? (lo >> 48) : (hi >> (16 * (k - 1)))) & 0xffff; }
};

struct DeadLut {
    uint64_t lo[64], hi[64]; // indexed by trymask | unknown << 5
    DeadLut() {
        for (uint32_t i = 0; i < 64; ++i) {
            const uint32_t tm = i & 31u, unk = i >> 5;
            lo[i] = 1ull | ((uint64_t) (unk ? 0 : 1) << 16) | ((uint64_t) unk << 32) | ((uint64_t) (tm & 1u) << 48);
            hi[i] = (uint64_t) ((tm >> 1) & 1u) | ((uint64_t) ((tm >> 2) & 1u) << 16) | ((uint64_t) ((tm >> 3) & 1u) << 32) |
                    ((uint64_t) ((tm >> 4) & 1u) << 48);
        }
    }
};
const DeadLut kDeadLut;

// dead positions in (lo, hi] that a skip-ahead hides (they are in the per-block totals K2 made);
// `rank` = dead entries of lo's tile in front of lo (LivePos::dead_rank)
struct HiddenTotals {
    uint32_t preambles = 0, bad = 0, unknown = 0, phase[5] = {0, 0, 0, 0, 0};
};

void count_dead(const SpanView &v, uint64_t lo, uint64_t hi, uint32_t rank, HiddenTotals &total) {
    if (hi <= lo)
        return;
    // tile t covers positions [t*kTile - kPosShift, (t+1)*kTile - kPosShift)
    const uint32_t t0 = (uint32_t) ((lo + kPosShift) / kTile), t1 = (uint32_t) ((hi + kPosShift) / kTile);
    DeadCount dc; // one frame body: at most 268 entries, the 16-bit fields cannot overflow
    for (uint32_t t = t0; t <= t1 && t < v.ntiles; ++t) {
        const TileOut &to = v.tiles[t];
        const uint32_t *d = v.dead + to.dead_off, *dend = d + to.ndead;
        const int64_t base = (int64_t) t * kTile - kPosShift;
        const uint32_t *it = (t == t0) ? d + rank : d;
        const int64_t last = (int64_t) hi - base; // last tile-local index counted
        for (; it != dend && (int64_t) (*it & 0x1fffu) <= last; ++it) {
            const uint32_t key = (*it >> 13) & 63u; // trymask | unknown << 5
            dc.lo += kDeadLut.lo[key];
            dc.hi += kDeadLut.hi[key];
        }
    }
    total.preambles += dc.preambles();
    total.bad += dc.bad();
    total.unknown += dc.unknown();
    for (int k = 0; k < 5; ++k)
        total.phase[k] += dc.phase(k);
}

} // namespace

void Resolver::resolve(const SpanView &v, std::vector<b200_message> &msgs, std::vector<b200_block_info> &blocks) {
    const uint64_t n = v.nsamples, B = v.block_samples;
    // ifileRun: full blocks, then (at end of stream) one short block, which is empty when the stream
    // length is a multiple of the block size (sdr_ifile.c:192-216)
    const uint64_t nfull = n / B;
    const uint64_t nblocks = nfull + (v.final_span ? 1 : 0);

    // room for every live position of the span: no reallocation (and no first-touch page faults after the
    // first span of this size) inside the walk
    msgs.reserve(msgs.size() + v.n_live + v.n_ac_hits + 64);
    skips_.clear();
    skips_.reserve((size_t) v.n_live + 64);
    // what skip-ahead hides is un-counted after the walk: the dead list it needs may still be arriving
    std::vector<Skip> &skips = skips_;
    // Mode A/C hits in stream order (the kernel appends them as it finds them)
    if (v.n_ac_hits > 1)
        std::sort(v.ac_hits, v.ac_hits + v.n_ac_hits, [](const AcHit &a, const AcHit &b) { return a.q < b.q; });
    uint32_t ac_i = 0;
    // The live positions and their records are two flat arrays in stream order, just written by the GPU: the
    // walk is sequential in both, which the hardware prefetcher follows.
    uint32_t live_i = 0;
    const uint32_t n_live = v.n_live;

    for (uint64_t k = 0; k < nblocks; ++k) {
        const uint64_t b0 = k * B, b1 = std::min(n, b0 + B), nk = b1 - b0;
        const uint64_t sample_counter = v.first_sample + b0;
        // sdr_ifile.c:187-190
        const uint64_t sampleTimestamp = (uint64_t) ((double) sample_counter * 12e6 / 2400000.0);
        const uint64_t sysTimestamp = sampleTimestamp / 12000U + startup_;

        // converter outputs of the block (convert.c:104-110 / 246-252)
        b200_block_info bi;
        if (v.format == B200_INPUT_UC8) {
            const unsigned long long sl = v.block_sums_u64[2 * k], sp = v.block_sums_u64[2 * k + 1];
            bi.mean_level = sl / 65536.0 / (unsigned) nk; // sic: 65536
            bi.mean_power = sp / 65535.0 / 65535.0 / (unsigned) nk;
        } else {
            bi.mean_level = (double) ((float) v.block_sums_f64[2 * k] / (float) (unsigned) nk);
            bi.mean_power = (double) ((float) v.block_sums_f64[2 * k + 1] / (float) (unsigned) nk);
        }
        blocks.push_back(bi);

        ifile_now_ = sysTimestamp; // demod_2400.c:253-255
        uint64_t sum_scaled_signal_power = 0;
        bool skipping = false;
        uint64_t skip_until = 0; // positions <= skip_until are skipped while `skipping`

        while (live_i < n_live && v.live[live_i].pos < b1) {
            const LivePos *lp = &v.live[live_i];
            const uint64_t p = lp->pos;
            ++live_i;
            if (skipping && p <= skip_until)
                continue;
            const uint32_t trymask = lp->info & 31u, nrec = (lp->info >> 8) & 7u;
            const LiveRec *recs = v.liverecs + lp->pad;

            // score_phase for every tried phase, in order (demod_2400.c:183-229, 306-330): every tried phase
            // counts; one without a record scores -2, which only ever wins as the very first phase tried
            // (a later phase must score strictly higher, and nothing scores below -2); the records are in
            // phase order, so walking them alone gives the same pick as walking the five phases
            for (int q = 0; q < 5; ++q)
                stats_.demod_preamblePhase[q] += (trymask >> q) & 1u;
            int bestscore = -42, bestphase = -1;
            const LiveRec *best = nullptr;
            if (trymask) {
                const int first = 4 + __builtin_ctz(trymask);
                if (nrec == 0 || (int) ((recs[0].w1 >> 24) & 15u) != first) {
                    bestscore = -2;
                    bestphase = first;
                }
                for (uint32_t ri = 0; ri < nrec; ++ri) {
                    const int sc = score(recs[ri]);
                    if (sc > bestscore) {
                        bestscore = sc;
                        bestphase = (int) ((recs[ri].w1 >> 24) & 15u);
                        best = &recs[ri];
                    }
                }
            }
            stats_.demod_preambles++; // demod_2400.c:339
            if (bestscore < 0) {      // demod_2400.c:342-348
                if (bestscore == -1)
                    stats_.demod_rejected_unknown_icao++;
                else
                    stats_.demod_rejected_bad++;
                continue;
            }

            msgs.emplace_back(); // built in place (room was reserved); taken back if the decoder rejects it
            b200_message &mm = msgs.back();
            memset(&mm, 0, sizeof(mm));
            const uint64_t j = p - b0;
            mm.timestampMsg = sampleTimestamp + j * 5 + (8 + 56) * 12 + (uint64_t) bestphase; // demod_2400.c:358
            mm.sysTimestampMsg = sysTimestamp + (mm.timestampMsg - sampleTimestamp) / 12000U; // :361
            ifile_now_ = mm.sysTimestampMsg;                                                   // :364-366
            mm.score = bestscore;
            mm.bestphase = (uint8_t) bestphase;

            const int result = decode(*best, mm); // demod_2400.c:372
            if (result < 0) {
                if (result == -1)
                    stats_.demod_rejected_unknown_icao++;
                else
                    stats_.demod_rejected_bad++;
                msgs.pop_back();
                continue;
            }
            stats_.demod_accepted[mm.correctedbits]++;
            stats_.demod_bestPhase[bestphase - 4]++;

            // demod_2400.c:387-408
            const int msglen = (best->msg[0] & 0x80) ? 112 : 56; // :350, from the uncorrected DF
            const int signal_len = msglen * 12 / 5;
            const uint64_t scaled = best->power;
            const double signal_power = scaled / 65535.0 / 65535.0;
            mm.signalLevel = signal_power / signal_len;
            stats_.signal_power_sum += signal_power;
            stats_.signal_power_count += (uint64_t) signal_len;
            sum_scaled_signal_power += scaled;
            if (mm.signalLevel > stats_.peak_signal_power)
                stats_.peak_signal_power = mm.signalLevel;
            if (mm.signalLevel > 0.50119)
                stats_.strong_signal_count++;

            // demod_2400.c:416: skip the frame body; the for loop ends at the block boundary
            skipping = true;
            skip_until = std::min<uint64_t>(p + (uint64_t) signal_len, b1 - 1);
            skips.push_back({p, skip_until, lp->dead_rank});

            stats_.messages_total++; // useModesMessage, mode_s.c:2149
            memset(mm.msg + mm.msgbits / 8, 0, 14 - mm.msgbits / 8);
            memset(mm.verbatim + mm.msgbits / 8, 0, 14 - mm.msgbits / 8);
        }

        // demodulate2400AC (readsb.c:831-833, demod_2400.c:522-708) runs after demodulate2400 on the same
        // mag_buf: the block's hits in order, skipping 69 samples past every reply that is taken (:707)
        {
            uint64_t next_ok = 0; // first data index that may be examined again
            for (; ac_i < v.n_ac_hits && (uint64_t) v.ac_hits[ac_i].q < (k + 1) * B; ++ac_i) {
                const AcHit h = v.ac_hits[ac_i];
                const uint64_t f1 = (uint64_t) h.q - k * B;
                if (f1 < next_ok)
                    continue;
                b200_message mm;
                memset(&mm, 0, sizeof(mm));
                mm.timestampMsg = sampleTimestamp + (h.f1_clock + 87 * 14) / 5;                           // :697, at F2
                mm.sysTimestampMsg = sysTimestamp + (mm.timestampMsg - sampleTimestamp) / 12000U; // :700
                mm.msgtype = 32;                                                                   // mode_ac.c:171
                mm.msgbits = 16;
                mm.msg[0] = mm.verbatim[0] = (uint8_t) (h.modeac >> 8);
                mm.msg[1] = mm.verbatim[1] = (uint8_t) h.modeac;
                mm.addr = (h.modeac & 0x0000FF7Fu) | (1u << 24);                                        // mode_ac.c:180
                msgs.push_back(mm);
                stats_.messages_total++; // useModesMessage, mode_s.c:2149
                ++modeac_;
                next_ok = f1 + 70; // f1_sample += 69, then the loop's ++
            }
        }

        // positions no message can come from: K2's per-block totals minus what skip-ahead hid
        if (nk) {
            const BlockDead &bd = v.block_dead[k];
            stats_.demod_preambles += bd.preambles;
            stats_.demod_rejected_bad += bd.rejected_bad;
            stats_.demod_rejected_unknown_icao += bd.rejected_unknown;
            for (int q = 0; q < 5; ++q)
                stats_.demod_preamblePhase[q] += bd.phase[q];
        }

        // demod_2400.c:423-427
        const double sum_signal_power = sum_scaled_signal_power / 65535.0 / 65535.0;
        stats_.noise_power_sum += (bi.mean_power * (uint32_t) nk - sum_signal_power);
        stats_.noise_power_count += nk;
        stats_.samples_processed += kOverlap + nk; // readsb.c:835

        filter_.expire(ifile_now_); // readsb.c:331
    }

    // dead positions hidden by skip-ahead were counted in K2's per-block totals: take them out again
    // (the counters are sums, so the order does not matter)
    if (v.dead_ready)
        v.dead_ready(v.dead_ctx); // also: the caller reuses the buffers once we return
    HiddenTotals hidden;
    auto dead_line = [&](const Skip &sk) { // where the count of a skip starts reading
        const uint32_t t = (uint32_t) ((sk.lo + kPosShift) / kTile);
        return v.dead + v.tiles[t].dead_off + sk.rank;
    };
    constexpr size_t kAhead = 8;
    for (size_t i = 0; i < skips.size(); ++i) {
        if (i + kAhead < skips.size())
            __builtin_prefetch(dead_line(skips[i + kAhead]));
        count_dead(v, skips[i].lo, skips[i].hi, skips[i].rank, hidden);
    }
    stats_.demod_preambles -= hidden.preambles;
    stats_.demod_rejected_bad -= hidden.bad;
    stats_.demod_rejected_unknown_icao -= hidden.unknown;
    for (int q = 0; q < 5; ++q)
        stats_.demod_preamblePhase[q] -= hidden.phase[q];
}

} // namespace b200
