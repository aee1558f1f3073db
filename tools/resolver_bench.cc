// Development aid: replays dumped resolver inputs (B200_DUMP_SPAN=dir) through the host resolver and
// times it.  g++ -O2 -Iinclude -Ireadsb_protobuf_b200/csrc tools/resolver_bench.cc
//     readsb_protobuf_b200/csrc/resolver.cc readsb_protobuf_b200/csrc/host_tables.cc -lpthread -o /tmp/resolver_bench
// Reads the current dump layout (2) and the r01 one (1, with the dead list: the hidden counts are then derived here).
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "resolver.h"

using namespace b200;

struct Span {
    uint64_t hdr[12];
    std::vector<TileOut> tiles;
    std::vector<uint32_t> dead;
    std::vector<LivePos> live;
    std::vector<LiveRec> recs;
    std::vector<LiveHidden> hidden;
    std::vector<BlockDead> bd;
    std::vector<unsigned long long> su;
    std::vector<double> sf;
};

int main(int argc, char **argv) {
    std::vector<Span> spans(argc - 1);
    std::vector<bool> in_s(1u << 24, false); // layout 1: the device-side address set, rebuilt from the records
    for (int i = 1; i < argc; ++i) {
        Span &s = spans[i - 1];
        FILE *f = fopen(argv[i], "rb");
        if (!f || fread(s.hdr, sizeof(s.hdr), 1, f) != 1)
            return 1;
        if (s.hdr[10] == 2) {
            // current layout: live positions, records, hidden-dead counts, block dead counters, block sums
            s.live.resize(s.hdr[7]);
            s.recs.resize(s.hdr[8]);
            s.hidden.resize(s.hdr[7]);
            s.bd.resize(s.hdr[9]);
            s.su.resize(2 * s.hdr[9]);
            s.sf.resize(2 * s.hdr[9]);
            size_t ok = fread(s.live.data(), sizeof(LivePos), s.live.size(), f);
            ok += fread(s.recs.data(), sizeof(LiveRec), s.recs.size(), f);
            ok += fread(s.hidden.data(), sizeof(LiveHidden), s.hidden.size(), f);
            ok += fread(s.bd.data(), sizeof(BlockDead), s.bd.size(), f);
            ok += fread(s.su.data(), 8, s.su.size(), f);
            ok += fread(s.sf.data(), 8, s.sf.size(), f);
            fclose(f);
            (void) ok;
            continue;
        }
        // layout 1 (r01: the dead list travelled to the host): derive the hidden counts here, the way
        // live_gather_kernel does on the device
        s.tiles.resize(s.hdr[5]);
        s.dead.resize(s.hdr[6]);
        s.live.resize(s.hdr[7]);
        s.recs.resize(s.hdr[8]);
        s.bd.resize(s.hdr[9]);
        s.su.resize(2 * s.hdr[9]);
        s.sf.resize(2 * s.hdr[9]);
        size_t ok = fread(s.tiles.data(), sizeof(TileOut), s.tiles.size(), f);
        ok += fread(s.dead.data(), 4, s.dead.size(), f);
        ok += fread(s.live.data(), sizeof(LivePos), s.live.size(), f);
        ok += fread(s.recs.data(), sizeof(LiveRec), s.recs.size(), f);
        ok += fread(s.bd.data(), sizeof(BlockDead), s.bd.size(), f);
        ok += fread(s.su.data(), 8, s.su.size(), f);
        ok += fread(s.sf.data(), 8, s.sf.size(), f);
        fclose(f);
        (void) ok;
        if (s.hdr[10] != 1)
            return 2;
        s.hidden.resize(s.live.size());
        const uint64_t n = s.hdr[0], B = s.hdr[2];
        for (size_t i = 0; i < s.live.size(); ++i) {
            const uint64_t p = s.live[i].pos;
            const uint64_t b1 = std::min<uint64_t>(n, (p / B + 1) * B);
            const uint64_t end_s = std::min<uint64_t>(p + 134, b1 - 1), end_l = std::min<uint64_t>(p + 268, b1 - 1);
            uint64_t lo = 0, hi = 0;
            LiveHidden h = {0, 0, 0, 0};
            bool snap = false;
            const uint32_t t0 = (uint32_t) ((p + kPosShift) / kTile), t1 = (uint32_t) ((end_l + kPosShift) / kTile);
            for (uint32_t t = t0; t <= t1 && t < s.tiles.size() && end_l > p; ++t) {
                const TileOut &to = s.tiles[t];
                const int64_t base = (int64_t) t * kTile - kPosShift;
                const uint32_t *it = s.dead.data() + to.dead_off + (t == t0 ? s.live[i].dead_rank : 0), *dend = s.dead.data() + to.dead_off + to.ndead;
                for (; it != dend; ++it) {
                    const int64_t pl = (int64_t) (*it & 0x1fffu);
                    if (pl > (int64_t) end_l - base)
                        break;
                    if (!snap && pl > (int64_t) end_s - base) {
                        h.short_lo = lo;
                        h.short_hi = hi;
                        snap = true;
                    }
                    const uint32_t tm = (*it >> 13) & 31u, unk = (*it >> 18) & 1u;
                    lo += 1ull | ((uint64_t) (unk ^ 1u) << 16) | ((uint64_t) unk << 32) | ((uint64_t) (tm & 1u) << 48);
                    hi += (uint64_t) ((tm >> 1) & 1u) | ((uint64_t) ((tm >> 2) & 1u) << 16) | ((uint64_t) ((tm >> 3) & 1u) << 32) |
                          ((uint64_t) ((tm >> 4) & 1u) << 48);
                }
            }
            if (!snap) {
                h.short_lo = lo;
                h.short_hi = hi;
            }
            h.long_lo = lo;
            h.long_hi = hi;
            s.hidden[i] = h;
        }
        // LiveRec::w0 bit 31 (key in S) as K2 sets it since layout 2: every clean DF17 / DF11-IID0 address of the
        // stream so far (such a frame is always live, so its record is here)
        for (const LiveRec &r : s.recs) {
            const uint32_t kind = (r.w0 >> 24) & 7u;
            if ((r.w0 & 0xffffffu) == 0 && ((kind == kKindES && (r.msg[0] >> 3) == 17) || kind == kKindDF11))
                in_s[r.w1 & 0xffffffu] = true;
        }
        for (LiveRec &r : s.recs)
            if (in_s[r.w1 & 0xffffffu] && !getenv("RB_NO_S"))
                r.w0 |= 0x80000000u;
            else if (getenv("RB_NO_S"))
                r.w0 |= 0x80000000u;
    }
    CrcTables crc(1);
    Resolver res(&crc, 0);
    MessageList msgs;
    std::vector<b200_block_info> blocks;
    const int reps = getenv("RB_REPS") ? atoi(getenv("RB_REPS")) : 20;
    std::vector<double> times;
    for (int rep = 0; rep < reps; ++rep) {
        res.reset();
        msgs.clear();
        blocks.clear();
        auto t0 = std::chrono::steady_clock::now();
        for (Span &s : spans) {
            SpanView v;
            v.nsamples = s.hdr[0];
            v.first_sample = s.hdr[1];
            v.block_samples = (uint32_t) s.hdr[2];
            v.final_span = s.hdr[3];
            v.format = (uint32_t) s.hdr[4];
            v.live = s.live.data();
            v.n_live = (uint32_t) s.live.size();
            v.liverecs = s.recs.data();
            v.hidden = s.hidden.data();
            v.block_dead = s.bd.data();
            v.block_sums_u64 = s.su.data();
            v.block_sums_f64 = s.sf.data();
            res.resolve(v, msgs, blocks);
        }
        double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        times.push_back(ms);
        if (reps <= 50 || rep % 100 == 0)
            printf("rep %d: %.3f ms, %zu msgs\n", rep, ms, msgs.size());
    }
    std::sort(times.begin(), times.end());
    printf("min %.3f ms, median %.3f ms over %d reps\n", times.front(), times[times.size() / 2], reps);
    // digest of everything the resolver produced (to compare two builds)
    auto fnv = [](const void *p, size_t n, uint64_t h) {
        const unsigned char *b = (const unsigned char *) p;
        for (size_t i = 0; i < n; ++i)
            h = (h ^ b[i]) * 1099511628211ull;
        return h;
    };
    uint64_t h = 1469598103934665603ull;
    h = fnv(msgs.data(), msgs.size() * sizeof(b200_message), h);
    h = fnv(blocks.data(), blocks.size() * sizeof(b200_block_info), h);
    const b200_demod_stats st = res.stats();
    h = fnv(&st, sizeof(st), h);
    printf("digest %016llx  (%zu msgs, %zu blocks, mismatches %llu, runs walked twice in the last pass %llu)\n", (unsigned long long) h,
           msgs.size(), blocks.size(), (unsigned long long) res.gpu_host_mismatches(), (unsigned long long) res.respeculated_runs());
    return 0;
}
