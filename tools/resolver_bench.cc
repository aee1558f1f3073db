// Development aid: replays dumped resolver inputs (B200_DUMP_SPAN=dir) through the host resolver and
// times it.  g++ -O2 -Iinclude -Ireadsb_protobuf_b200/csrc tools/resolver_bench.cc
//     readsb_protobuf_b200/csrc/resolver.cc readsb_protobuf_b200/csrc/host_tables.cc -o /tmp/resolver_bench
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
    std::vector<BlockDead> bd;
    std::vector<unsigned long long> su;
    std::vector<double> sf;
};

int main(int argc, char **argv) {
    std::vector<Span> spans(argc - 1);
    for (int i = 1; i < argc; ++i) {
        Span &s = spans[i - 1];
        FILE *f = fopen(argv[i], "rb");
        if (!f || fread(s.hdr, sizeof(s.hdr), 1, f) != 1)
            return 1;
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
        if (s.hdr[10] == 0) {
            // dump of the per-tile layout K2 writes: pack it in tile order, as order_live_kernel does
            std::vector<LivePos> live;
            std::vector<LiveRec> recs;
            for (const TileOut &to : s.tiles) {
                for (uint32_t i = 0; i < to.nlive; ++i) {
                    LivePos lp = s.live[to.live_off + i];
                    lp.pad = (uint32_t) recs.size() + (lp.info >> 16);
                    live.push_back(lp);
                }
                for (uint32_t i = 0; i < to.nliverec; ++i)
                    recs.push_back(s.recs[to.liverec_off + i]);
            }
            s.live.swap(live);
            s.recs.swap(recs);
        }
    }
    CrcTables crc(1);
    Resolver res(&crc, 0);
    std::vector<b200_message> msgs;
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
            v.ntiles = (uint32_t) s.hdr[5];
            v.tiles = s.tiles.data();
            v.dead = s.dead.data();
            v.live = s.live.data();
            v.n_live = (uint32_t) s.live.size();
            v.liverecs = s.recs.data();
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
    printf("digest %016llx  (%zu msgs, %zu blocks, mismatches %llu)\n", (unsigned long long) h, msgs.size(), blocks.size(),
           (unsigned long long) res.gpu_host_mismatches());
    return 0;
}
