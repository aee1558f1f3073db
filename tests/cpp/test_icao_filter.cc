// Property test of the rule the speculative resolver stands on (resolver.h, IcaoFilter::differs_only_unprobed /
// same_members): whenever the rule says "the run's result stands", a run replayed on the true state must get exactly
// the answers it got on the guessed state.  Random filter histories, random runs of test / add / expire.
//   g++ -std=c++17 -O1 -I include -I readsb_protobuf_b200/csrc tests/cpp/test_icao_filter.cc
//       readsb_protobuf_b200/csrc/resolver.cc readsb_protobuf_b200/csrc/host_tables.cc -lpthread
#include <stdio.h>
#include <stdlib.h>

#include <random>
#include <vector>

#include "resolver.h"

using b200::IcaoFilter;

struct Op {
    int kind; // 0 test, 1 add, 2 expire
    uint32_t addr;
    uint64_t now;
};

int main(int argc, char **argv) {
    const int trials = argc > 1 ? atoi(argv[1]) : 3000;
    std::mt19937_64 rng(12345);
    auto pick = [&](uint32_t n) { return (uint32_t) (rng() % n); };
    long kept_same = 0, kept_rule = 0, refused = 0, answers = 0;
    for (int t = 0; t < trials; ++t) {
        std::vector<uint32_t> pool(40 + pick(60));
        for (uint32_t &a : pool)
            a = pick(1u << 24);
        auto addr = [&]() { return pool[pick((uint32_t) pool.size())]; };
        // a common past ...
        IcaoFilter truth, guess;
        uint64_t now = 1000000 + pick(1000);
        for (int i = 0, n = (int) pick(200); i < n; ++i) {
            if (pick(10) == 0) {
                now += pick(40000);
                truth.expire(now);
                guess.expire(now);
            } else {
                const uint32_t a = addr();
                truth.add(a);
                guess.add(a);
            }
        }
        // ... then the two drift apart: inserts only one of them saw, now and then a flip only one of them took
        for (int i = 0, n = (int) pick(4); i < n; ++i)
            truth.add(addr());
        for (int i = 0, n = (int) pick(4); i < n; ++i)
            guess.add(addr());
        if (pick(6) == 0)
            truth.expire(now + 60000 + pick(5));
        if (pick(6) == 0)
            guess.expire(now + 60000 + pick(5));
        if (pick(3) == 0) { // re-inserts of members after a flip: tables change, answers do not
            for (int i = 0; i < 10; ++i) {
                const uint32_t a = addr();
                if (truth.test(a))
                    truth.add(a);
            }
        }
        const IcaoFilter::Snapshot s = guess.snapshot();
        // the run, walked on the guess with its probes recorded
        std::vector<Op> ops;
        uint64_t clock = now + pick(1000), last_now = 0;
        for (int i = 0, n = 20 + (int) pick(200); i < n; ++i) {
            const uint32_t k = pick(20);
            if (k == 0) {
                clock += pick(3) == 0 ? pick(70000) : pick(60);
                ops.push_back({2, 0, clock});
                last_now = clock;
            } else if (k < 6) {
                ops.push_back({1, addr(), 0});
            } else {
                ops.push_back({0, pick(4) == 0 ? pick(1u << 24) : addr(), 0});
            }
        }
        IcaoFilter walked;
        walked.track_probes(true);
        walked.load(s);
        std::vector<int> got;
        for (const Op &o : ops) {
            if (o.kind == 0)
                got.push_back(walked.test(o.addr));
            else if (o.kind == 1)
                walked.add(o.addr);
            else
                walked.expire(o.now);
        }
        const bool same = truth.same_members(s);
        const bool rule = !same && truth.differs_only_unprobed(s, walked, last_now);
        if (!same && !rule) {
            ++refused;
            continue;
        }
        (same ? kept_same : kept_rule)++;
        // the same run on the true state
        IcaoFilter replay;
        replay.load(truth.snapshot());
        size_t gi = 0;
        for (const Op &o : ops) {
            if (o.kind == 0) {
                const int want = replay.test(o.addr);
                if (want != got[gi]) {
                    fprintf(stderr, "trial %d: answer %zu differs (address %06x: guess %d, truth %d) although the rule kept the run (%s)\n", t, gi,
                            o.addr, got[gi], want, same ? "same members" : "differs only unprobed");
                    return 1;
                }
                ++gi;
                ++answers;
            } else if (o.kind == 1) {
                replay.add(o.addr);
            } else {
                replay.expire(o.now);
            }
        }
    }
    printf("%d trials: %ld kept as same members, %ld kept by the probe rule, %ld refused; %ld answers compared\n", trials, kept_same, kept_rule,
           refused, answers);
    // the test must not be vacuous
    return (kept_rule > trials / 50 && refused > trials / 50 && kept_same > 0) ? 0 : 2;
}
