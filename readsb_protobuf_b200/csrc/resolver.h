// resolver.h -- the order-dependent tail of demodulate2400 on the host.
//
// The kernels decide everything that does not depend on what was decoded before.  What is left
// is the part of the reference loop whose outcome depends on earlier messages of the same
// stream: the ICAO filter (icao_filter.c), best-phase pick (demod_2400.c:218-228), decode-time
// rejects (mode_s.c:445-555), skip-ahead (demod_2400.c:416) and the statistics that follow
// from them.  It walks only the positions K2 marked live.
#pragma once

#include <stdint.h>

#include <memory>
#include <vector>

#include "device_types.h"
#include "host_tables.h"
#include "readsb_b200.h"

namespace b200 {

// icao_filter.c: two open-addressed tables, flipped every 60 s of stream time
class IcaoFilter {
  public:
    IcaoFilter() { reset(); }
    void reset();
    void add(uint32_t addr);          // icaoFilterAdd, icao_filter.c:73-97
    bool test(uint32_t addr) const;   // icaoFilterTest, icao_filter.c:99-122
    void expire(uint64_t now_ms);     // icaoFilterExpire, icao_filter.c:150-164
    // every address currently stored (for seeding the device-side address set)
    void collect(std::vector<uint32_t> &out) const;

    // The filter as a value: the addresses each table received, in insertion order, which table is active and
    // when it flips next.  Replaying the insertions into empty tables rebuilds them slot for slot (a table only
    // receives addresses while it is the active one, from empty), which is how a copy is made (load) and how two
    // filters are compared.  Valid while no insert was ever dropped by a full table (replayable()).
    struct Snapshot {
        std::vector<uint32_t> seq_a, seq_b;
        bool a_active = true;
        uint64_t next_flip = 0;
        bool operator==(const Snapshot &o) const {
            return a_active == o.a_active && next_flip == o.next_flip && seq_a == o.seq_a && seq_b == o.seq_b;
        }
    };
    Snapshot snapshot() const;
    void load(const Snapshot &s);
    // same active table, same flip time, and the same SET of addresses in each table (whatever the insertion order):
    // a filter loaded from `s` then answers every test, and takes every expiry, exactly as this one
    bool same_members(const Snapshot &s) const;
    // Probe tracking (for a filter a run of mag_bufs is walked on speculatively): remember every address test() was
    // asked about since the last load().  differs_only_unprobed(s, walked, last_now): the answers `walked` got on its
    // way from state `s` are the answers it would have got from this filter's (the true) state, because
    //  * the two differ only in addresses the run never asked about (the difference between two filters only shrinks
    //    while the same inserts and flips are applied to both), and
    //  * either they flip at the same times, or the run's clock never reached a flip of either (last_now = the
    //    latest filter clock of the run) -- then test() only ever saw the union of the two tables, and which table
    //    is active or when it flips next had no say.
    // The run's result then stands although its start state was predicted wrong.
    void track_probes(bool on);
    bool differs_only_unprobed(const Snapshot &s, const IcaoFilter &walked, uint64_t last_now) const;
    // inserts since the last call that made test() change its answer (the address was in neither table)
    uint32_t take_new_members() {
        const uint32_t n = new_members_;
        new_members_ = 0;
        return n;
    }
    bool replayable() const { return !dropped_ && list_a_.size() < kReplayMax && list_b_.size() < kReplayMax; }

  private:
    static constexpr uint32_t kSize = 8192, kEmpty = 0xffffffffu;
    static constexpr size_t kReplayMax = 2048; // well inside what a table can take without ever filling up
    static uint32_t hash(uint32_t a);
    static bool probe(const uint32_t *t, uint32_t addr);
    bool test_tables(uint32_t addr) const;
    uint32_t a_[kSize], b_[kSize];
    uint32_t *active_;
    uint64_t next_flip_;
    // Shadow of the two tables for O(1) membership: one bit per 24-bit address and table, plus the
    // list of addresses each table holds (to clear its bits at a flip).  Exact as long as no insert
    // was ever dropped by a full table; after that the tables themselves answer.
    std::vector<uint64_t> bits_a_, bits_b_;
    std::vector<uint32_t> list_a_, list_b_;
    bool dropped_;
    uint32_t new_members_ = 0;
    bool track_ = false;
    mutable std::vector<uint64_t> probed_;      // one bit per 24-bit address, allocated when tracking is first switched on
    mutable std::vector<uint64_t> scratch_;     // differs_only_unprobed's marks
    mutable std::vector<uint32_t> probed_list_; // the addresses whose bit is set (to clear them again)
    bool probed(uint32_t addr) const { return !probed_.empty() && ((probed_[addr >> 6] >> (addr & 63u)) & 1ull); }
};

// The messages of a process call: a growable array that never zero-fills.  (std::vector::resize would write every
// new element on the caller's core just before the workers fill them in from theirs.)
class MessageList {
  public:
    MessageList() = default;
    MessageList(const MessageList &) = delete;
    MessageList &operator=(const MessageList &) = delete;
    ~MessageList();
    size_t size() const { return n_; }
    bool empty() const { return n_ == 0; }
    const b200_message *data() const { return p_; }
    b200_message *data() { return p_; }
    const b200_message &operator[](size_t i) const { return p_[i]; }
    void clear() { n_ = 0; }
    b200_message *grow(size_t extra); // room for `extra` more; returns the first of them, contents unspecified

  private:
    b200_message *p_ = nullptr;
    size_t n_ = 0, cap_ = 0;
};

struct SpanView {
    uint64_t nsamples;        // new samples == scan positions of the span
    uint64_t first_sample;    // stream sample index of the span's first new sample
    uint32_t block_samples;
    bool final_span;
    uint32_t format;
    // live positions of the span in stream order and their records (order_live packed K2's per-tile
    // lists): live[i].pad = index of the position's first record in liverecs; hidden[i] = what a frame
    // accepted at live[i] hides from the block's dead totals
    const LivePos *live;
    uint32_t n_live;
    const LiveRec *liverecs;
    const LiveHidden *hidden;
    const BlockDead *block_dead;               // [nblocks]
    const unsigned long long *block_sums_u64;  // [nblocks][2]
    const double *block_sums_f64;              // [nblocks][2]
    // Mode A/C hits of the span, unordered; the resolver sorts them in place
    AcHit *ac_hits = nullptr;
    uint32_t n_ac_hits = 0;
};

class WorkerPool;

class Resolver {
  public:
    Resolver(const CrcTables *crc, uint64_t startup_time_ms);
    ~Resolver();
    void reset();
    // Appends the span's messages and block infos; updates the running statistics.
    void resolve(const SpanView &v, MessageList &msgs, std::vector<b200_block_info> &blocks);
    const b200_demod_stats &stats() const { return stats_; }
    const IcaoFilter &filter() const { return filter_; }
    uint64_t gpu_host_mismatches() const { return mismatches_; }
    uint64_t modeac_count() const { return modeac_; } // Modes.stats_current.demod_modeac

    // one accepted frame (or Mode A/C reply) of the sequential walk, materialised afterwards: 24 bytes
    struct Accepted {
        uint32_t index;  // live position (Mode S) or hit (Mode A/C) of the span
        uint32_t rec;    // absolute index of the winning record
        int32_t score;
        uint32_t block;  // mag_buf of the span
        uint8_t phase, modeac, long_frame, pad;
        uint32_t pad2;
    };

    uint64_t respeculated_runs() const { return respeculated_; } // runs of mag_bufs walked twice (a wrong filter prediction)

  private:
    struct WalkOut;
    struct Potential;
    struct Run;
    static constexpr uint32_t kParallelWalkMinLive = 2048; // below this a span is walked in one run
    static int score(const IcaoFilter &f, const LiveRec &r);
    static int admit(IcaoFilter &f, const LiveRec &r, uint32_t *added); // the filter-dependent part of decodeModesMessage
    void walk(const SpanView &v, IcaoFilter &f, uint64_t k0, uint64_t k1, const std::vector<b200_block_info> &blocks, size_t block_base,
              WalkOut &out, bool log_adds) const;
    void prescan(const SpanView &v, uint64_t k0, uint64_t k1, std::vector<Potential> &adds, std::vector<uint64_t> &now_guess) const;
    // the rest of it; returns CRC disagreements, *signal_power = the frame's signal power (demod_2400.c:397)
    uint32_t build(const SpanView &v, const Accepted &a, b200_message &mm, double *signal_power) const;
    const CrcTables *crc_;
    uint64_t startup_;
    IcaoFilter filter_;
    b200_demod_stats stats_;
    uint64_t ifile_now_;
    uint64_t mismatches_;
    uint64_t modeac_;
    std::vector<Accepted> accepted_;
    std::unique_ptr<double[]> signal_power_; // per accepted frame of the span; never zero-filled
    size_t signal_power_cap_ = 0;
    std::vector<std::unique_ptr<Run>> runs_; // one per run of mag_bufs walked side by side (runs_[0]: the whole span)
    IcaoFilter sim_;                          // plays the potential adds through to predict the runs' start states
    uint64_t respeculated_ = 0;
    struct Trace { // B200_RESOLVER_TRACE: where the time of resolve() goes, printed when the resolver is destroyed
        uint64_t spans = 0, parallel_spans = 0, runs = 0, rewalks = 0, absolved = 0;
        double ms[7] = {0, 0, 0, 0, 0, 0, 0};
    } trace_;
    uint32_t min_live_;
    bool optimistic_ = true;        // runs start from the true state in front of the round (B200_RESOLVER_PREDICT=prescan: the predicted one)
    // new filter members of the previous span: while the population moves (a stream's first seconds) the runs'
    // start states are predicted from a prescan; once it is stable they all start from the true state
    uint32_t recent_new_members_ = 0xffffffffu;
    uint32_t optimistic_max_new_ = 2;
    static constexpr int kMaxRounds = 3;
    bool absolve_ = true; // keep a run whose start state was predicted wrong only in addresses it never asked about
    static constexpr size_t kMinNotesPerSlice = 256; // messages a worker assembles at least, when one run's notes are shared out
    uint32_t min_live_per_run_ = 512; // a run shorter than this costs more in hand-over than it saves
    uint64_t min_blocks_per_run_;
    WorkerPool *pool_;
};

} // namespace b200
