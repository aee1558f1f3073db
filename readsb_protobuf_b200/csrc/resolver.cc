// resolver.cc -- see resolver.h.
#include "resolver.h"

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <new>
#include <thread>

#include <chrono>

#include <sched.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

namespace b200 {

MessageList::~MessageList() {
    free(p_);
}

b200_message *MessageList::grow(size_t extra) {
    if (n_ + extra > cap_) {
        const size_t cap = std::max(n_ + extra, cap_ + cap_ / 2 + 1024);
        b200_message *p = static_cast<b200_message *>(realloc(p_, cap * sizeof(b200_message)));
        if (!p)
            throw std::bad_alloc();
        p_ = p;
        cap_ = cap;
    }
    b200_message *first = p_ + n_;
    n_ += extra;
    return first;
}

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
    // Already in the active table (the common case: the same aircraft again): both inserts below would find
    // their entries and change nothing.  The shadow bitmap knows without walking the probe chains.
    if (!dropped_ && addr < (1u << 24)) {
        const std::vector<uint64_t> &bits = (active_ == a_) ? bits_a_ : bits_b_;
        if ((bits[addr >> 6] >> (addr & 63u)) & 1ull)
            return;
    }
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
            if (!(((bits_a_[addr >> 6] | bits_b_[addr >> 6]) >> (addr & 63u)) & 1ull))
                ++new_members_; // in neither table before: test(addr) changes its answer
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
    if (!dropped_ && addr < (1u << 24)) {
        if (track_) {
            uint64_t &w = probed_[addr >> 6];
            const uint64_t bit = 1ull << (addr & 63u);
            if (!(w & bit)) {
                w |= bit;
                probed_list_.push_back(addr);
            }
        }
        return ((bits_a_[addr >> 6] | bits_b_[addr >> 6]) >> (addr & 63u)) & 1ull;
    }
    return test_tables(addr);
}

void IcaoFilter::track_probes(bool on) {
    for (uint32_t x : probed_list_)
        probed_[x >> 6] = 0;
    probed_list_.clear();
    if (on && probed_.empty())
        probed_.assign((1u << 24) / 64, 0);
    track_ = on;
}

bool IcaoFilter::differs_only_unprobed(const Snapshot &s, const IcaoFilter &walked, uint64_t last_now) const {
    if (dropped_ || walked.dropped_ || !walked.track_)
        return false;
    if (scratch_.empty())
        scratch_.assign((1u << 24) / 64, 0);
    auto bit = [](const std::vector<uint64_t> &b, uint32_t x) { return (b[x >> 6] >> (x & 63u)) & 1ull; };
    // Can a table flip while the run is walked, on either clock?  If not, test() only ever sees the union of the two
    // tables, and neither which of them is active nor when it flips next has any say in the run.
    const bool may_flip = last_now >= std::min(s.next_flip, next_flip_);
    if (may_flip && (s.a_active != (active_ == a_) || s.next_flip != next_flip_))
        return false;
    bool ok = true;
    for (int pass = 0; pass < (may_flip ? 2 : 1) && ok; ++pass) {
        // pass 0: table a (or the union), pass 1: table b
        const std::vector<uint32_t> *theirs[2] = {may_flip ? (pass ? &s.seq_b : &s.seq_a) : &s.seq_a, may_flip ? nullptr : &s.seq_b};
        const std::vector<uint32_t> *mine[2] = {may_flip ? (pass ? &list_b_ : &list_a_) : &list_a_, may_flip ? nullptr : &list_b_};
        // in s but not here
        for (const std::vector<uint32_t> *l : theirs)
            if (l)
                for (uint32_t x : *l) {
                    const bool here = may_flip ? bit(pass ? bits_b_ : bits_a_, x) : (bit(bits_a_, x) || bit(bits_b_, x));
                    if (!here && walked.probed(x))
                        ok = false;
                }
        // here but not in s
        for (const std::vector<uint32_t> *l : theirs)
            if (l)
                for (uint32_t x : *l)
                    scratch_[x >> 6] |= 1ull << (x & 63u);
        for (const std::vector<uint32_t> *l : mine)
            if (l && ok)
                for (uint32_t x : *l)
                    if (walked.probed(x) && !bit(scratch_, x)) {
                        ok = false;
                        break;
                    }
        for (const std::vector<uint32_t> *l : theirs)
            if (l)
                for (uint32_t x : *l)
                    scratch_[x >> 6] = 0;
    }
    return ok;
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

IcaoFilter::Snapshot IcaoFilter::snapshot() const {
    Snapshot s;
    s.seq_a = list_a_;
    s.seq_b = list_b_;
    s.a_active = active_ == a_;
    s.next_flip = next_flip_;
    return s;
}

void IcaoFilter::load(const Snapshot &s) {
    reset();
    const bool tracking = track_;
    track_probes(false); // forget the last walk's probes; the inserts below are not probes
    active_ = a_;
    for (uint32_t x : s.seq_a)
        add(x);
    active_ = b_;
    for (uint32_t x : s.seq_b)
        add(x);
    active_ = s.a_active ? a_ : b_;
    next_flip_ = s.next_flip;
    track_ = tracking;
}

bool IcaoFilter::same_members(const Snapshot &s) const {
    if (s.a_active != (active_ == a_) || s.next_flip != next_flip_ || s.seq_a.size() != list_a_.size() || s.seq_b.size() != list_b_.size() ||
        dropped_)
        return false;
    for (uint32_t x : s.seq_a)
        if (!((bits_a_[x >> 6] >> (x & 63u)) & 1ull))
            return false;
    for (uint32_t x : s.seq_b)
        if (!((bits_b_[x >> 6] >> (x & 63u)) & 1ull))
            return false;
    return true; // equal sizes, no duplicates within a sequence: the sets are equal
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
// worker pool: the message-assembly half of the resolve runs on a few host threads
// ------------------------------------------------------------------------------------------

// A handful of threads parked on a condition variable; run() hands them (and the caller) slices of an index
// range.  Only spans with thousands of accepted frames go through it: a sparse chunk is assembled inline.
static inline void spin_pause() {
#if defined(__x86_64__) || defined(__i386__)
    __builtin_ia32_pause();
#elif defined(__aarch64__)
    asm volatile("yield");
#endif
}

class WorkerPool {
  public:
    explicit WorkerPool(int nthreads) {
        for (int i = 1; i < nthreads; ++i)
            threads_.emplace_back([this, i] { loop(i); });
    }
    ~WorkerPool() {
        {
            std::lock_guard<std::mutex> g(m_);
            stop_ = true;
            generation_.fetch_add(1, std::memory_order_release);
        }
        cv_.notify_all();
        for (std::thread &t : threads_)
            t.join();
    }
    int size() const { return (int) threads_.size() + 1; }
    // fn(worker, begin, end) over [0, n) in slices of `grain`; returns when every slice is done.
    // fixed: item i always goes to worker i % size() (grain 1) -- a run of mag_bufs is then scanned, walked and
    // assembled by the same thread, whose cache holds its records.
    // The phases of one resolve() follow each other within microseconds and the resolves of a stream within a few
    // hundred, so a worker that has finished its share keeps polling the generation counter for kSpinUs before it
    // goes to sleep on the condition variable (a futex wake-up costs 30-60 us, as much as a whole phase of a sparse
    // chunk); the caller polls for completion, it has nothing else to do.
    void run(size_t n, size_t grain, const std::function<void(int, size_t, size_t)> &fn, bool fixed = false) {
        fn_ = &fn;
        n_ = n;
        grain_ = grain;
        fixed_ = fixed;
        next_.store(0, std::memory_order_relaxed);
        pending_.store((int) threads_.size(), std::memory_order_relaxed);
        bool wake;
        {
            std::lock_guard<std::mutex> g(m_); // a worker between its last poll and its sleep sees the new generation
            generation_.fetch_add(1, std::memory_order_release);
            wake = sleepers_ > 0;
        }
        if (wake)
            cv_.notify_all();
        work(0);
        for (uint32_t spins = 0; pending_.load(std::memory_order_acquire) != 0; ++spins) {
            if (spins < 4096)
                cpu_relax();
            else
                std::this_thread::yield(); // fewer cores than threads: let the workers run
        }
        fn_ = nullptr;
    }

  private:
    static constexpr double kSpinUs = 250.0;
    static void cpu_relax() {
#if defined(__x86_64__) || defined(__i386__)
        __builtin_ia32_pause();
#elif defined(__aarch64__)
        asm volatile("yield");
#endif
    }
    void work(int worker) {
        if (fixed_) {
            for (size_t i = (size_t) worker; i < n_; i += (size_t) size())
                (*fn_)(worker, i, i + 1);
            return;
        }
        for (;;) {
            const size_t b = next_.fetch_add(grain_, std::memory_order_relaxed);
            if (b >= n_)
                break;
            (*fn_)(worker, b, std::min(n_, b + grain_));
        }
    }
    void loop(int worker) {
        uint64_t seen = 0;
        for (;;) {
            // poll first ...
            const auto t0 = std::chrono::steady_clock::now();
            bool got = false;
            for (uint32_t spins = 0;; ++spins) {
                if (generation_.load(std::memory_order_acquire) != seen) {
                    got = true;
                    break;
                }
                cpu_relax();
                if ((spins & 63u) == 63u &&
                    std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count() > kSpinUs)
                    break;
            }
            if (!got) { // ... then sleep
                std::unique_lock<std::mutex> g(m_);
                ++sleepers_;
                cv_.wait(g, [&] { return generation_.load(std::memory_order_acquire) != seen; });
                --sleepers_;
            }
            seen = generation_.load(std::memory_order_acquire);
            if (stop_)
                return;
            work(worker);
            pending_.fetch_sub(1, std::memory_order_release);
        }
    }
    std::vector<std::thread> threads_;
    std::mutex m_;
    std::condition_variable cv_;
    const std::function<void(int, size_t, size_t)> *fn_ = nullptr;
    size_t n_ = 0, grain_ = 1;
    bool fixed_ = false;
    std::atomic<size_t> next_{0};
    std::atomic<int> pending_{0};
    std::atomic<uint64_t> generation_{0};
    int sleepers_ = 0;
    bool stop_ = false;
};

static int resolver_threads() {
    if (const char *e = getenv("B200_RESOLVER_THREADS")) {
        const int n = atoi(e);
        if (n >= 1)
            return std::min(n, 64);
    }
    cpu_set_t set;
    int cpus = 1;
    if (sched_getaffinity(0, sizeof(set), &set) == 0)
        cpus = CPU_COUNT(&set);
    // The CUDA runtime has threads of its own: on a roomy host leave them two of the cores this process may run on.
    // On a small share (the 8-GPU box gives a rank 4 CPUs) every core is worth more to the resolver than to them:
    // measured at N = 8, 2 / 3 / 4 threads on 4 CPUs: 1.61 / 1.35 / 1.25 ms per step.
    return std::max(1, cpus <= 8 ? cpus : std::min(16, cpus - 2));
}

// ------------------------------------------------------------------------------------------
// scoring and the CRC-dependent part of decode
// ------------------------------------------------------------------------------------------

// K2 marks a record whose address (key) is in the device-side set S of every address the filter could ever
// hold (LiveRec::w0 bit 31).  Outside S the filter's answer is "no" whatever its state, and the host does not
// have to touch its tables for the (random) addresses of the noise frames that share a live position.
static inline bool may_be_known(const LiveRec &r) {
    return (r.w0 >> 31) != 0;
}

// scoreModesMessage's values (mode_s.c:343-403) by [class][address known][repaired bits]: x / (errors + 1) spelled out,
// so that the walk does no integer division.  Row kKindDF11 is the IID == 0 case; kDf11Iid is IID != 0.
namespace {
constexpr int kDf11Iid = 5;
const int kScoreTable[6][2][4] = {
    {{-2, -2, -2, -2}, {-2, -2, -2, -2}},                                      // kKindBad
    {{-1, -1, -1, -1}, {1000, 1000, 1000, 1000}},                              // kKindAP
    {{-2, -2, -2, -2}, {1000, 1000, 1000, 1000}},                              // kKindAPCommB
    {{750 / 1, 750 / 2, 750 / 3, 750 / 4}, {1600 / 1, 1600 / 2, 1600 / 3, 1600 / 4}},   // kKindDF11, IID 0
    {{1400 / 1, 1400 / 2, 1400 / 3, 1400 / 4}, {1800 / 1, 1800 / 2, 1800 / 3, 1800 / 4}}, // kKindES
    {{-1, -1, -1, -1}, {1000 / 1, 1000 / 2, 1000 / 3, 1000 / 4}},              // kKindDF11, IID != 0
};
} // namespace

int Resolver::score(const IcaoFilter &filter_, const LiveRec &r) {
    const uint32_t crc = r.w0 & 0xffffffu, kind = (r.w0 >> 24) & 7u, errors = (r.w0 >> 28) & 3u;
    const uint32_t key = r.w1 & 0xffffffu; // the address the filter is asked about: the syndrome for Address/Parity frames
    const bool known = may_be_known(r) && filter_.test(key);
    if (kind > kKindES)
        return -2;
    const uint32_t row = (kind == kKindDF11 && (crc & 0x7fu) != 0) ? (uint32_t) kDf11Iid : kind;
    return kScoreTable[row][known ? 1 : 0][errors];
}

static inline uint32_t aa_field(const uint8_t *msg) { // getbits(msg, 9, 32)
    return ((uint32_t) msg[1] << 16) | ((uint32_t) msg[2] << 8) | (uint32_t) msg[3];
}

// The part of decodeModesMessage that can reject the frame or touches the filter (mode_s.c:445-555, 717-726),
// decided from what the kernels recorded about the frame: its syndrome, class, repair (error count and bit
// positions) and the address after repair (key).  Returns 0, or the reject code (-1 unknown ICAO, -2 bad).
// An all-zero frame (mode_s.c:434) never gets a class record, so that test cannot fire here.
int Resolver::admit(IcaoFilter &filter_, const LiveRec &r, uint32_t *added) {
    *added = 0xffffffffu;
    const uint32_t crc = r.w0 & 0xffffffu, kind = (r.w0 >> 24) & 7u, errors = (r.w0 >> 28) & 3u;
    const uint32_t key = r.w1 & 0xffffffu;
    switch (kind) {
        case kKindAP:      // DF0/4/5/16/24: mode_s.c:447-465
        case kKindAPCommB: // DF20/21: mode_s.c:548-554
            return (may_be_known(r) && filter_.test(key)) ? 0 : -1;
        case kKindDF11: { // mode_s.c:467-506
            if ((crc & 0xffff80u) && !(may_be_known(r) && filter_.test(key))) // a repaired all-call reply must come from a known aircraft
                return -1;
            if (!errors && (crc & 0x7fu) == 0) {
                filter_.add(key); // mode_s.c:717-726: a clean reply with IID 0
                *added = key;
            }
            return 0;
        }
        case kKindES: { // DF17/18: mode_s.c:508-546
            if (crc != 0) {
                // a repair that touched the address field (frame bits 8..31) needs the new address to be known
                const uint32_t b0 = r.errbits & 0xffu, b1 = (r.errbits >> 8) & 0xffu;
                const bool touched = (errors >= 1 && b0 >= 8 && b0 <= 31) || (errors >= 2 && b1 >= 8 && b1 <= 31);
                if (touched && !(may_be_known(r) && filter_.test(key)))
                    return -1;
            }
            if (!errors && (r.msg[0] >> 3) == 17) {
                filter_.add(key); // a clean DF17 (not DF18)
                *added = key;
            }
            return 0;
        }
        default:
            return -2;
    }
}

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

// Everything else decodeModesMessage and demodulate2400 put into the message (demod_2400.c:353-399,
// mode_s.c:424-562): pure functions of the winning record, the position and the block.  CRC and repair are
// recomputed on the host from the sliced bytes; a disagreement with the kernel's values (the ones admit()
// decided on) is counted, never hidden.
uint32_t Resolver::build(const SpanView &v, const Accepted &a, b200_message &mm, double *signal_power_out) const {
    memset(&mm, 0, sizeof(mm));
    *signal_power_out = 0.0;
    const uint64_t B = v.block_samples, b0 = (uint64_t) a.block * B;
    const uint64_t sampleTimestamp = (uint64_t) ((double) (v.first_sample + b0) * 12e6 / 2400000.0); // sdr_ifile.c:187-190
    const uint64_t sysTimestamp = sampleTimestamp / 12000U + startup_;
    if (a.modeac) {
        const AcHit h = v.ac_hits[a.index];
        mm.timestampMsg = sampleTimestamp + (h.f1_clock + 87 * 14) / 5;                       // demod_2400.c:697, at F2
        mm.sysTimestampMsg = sysTimestamp + (mm.timestampMsg - sampleTimestamp) / 12000U; // :700
        mm.msgtype = 32;                                                                   // mode_ac.c:171
        mm.msgbits = 16;
        mm.msg[0] = mm.verbatim[0] = (uint8_t) (h.modeac >> 8);
        mm.msg[1] = mm.verbatim[1] = (uint8_t) h.modeac;
        mm.addr = (h.modeac & 0x0000FF7Fu) | (1u << 24); // mode_ac.c:180
        return 0;
    }
    const LiveRec &r = v.liverecs[a.rec];
    const uint64_t j = (uint64_t) v.live[a.index].pos - b0;
    mm.timestampMsg = sampleTimestamp + j * 5 + (8 + 56) * 12 + (uint64_t) a.phase;       // demod_2400.c:358
    mm.sysTimestampMsg = sysTimestamp + (mm.timestampMsg - sampleTimestamp) / 12000U; // :361
    mm.score = a.score;
    mm.bestphase = a.phase;

    uint32_t bad = 0;
    memcpy(mm.msg, r.msg, 14);
    memcpy(mm.verbatim, r.msg, 14);
    uint8_t *msg = mm.msg;
    mm.msgtype = msg[0] >> 3;
    mm.msgbits = (mm.msgtype & 0x10) ? 112 : 56;
    mm.crc = crc_->checksum(msg, mm.msgbits);
    if (mm.crc != (r.w0 & 0xffffffu))
        ++bad;
    const uint32_t kind = (r.w0 >> 24) & 7u, errors = (r.w0 >> 28) & 3u;
    switch (mm.msgtype) {
        case 11:
            if (mm.crc & 0xffff80u) {
                const ErrorInfo *ei = crc_->diagnose(mm.crc & 0xffff80u, mm.msgbits);
                if (!ei || ei->errors > 1) {
                    ++bad;
                } else {
                    mm.correctedbits = (uint8_t) ei->errors;
                    CrcTables::fix(msg, ei);
                }
            }
            mm.addr = aa_field(msg);
            if (kind != kKindDF11)
                ++bad;
            break;
        case 17: case 18:
            if (mm.crc != 0) {
                const ErrorInfo *ei = crc_->diagnose(mm.crc, mm.msgbits);
                if (!ei) {
                    ++bad;
                } else {
                    mm.correctedbits = (uint8_t) ei->errors;
                    CrcTables::fix(msg, ei);
                }
            }
            mm.addr = aa_field(msg);
            if (kind != kKindES)
                ++bad;
            // decodeExtendedSquitter (mode_s.c:1373-1428) runs later in decodeModesMessage and may flag the address
            if (mm.msgtype == 18 && df18_non_icao(mm.msg))
                mm.addr |= 1u << 24;
            break;
        default: // Address/Parity frames: the address is the syndrome (mode_s.c:465, 554)
            mm.addr = mm.crc;
            if (kind != kKindAP && kind != kKindAPCommB)
                ++bad;
            break;
    }
    if (mm.correctedbits != errors || (mm.msgtype == 11 || mm.msgtype == 17 || mm.msgtype == 18 ? (mm.addr & 0xffffffu) != (r.w1 & 0xffffffu) : false))
        ++bad;

    // demod_2400.c:387-399
    const int signal_len = a.long_frame ? 268 : 134;
    const double signal_power = r.power / 65535.0 / 65535.0;
    mm.signalLevel = signal_power / signal_len;
    *signal_power_out = signal_power;
    memset(mm.msg + mm.msgbits / 8, 0, 14 - mm.msgbits / 8);
    memset(mm.verbatim + mm.msgbits / 8, 0, 14 - mm.msgbits / 8);
    return bad;
}

// ------------------------------------------------------------------------------------------
// the walk
// ------------------------------------------------------------------------------------------

namespace {

// sums of LiveHidden packs (their 16-bit fields hold one frame body each)
struct HiddenTotals {
    uint64_t preambles = 0, bad = 0, unknown = 0, phase[5] = {0, 0, 0, 0, 0};
    void add(uint64_t lo, uint64_t hi) {
        preambles += lo & 0xffff;
        bad += (lo >> 16) & 0xffff;
        unknown += (lo >> 32) & 0xffff;
        phase[0] += lo >> 48;
        phase[1] += hi & 0xffff;
        phase[2] += (hi >> 16) & 0xffff;
        phase[3] += (hi >> 32) & 0xffff;
        phase[4] += hi >> 48;
    }
    void add(const HiddenTotals &o) {
        preambles += o.preambles;
        bad += o.bad;
        unknown += o.unknown;
        for (int k = 0; k < 5; ++k)
            phase[k] += o.phase[k];
    }
};

} // namespace

// ------------------------------------------------------------------------------------------
// the walk over a run of mag_bufs
// ------------------------------------------------------------------------------------------

struct Resolver::Potential {
    uint32_t addr, block;
};

// what a run of mag_bufs adds to the statistics: sums, so the runs can be walked side by side (the two
// floating-point sums that depend on order are accumulated afterwards, in order, from per-block / per-frame terms)
struct Resolver::WalkOut {
    std::vector<Accepted> acc;
    std::vector<Potential> adds;     // icaoFilterAdd calls of the run, in order (the caller's filter follows them)
    std::vector<double> noise_terms; // per mag_buf: mean_power * n - sum_signal_power (demod_2400.c:423-426)
    std::vector<uint64_t> now;       // per mag_buf: Modes.ifile_now when icaoFilterExpire runs (readsb.c:331)
    uint32_t tried[32];              // positions walked, by try mask: demod_preamblePhase (demod_2400.c:184)
    uint32_t preambles, rejected_bad, rejected_unknown, accepted[3], best_phase[5], messages;
    uint64_t signal_power_count, modeac;
    void clear() {
        acc.clear();
        adds.clear();
        noise_terms.clear();
        now.clear();
        memset(tried, 0, sizeof(tried));
        preambles = rejected_bad = rejected_unknown = messages = 0;
        memset(accepted, 0, sizeof(accepted));
        memset(best_phase, 0, sizeof(best_phase));
        signal_power_count = modeac = 0;
    }
};

// first live position at or after scan position `pos`
static uint32_t first_live_at(const SpanView &v, uint64_t pos) {
    uint32_t lo = 0, hi = v.n_live;
    while (lo < hi) {
        const uint32_t mid = lo + (hi - lo) / 2;
        if ((uint64_t) v.live[mid].pos < pos)
            lo = mid + 1;
        else
            hi = mid;
    }
    return lo;
}

static uint32_t first_hit_at(const SpanView &v, uint64_t q) {
    uint32_t lo = 0, hi = v.n_ac_hits;
    while (lo < hi) {
        const uint32_t mid = lo + (hi - lo) / 2;
        if ((uint64_t) v.ac_hits[mid].q < q)
            lo = mid + 1;
        else
            hi = mid;
    }
    return lo;
}

// The exact sequential loop of demodulate2400 over the live positions of mag_bufs [k0, k1), reduced to what
// depends on order: skip-ahead, the ICAO filter `f` (scores, decode-time rejects, adds, the per-block expiry).
// It touches 16 bytes per position and the head of each record and leaves a 24-byte note per accepted frame.
void Resolver::walk(const SpanView &v, IcaoFilter &f, uint64_t k0, uint64_t k1, const std::vector<b200_block_info> &blocks, size_t block_base,
                    WalkOut &out, bool log_adds) const {
    const uint64_t n = v.nsamples, B = v.block_samples;
    out.clear();
    uint32_t live_i = first_live_at(v, k0 * B);
    uint32_t ac_i = first_hit_at(v, k0 * B);
    const uint32_t n_live = v.n_live;
    std::vector<Accepted> &acc = out.acc;

    for (uint64_t k = k0; k < k1; ++k) {
        const uint64_t b0 = k * B, b1 = std::min(n, b0 + B), nk = b1 - b0;
        const uint64_t sample_counter = v.first_sample + b0;
        // sdr_ifile.c:187-190
        const uint64_t sampleTimestamp = (uint64_t) ((double) sample_counter * 12e6 / 2400000.0);
        const uint64_t sysTimestamp = sampleTimestamp / 12000U + startup_;
        const b200_block_info &bi = blocks[block_base + k];

        uint64_t ifile_now = sysTimestamp; // demod_2400.c:253-255
        uint64_t sum_scaled_signal_power = 0;
        bool skipping = false;
        uint64_t skip_until = 0; // positions <= skip_until are skipped while `skipping`

        while (live_i < n_live && v.live[live_i].pos < b1) {
            const LivePos *lp = &v.live[live_i];
            const uint64_t p = lp->pos;
            const uint32_t li = live_i++;
            if (skipping && p <= skip_until)
                continue;
            const uint32_t trymask = lp->info & 31u, nrec = (lp->info >> 8) & 7u;
            const LiveRec *recs = v.liverecs + lp->pad;

            // score_phase for every tried phase, in order (demod_2400.c:183-229, 306-330): every tried phase
            // counts; one without a record scores -2, which only ever wins as the very first phase tried
            // (a later phase must score strictly higher, and nothing scores below -2); the records are in
            // phase order, so walking them alone gives the same pick as walking the five phases
            ++out.tried[trymask];
            int bestscore = -42, bestphase = -1;
            const LiveRec *best = nullptr;
            if (trymask) {
                const int first = 4 + __builtin_ctz(trymask);
                if (nrec == 0 || (int) ((recs[0].w1 >> 24) & 15u) != first) {
                    bestscore = -2;
                    bestphase = first;
                }
                for (uint32_t ri = 0; ri < nrec; ++ri) {
                    const int sc = score(f, recs[ri]);
                    if (sc > bestscore) {
                        bestscore = sc;
                        bestphase = (int) ((recs[ri].w1 >> 24) & 15u);
                        best = &recs[ri];
                    }
                }
            }
            out.preambles++;     // demod_2400.c:339
            if (bestscore < 0) { // demod_2400.c:342-348
                if (bestscore == -1)
                    out.rejected_unknown++;
                else
                    out.rejected_bad++;
                continue;
            }

            const uint64_t j = p - b0;
            const uint64_t timestampMsg = sampleTimestamp + j * 5 + (8 + 56) * 12 + (uint64_t) bestphase; // demod_2400.c:358
            ifile_now = sysTimestamp + (timestampMsg - sampleTimestamp) / 12000U;                         // :361, 364-366

            uint32_t added;
            const int result = admit(f, *best, &added); // demod_2400.c:372
            if (added != 0xffffffffu && log_adds)
                out.adds.push_back({added, (uint32_t) k});
            if (result < 0) {
                if (result == -1)
                    out.rejected_unknown++;
                else
                    out.rejected_bad++;
                continue;
            }
            out.accepted[(best->w0 >> 28) & 3u]++;
            out.best_phase[bestphase - 4]++;

            // (the floating-point side of :387-408 -- signal power and level, their running sum, peak and strong
            // count -- is done with the assembly: the divisions are per frame, only the sum is ordered)
            const bool long_frame = (best->msg[0] & 0x80) != 0; // :350, from the uncorrected DF
            const int signal_len = long_frame ? 268 : 134;
            out.signal_power_count += (uint64_t) signal_len;
            sum_scaled_signal_power += best->power;

            // demod_2400.c:416: skip the frame body; the for loop ends at the block boundary
            skipping = true;
            skip_until = std::min<uint64_t>(p + (uint64_t) signal_len, b1 - 1);

            out.messages++; // useModesMessage, mode_s.c:2149
            __builtin_prefetch(&v.hidden[li]); // the assembly reads it
            Accepted a;
            a.index = li;
            a.rec = (uint32_t) (best - v.liverecs);
            a.score = bestscore;
            a.block = (uint32_t) k;
            a.phase = (uint8_t) bestphase;
            a.modeac = 0;
            a.long_frame = long_frame ? 1 : 0;
            a.pad = 0;
            a.pad2 = 0;
            acc.push_back(a);
        }

        // demodulate2400AC (readsb.c:831-833, demod_2400.c:522-708) runs after demodulate2400 on the same
        // mag_buf: the block's hits in order, skipping 69 samples past every reply that is taken (:707)
        {
            uint64_t next_ok = 0; // first data index that may be examined again
            for (; ac_i < v.n_ac_hits && (uint64_t) v.ac_hits[ac_i].q < (k + 1) * B; ++ac_i) {
                const uint64_t f1 = (uint64_t) v.ac_hits[ac_i].q - k * B;
                if (f1 < next_ok)
                    continue;
                Accepted a;
                memset(&a, 0, sizeof(a));
                a.index = ac_i;
                a.block = (uint32_t) k;
                a.modeac = 1;
                acc.push_back(a);
                out.messages++; // useModesMessage, mode_s.c:2149
                ++out.modeac;
                next_ok = f1 + 70; // f1_sample += 69, then the loop's ++
            }
        }

        // demod_2400.c:423-427
        const double sum_signal_power = sum_scaled_signal_power / 65535.0 / 65535.0;
        out.noise_terms.push_back(bi.mean_power * (uint32_t) nk - sum_signal_power);
        out.now.push_back(ifile_now);
        f.expire(ifile_now); // readsb.c:331
    }
}

// ------------------------------------------------------------------------------------------
// speculation: runs of mag_bufs walked side by side
// ------------------------------------------------------------------------------------------

// Only the ICAO filter couples one mag_buf to the next.  What it will hold when a later run of mag_bufs starts
// can be predicted well: addresses enter it with clean DF17 / DF11-IID0 frames, which are accepted whatever the
// filter holds unless an earlier frame's body hides them.  So: (0) every run lists its potential adds and a guess
// of the filter clock at each block end; (1) the caller plays them through a copy of the filter and notes the
// predicted state in front of every run; (2) the runs are walked side by side, each from its predicted state with
// a filter of its own; (3) in order, a run's result stands if the state the previous run really left equals the
// prediction -- the insertion sequences of both tables, the active table and the flip time -- and is walked
// again from the true state otherwise.  The walk is a function of its input state, so the outcome is the
// sequential one; a wrong guess only costs time.

struct Resolver::Run {
    WalkOut out;
    IcaoFilter filter;              // the run's own filter: predicted state in, the state it leaves out
    std::vector<Potential> adds;    // clean DF17 / DF11-IID0 frames of the run, in order
    std::vector<uint64_t> now_guess; // per mag_buf: the filter clock at its end, guessed
    IcaoFilter::Snapshot predicted; // the state the run was started from
};

Resolver::Resolver(const CrcTables *crc, uint64_t startup_time_ms) : crc_(crc), startup_(startup_time_ms), pool_(nullptr) {
    const int n = resolver_threads();
    if (n > 1)
        pool_ = new WorkerPool(n);
    runs_.emplace_back(new Run());
    // development / test knobs: when a span is worth walking as several runs
    min_live_ = kParallelWalkMinLive;
    min_blocks_per_run_ = 4;
    if (const char *e = getenv("B200_RESOLVER_MIN_LIVE"))
        min_live_ = (uint32_t) strtoul(e, nullptr, 10);
    if (const char *e = getenv("B200_RESOLVER_PREDICT")) {
        optimistic_ = strcmp(e, "prescan") != 0;
        if (!strcmp(e, "always-optimistic")) // test knob: never mind how much the population moves
            optimistic_max_new_ = 0xffffffffu;
    }
    if (const char *e = getenv("B200_RESOLVER_ABSOLVE"))
        absolve_ = atoi(e) != 0;
    if (const char *e = getenv("B200_RESOLVER_MIN_LIVE_PER_RUN"))
        min_live_per_run_ = std::max<uint32_t>(1, (uint32_t) strtoul(e, nullptr, 10));
    if (const char *e = getenv("B200_RESOLVER_MIN_BLOCKS"))
        min_blocks_per_run_ = std::max<uint64_t>(1, strtoull(e, nullptr, 10));
    reset();
}

Resolver::~Resolver() {
    if (getenv("B200_RESOLVER_TRACE"))
        fprintf(stderr,
                "resolver: %llu spans, %llu as several runs (%llu runs, %llu walked twice, %llu kept although the filter prediction was off); ms: prescan %.2f predict %.2f walk %.2f "
                "validate %.2f merge %.2f assemble %.2f ordered sums %.2f\n",
                (unsigned long long) trace_.spans, (unsigned long long) trace_.parallel_spans, (unsigned long long) trace_.runs,
                (unsigned long long) trace_.rewalks, (unsigned long long) trace_.absolved, trace_.ms[0], trace_.ms[1], trace_.ms[2], trace_.ms[3], trace_.ms[4], trace_.ms[5], trace_.ms[6]);
    delete pool_;
}

static std::atomic<int> g_resolving{0}; // resolve() calls in progress in this process

static inline double trace_now() {
    using namespace std::chrono;
    return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

void Resolver::reset() {
    filter_.reset();
    memset(&stats_, 0, sizeof(stats_));
    ifile_now_ = 0;
    mismatches_ = 0;
    modeac_ = 0;
    recent_new_members_ = 0xffffffffu; // a new stream: nothing is known about its aircraft yet
    filter_.take_new_members();
}

// scoreModesMessage (mode_s.c:311-409) for a frame K1 already classified


void Resolver::prescan(const SpanView &v, uint64_t k0, uint64_t k1, std::vector<Potential> &adds, std::vector<uint64_t> &now_guess) const {
    const uint64_t n = v.nsamples, B = v.block_samples;
    adds.clear();
    now_guess.assign((size_t) (k1 - k0), 0);
    uint32_t li = first_live_at(v, k0 * B);
    for (uint64_t k = k0; k < k1; ++k) {
        const uint64_t b0 = k * B, b1 = std::min(n, b0 + B);
        const uint64_t sampleTimestamp = (uint64_t) ((double) (v.first_sample + b0) * 12e6 / 2400000.0);
        const uint64_t sysTimestamp = sampleTimestamp / 12000U + startup_;
        uint64_t now = sysTimestamp;
        for (; li < v.n_live && v.live[li].pos < b1; ++li) {
            const LivePos &lp = v.live[li];
            const uint32_t nrec = (lp.info >> 8) & 7u;
            const LiveRec *recs = v.liverecs + lp.pad;
            bool added = false;
            for (uint32_t ri = 0; ri < nrec; ++ri) {
                const LiveRec &r = recs[ri];
                const uint32_t kind = (r.w0 >> 24) & 7u, crc = r.w0 & 0xffffffu;
                const bool scores = kind == kKindES || (kind == kKindDF11 && (crc & 0x7fu) == 0); // >= 0 whatever the filter holds
                if (scores) {
                    const uint64_t ts = sampleTimestamp + ((uint64_t) lp.pos - b0) * 5 + (8 + 56) * 12 + ((r.w1 >> 24) & 15u);
                    now = sysTimestamp + (ts - sampleTimestamp) / 12000U;
                }
                if (!added && crc == 0 && ((kind == kKindES && (r.msg[0] >> 3) == 17) || kind == kKindDF11)) {
                    adds.push_back({r.w1 & 0xffffffu, (uint32_t) k});
                    added = true;
                }
            }
        }
        now_guess[(size_t) (k - k0)] = now;
    }
}

// Two halves.  The walk (above) is sequential in what it decides; runs of mag_bufs are walked side by side on
// predicted filter states when a span carries enough live positions.  The assembly of the messages from the
// walk's notes -- CRC recomputed and repaired on the host as a cross-check, timestamps, signal level, the
// un-counting of dead positions a frame body hides -- is independent per frame and is shared out over the
// worker pool as well.
void Resolver::resolve(const SpanView &v, MessageList &msgs, std::vector<b200_block_info> &blocks) {
    const uint64_t n = v.nsamples, B = v.block_samples;
    // ifileRun: full blocks, then (at end of stream) one short block, which is empty when the stream
    // length is a multiple of the block size (sdr_ifile.c:192-216)
    const uint64_t nfull = n / B;
    const uint64_t nblocks = nfull + (v.final_span ? 1 : 0);

    // Mode A/C hits in stream order (the kernel appends them as it finds them)
    if (v.n_ac_hits > 1)
        std::sort(v.ac_hits, v.ac_hits + v.n_ac_hits, [](const AcHit &a, const AcHit &b) { return a.q < b.q; });

    // converter outputs of every block (convert.c:104-110 / 246-252)
    const size_t block_base = blocks.size();
    for (uint64_t k = 0; k < nblocks; ++k) {
        const uint64_t b0 = k * B, nk = std::min(n, b0 + B) - b0;
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
    }

    // ---- the walk: one run, or several side by side ----
    // Several demodulators of one process (receiver streams on one GPU, a thread each) share the host's cores:
    // a resolve takes its share of the pool, not all of it.
    struct Busy {
        std::atomic<int> &n;
        int now;
        explicit Busy(std::atomic<int> &c) : n(c), now(c.fetch_add(1, std::memory_order_relaxed) + 1) {}
        ~Busy() { n.fetch_sub(1, std::memory_order_relaxed); }
    } busy(g_resolving);
    const int nworkers = pool_ ? std::max(1, pool_->size() / busy.now) : 1;
    int nruns = 1;
    // (walking side by side is about twice the work of one walk -- prescan, prediction, check: not worth it on two threads)
    const bool optimistic = optimistic_ && recent_new_members_ <= optimistic_max_new_;
    if (pool_ && nworkers >= (optimistic ? 2 : 3) && v.n_live >= min_live_ && nblocks >= 2 * min_blocks_per_run_ && filter_.replayable())
        nruns = (int) std::max<uint64_t>(1, std::min<uint64_t>({(uint64_t) nworkers, nblocks / min_blocks_per_run_, (uint64_t) v.n_live / min_live_per_run_}));
    if ((int) runs_.size() < std::max(nruns, 1)) {
        runs_.resize((size_t) std::max(nruns, 1));
        for (auto &r : runs_)
            if (!r)
                r.reset(new Run());
    }
    std::vector<uint64_t> cut((size_t) nruns + 1, nblocks); // run r walks mag_bufs [cut[r], cut[r + 1])
    cut[0] = 0;
    if (nruns > 1) {
        // equal shares of the live positions, cut at mag_buf boundaries
        for (int r = 1; r < nruns; ++r) {
            const uint32_t li = (uint32_t) ((uint64_t) v.n_live * (uint64_t) r / (uint64_t) nruns);
            uint64_t k = li < v.n_live ? (uint64_t) v.live[li].pos / B : nblocks;
            cut[(size_t) r] = std::min<uint64_t>(std::max<uint64_t>(k, cut[(size_t) r - 1]), nblocks);
        }
    }
    ++trace_.spans;
    bool force_prescan = false; // speculation from the true state kept failing on this span
    double t_mark = trace_now();
    auto lap = [&](int phase) {
        const double t = trace_now();
        trace_.ms[phase] += t - t_mark;
        t_mark = t;
    };
    if (nruns == 1) {
        walk(v, filter_, 0, nblocks, blocks, block_base, runs_[0]->out, false);
        lap(2);
    } else {
        ++trace_.parallel_spans;
        trace_.runs += (uint64_t) nruns;
      if (optimistic) {
        // Every run starts from the TRUE state in front of the first not yet confirmed run: an aircraft population
        // changes slowly, so the addresses the later runs ask about are almost always answered the same by that
        // state as by the one they would really meet (re-inserts of known aircraft change tables, not answers).
        // The runs are then confirmed in order -- same members, or a difference the run never asked about and no
        // flip inside it (differs_only_unprobed) -- while the caller's filter follows their inserts and expiries in
        // true order.  The first run that cannot be confirmed (a new aircraft it asked about, a table flip in front
        // of it) becomes the first run of the next round, started from the then-true state.  No prescan, no
        // prediction pass; a round confirms at least its first run.
        int base = 0, rounds = 0;
        while (base < nruns) {
            if (!filter_.replayable() || rounds >= kMaxRounds) {
                // the filter outgrew what a copy reproduces safely, or speculation keeps failing on this span: the
                // rest in one run, on the filter itself (and the next span predicts its start states again)

                ++respeculated_;
                walk(v, filter_, cut[(size_t) base], nblocks, blocks, block_base, runs_[(size_t) base]->out, false);
                nruns = base + 1;
                force_prescan = rounds >= kMaxRounds;
                break;
            }
            ++rounds;
            const IcaoFilter::Snapshot snap = filter_.snapshot();
            // A table flip is the one change of state that can be seen coming: it happens after the first mag_buf whose
            // clock reaches next_flip (icao_filter.c:150-164), and a mag_buf's clock lies within its 54.6 ms.  Runs
            // that start behind that mag_buf start from the flipped state: the inactive table emptied and made active.
            // (Its next_flip is a guess, which the confirmation tolerates as long as the run ends before either clock
            // flips again -- 60 s on.)
            uint64_t flip_block = nblocks;
            for (uint64_t k = cut[(size_t) base]; k < nblocks; ++k) {
                const uint64_t end_ts = (uint64_t) ((double) (v.first_sample + std::min(n, (k + 1) * B)) * 12e6 / 2400000.0) / 12000U + startup_;
                if (end_ts >= snap.next_flip) {
                    flip_block = k;
                    break;
                }
            }
            IcaoFilter::Snapshot flipped = snap;
            if (flip_block < nblocks) {
                (flipped.a_active ? flipped.seq_b : flipped.seq_a).clear();
                flipped.a_active = !flipped.a_active;
                flipped.next_flip = (uint64_t) ((double) (v.first_sample + std::min(n, (flip_block + 1) * B)) * 12e6 / 2400000.0) / 12000U + startup_ + 60000;
            }
            pool_->run((size_t) (nruns - base), 1, [&](int, size_t lo, size_t hi) {
                for (size_t i = lo; i < hi; ++i) {
                    const size_t r = (size_t) base + i;
                    Run &run = *runs_[r];
                    run.predicted = (cut[r] > flip_block) ? flipped : snap;
                    run.filter.track_probes(true);
                    run.filter.load(run.predicted);
                    walk(v, run.filter, cut[r], cut[r + 1], blocks, block_base, run.out, true);
                }
            }, true);
            lap(2);
            int r = base;
            for (; r < nruns; ++r) {
                Run &run = *runs_[(size_t) r];
                if (r > base) {
                    if (!filter_.replayable() || !run.filter.replayable())
                        break;
                    if (!filter_.same_members(run.predicted)) {
                        const uint64_t last_now = run.out.now.empty() ? 0 : *std::max_element(run.out.now.begin(), run.out.now.end());
                        if (!filter_.differs_only_unprobed(run.predicted, run.filter, last_now))
                            break;
                        ++trace_.absolved;
                    }
                }
                size_t ai = 0;
                for (uint64_t k = cut[(size_t) r]; k < cut[(size_t) r + 1]; ++k) {
                    for (; ai < run.out.adds.size() && run.out.adds[ai].block == (uint32_t) k; ++ai)
                        filter_.add(run.out.adds[ai].addr);
                    filter_.expire(run.out.now[(size_t) (k - cut[(size_t) r])]);
                }
            }
            lap(3);
            if (r < nruns) {
                trace_.rewalks += (uint64_t) (nruns - r);
                respeculated_ += (uint64_t) (nruns - r);
            }
            base = r;
        }
      } else {
        // (0) potential adds and filter-clock guesses of every run
        pool_->run((size_t) nruns, 1, [&](int, size_t lo, size_t hi) {
            for (size_t r = lo; r < hi; ++r)
                prescan(v, cut[r], cut[r + 1], runs_[r]->adds, runs_[r]->now_guess);
        }, true);
        lap(0);
        // (1) the predicted state in front of every run
        sim_.load(filter_.snapshot());
        for (int r = 0; r < nruns; ++r) {
            Run &run = *runs_[(size_t) r];
            run.predicted = sim_.snapshot();
            size_t ai = 0;
            for (uint64_t k = cut[(size_t) r]; k < cut[(size_t) r + 1]; ++k) {
                for (; ai < run.adds.size() && run.adds[ai].block == (uint32_t) k; ++ai)
                    sim_.add(run.adds[ai].addr);
                sim_.expire(run.now_guess[(size_t) (k - cut[(size_t) r])]);
            }
        }
        lap(1);
        // (2) every run from its predicted state
        pool_->run((size_t) nruns, 1, [&](int, size_t lo, size_t hi) {
            for (size_t r = lo; r < hi; ++r) {
                Run &run = *runs_[r];
                run.filter.track_probes(absolve_);
                run.filter.load(run.predicted);
                walk(v, run.filter, cut[r], cut[r + 1], blocks, block_base, run.out, true);
            }
        }, true);
        lap(2);
        // (3) in order: a run's result stands if the filter it started from had the true members; otherwise the run
        // is walked again from the true state.  The caller's own filter then follows the run's adds and expiries in
        // their true order, so its tables are laid out slot for slot as the sequential walk would leave them.
        for (int r = 0; r < nruns; ++r) {
            Run &run = *runs_[(size_t) r];
            if (!filter_.replayable()) {
                // the filter outgrew what a copy reproduces safely: the rest of the span in one run, on the filter itself
                ++respeculated_;
                walk(v, filter_, cut[(size_t) r], nblocks, blocks, block_base, run.out, false);
                nruns = r + 1;
                break;
            }
            if (!run.filter.replayable() ||
                (!filter_.same_members(run.predicted) &&
                 !(absolve_ && filter_.differs_only_unprobed(run.predicted, run.filter, run.out.now.empty() ? 0 : *std::max_element(run.out.now.begin(), run.out.now.end())) &&
                   ++trace_.absolved))) {
                ++respeculated_;
                ++trace_.rewalks;
                run.filter.load(filter_.snapshot());
                walk(v, run.filter, cut[(size_t) r], cut[(size_t) r + 1], blocks, block_base, run.out, true);
            }
            size_t ai = 0;
            for (uint64_t k = cut[(size_t) r]; k < cut[(size_t) r + 1]; ++k) {
                for (; ai < run.out.adds.size() && run.out.adds[ai].block == (uint32_t) k; ++ai)
                    filter_.add(run.out.adds[ai].addr);
                filter_.expire(run.out.now[(size_t) (k - cut[(size_t) r])]);
            }
        }
        lap(3);
      }
    }

    // how much the aircraft population moved in this span decides how the next one is speculated
    recent_new_members_ = filter_.take_new_members();
    if (force_prescan)
        recent_new_members_ = 0xffffffffu;

    // ---- assembly: one message per note, in place; every run by the thread that walked it.  A span walked in one
    // run is cut into one slice of its notes per worker instead. ----
    struct Slice {
        const Accepted *acc;
        size_t n, first; // its messages are msgs[base + first ...)
    };
    std::vector<Slice> slices;
    size_t count = 0;
    if (nruns == 1 && nworkers > 1 && runs_[0]->out.acc.size() >= 2 * (size_t) kMinNotesPerSlice) {
        const std::vector<Accepted> &acc = runs_[0]->out.acc;
        const size_t ns = std::min<size_t>((size_t) nworkers, acc.size() / kMinNotesPerSlice);
        for (size_t i = 0; i < ns; ++i) {
            const size_t lo = acc.size() * i / ns, hi = acc.size() * (i + 1) / ns;
            slices.push_back({acc.data() + lo, hi - lo, lo});
        }
        count = acc.size();
    } else {
        for (int r = 0; r < nruns; ++r) {
            const std::vector<Accepted> &acc = runs_[(size_t) r]->out.acc;
            slices.push_back({acc.data(), acc.size(), count});
            count += acc.size();
        }
    }
    const size_t nslices = slices.size();
    b200_message *out = msgs.grow(count);
    std::vector<HiddenTotals> hidden(nslices);
    std::vector<uint64_t> bad(nslices, 0);
    if (count > signal_power_cap_) {
        signal_power_cap_ = count + count / 2 + 1024;
        signal_power_.reset(new double[signal_power_cap_]);
    }
    double *power = signal_power_.get();
    std::atomic<int> turn{0}; // whose frames the ordered statistics take next
    auto assemble = [&](int, size_t lo, size_t hi) {
        for (size_t r = lo; r < hi; ++r) {
            const Accepted *acc = slices[r].acc;
            const size_t na = slices[r].n;
            HiddenTotals h; // thread-local until the slice is done: the slices' slots share cache lines
            uint64_t nbad = 0;
            b200_message *o = out + slices[r].first;
            double *pw = power + slices[r].first;
            for (size_t i = 0; i < na; ++i) {
                const Accepted &a = acc[i];
                nbad += build(v, a, o[i], &pw[i]);
                if (!a.modeac) {
                    // dead positions the frame body hides were counted in K2's per-block totals
                    const LiveHidden &lh = v.hidden[a.index];
                    if (a.long_frame)
                        h.add(lh.long_lo, lh.long_hi);
                    else
                        h.add(lh.short_lo, lh.short_hi);
                }
            }
            hidden[r] = h;
            bad[r] = nbad;
            // demod_2400.c:398-407 in message order: the running sum of doubles is the one thing here that depends on
            // it, so the slices take turns, each adding its own (cache-warm) terms
            for (uint32_t spins = 0; turn.load(std::memory_order_acquire) != (int) r; ++spins) {
                if (spins < 2048)
                    spin_pause();
                else
                    std::this_thread::yield(); // the slice before this one may sit on a core that was taken away
            }
            for (size_t i = 0; i < na; ++i) {
                if (acc[i].modeac)
                    continue;
                stats_.signal_power_sum += pw[i];
                const double level = o[i].signalLevel;
                if (level > stats_.peak_signal_power)
                    stats_.peak_signal_power = level;
                if (level > 0.50119)
                    stats_.strong_signal_count++;
            }
            turn.store((int) r + 1, std::memory_order_release);
        }
    };
    if (nslices > 1)
        pool_->run(nslices, 1, assemble, true);
    else if (nslices == 1)
        assemble(0, 0, 1);
    lap(5);

    // ---- the runs' counters, in order ----
    for (int r = 0; r < nruns; ++r) {
        const WalkOut &o = runs_[(size_t) r]->out;
        stats_.demod_preambles += o.preambles;
        stats_.demod_rejected_bad += o.rejected_bad;
        stats_.demod_rejected_unknown_icao += o.rejected_unknown;
        for (int q = 0; q < 3; ++q)
            stats_.demod_accepted[q] += o.accepted[q];
        for (int q = 0; q < 5; ++q)
            stats_.demod_bestPhase[q] += o.best_phase[q];
        stats_.messages_total += o.messages;
        stats_.signal_power_count += o.signal_power_count;
        modeac_ += o.modeac;
        for (uint32_t mask = 1; mask < 32; ++mask)
            for (int q = 0; q < 5; ++q)
                if ((mask >> q) & 1u)
                    stats_.demod_preamblePhase[q] += o.tried[mask];
        for (double t : o.noise_terms)
            stats_.noise_power_sum += t; // in block order
        if (!o.now.empty())
            ifile_now_ = o.now.back();
    }
    for (uint64_t b : bad)
        mismatches_ += b;
    for (uint64_t k = 0; k < nblocks; ++k) {
        const uint64_t b0 = k * B, nk = std::min(n, b0 + B) - b0;
        // positions no message can come from: K2's per-block totals (what skip-ahead hid of them is taken
        // out again below)
        if (nk) {
            const BlockDead &bd = v.block_dead[k];
            stats_.demod_preambles += bd.preambles;
            stats_.demod_rejected_bad += bd.rejected_bad;
            stats_.demod_rejected_unknown_icao += bd.rejected_unknown;
            for (int q = 0; q < 5; ++q)
                stats_.demod_preamblePhase[q] += bd.phase[q];
        }
        stats_.noise_power_count += nk;
        stats_.samples_processed += kOverlap + nk; // readsb.c:835
    }
    HiddenTotals total;
    for (const HiddenTotals &h : hidden)
        total.add(h);
    stats_.demod_preambles -= (uint32_t) total.preambles;
    stats_.demod_rejected_bad -= (uint32_t) total.bad;
    stats_.demod_rejected_unknown_icao -= (uint32_t) total.unknown;
    for (int q = 0; q < 5; ++q)
        stats_.demod_preamblePhase[q] -= (uint32_t) total.phase[q];
    lap(6);
}

} // namespace b200
