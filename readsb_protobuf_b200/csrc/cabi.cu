// cabi.cu -- the C ABI of include/readsb_b200.h: context, device buffers, launch sequence.
//
// A process call cuts the span into chunks of whole mag_bufs and runs them as a pipeline:
//   copy stream : H2D of chunks i+1 ... .................. (host-buffer entry only; all copies queued up front)
//   exec stream : [K1a scan -> K1b slice/CRC -> K2 classify -> order_live (-> Mode A/C) -> one small download]
//                 of chunks i .. i+5 (six ChunkSets: the host issues that far ahead of the chunk it resolves)
//                 (order_live packs the live positions, what each would hide of the dead list and their records in
//                 stream order in device memory; the dead list itself never leaves the device)
//   d2h stream  : one DMA of exactly the packed bytes of chunk i (and of chunk i+1 when its kernels are through)
//   host        : order-dependent resolve (resolver.cc: speculative, on a pool of host threads) of chunk i
// Chunks are exact: K2 of chunk i only needs the address set of chunks <= i, which is what the ICAO
// filter can hold when the host resolves chunk i.  There is no CPU implementation of the kernels;
// without a usable sm_100 device every entry point returns B200_ERR_CUDA.

#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <memory>
#include <new>
#include <string>
#include <utility>
#include <vector>

#include "device_types.h"
#include "host_tables.h"
#include "kernels.cuh"
#include "readsb_b200.h"
#include "resolver.h"

using namespace b200;

static thread_local std::string g_last_error;

static int fail(int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}

#define CUDA_TRY(expr)                                                                                   \
    do {                                                                                                 \
        cudaError_t _e = (expr);                                                                         \
        if (_e != cudaSuccess)                                                                           \
            return fail(B200_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)

namespace {

template <typename T>
struct DevBuf {
    T *p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t n) {
        if (n <= cap)
            return cudaSuccess;
        if (p)
            cudaFree(p);
        p = nullptr;
        cap = 0;
        cudaError_t e = cudaMalloc(&p, n * sizeof(T));
        if (e == cudaSuccess)
            cap = n;
        return e;
    }
    void release() {
        if (p)
            cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

template <typename T>
struct PinnedBuf {
    T *p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t n) {
        if (n <= cap)
            return cudaSuccess;
        if (p)
            cudaFreeHost(p);
        p = nullptr;
        cap = 0;
        cudaError_t e = cudaHostAlloc(&p, n * sizeof(T), cudaHostAllocDefault);
        if (e == cudaSuccess)
            cap = n;
        return e;
    }
    void release() {
        if (p)
            cudaFreeHost(p);
        p = nullptr;
        cap = 0;
    }
};

double now_ms() {
    using namespace std::chrono;
    return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

// development aid (B200_SPAN_TRACE=1): host-side time marks of a span's chunk pipeline, printed when the span ends
struct SpanTrace {
    bool on;
    double t0;
    std::vector<std::pair<const char *, double>> marks;
    SpanTrace() : on(getenv("B200_SPAN_TRACE") != nullptr), t0(0) {}
    void begin() {
        if (on) {
            marks.clear();
            t0 = now_ms();
        }
    }
    void mark(const char *what) {
        if (on)
            marks.emplace_back(what, now_ms() - t0);
    }
    void print(uint64_t nsamples) {
        if (!on)
            return;
        fprintf(stderr, "span of %llu samples:", (unsigned long long) nsamples);
        for (auto &m : marks)
            fprintf(stderr, " %s@%.3f", m.first, m.second);
        fprintf(stderr, "\n");
    }
};
static thread_local SpanTrace g_trace;

// a typed window into a larger allocation (not owning)
template <typename T>
struct View {
    T *p = nullptr;
};

// per-tile slabs K1 writes into on the first attempt: ~6x the candidate / record density of noise at
// the default threshold; a chunk that needs more is run again with slabs placed exactly
const uint32_t kCandSlab = 640, kRecSlab = 448;
// samples per pipeline chunk (rounded to whole mag_bufs)
const uint64_t kChunkTarget = 32ull << 20;

// everything one in-flight chunk owns
struct ChunkSet {
    DevBuf<uint32_t> d_cand, d_tile_off, d_dead;
    DevBuf<uint16_t> d_magbuf; // K1a's magnitudes of the chunk
    DevBuf<uint16_t> d_step_off;
    DevBuf<PhaseRec> d_recs;
    DevBuf<TileDesc> d_tiles;
    // the small per-chunk outputs share one device allocation and one pinned mirror, so that a chunk
    // costs one memset and one download: [counters | block sums | block dead counters || tile outputs]
    DevBuf<uint8_t> d_small;
    PinnedBuf<uint8_t> h_small;
    size_t small_zero_bytes = 0, small_bytes = 0;
    View<ScanCounters> d_counters, h_counters;
    View<unsigned long long> d_sums_u64, h_sums_u64;
    View<double> d_sums_f64, h_sums_f64;
    View<BlockDead> d_block_dead, h_block_dead;
    View<TileOut> d_tiles_out, h_tiles_out;
    // K2 leaves the live positions and records in device memory, tile by tile in the order its warps reserved
    // them; order_live packs them into stream order (d_packed below), together with the dead-position counts a
    // frame accepted at each live position would hide.  The (much longer) dead list itself stays in device memory.
    DevBuf<LivePos> d_live;
    DevBuf<LiveRec> d_liverecs;
    DevBuf<uint2> d_live_base;
    // order_live's output, packed [n_live LivePos | n_live LiveHidden | n_liverec LiveRec] in device memory; the copy
    // engine brings exactly those bytes to the pinned mirror once the host has read the chunk's counters (the SMs'
    // own posted writes over PCIe cost the GPU 20-55 us per chunk; the DMA runs under the next chunk's kernels)
    DevBuf<uint8_t> d_packed;
    PinnedBuf<uint8_t> h_packed;
    size_t live_cap = 0, liverec_cap = 0;
    cudaEvent_t ev_packed = nullptr;
    bool packed_issued = false;
    bool in_flight = false; // issued, not yet resolved
    // Mode A/C (only with cfg.mode_ac): per-block noise levels, unordered hit list in pinned host memory
    DevBuf<uint32_t> d_ac_noise;
    PinnedBuf<AcHit> h_ac_hits;
    cudaEvent_t ev_begin = nullptr, ev_k1 = nullptr, ev_k1b = nullptr, ev_k2 = nullptr, ev_small = nullptr;

    // what is in flight
    uint64_t start = 0, nsamples = 0;
    bool final_chunk = false;
    uint32_t head_valid = 0;
    const uint8_t *iq = nullptr, *head = nullptr;
    const double *span_fsums = nullptr; // float formats, device-resident span: this chunk's slice of the span's block sums
    size_t small_d2h_bytes = 0;

    void release() {
        d_cand.release(); d_tile_off.release(); d_recs.release(); d_tiles.release(); d_magbuf.release(); d_step_off.release();
        d_small.release(); h_small.release(); d_dead.release();
        h_packed.release(); d_packed.release(); live_cap = liverec_cap = 0; d_live.release(); d_liverecs.release(); d_live_base.release(); d_ac_noise.release(); h_ac_hits.release();
        for (cudaEvent_t *e : {&ev_begin, &ev_k1, &ev_k1b, &ev_k2, &ev_small, &ev_packed})
            if (*e) {
                cudaEventDestroy(*e);
                *e = nullptr;
            }
    }
};

} // namespace

struct b200_demod {
    b200_demod_config cfg;
    int bytes_per_sample = 2; // of the caller's IQ
    // what K1a / K2 / the resolver see: the caller's format, or -- behind the DC-filter front end -- format 3,
    // a stream of u16 magnitudes
    uint32_t eff_format = 0;
    int eff_bps = 2;
    float dc_a = 0, dc_b = 1;           // init_converter, convert.c:476-488
    DevBuf<float> d_dc_aI, d_dc_aQ, d_dc_state;
    DevBuf<uint16_t> d_dc_mag;          // the DC-filtered magnitude stream of the span
    DevBuf<uint8_t> d_dc_raw;           // host-buffer entry: the caller's IQ on the device
    bool dc_sums_ready = false;         // d_span_fsums already holds the span's block sums
    int sm_count = 0;
    int scan_grid = 0, slice_grid = 0;
    size_t ac_hit_cap = 0; // grown when a chunk's Mode A/C hits did not fit
    std::unique_ptr<CrcTables> crc;
    std::unique_ptr<Resolver> resolver;

    cudaStream_t stream = nullptr;      // exec stream of the host-buffer entry
    cudaStream_t copy_stream = nullptr; // H2D
    cudaStream_t d2h_stream = nullptr;  // the chunks' packed live data
    cudaEvent_t ev_h2d_begin = nullptr, ev_h2d_end = nullptr;
    std::vector<cudaEvent_t> ev_chunk_h2d;

    // tables
    DevBuf<uint16_t> d_lut;     // uc8 table, plain + bank-swizzled copy
    DevBuf<uint16_t> d_lut_q11; // format 4: the sc16q11 table (padded to 65536 entries), plain + swizzled
    DevBuf<ErrorInfo> d_tab_short, d_tab_long;
    DevBuf<uint32_t> d_bitmap;
    std::vector<uint16_t> h_lut;

    // stream state
    DevBuf<uint8_t> d_head, d_head_tmp;
    uint32_t head_valid = 0;
    uint64_t first_sample = 0;
    bool finished = false;
    bool failed = false; // a process call failed half-way: the stream state is inconsistent until b200_demod_reset

    // span buffers
    DevBuf<uint8_t> d_iq;
    // chunks in flight: the host issues up to kSets chunks ahead of the one it is resolving, so that in a span of a
    // few chunks every launch is queued while the GPU works on the first one and the host's part of the pipeline is
    // the resolver alone
    static constexpr int kSets = 6;
    ChunkSet sets[kSets];
    DevBuf<uint8_t> d_dbg_masks;
    DevBuf<double> d_span_fsums; // float formats: [mag_bufs of the span][2]
    DevBuf<uint16_t> d_mag;
    DevBuf<uint8_t> d_frames;
    DevBuf<uint32_t> d_syn;
    DevBuf<int8_t> d_err, d_bits;
    DevBuf<unsigned long long> d_csum_u64;
    DevBuf<double> d_csum_f64;

    // results of the last call
    MessageList msgs;
    std::vector<b200_block_info> blocks;
    b200_timing timing;

    ~b200_demod() {
        cudaSetDevice(cfg.device);
        d_lut.release(); d_lut_q11.release(); d_tab_short.release(); d_tab_long.release(); d_bitmap.release();
        d_head.release(); d_head_tmp.release(); d_iq.release();
        for (ChunkSet &c : sets)
            c.release();
        d_span_fsums.release(); d_dbg_masks.release(); d_mag.release(); d_frames.release(); d_syn.release(); d_err.release(); d_bits.release();
        d_csum_u64.release(); d_csum_f64.release();
        d_dc_aI.release(); d_dc_aQ.release(); d_dc_state.release(); d_dc_mag.release(); d_dc_raw.release();
        for (cudaEvent_t e : ev_chunk_h2d)
            cudaEventDestroy(e);
        if (ev_h2d_begin)
            cudaEventDestroy(ev_h2d_begin);
        if (ev_h2d_end)
            cudaEventDestroy(ev_h2d_end);
        for (cudaStream_t s : {stream, copy_stream, d2h_stream})
            if (s)
                cudaStreamDestroy(s);
    }
};

extern "C" const char *b200_last_error(void) {
    return g_last_error.c_str();
}

static int ensure_chunk_buffers(b200_demod *d, ChunkSet &c, uint64_t nsamples, size_t cand_total, size_t rec_total, size_t dead_cap,
                                size_t live_cap, size_t liverec_cap) {
    const size_t ntiles = tiles_for(nsamples);
    const size_t nblocks = (size_t) (nsamples / d->cfg.block_samples + 2);
    CUDA_TRY(c.d_cand.ensure(cand_total + 1));
    CUDA_TRY(c.d_recs.ensure(rec_total + 1));
    CUDA_TRY(c.d_dead.ensure(dead_cap));
    c.live_cap = std::max(c.live_cap, live_cap);
    c.liverec_cap = std::max(c.liverec_cap, liverec_cap);
    {
        const size_t bytes = c.live_cap * (sizeof(LivePos) + sizeof(LiveHidden)) + c.liverec_cap * sizeof(LiveRec) + 64;
        CUDA_TRY(c.d_packed.ensure(bytes));
        CUDA_TRY(c.h_packed.ensure(bytes));
    }
    CUDA_TRY(c.d_live.ensure(c.live_cap));
    CUDA_TRY(c.d_liverecs.ensure(c.liverec_cap));
    CUDA_TRY(c.d_live_base.ensure(ntiles + 1));
    CUDA_TRY(c.d_tiles.ensure(ntiles + 1));
    CUDA_TRY(c.d_magbuf.ensure(ntiles * (size_t) kTile + kMagSlack));
    CUDA_TRY(c.d_step_off.ensure((ntiles + 1) * (size_t) kScanSteps));
    if (d->cfg.mode_ac) {
        CUDA_TRY(c.d_ac_noise.ensure(nblocks));
        CUDA_TRY(c.h_ac_hits.ensure(std::max<size_t>(c.h_ac_hits.cap, std::max<size_t>(d->ac_hit_cap, (size_t) (nsamples / 256 + 4096)))));
    }
    {
        auto up = [](size_t x) { return (x + 63) & ~(size_t) 63; };
        const size_t o_cnt = 0, o_su = up(sizeof(ScanCounters)), o_sf = o_su + up(2 * nblocks * sizeof(unsigned long long)),
                     o_bd = o_sf + up(2 * nblocks * sizeof(double)), o_to = o_bd + up(nblocks * sizeof(BlockDead)),
                     total = o_to + up((ntiles + 1) * sizeof(TileOut));
        CUDA_TRY(c.d_small.ensure(total));
        CUDA_TRY(c.h_small.ensure(total));
        c.small_zero_bytes = o_to;
        c.small_bytes = total;
        auto bind = [&](uint8_t *base, View<ScanCounters> &cnt, View<unsigned long long> &su, View<double> &sf, View<BlockDead> &bd,
                        View<TileOut> &to) {
            cnt.p = reinterpret_cast<ScanCounters *>(base + o_cnt);
            su.p = reinterpret_cast<unsigned long long *>(base + o_su);
            sf.p = reinterpret_cast<double *>(base + o_sf);
            bd.p = reinterpret_cast<BlockDead *>(base + o_bd);
            to.p = reinterpret_cast<TileOut *>(base + o_to);
        };
        bind(c.d_small.p, c.d_counters, c.d_sums_u64, c.d_sums_f64, c.d_block_dead, c.d_tiles_out);
        bind(c.h_small.p, c.h_counters, c.h_sums_u64, c.h_sums_f64, c.h_block_dead, c.h_tiles_out);
    }
    if (!c.ev_begin) {
        CUDA_TRY(cudaEventCreate(&c.ev_begin));
        CUDA_TRY(cudaEventCreate(&c.ev_k1));
        CUDA_TRY(cudaEventCreate(&c.ev_k1b));
        CUDA_TRY(cudaEventCreate(&c.ev_k2));
        CUDA_TRY(cudaEventCreate(&c.ev_small));
        CUDA_TRY(cudaEventCreateWithFlags(&c.ev_packed, cudaEventDisableTiming));
    }
    return B200_OK;
}

extern "C" int b200_demod_create(const b200_demod_config *cfg, b200_demod **out) {
    if (!cfg || !out)
        return fail(B200_ERR_ARG, "null argument");
    *out = nullptr;
    if (cfg->abi_version != B200_ABI_VERSION)
        return fail(B200_ERR_ARG, "abi_version %d != %d", cfg->abi_version, B200_ABI_VERSION);
    if (cfg->input_format < B200_INPUT_UC8 || cfg->input_format > B200_INPUT_SC16Q11)
        return fail(B200_ERR_ARG, "no suitable converter for format=%d", cfg->input_format); // convert.c:460-464
    if (cfg->nfix_crc < 0 || cfg->nfix_crc > 2)
        return fail(B200_ERR_ARG, "nfix_crc must be 0, 1 or 2");
    if (cfg->preamble_threshold < 1 || cfg->preamble_threshold > 6000)
        return fail(B200_ERR_ARG, "preamble_threshold out of range");
    if (cfg->sc16q11_table_bits < 0 || cfg->sc16q11_table_bits > 11)
        return fail(B200_ERR_ARG, "sc16q11_table_bits must be 0 (float path) or 1..11 (SC16Q11_TABLE_BITS, convert.c:264-268)");

    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(B200_ERR_CUDA, "no CUDA device: %s (this library has no CPU fallback)",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
    if (cfg->device < 0 || cfg->device >= ndev)
        return fail(B200_ERR_ARG, "device %d out of range (%d devices)", cfg->device, ndev);
    CUDA_TRY(cudaSetDevice(cfg->device));
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, cfg->device));
    if (prop.major != 10)
        return fail(B200_ERR_CUDA, "device %d is sm_%d%d; this library carries sm_100a code only", cfg->device, prop.major,
                    prop.minor);

    std::unique_ptr<b200_demod> d(new (std::nothrow) b200_demod());
    if (!d)
        return fail(B200_ERR_NOMEM, "out of memory");
    d->cfg = *cfg;
    if (d->cfg.block_samples == 0)
        d->cfg.block_samples = B200_DEFAULT_BLOCK_SAMPLES;
    if (d->cfg.block_samples % 8 != 0)
        return fail(B200_ERR_ARG, "block_samples must be a multiple of 8");
    if (d->cfg.max_span_samples == 0)
        d->cfg.max_span_samples = 64ull << 20;
    d->bytes_per_sample = (cfg->input_format == B200_INPUT_UC8) ? 2 : 4;
    d->eff_format = (uint32_t) cfg->input_format;
    d->eff_bps = d->bytes_per_sample;
    if (cfg->filter_dc) {
        // init DC block @ 1Hz (convert.c:476-480): float fields assigned from double expressions; the
        // demodulator is the 2.4 MS/s one (Modes.sample_rate, readsb.c:141)
        d->dc_b = (float) exp(-2.0 * M_PI * 1.0 / 2400000.0);
        d->dc_a = (float) (1.0 - d->dc_b);
        d->eff_format = 3;
        d->eff_bps = 2;
    } else if (cfg->input_format == B200_INPUT_SC16Q11 && cfg->sc16q11_table_bits) {
        d->eff_format = 4; // converters_table order (convert.c:432-437): the table converter when it is compiled in
    }
    d->sm_count = prop.multiProcessorCount;
    memset(&d->timing, 0, sizeof(d->timing));

    d->crc.reset(new CrcTables(cfg->nfix_crc));
    d->resolver.reset(new Resolver(d->crc.get(), cfg->startup_time_ms));

    CUDA_TRY(cudaStreamCreateWithFlags(&d->stream, cudaStreamNonBlocking));
    CUDA_TRY(cudaStreamCreateWithFlags(&d->copy_stream, cudaStreamNonBlocking));
    CUDA_TRY(cudaStreamCreateWithFlags(&d->d2h_stream, cudaStreamNonBlocking));
    CUDA_TRY(cudaEventCreate(&d->ev_h2d_begin));
    CUDA_TRY(cudaEventCreate(&d->ev_h2d_end));
    CUDA_TRY(scan_configure());
    CUDA_TRY(scan2_configure());
    CUDA_TRY(scan3_configure());
    CUDA_TRY(slice_configure());
    CUDA_TRY(upload_constants(d->crc->bit_syndromes()));

    d->h_lut.resize(65536);
    build_uc8_table(d->h_lut.data());
    // second copy in the bank-swizzled shared-memory layout of K1 (32-bit word j of row Q at j ^ (Q & 31)),
    // so that every CTA stages the table with straight 16-byte copies
    // ... and a third one in scan2_kernel's layout: entry i at i ^ ((i >> 5) & 0x38), i.e. 32-bit word j at
    // j ^ ((j >> 5) & 0x1c)
    {
        std::vector<uint32_t> all(3 * 32768);
        const uint32_t *words = reinterpret_cast<const uint32_t *>(d->h_lut.data());
        for (uint32_t i = 0; i < 32768; ++i) {
            all[i] = words[i];
            all[32768 + (i ^ ((i >> 7) & 31u))] = words[i];
            all[65536 + (i ^ ((i >> 5) & 0x1cu))] = words[i];
        }
        CUDA_TRY(d->d_lut.ensure(3 * 65536));
        CUDA_TRY(cudaMemcpy(d->d_lut.p, all.data(), 3 * 65536 * sizeof(uint16_t), cudaMemcpyHostToDevice));
    }

    if (d->eff_format == 4) {
        // up to 8 bits the table (<= 65536 entries) takes the uc8 table's place in K1a's shared memory, staged from
        // its swizzled copy; with 9..11 bits (512 KiB .. 8 MiB) it stays in global memory and is read through L2
        const int bits = cfg->sc16q11_table_bits;
        const size_t entries = std::max<size_t>(65536, (size_t) 1 << (2 * bits));
        std::vector<uint16_t> table(entries, 0);
        build_sc16q11_table(bits, table.data());
        if (bits <= 8) {
            std::vector<uint32_t> both(65536);
            const uint32_t *words = reinterpret_cast<const uint32_t *>(table.data());
            for (uint32_t i = 0; i < 32768; ++i) {
                both[i] = words[i];
                both[32768 + (i ^ ((i >> 7) & 31u))] = words[i];
            }
            CUDA_TRY(d->d_lut_q11.ensure(2 * 65536));
            CUDA_TRY(cudaMemcpy(d->d_lut_q11.p, both.data(), 2 * 65536 * sizeof(uint16_t), cudaMemcpyHostToDevice));
        } else {
            CUDA_TRY(d->d_lut_q11.ensure(entries));
            CUDA_TRY(cudaMemcpy(d->d_lut_q11.p, table.data(), entries * sizeof(uint16_t), cudaMemcpyHostToDevice));
        }
    }

    const auto &ts = d->crc->short_table();
    const auto &tl = d->crc->long_table();
    CUDA_TRY(d->d_tab_short.ensure(ts.size() + 1));
    CUDA_TRY(d->d_tab_long.ensure(tl.size() + 1));
    if (!ts.empty())
        CUDA_TRY(cudaMemcpy(d->d_tab_short.p, ts.data(), ts.size() * sizeof(ErrorInfo), cudaMemcpyHostToDevice));
    if (!tl.empty())
        CUDA_TRY(cudaMemcpy(d->d_tab_long.p, tl.data(), tl.size() * sizeof(ErrorInfo), cudaMemcpyHostToDevice));

    CUDA_TRY(d->d_bitmap.ensure((1u << 24) / 32));
    CUDA_TRY(cudaMemsetAsync(d->d_bitmap.p, 0, (1u << 24) / 8, d->stream));
    CUDA_TRY(d->d_head.ensure((size_t) kHead * 4));
    CUDA_TRY(d->d_head_tmp.ensure((size_t) kHead * 4));
    CUDA_TRY(cudaMemsetAsync(d->d_head.p, 0, (size_t) kHead * 4, d->stream));
    CUDA_TRY(d->d_dc_state.ensure(4));
    CUDA_TRY(cudaMemsetAsync(d->d_dc_state.p, 0, 4 * sizeof(float), d->stream)); // z1_I = z1_Q = 0, convert.c:473-474
    CUDA_TRY(cudaStreamSynchronize(d->stream));

    // one persistent CTA per SM (the uc8 table takes most of an SM's shared memory)
    d->scan_grid = d->sm_count;
    d->slice_grid = d->sm_count;
    *out = d.release();
    return B200_OK;
}

extern "C" void b200_demod_destroy(b200_demod *d) {
    delete d;
}

extern "C" int b200_demod_reset(b200_demod *d) {
    if (!d)
        return fail(B200_ERR_ARG, "null context");
    CUDA_TRY(cudaSetDevice(d->cfg.device));
    CUDA_TRY(cudaMemsetAsync(d->d_bitmap.p, 0, (1u << 24) / 8, d->stream));
    CUDA_TRY(cudaMemsetAsync(d->d_head.p, 0, (size_t) kHead * 4, d->stream));
    CUDA_TRY(cudaMemsetAsync(d->d_dc_state.p, 0, 4 * sizeof(float), d->stream));
    CUDA_TRY(cudaStreamSynchronize(d->stream));
    d->head_valid = 0;
    d->first_sample = 0;
    d->finished = false;
    d->failed = false;
    d->resolver->reset();
    d->msgs.clear();
    d->blocks.clear();
    memset(&d->timing, 0, sizeof(d->timing));
    return B200_OK;
}

static ScanArgs make_scan_args(b200_demod *d, ChunkSet &c, const uint8_t *d_iq, const uint8_t *d_head, uint64_t nsamples,
                               uint32_t head_valid, uint32_t cand_slab, uint32_t rec_slab, const uint32_t *tile_off) {
    ScanArgs a;
    memset(&a, 0, sizeof(a));
    a.iq = d_iq;
    a.head = d_head;
    a.head_valid = head_valid;
    a.format = d->eff_format;
    a.nsamples = nsamples;
    a.threshold = d->cfg.preamble_threshold;
    a.block_samples = d->cfg.block_samples;
    a.ntiles = tiles_for(nsamples);
    a.lut = (d->eff_format == 4) ? d->d_lut_q11.p : d->d_lut.p;
    a.lut_swz = (d->eff_format == 4 && d->cfg.sc16q11_table_bits > 8) ? a.lut : a.lut + 65536; // no staged copy of a big table
    a.lut_swz2 = d->d_lut.p + 2 * 65536;
    a.fast_lo = a.fast_hi = 0;
    a.table_bits = d->cfg.sc16q11_table_bits;
    a.tab_short = d->d_tab_short.p;
    a.tab_long = d->d_tab_long.p;
    a.n_short = (int32_t) d->crc->short_table().size();
    a.n_long = (int32_t) d->crc->long_table().size();
    a.addr_bitmap = d->d_bitmap.p;
    a.cand = c.d_cand.p;
    a.recs = c.d_recs.p;
    a.tiles = c.d_tiles.p;
    a.mag = c.d_magbuf.p;
    a.step_off = c.d_step_off.p;
    a.cand_slab = cand_slab;
    a.rec_slab = rec_slab;
    a.tile_off = tile_off;
    a.counters = c.d_counters.p;
    a.block_sums_u64 = c.d_sums_u64.p;
    a.block_sums_f64 = c.d_sums_f64.p;
    a.dbg_masks = nullptr;
    // the interior tiles of a uc8 span go through the register-window kernel (scan2.cu), the edge tiles (carried
    // head, ragged tail) through scan_kernel
    if (scan2_supports(a) && !getenv("B200_NO_SCAN2"))
        scan2_tile_range(nsamples, a.fast_lo, a.fast_hi);
    return a;
}

// K1a: scan2_kernel when the span has interior uc8 tiles for it (it then takes the edge tiles along), else scan_kernel
static cudaError_t launch_k1a(const ScanArgs &sa, int mode, int grid, cudaStream_t s) {
    if (sa.fast_hi > sa.fast_lo)
        return launch_scan2(sa, mode, grid, s);
    return launch_scan(sa, mode, grid, s);
}

static SliceArgs make_slice_args(const ScanArgs &sa) {
    SliceArgs b;
    memset(&b, 0, sizeof(b));
    b.mag = sa.mag;
    b.cand = sa.cand;
    b.step_off = sa.step_off;
    b.recs = sa.recs;
    b.tiles = sa.tiles;
    b.ntiles = sa.ntiles;
    b.cand_slab = sa.cand_slab;
    b.rec_slab = sa.rec_slab;
    b.tile_off = sa.tile_off;
    b.tab_short = sa.tab_short;
    b.tab_long = sa.tab_long;
    b.n_short = sa.n_short;
    b.n_long = sa.n_long;
    b.addr_bitmap = sa.addr_bitmap;
    b.counters = sa.counters;
    return b;
}

static int zero_chunk_outputs(b200_demod *, ChunkSet &c, uint64_t, cudaStream_t s) {
    CUDA_TRY(cudaMemsetAsync(c.d_small.p, 0, c.small_zero_bytes, s));
    return B200_OK;
}

// after a span: head <- last kHead samples of (head ++ span)
static int carry_head(b200_demod *d, const uint8_t *d_iq, uint64_t nsamples, cudaStream_t s) {
    const size_t bps = (size_t) d->eff_bps;
    if (nsamples >= (uint64_t) kHead) {
        CUDA_TRY(cudaMemcpyAsync(d->d_head.p, d_iq + (nsamples - kHead) * bps, kHead * bps, cudaMemcpyDeviceToDevice, s));
        d->head_valid = kHead;
    } else if (nsamples > 0) {
        const size_t keep = kHead - (size_t) nsamples;
        CUDA_TRY(cudaMemcpyAsync(d->d_head_tmp.p, d->d_head.p + nsamples * bps, keep * bps, cudaMemcpyDeviceToDevice, s));
        CUDA_TRY(cudaMemcpyAsync(d->d_head.p, d->d_head_tmp.p, keep * bps, cudaMemcpyDeviceToDevice, s));
        CUDA_TRY(cudaMemcpyAsync(d->d_head.p + keep * bps, d_iq, nsamples * bps, cudaMemcpyDeviceToDevice, s));
        d->head_valid = (uint32_t) std::min<uint64_t>(kHead, d->head_valid + nsamples);
    }
    return B200_OK;
}

// enqueue K1, K2 and the small D2H of a chunk on stream s
static int issue_chunk(b200_demod *d, ChunkSet &c, cudaStream_t s, bool exact, size_t cand_total, size_t rec_total, size_t dead_cap,
                       size_t live_cap, size_t liverec_cap, uint32_t *launches) {
    const uint64_t n = c.nsamples;
    const uint32_t B = d->cfg.block_samples;
    int rc = ensure_chunk_buffers(d, c, n, cand_total, rec_total, dead_cap, live_cap, liverec_cap);
    if (rc != B200_OK)
        return rc;
    rc = zero_chunk_outputs(d, c, n, s);
    if (rc != B200_OK)
        return rc;
    const uint32_t ntiles = tiles_for(n);
    const size_t nblocks = (size_t) (n / B + (c.final_chunk ? 1 : 0));

    ScanArgs sa = make_scan_args(d, c, c.iq, c.head, n, c.head_valid, kCandSlab, kRecSlab, exact ? c.d_tile_off.p : nullptr);
    const bool float_format = d->eff_format != B200_INPUT_UC8 && d->eff_format != 4; // the table converters sum integers
    if (float_format && n) {
        // mean_level / mean_power of the float converters are sequential float sums (convert.c:228,241-242):
        // K1a leaves them alone, float_block_sums_kernel walks every mag_buf's chain in order.  That kernel
        // is latency-bound (a dependent add per sample) and runs best alone: for a device-resident span it
        // ran once over all mag_bufs before the first chunk (c.span_fsums), for host buffers it runs here,
        // behind the chunk's H2D, where the GPU would otherwise wait for the next copy.
        sa.block_sums_f64 = nullptr;
        const uint32_t nb = (uint32_t) ((n + B - 1) / B);
        if (c.span_fsums)
            CUDA_TRY(cudaMemcpyAsync(c.d_sums_f64.p, c.span_fsums, 2 * (size_t) nb * sizeof(double), cudaMemcpyDeviceToDevice, s));
        else
            CUDA_TRY(launch_float_block_sums(c.iq, sa.format, n, B, nb, c.d_sums_f64.p, s));
        if (launches)
            *launches += c.span_fsums ? 0 : 1;
    }
    CUDA_TRY(cudaEventRecord(c.ev_begin, s));
    CUDA_TRY(launch_k1a(sa, 1, d->scan_grid, s));
    CUDA_TRY(cudaEventRecord(c.ev_k1, s));
    CUDA_TRY(launch_slice(make_slice_args(sa), d->slice_grid, s));
    CUDA_TRY(cudaEventRecord(c.ev_k1b, s));

    ClassifyArgs ca;
    memset(&ca, 0, sizeof(ca));
    ca.iq = c.iq;
    ca.head = c.head;
    ca.head_valid = c.head_valid;
    ca.format = sa.format;
    ca.nsamples = n;
    ca.block_samples = B;
    ca.ntiles = ntiles;
    ca.lut = sa.lut;
    ca.table_bits = sa.table_bits;
    ca.tab_short = sa.tab_short;
    ca.tab_long = sa.tab_long;
    ca.n_short = sa.n_short;
    ca.n_long = sa.n_long;
    ca.addr_bitmap = d->d_bitmap.p;
    ca.cand = c.d_cand.p;
    ca.recs = c.d_recs.p;
    ca.tiles = c.d_tiles.p;
    ca.mag = c.d_magbuf.p;
    ca.max_cand_per_tile = exact ? 0xffffffffu : kCandSlab;
    ca.dead = c.d_dead.p;
    ca.live = c.d_live.p;
    ca.liverecs = c.d_liverecs.p;
    ca.tiles_out = c.d_tiles_out.p;
    ca.dead_cap = (uint32_t) std::min<size_t>(c.d_dead.cap, 0xffffffffu);
    ca.live_cap = (uint32_t) std::min<size_t>(c.live_cap, 0xffffffffu);
    ca.liverec_cap = (uint32_t) std::min<size_t>(c.liverec_cap, 0xffffffffu);
    ca.counters = c.d_counters.p;
    ca.block_dead = c.d_block_dead.p;
    // K2 looks at K1's overflow flag itself and does nothing when K1 did not fit
    CUDA_TRY(launch_classify(ca, s));
    // live positions / hidden counts / records in stream order, packed back to back in device memory
    CUDA_TRY(launch_order_live(c.d_tiles_out.p, ntiles, c.d_counters.p, c.d_live_base.p, c.d_live.p, c.d_liverecs.p, c.d_dead.p, n, B,
                               c.d_packed.p, s));
    c.packed_issued = false;
    c.in_flight = true;
    CUDA_TRY(cudaEventRecord(c.ev_k2, s));
    if (d->cfg.mode_ac && n) {
        // demodulate2400AC (readsb.c:831-833) over the same magnitudes
        ModeacArgs ma;
        memset(&ma, 0, sizeof(ma));
        ma.mag = c.d_magbuf.p;
        ma.nsamples = n;
        ma.block_samples = B;
        ma.format = sa.format;
        ma.nblocks = (uint32_t) ((n + B - 1) / B);
        ma.sums_u64 = c.d_sums_u64.p;
        ma.sums_f64 = c.d_sums_f64.p;
        ma.noise_level = c.d_ac_noise.p;
        ma.hits = c.h_ac_hits.p;
        ma.hit_cap = (uint32_t) std::min<size_t>(c.h_ac_hits.cap, 0xffffffffu);
        ma.counters = c.d_counters.p;
        CUDA_TRY(launch_modeac(ma, s));
        if (launches)
            *launches += 2;
    }
    if (launches)
        *launches += ntiles ? 5 : 0;
    CUDA_TRY(cudaMemcpyAsync(c.h_small.p, c.d_small.p, c.small_bytes, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaEventRecord(c.ev_small, s));
    c.small_d2h_bytes = sizeof(ScanCounters) + ntiles * sizeof(TileOut) +
                        nblocks * (2 * sizeof(unsigned long long) + 2 * sizeof(double) + sizeof(BlockDead));
    return B200_OK;
}

// the chunk's packed live data -> its pinned mirror, on the download stream (the chunk's kernels are known to be through)
static int download_packed(b200_demod *d, ChunkSet &c, const ScanCounters &cnt) {
    if (c.packed_issued)
        return B200_OK;
    const size_t bytes = (size_t) cnt.n_live * (sizeof(LivePos) + sizeof(LiveHidden)) + (size_t) cnt.n_liverec * sizeof(LiveRec);
    if (bytes)
        CUDA_TRY(cudaMemcpyAsync(c.h_packed.p, c.d_packed.p, bytes, cudaMemcpyDeviceToHost, d->d2h_stream));
    CUDA_TRY(cudaEventRecord(c.ev_packed, d->d2h_stream));
    c.packed_issued = true;
    return B200_OK;
}

// wait for a chunk's kernels, fetch its survivors, resolve it on the host
static int finish_chunk(b200_demod *d, ChunkSet &c, cudaStream_t exec, uint32_t *launches, b200_timing &t) {
    const uint64_t n = c.nsamples;
    const uint32_t ntiles = tiles_for(n);
    size_t dead_cap = c.d_dead.cap, live_cap = c.live_cap, liverec_cap = c.liverec_cap;

    g_trace.mark("wait");
    CUDA_TRY(cudaEventSynchronize(c.ev_small));
    g_trace.mark("ready");
    ScanCounters cnt = *c.h_counters.p;
    {
        float ms = 0;
        cudaEventElapsedTime(&ms, c.ev_begin, c.ev_k1);
        t.scan_ms += ms;
        cudaEventElapsedTime(&ms, c.ev_k1, c.ev_k1b);
        t.slice_ms += ms;
        cudaEventElapsedTime(&ms, c.ev_k1b, c.ev_k2);
        t.classify_ms += ms;
    }
    bool prev_exact = false; // did the run that overflowed already use exact slabs?
    for (int attempt = 0; cnt.overflow; ++attempt) {
        if (attempt >= 3)
            return fail(B200_ERR_CAPACITY, "candidate buffers overflowed after %d attempts (flags 0x%x)", attempt + 1, cnt.overflow);
        bool exact = false;
        size_t cand_total = (size_t) ntiles * kCandSlab, rec_total = (size_t) ntiles * kRecSlab;
        if (cnt.overflow & 3u) {
            // a tile outgrew its slab: every tile reported its true counts, place the slabs exactly
            std::vector<TileDesc> tiles(ntiles);
            CUDA_TRY(cudaMemcpy(tiles.data(), c.d_tiles.p, ntiles * sizeof(TileDesc), cudaMemcpyDeviceToHost));
            std::vector<uint32_t> off(2 * ((size_t) ntiles + 1));
            uint64_t co = 0, ro = 0;
            for (uint32_t i = 0; i < ntiles; ++i) {
                off[2 * i] = (uint32_t) co;
                off[2 * i + 1] = (uint32_t) ro;
                co += tiles[i].ncand;
                // K1b saw only the candidates that fit: a truncated tile gets the upper bound (5 phases each)
                ro += (tiles[i].ncand > kCandSlab && !prev_exact) ? (uint64_t) 5 * tiles[i].ncand : tiles[i].nrec;
            }
            off[2 * (size_t) ntiles] = (uint32_t) co;
            off[2 * (size_t) ntiles + 1] = (uint32_t) ro;
            if (co > 0xfffffff0ull || ro > 0xfffffff0ull)
                return fail(B200_ERR_CAPACITY, "chunk too dense for 32-bit record indices");
            CUDA_TRY(c.d_tile_off.ensure(off.size()));
            CUDA_TRY(cudaMemcpy(c.d_tile_off.p, off.data(), off.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
            cand_total = std::max<size_t>((size_t) co, 1);
            rec_total = std::max<size_t>((size_t) ro, 1);
            exact = true;
            // K2 did not run: size its lists from K1's counts so that the retry cannot fail there
            dead_cap = std::max<size_t>(dead_cap, cand_total);
            live_cap = std::max<size_t>(live_cap, cand_total);
            liverec_cap = std::max<size_t>(liverec_cap, rec_total);
        } else if (cnt.overflow & 32u) {
            // the Mode A/C hit list was too small; its counter kept counting
            d->ac_hit_cap = (size_t) cnt.n_modeac_hits + 4096;
        } else {
            // K2's lists were too small; its counters kept counting past the capacity
            dead_cap = std::max<size_t>(dead_cap, (size_t) cnt.n_dead + 4096);
            live_cap = std::max<size_t>(live_cap, (size_t) cnt.n_live + 4096);
            liverec_cap = std::max<size_t>(liverec_cap, (size_t) cnt.n_liverec + 4096);
        }
        // later chunks may be queued or running on `exec`; this set is idle, so re-issuing it (behind them) is safe
        int rc = issue_chunk(d, c, exec, exact, cand_total, rec_total, dead_cap, live_cap, liverec_cap, launches);
        if (rc != B200_OK)
            return rc;
        prev_exact = exact;
        CUDA_TRY(cudaEventSynchronize(c.ev_small));
        cnt = *c.h_counters.p;
        float ms = 0;
        cudaEventElapsedTime(&ms, c.ev_begin, c.ev_k1);
        t.scan_ms += ms;
        cudaEventElapsedTime(&ms, c.ev_k1, c.ev_k1b);
        t.slice_ms += ms;
        cudaEventElapsedTime(&ms, c.ev_k1b, c.ev_k2);
        t.classify_ms += ms;
    }

    // order_live packed the live positions, their hidden-dead counts and their records in device memory: bring
    // exactly those bytes over (copy engine), and -- while waiting -- do the same for the chunks behind this one
    // whose kernels are already through, so that their bytes are there when their turn comes
    {
        const int rc = download_packed(d, c, cnt);
        if (rc != B200_OK)
            return rc;
    }
    static const int lookahead = [] {
        const char *e = getenv("B200_D2H_LOOKAHEAD");
        const int v = e ? atoi(e) : 1; // one chunk ahead hides the DMA; five at once (50 MB on a dense stream) slow the resolver's own memory traffic by 10 %
        return v < 0 ? 0 : (v > b200_demod::kSets - 1 ? b200_demod::kSets - 1 : v);
    }();
    for (int ahead = 1; ahead <= lookahead; ++ahead) {
        ChunkSet &nx = d->sets[(size_t) ((&c - d->sets) + ahead) % b200_demod::kSets];
        if (!nx.in_flight || nx.packed_issued || cudaEventQuery(nx.ev_small) != cudaSuccess)
            continue;
        const ScanCounters nc = *nx.h_counters.p;
        if (!nc.overflow) {
            const int rc = download_packed(d, nx, nc);
            if (rc != B200_OK)
                return rc;
        }
    }
    cudaGetLastError(); // cudaEventQuery's cudaErrorNotReady is not an error
    g_trace.mark("d2h");
    CUDA_TRY(cudaEventSynchronize(c.ev_packed));
    c.in_flight = false;
    // host: the order-dependent tail
    const double t_res0 = now_ms();
    SpanView v;
    v.nsamples = n;
    v.first_sample = d->first_sample + c.start;
    v.block_samples = d->cfg.block_samples;
    v.final_span = c.final_chunk;
    // the resolver only tells integer block sums (the table converters) from float ones
    v.format = (d->eff_format == 4) ? (uint32_t) B200_INPUT_UC8 : d->eff_format;
    v.live = reinterpret_cast<const LivePos *>(c.h_packed.p);
    v.n_live = (uint32_t) cnt.n_live;
    v.hidden = reinterpret_cast<const LiveHidden *>(c.h_packed.p + (size_t) cnt.n_live * sizeof(LivePos));
    v.liverecs = reinterpret_cast<const LiveRec *>(c.h_packed.p + (size_t) cnt.n_live * (sizeof(LivePos) + sizeof(LiveHidden)));
    v.block_dead = c.h_block_dead.p;
    v.block_sums_u64 = c.h_sums_u64.p;
    v.block_sums_f64 = c.h_sums_f64.p;
    v.ac_hits = d->cfg.mode_ac ? c.h_ac_hits.p : nullptr;
    v.n_ac_hits = d->cfg.mode_ac ? cnt.n_modeac_hits : 0;
    if (const char *dump = getenv("B200_DUMP_SPAN")) {
        // development aid: write the resolver's inputs of this chunk to a file (tools/resolver_bench.cc,
        // b200_host_resolve_dumps)
        char path[512];
        snprintf(path, sizeof(path), "%s/span_%llu.bin", dump, (unsigned long long) v.first_sample); // stream position of the chunk
        if (FILE *f = fopen(path, "wb")) {
            const uint64_t nblocks = n / v.block_samples + 2;
            uint64_t hdr[12] = {n, v.first_sample, v.block_samples, v.final_span, v.format /* as the resolver sees it */, 0, 0, cnt.n_live, cnt.n_liverec, nblocks, 2 /* layout version: live lists in stream order + hidden counts */, 0};
            fwrite(hdr, sizeof(hdr), 1, f);
            fwrite(v.live, sizeof(LivePos), cnt.n_live, f);
            fwrite(v.liverecs, sizeof(LiveRec), cnt.n_liverec, f);
            fwrite(v.hidden, sizeof(LiveHidden), cnt.n_live, f);
            fwrite(v.block_dead, sizeof(BlockDead), nblocks, f);
            fwrite(v.block_sums_u64, sizeof(unsigned long long), 2 * nblocks, f);
            fwrite(v.block_sums_f64, sizeof(double), 2 * nblocks, f);
            fclose(f);
        }
    }
    try {
        d->resolver->resolve(v, d->msgs, d->blocks);
    } catch (const std::bad_alloc &) {
        return fail(B200_ERR_NOMEM, "out of host memory while resolving a chunk");
    }
    t.resolve_ms += (float) (now_ms() - t_res0);
    g_trace.mark("resolved");
    t.d2h_bytes += c.small_d2h_bytes + (size_t) cnt.n_live * (sizeof(LivePos) + sizeof(LiveHidden)) + (size_t) cnt.n_liverec * sizeof(LiveRec);
    t.n_candidates += cnt.n_cand;
    t.n_phase_records += cnt.n_rec;
    t.n_live += cnt.n_live;
    return B200_OK;
}

// The span is resident (or arriving, chunk by chunk, on the copy stream) at d_iq.
static int run_span_impl(b200_demod *d, const uint8_t *d_iq, uint64_t nsamples, uint32_t flags, cudaStream_t exec, const void *host_src);

// A failure after the first chunk was issued leaves the resolver's filter and counters ahead of first_sample, the
// head carry and the message list: quiesce the streams and refuse further spans until b200_demod_reset.
static int run_span(b200_demod *d, const uint8_t *d_iq, uint64_t nsamples, uint32_t flags, cudaStream_t exec, const void *host_src) {
    const int rc = run_span_impl(d, d_iq, nsamples, flags, exec, host_src);
    if (rc != B200_OK) {
        const std::string why = g_last_error; // keep the first error: the synchronisation below may add its own
        cudaStreamSynchronize(exec);
        cudaStreamSynchronize(d->copy_stream);
        cudaStreamSynchronize(d->d2h_stream);
        for (ChunkSet &c : d->sets)
            c.in_flight = false;
        cudaGetLastError();
        g_last_error = why;
        d->failed = true;
    }
    return rc;
}

static int run_span_impl(b200_demod *d, const uint8_t *d_iq, uint64_t nsamples, uint32_t flags, cudaStream_t exec, const void *host_src) {
    const bool final_span = (flags & B200_FLAG_FINAL) != 0;
    const uint32_t B = d->cfg.block_samples;
    const size_t bps = (size_t) d->eff_bps;
    const double t_start = now_ms();

    // whole mag_bufs, and close to two tiles per resident K1a warp: a chunk is then two full waves
    // of the scan kernel instead of one full wave and a ragged one
    const uint64_t k1a_warps = (uint64_t) k1a_warps_per_cta(d->eff_format);
    const uint64_t two_waves = (uint64_t) 2 * d->scan_grid * k1a_warps * kTile - kTile; // tiles_for() adds one tile for the tail
    const uint64_t chunk_target = (d->sm_count > 0 && two_waves >= (16ull << 20) && two_waves <= (64ull << 20)) ? two_waves : kChunkTarget;
    const uint64_t chunk = std::max<uint64_t>(1, chunk_target / B) * B;
    // chunk i = [starts[i], starts[i + 1]), whole mag_bufs, as equal as possible (a short last chunk would run
    // the kernels nearly empty).  For host buffers the pipeline is bound by the H2D copies, so what counts
    // is the work left when the last byte has landed: the end of the span is cut into chunks that shrink
    // (main chunks of at most 32 M samples, then 12 M and 4 M), each small enough to be through the kernels
    // and the resolver while the rest is still being copied (processing a chunk takes about half as long
    // as copying it).
    std::vector<uint64_t> starts;
    {
        const uint64_t nb = (nsamples + B - 1) / B, big = std::max<uint64_t>(1, chunk / B);
        uint64_t tails[2] = {std::max<uint64_t>(1, (12ull << 20) / B), std::max<uint64_t>(1, (4ull << 20) / B)};
        const uint64_t tail_total = tails[0] + tails[1];
        const bool ramp = host_src && nb > tail_total + big / 2;
        const uint64_t main_blocks = ramp ? nb - tail_total : nb;
        const uint64_t main_big = ramp ? std::min<uint64_t>(big, std::max<uint64_t>(1, (32ull << 20) / B)) : big;
        const uint64_t nmain = std::max<uint64_t>(1, (main_blocks + main_big - 1) / main_big);
        uint64_t at = 0; // in mag_bufs
        // one wave of K1a = one tile per resident warp (tiles_for() adds a tile for the tail of a chunk)
        const uint64_t wave = ((uint64_t) d->scan_grid * k1a_warps - 1) * kTile / B;
        if (!ramp && wave > 0 && nb > 4 * wave) {
            // device-resident span: whole waves per chunk, so that only the last launch ends on a partial
            // wave -- one wave first (the host resolver starts early), then three at a time (every chunk costs
            // about 25 us of launch gaps and of order_live's fixed latency; B200_CHUNK_WAVES, measured 2 -> 3:
            // 1.15 -> 1.06 ms for the 144 M-sample stream)
            static const uint64_t per_chunk = [] {
                const char *e = getenv("B200_CHUNK_WAVES");
                const long v = e ? atol(e) : 3;
                return (uint64_t) (v >= 1 && v <= 16 ? v : 3);
            }();
            static const uint64_t first_waves = [] {
                const char *e = getenv("B200_CHUNK_FIRST_WAVES");
                const long v = e ? atol(e) : 1;
                return (uint64_t) (v >= 1 && v <= 16 ? v : 1);
            }();
            starts.push_back(0);
            at = first_waves * wave;
            while (nb - at > (per_chunk + 1) * wave) {
                starts.push_back(at * B);
                at += per_chunk * wave;
            }
            starts.push_back(at * B);
            // ... and optionally a short last chunk (B200_CHUNK_TAIL_DIV: rest / div, a wave at most).  Nothing
            // overlaps the host's resolve of the last chunk, which argued for a small one while the resolver was one
            // thread; now a separate tail costs more in launches and a ragged K1a wave than it hides (measured).
            static const uint64_t tail_div = [] {
                const char *e = getenv("B200_CHUNK_TAIL_DIV");
                return (uint64_t) (e ? atol(e) : 0);
            }();
            const uint64_t rest = nb - at, tail = tail_div ? std::min<uint64_t>(rest / tail_div, wave) : 0;
            if (tail >= 1 && rest > tail)
                starts.push_back((nb - tail) * B);
        } else {
            for (uint64_t i = 0; i < nmain; ++i) {
                starts.push_back(std::min(nsamples, at * B));
                at += main_blocks / nmain + (i < main_blocks % nmain ? 1 : 0);
            }
        }
        if (ramp)
            for (uint64_t tsz : tails) {
                starts.push_back(std::min(nsamples, at * B));
                at += tsz;
            }
        starts.push_back(nsamples);
    }
    const uint64_t nchunks = starts.size() - 1;
    b200_timing t;
    memset(&t, 0, sizeof(t));
    uint32_t launches = 0;
    d->msgs.clear();
    d->blocks.clear();

    // host-buffer entry: all chunk copies are queued up front on the copy stream
    if (host_src) {
        while (d->ev_chunk_h2d.size() < nchunks) {
            cudaEvent_t e;
            CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            d->ev_chunk_h2d.push_back(e);
        }
        CUDA_TRY(cudaEventRecord(d->ev_h2d_begin, d->copy_stream));
        for (uint64_t i = 0; i < nchunks; ++i) {
            const uint64_t s0 = starts[i], s1 = starts[i + 1];
            if (s1 > s0)
                CUDA_TRY(cudaMemcpyAsync(const_cast<uint8_t *>(d_iq) + s0 * bps, (const uint8_t *) host_src + s0 * bps, (s1 - s0) * bps,
                                         cudaMemcpyHostToDevice, d->copy_stream));
            CUDA_TRY(cudaEventRecord(d->ev_chunk_h2d[i], d->copy_stream));
        }
        CUDA_TRY(cudaEventRecord(d->ev_h2d_end, d->copy_stream));
    }

    // float formats, device-resident span: every mag_buf's sequential float sums in one launch up front
    const bool span_sums = !host_src && d->eff_format != B200_INPUT_UC8 && d->eff_format != 4 && nsamples > 0;
    if (span_sums && !d->dc_sums_ready) {
        const uint32_t nb = (uint32_t) ((nsamples + B - 1) / B);
        CUDA_TRY(d->d_span_fsums.ensure(2 * (size_t) nb + 2));
        CUDA_TRY(launch_float_block_sums(d_iq, (uint32_t) d->cfg.input_format, nsamples, B, nb, d->d_span_fsums.p, exec));
        launches += 1;
    }

    auto setup = [&](uint64_t i) -> int {
        ChunkSet &c = d->sets[i % b200_demod::kSets];
        c.start = starts[i];
        c.nsamples = starts[i + 1] - starts[i];
        c.final_chunk = final_span && (i + 1 == nchunks);
        c.iq = d_iq + c.start * bps;
        c.span_fsums = span_sums ? d->d_span_fsums.p + 2 * (c.start / B) : nullptr;
        if (i == 0) {
            c.head = d->d_head.p;
            c.head_valid = d->head_valid;
        } else {
            c.head = c.iq - (size_t) kHead * bps; // the previous chunk's tail, contiguous in the span buffer
            c.head_valid = kHead;
        }
        if (host_src)
            CUDA_TRY(cudaStreamWaitEvent(exec, d->ev_chunk_h2d[i], 0));
        const uint32_t ntiles = tiles_for(c.nsamples);
        return issue_chunk(d, c, exec, false, (size_t) ntiles * kCandSlab, (size_t) ntiles * kRecSlab,
                           std::max<size_t>(c.d_dead.cap, (size_t) (c.nsamples / 24 + 4096)),
                           std::max<size_t>(c.live_cap, (size_t) (c.nsamples / 128 + 4096)),
                           std::max<size_t>(c.liverec_cap, (size_t) (c.nsamples / 64 + 4096)), &launches);
    };

    g_trace.begin();
    int rc = B200_OK;
    uint64_t issued = 0;
    for (uint64_t i = 0; i < nchunks; ++i) {
        // the GPU works through chunks i .. i + kSets - 1 while the host resolves chunk i
        for (; issued < nchunks && issued < i + b200_demod::kSets; ++issued) {
            rc = setup(issued);
            if (rc != B200_OK)
                return rc;
            g_trace.mark("issued");
        }
        rc = finish_chunk(d, d->sets[i % b200_demod::kSets], exec, &launches, t);
        if (rc != B200_OK)
            return rc;
    }
    rc = carry_head(d, d_iq, nsamples, exec);
    if (rc != B200_OK)
        return rc;
    CUDA_TRY(cudaStreamSynchronize(exec));
    g_trace.mark("end");
    g_trace.print(nsamples);

    d->first_sample += nsamples;
    if (final_span)
        d->finished = true;
    if (host_src) {
        CUDA_TRY(cudaEventSynchronize(d->ev_h2d_end));
        cudaEventElapsedTime(&t.h2d_ms, d->ev_h2d_begin, d->ev_h2d_end);
    }
    t.total_ms = (float) (now_ms() - t_start);
    t.scan_launches = launches;
    t.chunks = (uint32_t) nchunks;
    d->timing = t;
    return B200_OK;
}

static int check_span(b200_demod *d, const void *iq, uint64_t nsamples, uint32_t flags) {
    if (!d)
        return fail(B200_ERR_ARG, "null context");
    if (!iq && nsamples)
        return fail(B200_ERR_ARG, "null sample buffer");
    if (d->failed)
        return fail(B200_ERR_STATE, "an earlier span failed half-way (%s); call b200_demod_reset", g_last_error.c_str());
    if (d->finished)
        return fail(B200_ERR_STATE, "the stream already ended (B200_FLAG_FINAL); call b200_demod_reset");
    if (nsamples > d->cfg.max_span_samples)
        return fail(B200_ERR_CAPACITY, "span of %llu samples exceeds max_span_samples %llu", (unsigned long long) nsamples,
                    (unsigned long long) d->cfg.max_span_samples);
    if (!(flags & B200_FLAG_FINAL) && nsamples % d->cfg.block_samples != 0)
        return fail(B200_ERR_ARG, "a non-final span must hold whole blocks of %u samples", d->cfg.block_samples);
    return B200_OK;
}

// --dcfilter: the span's IQ (on the device) -> DC-filtered u16 magnitudes + per-mag_buf float sums, then the
// usual pipeline over the magnitude stream (format 3).  The filter state runs on from span to span.
static int run_span_dc(b200_demod *d, const uint8_t *d_raw, uint64_t nsamples, uint32_t flags, cudaStream_t exec) {
    const uint32_t B = d->cfg.block_samples;
    const size_t padded = (size_t) ((nsamples + 1023) / 1024 * 1024 + 1024);
    CUDA_TRY(d->d_dc_aI.ensure(padded));
    CUDA_TRY(d->d_dc_aQ.ensure(padded));
    CUDA_TRY(d->d_dc_mag.ensure((size_t) nsamples + 128));
    const uint32_t nb = (uint32_t) ((nsamples + B - 1) / B);
    CUDA_TRY(d->d_span_fsums.ensure(2 * (size_t) nb + 2));
    CUDA_TRY(launch_dc_front_end(d_raw, (uint32_t) d->cfg.input_format, nsamples, B, d->dc_a, d->dc_b, d->d_dc_aI.p, d->d_dc_aQ.p,
                                 d->d_dc_state.p, d->d_dc_mag.p, d->d_span_fsums.p, exec));
    d->dc_sums_ready = true;
    const int rc = run_span(d, reinterpret_cast<const uint8_t *>(d->d_dc_mag.p), nsamples, flags, exec, nullptr);
    d->dc_sums_ready = false;
    if (rc == B200_OK)
        d->timing.scan_launches += nsamples ? 3 : 0;
    return rc;
}

extern "C" int b200_demod_process(b200_demod *d, const void *iq, uint64_t nsamples, uint32_t flags) {
    int rc = check_span(d, iq, nsamples, flags);
    if (rc != B200_OK)
        return rc;
    CUDA_TRY(cudaSetDevice(d->cfg.device));
    const size_t bytes = (size_t) nsamples * d->bytes_per_sample;
    if (d->cfg.filter_dc) {
        // the filter is one sequential chain over the span: nothing to overlap the copy with
        CUDA_TRY(d->d_dc_raw.ensure(bytes + 256));
        if (bytes)
            CUDA_TRY(cudaMemcpyAsync(d->d_dc_raw.p, iq, bytes, cudaMemcpyHostToDevice, d->stream));
        return run_span_dc(d, d->d_dc_raw.p, nsamples, flags, d->stream);
    }
    CUDA_TRY(d->d_iq.ensure(bytes + 256));
    return run_span(d, d->d_iq.p, nsamples, flags, d->stream, nsamples ? iq : nullptr);
}

extern "C" void *b200_host_alloc(size_t bytes) {
    void *p = nullptr;
    if (bytes == 0 || cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) {
        cudaGetLastError(); // not sticky: the caller falls back to pageable memory
        return nullptr;
    }
    return p;
}

extern "C" void b200_host_free(void *p) {
    if (p)
        cudaFreeHost(p);
}

extern "C" int b200_demod_process_device(b200_demod *d, const void *d_iq, uint64_t nsamples, uint32_t flags, void *cuda_stream) {
    int rc = check_span(d, d_iq, nsamples, flags);
    if (rc != B200_OK)
        return rc;
    if ((uintptr_t) d_iq & 15u)
        return fail(B200_ERR_ARG, "device span must be 16-byte aligned");
    CUDA_TRY(cudaSetDevice(d->cfg.device));
    cudaStream_t s = cuda_stream ? (cudaStream_t) cuda_stream : d->stream;
    if (d->cfg.filter_dc)
        return run_span_dc(d, (const uint8_t *) d_iq, nsamples, flags, s);
    return run_span(d, (const uint8_t *) d_iq, nsamples, flags, s, nullptr);
}

extern "C" uint64_t b200_demod_message_count(const b200_demod *d) {
    return d ? d->msgs.size() : 0;
}

extern "C" const b200_message *b200_demod_messages(const b200_demod *d) {
    return d ? d->msgs.data() : nullptr;
}

extern "C" uint64_t b200_demod_block_count(const b200_demod *d) {
    return d ? d->blocks.size() : 0;
}

extern "C" const b200_block_info *b200_demod_blocks(const b200_demod *d) {
    return d ? d->blocks.data() : nullptr;
}

extern "C" int b200_demod_get_stats(const b200_demod *d, b200_demod_stats *out) {
    if (!d || !out)
        return fail(B200_ERR_ARG, "null argument");
    *out = d->resolver->stats();
    // reserved[0]: kernel/host CRC disagreements (must stay 0)
    out->reserved[0] = (double) d->resolver->gpu_host_mismatches();
    return B200_OK;
}

extern "C" uint64_t b200_demod_modeac_count(const b200_demod *d) {
    return d ? d->resolver->modeac_count() : 0;
}

extern "C" int b200_demod_get_timing(const b200_demod *d, b200_timing *out) {
    if (!d || !out)
        return fail(B200_ERR_ARG, "null argument");
    *out = d->timing;
    return B200_OK;
}

extern "C" int b200_scan_device(b200_demod *d, const void *d_iq, uint64_t nsamples, int mode, void *cuda_stream, float *ms_out,
                                uint64_t *n_candidates_out) {
    if (!d || !d_iq)
        return fail(B200_ERR_ARG, "null argument");
    if ((uintptr_t) d_iq & 15u)
        return fail(B200_ERR_ARG, "device span must be 16-byte aligned");
    if (nsamples > d->cfg.max_span_samples)
        return fail(B200_ERR_CAPACITY, "span exceeds max_span_samples");
    if (d->cfg.filter_dc)
        return fail(B200_ERR_ARG, "the bare scan entry reads raw IQ: not available with filter_dc");
    CUDA_TRY(cudaSetDevice(d->cfg.device));
    cudaStream_t s = cuda_stream ? (cudaStream_t) cuda_stream : d->stream;
    ChunkSet &c = d->sets[0];
    const size_t nt = tiles_for(nsamples);
    int rc = ensure_chunk_buffers(d, c, nsamples, nt * kCandSlab, nt * kRecSlab, std::max<size_t>(c.d_dead.cap, 4096),
                                  std::max<size_t>(c.live_cap, 4096), std::max<size_t>(c.liverec_cap, 4096));
    if (rc != B200_OK)
        return rc;
    rc = zero_chunk_outputs(d, c, nsamples, s);
    if (rc != B200_OK)
        return rc;
    ScanArgs sa = make_scan_args(d, c, (const uint8_t *) d_iq, d->d_head.p, nsamples, 0, kCandSlab, kRecSlab, nullptr);
    CUDA_TRY(cudaEventRecord(c.ev_begin, s));
    CUDA_TRY(launch_k1a(sa, mode ? 1 : 0, d->scan_grid, s));
    if (mode >= 2)
        CUDA_TRY(launch_slice(make_slice_args(sa), d->slice_grid, s));
    CUDA_TRY(cudaEventRecord(c.ev_k1, s));
    CUDA_TRY(cudaMemcpyAsync(c.h_counters.p, c.d_counters.p, sizeof(ScanCounters), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    float ms = 0;
    CUDA_TRY(cudaEventElapsedTime(&ms, c.ev_begin, c.ev_k1));
    if (ms_out)
        *ms_out = ms;
    if (n_candidates_out)
        *n_candidates_out = c.h_counters.p->n_cand;
    return B200_OK;
}

extern "C" int b200_convert(b200_demod *d, const void *iq, uint32_t nsamples, uint16_t *mag, double *mean_level, double *mean_power) {
    if (!d || (!iq && nsamples) || (!mag && nsamples))
        return fail(B200_ERR_ARG, "null argument");
    CUDA_TRY(cudaSetDevice(d->cfg.device));
    const size_t bytes = (size_t) nsamples * d->bytes_per_sample;
    CUDA_TRY(d->d_frames.ensure(bytes + 16));
    CUDA_TRY(d->d_mag.ensure((size_t) nsamples + 8));
    CUDA_TRY(d->d_csum_u64.ensure(4));
    CUDA_TRY(d->d_csum_f64.ensure(4));
    cudaStream_t s = d->stream;
    CUDA_TRY(cudaMemsetAsync(d->d_csum_u64.p, 0, 2 * sizeof(unsigned long long), s));
    CUDA_TRY(cudaMemsetAsync(d->d_csum_f64.p, 0, 2 * sizeof(double), s));
    if (bytes)
        CUDA_TRY(cudaMemcpyAsync(d->d_frames.p, iq, bytes, cudaMemcpyHostToDevice, s));
    if (d->cfg.filter_dc) {
        // convert_*_generic (convert.c:113-213, 374-423): one call, the filter state carried to the next one
        const size_t padded = ((size_t) nsamples + 1023) / 1024 * 1024 + 1024;
        CUDA_TRY(d->d_dc_aI.ensure(padded));
        CUDA_TRY(d->d_dc_aQ.ensure(padded));
        CUDA_TRY(launch_dc_front_end(d->d_frames.p, (uint32_t) d->cfg.input_format, nsamples, (nsamples + 7u) & ~7u, d->dc_a, d->dc_b,
                                     d->d_dc_aI.p, d->d_dc_aQ.p, d->d_dc_state.p, d->d_mag.p, d->d_csum_f64.p, s));
    } else {
        CUDA_TRY(launch_convert(d->d_frames.p, d->eff_format, nsamples, d->eff_format == 4 ? d->d_lut_q11.p : d->d_lut.p,
                                d->cfg.sc16q11_table_bits, d->d_mag.p, d->d_csum_u64.p, d->d_csum_f64.p, s));
        if (d->eff_format != B200_INPUT_UC8 && d->eff_format != 4 && nsamples) // the float converters' sums in the reference's order
            CUDA_TRY(launch_float_block_sums(d->d_frames.p, (uint32_t) d->cfg.input_format, nsamples, (nsamples + 7u) & ~7u, 1, d->d_csum_f64.p, s));
    }
    unsigned long long su[2] = {0, 0};
    double sf[2] = {0, 0};
    if (nsamples)
        CUDA_TRY(cudaMemcpyAsync(mag, d->d_mag.p, (size_t) nsamples * sizeof(uint16_t), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaMemcpyAsync(su, d->d_csum_u64.p, sizeof(su), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaMemcpyAsync(sf, d->d_csum_f64.p, sizeof(sf), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    if (d->eff_format == B200_INPUT_UC8 || d->eff_format == 4) { // integer sums: convert.c:104-110, 321-327
        if (mean_level)
            *mean_level = su[0] / 65536.0 / nsamples; // convert.c:105
        if (mean_power)
            *mean_power = su[1] / 65535.0 / 65535.0 / nsamples; // convert.c:109
    } else {
        if (mean_level)
            *mean_level = (double) ((float) sf[0] / (float) nsamples); // convert.c:246-252
        if (mean_power)
            *mean_power = (double) ((float) sf[1] / (float) nsamples);
    }
    return B200_OK;
}

extern "C" int b200_uc8_table(b200_demod *d, uint16_t *table65536) {
    if (!d || !table65536)
        return fail(B200_ERR_ARG, "null argument");
    CUDA_TRY(cudaSetDevice(d->cfg.device));
    // what the kernels actually read, not the host copy
    CUDA_TRY(cudaMemcpy(table65536, d->d_lut.p, 65536 * sizeof(uint16_t), cudaMemcpyDeviceToHost));
    return B200_OK;
}

extern "C" int b200_debug_scan(b200_demod *d, const void *iq, uint64_t nsamples, uint8_t *try_masks, b200_phase_record *records,
                               uint64_t record_cap, uint64_t *n_records) {
    if (!d || (!iq && nsamples))
        return fail(B200_ERR_ARG, "null argument");
    if (nsamples > d->cfg.max_span_samples)
        return fail(B200_ERR_CAPACITY, "span exceeds max_span_samples");
    if (d->cfg.filter_dc)
        return fail(B200_ERR_ARG, "the debug scan reads raw IQ: not available with filter_dc");
    CUDA_TRY(cudaSetDevice(d->cfg.device));
    cudaStream_t s = d->stream;
    ChunkSet &c = d->sets[0];
    const size_t bytes = (size_t) nsamples * d->bytes_per_sample;
    CUDA_TRY(d->d_iq.ensure(bytes + 256));
    CUDA_TRY(d->d_dbg_masks.ensure((size_t) nsamples + 16));
    // slabs that can hold every position of a tile as a candidate with five records
    const size_t nt = tiles_for(nsamples);
    int rc = ensure_chunk_buffers(d, c, nsamples, nt * kTile, nt * kTile * 5, std::max<size_t>(c.d_dead.cap, 4096),
                                  std::max<size_t>(c.live_cap, 4096), std::max<size_t>(c.liverec_cap, 4096));
    if (rc != B200_OK)
        return rc;
    rc = zero_chunk_outputs(d, c, nsamples, s);
    if (rc != B200_OK)
        return rc;
    if (bytes)
        CUDA_TRY(cudaMemcpyAsync(d->d_iq.p, iq, bytes, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemsetAsync(d->d_dbg_masks.p, 0, (size_t) nsamples + 16, s));
    ScanArgs sa = make_scan_args(d, c, d->d_iq.p, d->d_head.p, nsamples, 0, kTile, kTile * 5, nullptr);
    sa.dbg_masks = d->d_dbg_masks.p;
    CUDA_TRY(launch_k1a(sa, 1, d->scan_grid, s));
    CUDA_TRY(launch_slice(make_slice_args(sa), d->slice_grid, s));
    CUDA_TRY(cudaMemcpyAsync(c.h_counters.p, c.d_counters.p, sizeof(ScanCounters), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    if (c.h_counters.p->overflow)
        return fail(B200_ERR_CAPACITY, "debug scan overflowed its buffers");
    if (try_masks && nsamples)
        CUDA_TRY(cudaMemcpy(try_masks, d->d_dbg_masks.p, (size_t) nsamples, cudaMemcpyDeviceToHost));
    const uint32_t ntiles = sa.ntiles;
    std::vector<TileDesc> tiles(ntiles);
    if (ntiles)
        CUDA_TRY(cudaMemcpy(tiles.data(), c.d_tiles.p, ntiles * sizeof(TileDesc), cudaMemcpyDeviceToHost));
    uint64_t total = 0;
    std::vector<PhaseRec> tmp;
    for (uint32_t t = 0; t < ntiles; ++t) {
        const TileDesc &td = tiles[t];
        if (!td.nrec)
            continue;
        tmp.resize(td.nrec);
        CUDA_TRY(cudaMemcpy(tmp.data(), c.d_recs.p + td.rec_off, td.nrec * sizeof(PhaseRec), cudaMemcpyDeviceToHost));
        for (uint32_t i = 0; i < td.nrec; ++i) {
            if (records && total < record_cap) {
                b200_phase_record &o = records[total];
                o.position = tmp[i].pos;
                o.crc = tmp[i].w0 & 0xffffffu;
                o.key = tmp[i].w1 & 0xffffffu;
                o.phase = (uint8_t) ((tmp[i].w1 >> 24) & 15u);
                o.kind = (uint8_t) ((tmp[i].w0 >> 24) & 7u);
                o.errors = (uint8_t) ((tmp[i].w0 >> 28) & 3u);
                o.reserved = 0;
            }
            ++total;
        }
    }
    if (n_records)
        *n_records = total;
    return B200_OK;
}

extern "C" int b200_crc_batch(b200_demod *d, const uint8_t *frames14, uint32_t n, uint32_t *syndromes, int8_t *errors, int8_t *bits2) {
    if (!d || (!frames14 && n) || !syndromes || !errors || !bits2)
        return fail(B200_ERR_ARG, "null argument");
    CUDA_TRY(cudaSetDevice(d->cfg.device));
    cudaStream_t s = d->stream;
    CUDA_TRY(d->d_frames.ensure((size_t) n * 14 + 16));
    CUDA_TRY(d->d_syn.ensure((size_t) n + 1));
    CUDA_TRY(d->d_err.ensure((size_t) n + 1));
    CUDA_TRY(d->d_bits.ensure((size_t) 2 * n + 2));
    if (n == 0)
        return B200_OK;
    CUDA_TRY(cudaMemcpyAsync(d->d_frames.p, frames14, (size_t) n * 14, cudaMemcpyHostToDevice, s));
    CUDA_TRY(launch_crc_batch(d->d_frames.p, n, d->d_tab_short.p, (int) d->crc->short_table().size(), d->d_tab_long.p,
                              (int) d->crc->long_table().size(), d->d_syn.p, d->d_err.p, d->d_bits.p, s));
    CUDA_TRY(cudaMemcpyAsync(syndromes, d->d_syn.p, (size_t) n * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaMemcpyAsync(errors, d->d_err.p, (size_t) n, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaMemcpyAsync(bits2, d->d_bits.p, (size_t) 2 * n, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    return B200_OK;
}

extern "C" int b200_error_table(b200_demod *d, int bits, b200_errorinfo *out, int cap) {
    if (!d || (bits != 56 && bits != 112))
        return fail(B200_ERR_ARG, "bits must be 56 or 112");
    const auto &t = (bits == 56) ? d->crc->short_table() : d->crc->long_table();
    for (size_t i = 0; i < t.size() && (int) i < cap && out; ++i) {
        out[i].syndrome = t[i].syndrome;
        out[i].errors = t[i].errors;
        out[i].bit[0] = t[i].bit[0];
        out[i].bit[1] = t[i].bit[1];
        out[i].padding = 0;
    }
    return (int) t.size();
}

// ---- host-only helpers ----

extern "C" int b200_abi_sizeof(int which) {
    switch (which) {
        case 0: return (int) sizeof(b200_message);
        case 1: return (int) sizeof(b200_demod_stats);
        case 2: return (int) sizeof(b200_block_info);
        case 3: return (int) sizeof(b200_timing);
        case 4: return (int) sizeof(b200_phase_record);
        case 5: return (int) sizeof(b200_errorinfo);
        case 6: return (int) sizeof(b200_demod_config);
        default: return -1;
    }
}

extern "C" uint32_t b200_host_checksum(const uint8_t *msg, int bits) {
    static const CrcTables tables(0);
    return tables.checksum(msg, bits);
}

extern "C" int b200_host_error_table(int nfix, int bits, b200_errorinfo *out, int cap) {
    if (nfix < 0 || nfix > 2 || (bits != 56 && bits != 112))
        return fail(B200_ERR_ARG, "bad nfix/bits");
    CrcTables tables(nfix);
    const auto &t = (bits == 56) ? tables.short_table() : tables.long_table();
    for (size_t i = 0; i < t.size() && (int) i < cap && out; ++i) {
        out[i].syndrome = t[i].syndrome;
        out[i].errors = t[i].errors;
        out[i].bit[0] = t[i].bit[0];
        out[i].bit[1] = t[i].bit[1];
        out[i].padding = 0;
    }
    return (int) t.size();
}

extern "C" void b200_host_uc8_table(uint16_t *table65536) {
    build_uc8_table(table65536);
}

extern "C" int b200_host_resolve_dumps(const char *const *paths, uint32_t npaths, int nfix_crc, b200_message *msgs, uint64_t msg_cap,
                                       uint64_t *n_msgs, b200_block_info *blocks, uint64_t block_cap, uint64_t *n_blocks,
                                       b200_demod_stats *stats) {
    if (!paths || !n_msgs || !n_blocks || !stats || nfix_crc < 0 || nfix_crc > 2)
        return fail(B200_ERR_ARG, "bad argument");
    CrcTables crc(nfix_crc);
    Resolver res(&crc, 0);
    MessageList out_msgs;
    std::vector<b200_block_info> out_blocks;
    for (uint32_t i = 0; i < npaths; ++i) {
        // the layout finish_chunk writes: 12-word header, live positions, live records, hidden-dead counts,
        // block dead counters, block sums (u64 and f64)
        FILE *f = fopen(paths[i], "rb");
        uint64_t hdr[12];
        if (!f || fread(hdr, sizeof(hdr), 1, f) != 1) {
            if (f)
                fclose(f);
            return fail(B200_ERR_ARG, "cannot read %s", paths[i]);
        }
        if (hdr[10] != 2) {
            fclose(f);
            return fail(B200_ERR_ARG, "%s is a span dump of another library version (layout %llu, this one reads 2)", paths[i],
                        (unsigned long long) hdr[10]);
        }
        {
            // the counts must describe this file, byte for byte, before anything is sized from them
            long fsize_l = -1;
            if (fseek(f, 0, SEEK_END) != 0 || (fsize_l = ftell(f)) < 0 || fseek(f, (long) sizeof(hdr), SEEK_SET) != 0) {
                fclose(f);
                return fail(B200_ERR_ARG, "cannot size %s", paths[i]);
            }
            const unsigned long long fsize = (unsigned long long) fsize_l;
            const unsigned long long limit = fsize / 4 + 1; // no record is smaller than 4 bytes
            // everything the resolver narrows to 32 bits must fit (a zero or 2^32 block size would divide by zero)
            bool sane = hdr[2] > 0 && hdr[2] <= 0xffffffffull && hdr[0] <= 0xffffffffull && hdr[4] <= 4 && hdr[3] <= 1;
            for (int k = 7; k <= 9; ++k)
                sane = sane && hdr[k] <= limit;
            const unsigned long long want = sizeof(hdr) + hdr[7] * (sizeof(LivePos) + sizeof(LiveHidden)) + hdr[8] * sizeof(LiveRec) +
                                            hdr[9] * (sizeof(BlockDead) + 2 * sizeof(unsigned long long) + 2 * sizeof(double));
            if (!sane || want != fsize) {
                fclose(f);
                return fail(B200_ERR_ARG, "%s is not a span dump of this library version", paths[i]);
            }
        }
        std::vector<LivePos> live(hdr[7]);
        std::vector<LiveRec> recs(hdr[8]);
        std::vector<LiveHidden> hidden(hdr[7]);
        std::vector<BlockDead> bd(hdr[9]);
        std::vector<unsigned long long> su(2 * hdr[9]);
        std::vector<double> sf(2 * hdr[9]);
        size_t got = fread(live.data(), sizeof(LivePos), live.size(), f);
        got += fread(recs.data(), sizeof(LiveRec), recs.size(), f);
        got += fread(hidden.data(), sizeof(LiveHidden), hidden.size(), f);
        got += fread(bd.data(), sizeof(BlockDead), bd.size(), f);
        got += fread(su.data(), sizeof(unsigned long long), su.size(), f);
        got += fread(sf.data(), sizeof(double), sf.size(), f);
        fclose(f);
        if (got != 2 * live.size() + recs.size() + bd.size() + su.size() + sf.size())
            return fail(B200_ERR_ARG, "%s is truncated", paths[i]);
        // every index the resolver follows must stay inside the arrays just read
        bool ok = hdr[9] >= hdr[0] / hdr[2] + 1;
        uint32_t prev_pos = 0;
        for (const LivePos &lp : live) {
            ok = ok && lp.pos < hdr[0] && lp.pos >= prev_pos && (uint64_t) lp.pad + ((lp.info >> 8) & 7u) <= recs.size();
            prev_pos = lp.pos;
        }
        if (!ok)
            return fail(B200_ERR_ARG, "%s is inconsistent", paths[i]);
        SpanView v;
        v.nsamples = hdr[0];
        v.first_sample = hdr[1];
        v.block_samples = (uint32_t) hdr[2];
        v.final_span = hdr[3] != 0;
        v.format = (uint32_t) hdr[4];
        v.live = live.data();
        v.n_live = (uint32_t) live.size();
        v.liverecs = recs.data();
        v.hidden = hidden.data();
        v.block_dead = bd.data();
        v.block_sums_u64 = su.data();
        v.block_sums_f64 = sf.data();
        res.resolve(v, out_msgs, out_blocks);
    }
    *n_msgs = out_msgs.size();
    *n_blocks = out_blocks.size();
    *stats = res.stats();
    stats->reserved[0] = (double) res.gpu_host_mismatches();
    if (out_msgs.size() > msg_cap || out_blocks.size() > block_cap)
        return fail(B200_ERR_CAPACITY, "%zu messages / %zu blocks do not fit the caller's arrays", out_msgs.size(), out_blocks.size());
    if (msgs && !out_msgs.empty())
        memcpy(msgs, out_msgs.data(), out_msgs.size() * sizeof(b200_message));
    if (blocks && !out_blocks.empty())
        memcpy(blocks, out_blocks.data(), out_blocks.size() * sizeof(b200_block_info));
    return B200_OK;
}

extern "C" int b200_host_filter_script(const uint8_t *ops, const uint64_t *args, uint32_t n, uint8_t *results) {
    if (!ops || !args || !results)
        return fail(B200_ERR_ARG, "null argument");
    std::unique_ptr<IcaoFilter> f(new IcaoFilter());
    for (uint32_t i = 0; i < n; ++i) {
        results[i] = 0;
        switch (ops[i]) {
            case 0: f->add((uint32_t) args[i]); break;
            case 1: results[i] = f->test((uint32_t) args[i]) ? 1 : 0; break;
            case 2: f->expire(args[i]); break;
            default: return fail(B200_ERR_ARG, "bad filter op");
        }
    }
    return B200_OK;
}
