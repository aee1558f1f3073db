"""GPU (-m gpu): the CUDA path, called through the C ABI, against the oracle and the goldens.

Bar: bit-exact for everything -- message bytes, CRC, corrected bits, score, 12 MHz and ms
timestamps, every demod counter, the block means and signalLevel (an integer sum divided twice;
north_star allows 1e-5).  The float-path converters' (sc16 / sc16q11) per-block mean_level /
mean_power are sequential float32 sums in the reference (convert.c:228,241-242); the library walks
them in the same order (float_block_sums_kernel), so they and the noise_power_sum derived from them
are bit-exact too.
"""
import numpy as np
import pytest

from conftest import GOLDEN_NAMES, load_golden
from oracle import port, ref
from readsb_protobuf_b200 import api, results, synth

pytestmark = pytest.mark.gpu

FLOAT_SUM_RTOL = 0.0  # sc16 / sc16q11 block means and noise_power_sum: exact as well


def rtol_for(fmt):
    return 0.0 if fmt == "uc8" else FLOAT_SUM_RTOL


def run_gpu(iq, fmt="uc8", span_samples=None, **flags):
    with api.Demodulator(fmt=fmt, **flags) as d:
        got = d.run(iq, span_samples=span_samples)
        assert d.crc_mismatches() == 0
    return got


def assert_parity(got, want, fmt):
    diffs = results.compare_results(got, want, float_rtol=rtol_for(fmt), signal_atol=0.0)
    assert diffs == [], diffs


# ------------------------------------------------------------------------------------------
# whole path
# ------------------------------------------------------------------------------------------

@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_matches_reference_golden(name):
    iq, want, meta = load_golden(name)
    got = run_gpu(iq, meta["fmt"], **meta["flags"])
    assert len(got.msgs) == len(want.msgs) > 0
    assert_parity(got, want, meta["fmt"])
    # ... and so are the bytes the Beast / raw services would send for them (net_io.c:769-896)
    assert api.format_beast(got.msgs) == port.format_beast(want.msgs)
    assert api.format_raw(got.msgs, mlat=True) == port.format_raw(want.msgs, mlat=True)


CASES = [
    dict(cfg=synth.SynthConfig(seed=1, nsamples=2_400_000, frames_per_s=200)),  # BASELINE configs[0]
    dict(cfg=synth.SynthConfig(seed=51, nsamples=3_000_000, frames_per_s=5000, frac_biterror=0.2)),  # dense, configs[3] shape
    dict(cfg=synth.SynthConfig(seed=52, nsamples=1_500_000, fmt="sc16", frames_per_s=2000, frac_biterror=0.2)),
    dict(cfg=synth.SynthConfig(seed=53, nsamples=1_500_000, fmt="sc16q11", frames_per_s=2000, frac_biterror=0.2)),
    dict(cfg=synth.SynthConfig(seed=54, nsamples=1_000_000, frames_per_s=5000, frac_biterror=0.5), nfix=2),
    dict(cfg=synth.SynthConfig(seed=55, nsamples=1_000_000, frames_per_s=5000, frac_biterror=0.5), nfix=0),
    dict(cfg=synth.SynthConfig(seed=56, nsamples=1_000_003, frames_per_s=3000, frac_biterror=0.3), block_samples=50000),
    dict(cfg=synth.SynthConfig(seed=57, nsamples=700_000, frames_per_s=3000), threshold=40),
    dict(cfg=synth.SynthConfig(seed=58, nsamples=700_000, frames_per_s=3000), threshold=120),
    dict(cfg=synth.SynthConfig(seed=59, nsamples=700_000, frames_per_s=3000), threshold=400),
    dict(cfg=synth.SynthConfig(seed=60, nsamples=4 * 131072, frames_per_s=3000)),  # whole blocks: empty final block
    dict(cfg=synth.SynthConfig(seed=61, nsamples=900_000, frames_per_s=8000, noise_sigma=0.1, amp_max=1.4)),  # clipping, overlaps
    dict(cfg=synth.SynthConfig(seed=62, nsamples=600_000, frames_per_s=2000, noise_sigma=0.001)),  # near silence
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: f"{c['cfg'].fmt}-s{c['cfg'].seed}")
def test_matches_oracle(case):
    case = dict(case)
    cfg = case.pop("cfg")
    iq, _ = synth.generate(cfg)
    want = port.run(iq, cfg.fmt, **case)
    got = run_gpu(iq, cfg.fmt, **case)
    assert_parity(got, want, cfg.fmt)


@pytest.mark.parametrize("n", [0, 1, 7, 18, 19, 100, 325, 326, 327, 328, 329, 600, 8191, 8192, 8193, 8520, 16384 + 5])
def test_ragged_and_empty_inputs(n):
    cfg = synth.SynthConfig(seed=70 + n % 13, nsamples=max(n, 1), frames_per_s=20000, noise_sigma=0.05)
    iq = synth.generate(cfg)[0][: 2 * n]
    want = port.run(iq, "uc8")
    got = run_gpu(iq, "uc8")
    assert_parity(got, want, "uc8")


def test_span_split_is_invisible():
    """Feeding the stream span by span (overlap carry, filter state, counters) equals one shot."""
    cfg = synth.SynthConfig(seed=81, nsamples=2_000_000, frames_per_s=5000, frac_biterror=0.2)
    iq, _ = synth.generate(cfg)
    want = port.run(iq, "uc8")
    for span in (131072, 131072 * 4):
        assert_parity(run_gpu(iq, "uc8", span_samples=span), want, "uc8")
    # small blocks, spans shorter than the 326-sample overlap history
    want = port.run(iq[:200_000], "uc8", block_samples=256)
    assert_parity(run_gpu(iq[:200_000], "uc8", span_samples=256, block_samples=256), want, "uc8")
    # a float format, one mag_buf per call like a live SDR, with Mode A/C on: the per-block float sums and
    # the thresholds derived from them must not depend on how the stream is cut
    cfg = synth.SynthConfig(seed=85, nsamples=1_500_000, fmt="sc16q11", frames_per_s=3000, frac_biterror=0.2, modeac_per_s=2000)
    iq, _ = synth.generate(cfg)
    want = port.run(iq, "sc16q11", modeac=True)
    for span in (131072, 131072 * 3):
        assert_parity(run_gpu(iq, "sc16q11", span_samples=span, modeac=True), want, "sc16q11")


@pytest.mark.parametrize("fmt", ["sc16", "sc16q11"])
@pytest.mark.parametrize("kind", ["ties", "constant", "silence_then_full_scale", "random_few_bits"])
def test_float_block_sums_on_inputs_built_to_break_them(fmt, kind):
    """The float converters' sequential sums (convert.c:228,241-242) are reproduced batch-parallel
    (float_block_sums_kernel: whole batches rounded to the running sum's ulp at once); the batches it has to
    walk one by one are those with a tie -- a sample exactly half-way between two ulps of the sum -- and those
    that carry the sum into the next binade.  These inputs are made of them."""
    rng = np.random.default_rng(5)
    n = 3 * 131072 + 4001
    scale = 32768 if fmt == "sc16" else 2048
    I = np.zeros(n, dtype=np.int64)
    Q = np.zeros(n, dtype=np.int64)
    if kind == "ties":
        # Q = 0: mag = |I| / scale exactly, an odd multiple of 1 / scale -- once the level sum is in the binade whose
        # half-ulp is 1 / scale every sample is a tie (sc16: [512, 1024), reached after ~8 K samples of mag 0.06)
        I = (2 * rng.integers(scale // 40, scale // 12, size=n) + 1) * rng.choice([-1, 1], size=n)
    elif kind == "constant":
        I[:] = scale // 8 + 1
        Q[:] = -(scale // 16)
    elif kind == "silence_then_full_scale":
        I[n // 3:] = scale - 1  # sums stay 0 for a block, then cross a binade every few batches
        Q[n // 3:] = -(scale - 1)
        I[2 * n // 3:] = 3
        Q[2 * n // 3:] = 0
    else:
        I = rng.integers(-8, 9, size=n) * (scale // 64)
        Q = rng.integers(-8, 9, size=n) * (scale // 64)
    iq = np.empty(2 * n, dtype="<i2")
    iq[0::2] = I
    iq[1::2] = Q
    iq = iq.view(np.uint8)
    want = port.run(iq, fmt)
    got = run_gpu(iq, fmt)
    assert_parity(got, want, fmt)
    assert np.array_equal(got.blocks["mean_level"], want.blocks["mean_level"])
    assert np.array_equal(got.blocks["mean_power"], want.blocks["mean_power"])
    # one mag_buf per call (the host-buffer path runs the kernel per chunk)
    assert_parity(run_gpu(iq, fmt, span_samples=131072), want, fmt)


@pytest.mark.parametrize("seed,nsamples,block", [(311, 1_000_000, 131072), (312, 600_001, 50000), (313, 4 * 131072, 131072)])
def test_modeac_matches_oracle(seed, nsamples, block):
    """--modeac (demodulate2400AC, demod_2400.c:522-708): a block's Mode A/C replies follow its Mode S
    messages; bit-exact for uc8 (the noise level comes from the exact integer block sums)."""
    cfg = synth.SynthConfig(seed=seed, nsamples=nsamples, frames_per_s=2000, frac_biterror=0.2, modeac_per_s=4000)
    iq, _ = synth.generate(cfg)
    want = port.run(iq, "uc8", block_samples=block, modeac=True)
    assert int(np.sum(want.msgs["msgtype"] == 32)) > 100
    with api.Demodulator(block_samples=block, modeac=True) as d:
        got = d.run(iq)
        assert d.modeac_count() == int(np.sum(want.msgs["msgtype"] == 32))
    assert_parity(got, want, "uc8")
    # span by span (one mag_buf per call, like a live SDR) gives the same list
    with api.Demodulator(block_samples=block, modeac=True) as d:
        assert_parity(d.run(iq, span_samples=block), want, "uc8")


@pytest.mark.parametrize("fmt", ["sc16", "sc16q11"])
def test_modeac_float_formats(fmt):
    """Mode A/C thresholds derive from the block's mean level / power, which for the float converters
    are order-dependent float sums: exact only because the library sums in the reference's order."""
    cfg = synth.SynthConfig(seed=314, nsamples=800_000, fmt=fmt, frames_per_s=1500, modeac_per_s=4000, frac_biterror=0.1)
    iq, _ = synth.generate(cfg)
    want = port.run(iq, fmt, modeac=True)
    assert int(np.sum(want.msgs["msgtype"] == 32)) > 100
    assert_parity(run_gpu(iq, fmt, modeac=True), want, fmt)


def test_modeac_off_by_default_and_dense_hits():
    """Without the flag no Mode A/C kernel runs; with it, a pulse train that decodes almost everywhere
    exercises the grow-and-retry of the hit list and the 69-sample skip."""
    n = 400_000
    i = np.full(n, 128, dtype=np.uint8)
    # F1 and F2 pulses every 87 * 14 / 25 = 48.72 samples: period 1218 ticks of 60 MHz
    t = (np.arange(0, n * 25 - 2000, 1218) // 25).astype(np.int64)
    i[t] = 250
    i[t + 1] = 200
    iq = np.stack([i, np.full(n, 128, dtype=np.uint8)], axis=1).reshape(-1)
    want = port.run(iq, "uc8", modeac=True)
    assert int(np.sum(want.msgs["msgtype"] == 32)) > 1000
    assert_parity(run_gpu(iq, "uc8", modeac=True), want, "uc8")
    plain = run_gpu(iq, "uc8")
    assert int(np.sum(plain.msgs["msgtype"] == 32)) == 0
    assert_parity(plain, port.run(iq, "uc8"), "uc8")


def with_dc_offset(iq, fmt, di, dq):
    """The same stream seen through a receiver with a DC offset on both rails (what --dcfilter is for)."""
    if fmt == "uc8":
        v = iq.astype(np.int32).reshape(-1, 2) + [di, dq]
        return np.clip(v, 0, 255).astype(np.uint8).reshape(-1)
    full = 32767 if fmt == "sc16" else 2047
    v = iq.view("<i2").astype(np.int32).reshape(-1, 2) + [di, dq]
    return np.clip(v, -full - 1, full).astype("<i2").reshape(-1).view(np.uint8)


@pytest.mark.parametrize("fmt,di,dq", [("uc8", 9, -6), ("sc16", 1500, -900), ("sc16q11", -120, 75)])
def test_dcfilter_matches_oracle(fmt, di, dq):
    """--dcfilter: convert_*_generic (convert.c:113-213, 374-423).  The DC block is a float
    recurrence over the whole stream; the library walks the same chain, so magnitudes, block means and the
    message list are bit-exact, however the stream is cut into spans, with Mode A/C on top."""
    cfg = synth.SynthConfig(seed=401, nsamples=1_800_000, fmt=fmt, frames_per_s=3000, frac_biterror=0.2, modeac_per_s=1500)
    iq = with_dc_offset(synth.generate(cfg)[0], fmt, di, dq)
    want = port.run(iq, fmt, dcfilter=True, modeac=True)
    assert len(want.msgs) > 500
    assert results.compare_results(want, port.run(iq, fmt, modeac=True)) != []  # the filter changes the outcome
    assert_parity(run_gpu(iq, fmt, dcfilter=True, modeac=True), want, fmt)
    for span in (131072, 131072 * 5):
        assert_parity(run_gpu(iq, fmt, span_samples=span, dcfilter=True, modeac=True), want, fmt)
    assert_parity(run_gpu(iq, fmt, dcfilter=True), port.run(iq, fmt, dcfilter=True), fmt)


@pytest.mark.parametrize("n,block", [(0, 131072), (1, 131072), (5, 8), (1023, 256), (1024, 1024), (1025, 512), (40_001, 131072),
                                     (262_144, 131072)])
def test_dcfilter_ragged_lengths(n, block):
    cfg = synth.SynthConfig(seed=410 + n % 7, nsamples=max(n, 1), frames_per_s=20000, noise_sigma=0.05)
    iq = with_dc_offset(synth.generate(cfg)[0][: 2 * n], "uc8", 7, 3)
    want = port.run(iq, "uc8", dcfilter=True, block_samples=block)
    assert_parity(run_gpu(iq, "uc8", dcfilter=True, block_samples=block), want, "uc8")


@pytest.mark.parametrize("fmt", ["uc8", "sc16", "sc16q11"])
def test_dcfilter_converter_bit_exact(fmt):
    """The iq_convert_fn boundary with filter_dc: magnitudes and means of consecutive calls (the filter
    state runs on from call to call, struct converter_state convert.c:25-30)."""
    rng = np.random.default_rng(19)
    calls = [131072, 131072, 50_000, 1, 1030, 7]
    n = sum(calls)
    if fmt == "uc8":
        iq = rng.integers(0, 256, 2 * n, dtype=np.uint8)
        iq[:512] = np.repeat(np.array([0, 255, 127, 128], dtype=np.uint8), 128)
    else:
        full = 32767 if fmt == "sc16" else 2047
        v = rng.integers(-full - 1, full + 1, 2 * n).astype("<i2")
        v[:8] = [full, full, -full - 1, -full - 1, 0, 0, full, 0]
        iq = v.view(np.uint8)
    want_mag, want_means = port.convert_dc(iq, fmt, calls)
    bps = 2 if fmt == "uc8" else 4
    with api.Demodulator(fmt=fmt, dcfilter=True) as d:
        at = 0
        for c, (wl, wp) in zip(calls, want_means):
            mag, ml, mp = d.convert(iq[at * bps: (at + c) * bps])
            assert np.array_equal(mag, want_mag[at: at + c])
            assert ml == wl and mp == wp
            at += c
        # a reset starts the filter from zero again (init_converter, convert.c:473-474)
        d.reset()
        mag, _, _ = d.convert(iq[: 1000 * bps])
        assert np.array_equal(mag, want_mag[:1000])


@pytest.mark.parametrize("bits", [8, 7, 4, 9, 10, 11])
def test_sc16q11_table_converter_matches_oracle(bits):
    """sc16q11 through the magnitude table of a -DSC16Q11_TABLE_BITS build
    (convert_sc16q11_table, convert.c:264-328; the armhf package uses 8 bits): integer block sums,
    table in K1a's shared memory like the uc8 one; with 9..11 bits (the reference's convert_benchmark rows) the
    table no longer fits there and is read through L2."""
    cfg = synth.SynthConfig(seed=420 + bits, nsamples=1_500_000, fmt="sc16q11", frames_per_s=3000, frac_biterror=0.2,
                            modeac_per_s=1500, amp_max=1.3)
    iq, _ = synth.generate(cfg)
    want = port.run(iq, "sc16q11", table_bits=bits, modeac=True)
    assert len(want.msgs) > (500 if bits >= 7 else 0)
    assert_parity(run_gpu(iq, "sc16q11", table_bits=bits, modeac=True), want, "sc16q11")
    assert_parity(run_gpu(iq, "sc16q11", table_bits=bits, modeac=True, span_samples=131072 * 3), want, "sc16q11")
    # --dcfilter picks the generic float converter in a table build too (converters_table, convert.c:432-437)
    assert_parity(run_gpu(iq, "sc16q11", table_bits=bits, dcfilter=True), port.run(iq, "sc16q11", table_bits=bits, dcfilter=True),
                  "sc16q11")
    # the converter boundary: every table entry, the sign folds, -32768 and the & 2047 wrap
    v = np.zeros((70_000, 2), dtype="<i2")
    k = np.arange(65536)
    v[:65536, 0] = (k >> 8) << 3
    v[:65536, 1] = (k & 255) << 3
    if bits > 8:  # the low bits matter now: walk them too
        v[:65536, 0] |= (k * 5) & 7
        v[:65536, 1] |= (k * 3) & 7
    v[65536:65540] = [[-32768, 32767], [-2048, 2048], [-1, -2047], [4095, -4096]]
    v[65540:] = np.random.default_rng(bits).integers(-32768, 32768, (70_000 - 65540, 2))
    with api.Demodulator(fmt="sc16q11", table_bits=bits) as d:
        mag, ml, mp = d.convert(v.reshape(-1).view(np.uint8))
    wmag, wl, wp = port.convert_sc16q11_table(v.reshape(-1).view(np.uint8), bits)
    assert np.array_equal(mag, wmag) and ml == wl and mp == wp


def test_dcfilter_device_resident_input():
    import torch
    cfg = synth.SynthConfig(seed=402, nsamples=1_000_000, fmt="sc16", frames_per_s=4000, frac_biterror=0.2)
    iq = with_dc_offset(synth.generate(cfg)[0], "sc16", 800, 800)
    want = port.run(iq, "sc16", dcfilter=True)
    dev = torch.from_numpy(iq).cuda()
    with api.Demodulator(fmt="sc16", dcfilter=True) as d:
        r = d.process_device(dev.data_ptr(), cfg.nsamples, final=True, stream=torch.cuda.current_stream().cuda_stream)
        st = d.stats().copy()
    st["convert_cpu_s"] = st["demod_cpu_s"] = 0
    assert_parity(results.DemodResult(r.msgs, st, r.blocks, cfg.nsamples), want, "sc16")


@pytest.mark.parametrize("seed", range(24))
def test_randomized_configurations(seed):
    """Random corners of the parameter space: format, repair depth, threshold, mag_buf size, stream
    length (ragged), traffic density, noise level, Mode A/C, span size -- all against the oracle, bit-exact."""
    rng = np.random.default_rng(1000 + seed)
    fmt = ["uc8", "uc8", "sc16", "sc16q11"][int(rng.integers(0, 4))]
    block = int(rng.integers(125, 25000)) * 8
    nsamples = int(rng.integers(1, 1_200_000))
    modeac = bool(rng.integers(0, 2))
    cfg = synth.SynthConfig(seed=2000 + seed, nsamples=nsamples, fmt=fmt, frames_per_s=float(rng.choice([50, 500, 3000, 9000])),
                            frac_biterror=float(rng.choice([0.0, 0.2, 0.6])), noise_sigma=float(rng.choice([0.002, 0.02, 0.08])),
                            amp_max=float(rng.choice([0.3, 0.9, 1.3])), n_icao=int(rng.choice([3, 200])),
                            modeac_per_s=float(rng.choice([0, 2000])) if modeac else 0.0)
    flags = dict(nfix=int(rng.integers(0, 3)), threshold=int(rng.choice([40, 58, 75, 130, 400])), block_samples=block)
    iq, _ = synth.generate(cfg)
    want = port.run(iq, fmt, modeac=modeac, **flags)
    span = None if rng.integers(0, 2) else block * int(rng.integers(1, 6))
    got = run_gpu(iq, fmt, span_samples=span, modeac=modeac, **flags)
    assert_parity(got, want, fmt)
    if seed % 3 == 0:  # the same corner behind the DC-filter front end
        want = port.run(iq, fmt, modeac=modeac, dcfilter=True, **flags)
        assert_parity(run_gpu(iq, fmt, span_samples=span, modeac=modeac, dcfilter=True, **flags), want, fmt)


def test_icao_filter_flips_across_minutes():
    """> 120 s of stream so that addresses age out (icao_filter.c:150-164) -- sparse, to stay fast."""
    cfg = synth.SynthConfig(seed=82, nsamples=int(130 * 2.4e6), frames_per_s=20, n_icao=5, noise_sigma=0.004)
    iq, _ = synth.generate(cfg)
    want = port.run(iq, "uc8")
    got = run_gpu(iq, "uc8", max_span_samples=cfg.nsamples + 1024)
    assert len(want.msgs) > 500
    assert_parity(got, want, "uc8")


def test_adversarial_candidate_density():
    """A pulse train that makes most positions preamble candidates: exercises K1's slow rounds and the
    grow-and-retry of the candidate buffers."""
    n = 300_000
    pattern = np.array([200, 128, 128, 128, 140, 128, 140, 128, 140, 250, 250], dtype=np.uint8)  # ~27 % of positions
    i = np.tile(pattern, n // len(pattern) + 1)[:n]
    i[100_000:200_000] = 128  # a quiet stretch in the middle: fast and slow tiles in one span
    q = np.full(n, 128, dtype=np.uint8)
    iq = np.stack([i, q], axis=1).reshape(-1)
    want = port.run(iq, "uc8")
    assert int(want.stats["demod_preambles"]) > n // 8
    assert_parity(run_gpu(iq, "uc8"), want, "uc8")


def test_device_resident_input_equals_host_input():
    import torch
    cfg = synth.SynthConfig(seed=83, nsamples=1_200_000, frames_per_s=4000, frac_biterror=0.2)
    iq, _ = synth.generate(cfg)
    want = port.run(iq, "uc8")
    dev = torch.from_numpy(iq).cuda()
    with api.Demodulator() as d:
        r = d.process_device(dev.data_ptr(), cfg.nsamples, final=True, stream=torch.cuda.current_stream().cuda_stream)
        st = d.stats().copy()
        t = r.timing
    st["convert_cpu_s"] = st["demod_cpu_s"] = 0
    got = results.DemodResult(r.msgs, st, r.blocks, cfg.nsamples)
    assert_parity(got, want, "uc8")
    assert t["scan_launches"] == 5 and t["scan_ms"] > 0 and t["n_candidates"] > 0


def require_ref():
    """The reference-dependent tests must not pass by skipping: oracle/_ref travels with the snapshot (it is
    git-ignored, not gpurun-ignored), so its absence on a GPU box is a packaging error."""
    assert ref.available(), "oracle/_ref/ref_demod is missing on the GPU box: run __graft_entry__.build() before the snapshot"


def test_full_size_stream_matches_the_reference_itself():
    """BASELINE configs[1] at full size (60 s, 144 M samples) against the unmodified reference."""
    require_ref()
    cfg = synth.baseline_config(1)
    iq, frames = synth.generate(cfg)
    want = ref.run(iq, "uc8")
    got = run_gpu(iq, "uc8", max_span_samples=cfg.nsamples + 1024)
    assert len(want.msgs) > 0.7 * len(frames)
    assert_parity(got, want, "uc8")
    # size-independent property: span-by-span equals one shot (checksum of checksums)
    again = run_gpu(iq, "uc8", span_samples=131072 * 64)
    assert np.bitwise_xor.reduce(again.msgs["crc"] ^ again.msgs["timestampMsg"].astype(np.uint32)) == \
        np.bitwise_xor.reduce(got.msgs["crc"] ^ got.msgs["timestampMsg"].astype(np.uint32))
    assert_parity(again, want, "uc8")


@pytest.mark.parametrize("fmt", ["sc16", "sc16q11"])
def test_full_size_sc16_stream_matches_the_reference_itself(fmt):
    """BASELINE configs[2] at full size (60 s of sc16 / sc16q11, 576 MB) against the unmodified reference,
    block means included (the float sums are order-dependent)."""
    require_ref()
    import dataclasses
    cfg = dataclasses.replace(synth.baseline_config(2), fmt=fmt)
    iq, frames = synth.generate(cfg)
    want = ref.run(iq, cfg.fmt)
    got = run_gpu(iq, cfg.fmt, max_span_samples=cfg.nsamples + 1024)
    assert len(want.msgs) > 0.7 * len(frames)
    assert_parity(got, want, cfg.fmt)


def test_full_size_dense_stream_matches_the_reference_itself():
    """BASELINE configs[3] at full size: 600 s of uc8 (1.44 G samples, 2.88 GB), 5000 frames/s with overlaps,
    20 % of them with one flipped bit -- several hundred mag_buf-aligned pipeline chunks, ten ICAO-filter
    flips, millions of skip-aheads -- against the unmodified reference at zero tolerance.  Then the
    size-independent property: the same stream fed span by span gives the same messages."""
    require_ref()
    cfg = synth.baseline_config(3)
    iq, frames = synth.generate(cfg)
    want = ref.run(iq, "uc8")
    got = run_gpu(iq, "uc8", max_span_samples=cfg.nsamples + 1024)
    assert len(want.msgs) > 0.4 * len(frames) and int(want.stats["demod_accepted"][1]) > 200_000
    assert_parity(got, want, "uc8")
    again = run_gpu(iq, "uc8", span_samples=131072 * 512)
    assert_parity(again, want, "uc8")


def test_full_size_aggressive_fix_matches_the_reference_itself():
    """configs[1] at full size with --aggressive (nfix 2: the 1326 / 3831-entry syndrome tables)."""
    require_ref()
    cfg = synth.SynthConfig(seed=6, nsamples=int(60 * 2.4e6), frames_per_s=1000, frac_biterror=0.5)
    iq, frames = synth.generate(cfg)
    want = ref.run(iq, "uc8", nfix=2)
    got = run_gpu(iq, "uc8", nfix=2, max_span_samples=cfg.nsamples + 1024)
    assert int(want.stats["demod_accepted"][2]) > 0
    assert_parity(got, want, "uc8")


def test_dense_traffic_two_minutes():
    """BASELINE configs[3] shape (5000 frames/s with overlaps, 20 % with a flipped bit) over 120 s of
    stream: several pipeline chunks, two ICAO-filter flips, the 1-bit repair path on thousands of frames."""
    cfg = synth.SynthConfig(seed=4, nsamples=int(120 * 2.4e6), frames_per_s=5000, frac_biterror=0.2)
    iq, frames = synth.generate(cfg)
    want = port.run(iq, "uc8")
    got = run_gpu(iq, "uc8", max_span_samples=cfg.nsamples + 1024)
    assert len(want.msgs) > 0.4 * len(frames) and int(want.stats["demod_accepted"][1]) > 20_000
    assert_parity(got, want, "uc8")


# ------------------------------------------------------------------------------------------
# kernel-level parity
# ------------------------------------------------------------------------------------------

def test_uc8_table_exhaustive():
    with api.Demodulator() as d:
        assert np.array_equal(d.uc8_table(), port.uc8_table())


@pytest.mark.parametrize("fmt", ["uc8", "sc16", "sc16q11"])
def test_converter_bit_exact(fmt):
    rng = np.random.default_rng(9)
    n = 300_001
    if fmt == "uc8":
        iq = rng.integers(0, 256, 2 * n, dtype=np.uint8)
        iq[:512] = np.repeat(np.array([0, 255, 127, 128], dtype=np.uint8), 128)  # corners / centre
    else:
        full = 32767 if fmt == "sc16" else 2047
        v = rng.integers(-full - 1, full + 1, 2 * n).astype("<i2")
        v[:8] = [full, full, -full - 1, -full - 1, 0, 0, full, 0]  # clamp at 1.0, zero
        iq = v.view(np.uint8)
    with api.Demodulator(fmt=fmt) as d:
        mag, ml, mp = d.convert(iq)
    want, wl, wp = port.convert(iq, fmt)
    assert np.array_equal(mag, want)
    assert ml == wl and mp == wp  # the float formats too: sums taken in the reference's order


def test_try_masks_and_phase_records():
    """K1's per-position try mask and its (position, phase) class records against the oracle's
    preamble test, slicer and CRC on every candidate."""
    cfg = synth.SynthConfig(seed=84, nsamples=400_000, frames_per_s=5000, frac_biterror=0.3)
    iq, _ = synth.generate(cfg)
    with api.Demodulator() as d:
        masks, recs = d.debug_scan(iq)
        tab_short, tab_long = d.error_table(56), d.error_table(112)
    mag, _, _ = port.convert(iq, "uc8")
    m = np.concatenate([np.zeros(326, np.uint16), mag, np.zeros(64, np.uint16)])  # stream start: zero overlap
    want_masks = port.try_masks(m[: cfg.nsamples + 18])
    assert np.array_equal(masks, want_masks[: cfg.nsamples])
    assert 0.002 < (masks != 0).mean() < 0.1

    short_syn = {int(e["syndrome"]): e for e in tab_short}
    long_syn = {int(e["syndrome"]): e for e in tab_long}
    got = {(int(r["position"]), int(r["phase"])): r for r in recs}
    assert len(got) == len(recs)
    n_expected = 0
    for j in np.nonzero(masks)[0]:
        for ph in range(4, 9):
            if not (masks[j] >> (ph - 4)) & 1:
                continue
            b0 = port.slice_bytes(m, int(j), ph, 1)[0]
            df = b0 >> 3
            nbytes = 7 if df in (0, 4, 5, 11) else 14 if df in (16, 17, 18, 20, 21, 24) else 0
            kind, key, errors = 0, None, 0
            if nbytes:
                msg = port.slice_bytes(m, int(j), ph, nbytes)
                crc = port.checksum(msg)
                aa = int.from_bytes(msg[1:4], "big")
                if any(msg):
                    if df in (0, 4, 5, 16, 24):
                        kind, key = api.KIND_AP, crc
                    elif df in (20, 21):
                        kind, key = api.KIND_AP_COMMB, crc
                    else:
                        syn = crc & 0xFFFF80 if df == 11 else crc
                        tab = short_syn if df == 11 else long_syn
                        e = None if syn == 0 else tab.get(syn)
                        if syn == 0 or (e is not None and (df != 11 or e["errors"] <= 1)):
                            kind = api.KIND_DF11 if df == 11 else api.KIND_ES
                            errors = 0 if syn == 0 else int(e["errors"])
                            key = aa
                            for b in ([] if syn == 0 else [int(x) for x in e["bit"][:errors]]):
                                if 8 <= b <= 31:
                                    key ^= 1 << (31 - b)
            if kind:
                n_expected += 1
                r = got.get((int(j), ph))
                assert r is not None, (j, ph)
                assert (int(r["crc"]), int(r["kind"]), int(r["key"]), int(r["errors"])) == (crc, kind, key, errors)
    assert n_expected == len(recs) > 100


@pytest.mark.parametrize("nfix", [0, 1, 2])
def test_crc_batch_and_tables(nfix):
    rng = np.random.default_rng(10 + nfix)
    with api.Demodulator(nfix=nfix) as d:
        for bits in (56, 112):
            a, b = d.error_table(bits), port.error_table(nfix, bits)
            assert all(np.array_equal(a[f], b[f]) for f in ("syndrome", "errors", "bit"))
        frames = rng.integers(0, 256, (4000, 14), dtype=np.uint8)
        # plant valid frames with 0, 1 and 2 flipped bits
        for i in range(0, 3000):
            nb = 14 if frames[i, 0] & 0x80 else 7
            crc = port.checksum(bytes(frames[i, :nb - 3]) + b"\0\0\0")
            frames[i, nb - 3:nb] = [crc >> 16, (crc >> 8) & 255, crc & 255]
            for _ in range(i % 3):
                bit = int(rng.integers(5, nb * 8))
                frames[i, bit >> 3] ^= 0x80 >> (bit & 7)
        syn, err, bits2 = d.crc_batch(frames)
    tabs = {56: {int(e["syndrome"]): e for e in port.error_table(nfix, 56)},
            112: {int(e["syndrome"]): e for e in port.error_table(nfix, 112)}}
    for i in range(len(frames)):
        nb = 14 if frames[i, 0] & 0x80 else 7
        want = port.checksum(bytes(frames[i, :nb]))
        assert int(syn[i]) == want
        if want == 0:
            assert err[i] == 0
        else:
            e = tabs[nb * 8].get(want)
            assert err[i] == (-1 if e is None else e["errors"])
            if e is not None:
                assert list(bits2[i]) == list(e["bit"])
    assert (err[:3000:3] == 0).all()


# ------------------------------------------------------------------------------------------
# drop-in: the reference's own program on top of the shim
# ------------------------------------------------------------------------------------------

def test_reference_program_with_the_shim_across_filter_flips():
    """The drop-in binary over 130 s of stream: two ICAO-filter generations pass, readsb's own copy of the filter
    (expired by its main loop) must never make it drop a frame the library accepted; output equals the oracle's."""
    import os
    import subprocess
    import tempfile
    from readsb_protobuf_b200 import build
    exe = build.ORACLE / "_ref" / "readsb_b200"
    assert exe.exists(), "oracle/_ref/readsb_b200 is missing on the GPU box (built by __graft_entry__.build() where /root/reference exists)"
    cfg = synth.SynthConfig(seed=93, nsamples=int(130 * 2.4e6), frames_per_s=400, frac_biterror=0.2, n_icao=60)
    iq, _ = synth.generate(cfg)
    want = port.run(iq, "uc8")
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "in.bin")
        iq.tofile(path)
        out = subprocess.run([str(exe), "--device-type", "ifile", "--ifile", path, "--preamble-threshold", "58", "--raw", "--mlat"],
                             capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("@")]
    expect = ["@%012X%s;" % (int(m["timestampMsg"]), bytes(m["msg"][: m["msgbits"] // 8]).hex()) for m in want.msgs]
    assert len(expect) > 20000
    assert lines == expect
    assert "disagrees" not in out.stderr


@pytest.mark.parametrize("modeac,dcfilter", [(False, False), (True, False), (True, True)], ids=["modes", "modeac", "modeac-dcfilter"])
def test_reference_program_with_the_shim_prints_the_same_messages(modeac, dcfilter):
    """oracle/_ref/readsb_b200 = readsb's main(), FIFO, CRC, field decoder and tracker objects linked
    with readsb_protobuf_b200/shim/readsb_b200_shim.c + libreadsb_b200.so instead of convert.o,
    demod_2400.o and sdr_ifile.o.  Its --raw --mlat output must equal the reference path's."""
    import os
    import subprocess
    import tempfile
    from readsb_protobuf_b200 import build

    exe = build.ORACLE / "_ref" / "readsb_b200"
    assert exe.exists(), "oracle/_ref/readsb_b200 is missing on the GPU box (built by __graft_entry__.build() where /root/reference exists)"
    cfg = synth.SynthConfig(seed=91, nsamples=3_000_000, frames_per_s=3000, frac_biterror=0.2,
                            modeac_per_s=3000 if modeac else 0)
    iq, _ = synth.generate(cfg)
    if dcfilter:
        iq = with_dc_offset(iq, "uc8", 5, -4)
    want = port.run(iq, "uc8", modeac=modeac, dcfilter=dcfilter)
    extra = (["--modeac"] if modeac else []) + (["--dcfilter"] if dcfilter else [])
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "in.bin")
        iq.tofile(path)
        out = subprocess.run([str(exe), "--device-type", "ifile", "--ifile", path, "--preamble-threshold", "58",
                              "--raw", "--mlat", *extra], capture_output=True, text=True, timeout=300)
        assert out.returncode == 0, out.stderr[-2000:]
        stats = subprocess.run([str(exe), "--device-type", "ifile", "--ifile", path, "--preamble-threshold", "58",
                                "--quiet", "--stats", *extra], capture_output=True, text=True, timeout=300)
    lines = [l for l in out.stdout.splitlines() if l.startswith("@")]
    expect = ["@%012X%s;" % (int(m["timestampMsg"]), bytes(m["msg"][: m["msgbits"] // 8]).hex()) for m in want.msgs]
    assert len(expect) > 1000
    assert lines == expect
    assert "disagrees" not in out.stderr
    # the --stats block readsb prints at exit (stats.c:80-125) carries the same demodulator counters
    text = stats.stdout
    st = want.stats
    for line in (f"{int(st['samples_processed'])} samples processed",
                 f"{int(st['demod_preambles'])} Mode-S message preambles received",
                 f"{int(st['demod_rejected_bad'])} with bad message format or invalid CRC",
                 f"{int(st['demod_rejected_unknown_icao'])} with unrecognized ICAO address",
                 f"{int(st['demod_accepted'][0])} accepted with correct CRC",
                 f"{int(st['demod_accepted'][1])} accepted with 1-bit error repaired",
                 f"{int(st['messages_total'])} total usable messages",
                 f"{int(np.sum(want.msgs['msgtype'] == 32))} Mode A/C messages received"):
        assert line in text, (line, text[:1500])
    if modeac:
        assert int(np.sum(want.msgs["msgtype"] == 32)) > 100
    phase_rows = [" ".join(str(int(x)) for x in st["demod_preamblePhase"]), " ".join(str(int(x)) for x in st["demod_bestPhase"])]
    squashed = " ".join(text.split())
    for row in phase_rows:
        assert row in squashed, (row, text[:1500])
