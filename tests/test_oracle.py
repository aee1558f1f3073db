"""CPU: the oracle (C restatement) against the reference-generated goldens, the reference's own
known answers, and -- where /root/reference is mounted -- the live reference binary."""
import hashlib

import numpy as np
import pytest

from conftest import GOLDEN as GOLDEN_DIR, GOLDEN_NAMES, load_golden
from oracle import port, ref
from readsb_protobuf_b200 import results, synth


@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_port_matches_reference_golden(name):
    iq, want, meta = load_golden(name)
    assert hashlib.sha256(iq.tobytes()).hexdigest() == meta["sha256"]
    got = port.run(iq, meta["fmt"], **meta["flags"])
    assert len(want.msgs) > 0
    # bit-exact everywhere, float sums included (same operations in the same order)
    assert results.compare_results(got, want, float_rtol=0.0, signal_atol=0.0) == []


def test_known_answer_frame():
    # the only Mode S frame quoted in the reference tree (net_io.c:1645): DF17, CRC 0;
    # first sight scores 1400, second sight 1800 (mode_s.c:288-289)
    iq, want, meta = load_golden("kat_frame")
    got = port.run(iq, "uc8")
    assert [bytes(m["msg"]).hex().upper() for m in got.msgs] == [meta["frame"], meta["frame"]]
    assert [int(m["score"]) for m in got.msgs] == [1400, 1800]
    assert [int(m["crc"]) for m in got.msgs] == [0, 0]
    # the frame starts at tick 100003; the reference reports 0x19000 (+326*5 bias)
    assert int(got.msgs[0]["timestampMsg"]) == 0x19000
    assert port.checksum(bytes.fromhex(meta["frame"])) == 0


def test_crc_known_answers():
    # crc.c:59-64 / SURVEY.md 8a: first and last single-bit syndromes
    assert port.lib().mo_single_bit_syndrome(0) == 0x3935EA
    assert port.lib().mo_single_bit_syndrome(111) == 0x000001
    # table sizes printed by the reference's own `crctests` (crc.c -DCRCDEBUG)
    assert len(port.error_table(1, 56)) == 51 and len(port.error_table(1, 112)) == 107
    assert len(port.error_table(2, 56)) == 1326 and len(port.error_table(2, 112)) == 3831
    assert len(port.error_table(0, 112)) == 0
    # every entry regenerates its own syndrome (the check crc.c:309-333 performs)
    for nfix, bits in ((1, 56), (1, 112), (2, 56), (2, 112)):
        t = port.error_table(nfix, bits)
        assert np.all(np.diff(t["syndrome"].astype(np.int64)) > 0)
        for e in t[:: max(1, len(t) // 97)]:
            msg = bytearray(bits // 8)
            for b in (int(x) for x in e["bit"][: int(e["errors"])]):
                assert 5 <= b < bits
                msg[b >> 3] ^= 0x80 >> (b & 7)
            assert port.checksum(bytes(msg)) == e["syndrome"]


def test_checksum_is_linear():
    rng = np.random.default_rng(5)
    for bits in (56, 112):
        for _ in range(200):
            m = rng.integers(0, 256, bits // 8, dtype=np.uint8)
            x = 0
            for b in range(bits):
                if (m[b >> 3] >> (7 - (b & 7))) & 1:
                    x ^= port.lib().mo_single_bit_syndrome(b + 112 - bits)
            assert x == port.checksum(m.tobytes())


def test_uc8_table_known_answers():
    t = port.uc8_table().reshape(256, 256)
    # SURVEY.md section 7.2 probes of the reference table
    assert t[0, 0] == 65535 and t[127, 127] == 363 and t[128, 128] == 363 and t[255, 128] == 65535
    assert t.min() == 363
    assert np.array_equal(t, t[::-1, :]) and np.array_equal(t, t[:, ::-1]) and np.array_equal(t, t.T)
    grid = t.astype(np.float64)
    assert abs(grid.mean() / 65536.0 - 0.740231129) < 1e-9
    assert abs((grid ** 2).mean() / 65535.0 ** 2 - 0.610363797) < 1e-9


@pytest.mark.skipif(not ref.available(), reason="reference binary not built (no /root/reference here)")
@pytest.mark.parametrize("case", [
    dict(cfg=synth.SynthConfig(seed=21, nsamples=700_000, frames_per_s=4000, frac_biterror=0.3)),
    dict(cfg=synth.SynthConfig(seed=22, nsamples=500_000, fmt="sc16", frames_per_s=3000, frac_biterror=0.2)),
    dict(cfg=synth.SynthConfig(seed=23, nsamples=500_000, fmt="sc16q11", frames_per_s=3000, frac_biterror=0.2)),
    dict(cfg=synth.SynthConfig(seed=24, nsamples=400_000, frames_per_s=5000, frac_biterror=0.5), nfix=2),
    dict(cfg=synth.SynthConfig(seed=25, nsamples=400_000, frames_per_s=5000, frac_biterror=0.5), nfix=0),
    dict(cfg=synth.SynthConfig(seed=26, nsamples=300_003, frames_per_s=5000), block_samples=50000, threshold=40),
    dict(cfg=synth.SynthConfig(seed=27, nsamples=4 * 65536, frames_per_s=5000), block_samples=65536),
    dict(cfg=synth.SynthConfig(seed=28, nsamples=100, frames_per_s=0)),
    dict(cfg=synth.SynthConfig(seed=29, nsamples=0, frames_per_s=0)),
])
def test_port_matches_live_reference(case):
    case = dict(case)
    cfg = case.pop("cfg")
    iq, _ = synth.generate(cfg)
    got = port.run(iq, cfg.fmt, **case)
    want = ref.run(iq, cfg.fmt, **case)
    assert results.compare_results(got, want, float_rtol=0.0, signal_atol=0.0) == []


@pytest.mark.skipif(not ref.available(), reason="reference binary not built (no /root/reference here)")
def test_port_converter_matches_live_reference():
    for fmt in ("uc8", "sc16", "sc16q11"):
        cfg = synth.SynthConfig(seed=31, nsamples=140_000, fmt=fmt, frames_per_s=3000, amp_max=1.5)
        iq, _ = synth.generate(cfg)
        want = ref.magnitudes(iq, fmt)
        bps = 2 if fmt == "uc8" else 4
        got = np.concatenate([port.convert(iq[o * bps: (o + 131072) * bps], fmt)[0] for o in range(0, cfg.nsamples, 131072)])
        assert np.array_equal(got, want)


@pytest.mark.skipif(not ref.available(), reason="reference binary not built (no /root/reference here)")
@pytest.mark.parametrize("fmt,seed", [("uc8", 41), ("sc16", 42), ("sc16q11", 43)])
def test_port_dcfilter_matches_live_reference(fmt, seed):
    """--dcfilter: the restated convert_*_generic against the reference's own, through
    the whole path (messages, stats, block means) and magnitude by magnitude."""
    cfg = synth.SynthConfig(seed=seed, nsamples=600_000, fmt=fmt, frames_per_s=3000, frac_biterror=0.2, modeac_per_s=1500,
                            amp_max=1.3)
    iq, _ = synth.generate(cfg)
    for modeac in (False, True):
        got = port.run(iq, fmt, dcfilter=True, modeac=modeac)
        want = ref.run(iq, fmt, dcfilter=True, modeac=modeac)
        assert len(want.msgs) > 300
        assert results.compare_results(got, want, float_rtol=0.0, signal_atol=0.0) == []
    assert results.compare_results(got, port.run(iq, fmt, modeac=True), float_rtol=0.0, signal_atol=0.0) != []
    calls = [131072] * (cfg.nsamples // 131072) + [cfg.nsamples % 131072]
    mag, _ = port.convert_dc(iq, fmt, calls)
    assert np.array_equal(mag, ref.magnitudes(iq, fmt, dcfilter=True))


@pytest.mark.skipif(not ref.available(), reason="reference binary not built (no /root/reference here)")
def test_port_sc16q11_table_matches_live_reference():
    """convert_sc16q11_table (convert.c:264-328) against the reference built the way its
    armhf package is (-DSC16Q11_TABLE_BITS=8, oracle/_ref/ref_demod_tb8): whole path and magnitudes."""
    cfg = synth.SynthConfig(seed=44, nsamples=600_000, fmt="sc16q11", frames_per_s=3000, frac_biterror=0.2, modeac_per_s=1500,
                            amp_max=1.3)
    iq, _ = synth.generate(cfg)
    for kw in (dict(), dict(modeac=True), dict(dcfilter=True)):  # --dcfilter picks the generic converter in that build too
        got = port.run(iq, "sc16q11", table_bits=8, **kw)
        want = ref.run(iq, "sc16q11", table_bits=8, **kw)
        assert len(want.msgs) > 300
        assert results.compare_results(got, want, float_rtol=0.0, signal_atol=0.0) == []
    assert results.compare_results(port.run(iq, "sc16q11", table_bits=8), port.run(iq, "sc16q11"), float_rtol=0.0, signal_atol=0.0) != []
    mag = np.concatenate([port.convert_sc16q11_table(iq[o * 4: (o + 131072) * 4], 8)[0] for o in range(0, cfg.nsamples, 131072)])
    assert np.array_equal(mag, ref.magnitudes(iq, "sc16q11", table_bits=8))
    # the table itself: 0 at the origin, clamped to full scale beyond the unit circle, symmetric
    t = port.sc16q11_table(8).reshape(256, 256)
    assert t[0, 0] == 0 and t[255, 255] == 65535 and t[181, 181] == 65528 and t[182, 182] == 65535 and np.array_equal(t, t.T)
    # the larger tables of the reference's oneoff/convert_benchmark.c (oracle/_ref/ref_demod_tb9 .. tb11)
    for bits in (9, 10, 11):
        assert results.compare_results(port.run(iq, "sc16q11", table_bits=bits), ref.run(iq, "sc16q11", table_bits=bits),
                                       float_rtol=0.0, signal_atol=0.0) == []
    # other formats are untouched by the build flag
    iq2, _ = synth.generate(synth.SynthConfig(seed=45, nsamples=200_000, fmt="sc16", frames_per_s=3000))
    assert results.compare_results(port.run(iq2, "sc16", table_bits=8), ref.run(iq2, "sc16", table_bits=8), float_rtol=0.0,
                                   signal_atol=0.0) == []


def test_generator_is_deterministic_and_chunk_invariant():
    cfg = synth.SynthConfig(seed=7, nsamples=300_000, frames_per_s=2000)
    frames = synth.plan(cfg)
    whole = synth.render(cfg, frames)
    parts = np.concatenate([synth.render(cfg, frames, first=a, count=b - a)
                            for a, b in ((0, 1000), (1000, 70_001), (70_001, 300_000))])
    assert np.array_equal(whole, parts)
    assert np.array_equal(whole, synth.render(cfg, synth.plan(cfg)))
    # every planted frame has a valid parity field for its type
    for f in frames[:200]:
        msg = bytes(f["msg"][: f["nbytes"]])
        if f["errbit"] < 0 and f["df"] in (11, 17):
            assert port.checksum(msg) == 0


# ------------------------------------------------------------------------------------------
# Mode A/C (demodulate2400AC), SURVEY.md 8f row 2
# ------------------------------------------------------------------------------------------

def test_modeac_golden_holds_replies():
    iq, want, meta = load_golden("uc8_modeac")
    ac = want.msgs[want.msgs["msgtype"] == 32]
    assert len(ac) > 50 and np.all(ac["msgbits"] == 16) and np.all(ac["addr"] >> 24 == 1)
    # decodeModeAMessage (mode_ac.c:168-181): msg[0..1] is the code, the address drops the SPI bit
    code = (ac["msg"][:, 0].astype(np.uint32) << 8) | ac["msg"][:, 1]
    assert np.array_equal(ac["addr"] & 0xffff, code & 0xff7f)


@pytest.mark.skipif(not ref.available(), reason="needs /root/reference or a prebuilt oracle/_ref")
@pytest.mark.parametrize("fmt,seed", [("uc8", 301), ("uc8", 302), ("sc16", 303)])
def test_port_modeac_matches_live_reference(fmt, seed):
    cfg = synth.SynthConfig(seed=seed, nsamples=700_000, fmt=fmt, frames_per_s=1500, frac_biterror=0.1,
                            modeac_per_s=3000, noise_sigma=0.01 if seed == 302 else 0.02)
    iq, frames = synth.generate(cfg)
    want = ref.run(iq, fmt, modeac=True)
    got = port.run(iq, fmt, modeac=True)
    assert int(np.sum(want.msgs["msgtype"] == 32)) > 100 and int(np.sum(want.msgs["msgtype"] != 32)) > 100
    assert results.compare_results(got, want, float_rtol=0.0, signal_atol=0.0) == []
    # without the flag nothing changes for Mode S
    plain = port.run(iq, fmt)
    assert np.array_equal(plain.msgs, got.msgs[got.msgs["msgtype"] != 32])


def test_df18_non_icao_address_flag():
    """decodeExtendedSquitter flags mm->addr with MODES_NON_ICAO_ADDRESS for DF18 depending on CF and the
    ME field's IMF bit (mode_s.c:1373-1470); golden uc8_df18 holds every CF x ME-type combination."""
    iq, want, meta = load_golden("uc8_df18")
    assert len(want.msgs) == meta["n_frames"] and np.all(want.msgs["msgtype"] == 18)
    flagged = (want.msgs["addr"] >> 24) & 1
    assert 0 < int(flagged.sum()) < len(flagged)
    cf = want.msgs["msg"][:, 0] & 7
    assert np.all(flagged[cf == 0] == 0) and np.all(flagged[np.isin(cf, (1, 4, 5, 7))] == 1)
    assert 0 < int(flagged[np.isin(cf, (2, 3, 6))].sum()) < int(np.isin(cf, (2, 3, 6)).sum())
    got = port.run(iq, "uc8")
    assert results.compare_results(got, want, float_rtol=0.0, signal_atol=0.0) == []


# ------------------------------------------------------------------------------------------
# Beast / raw output writers (net_io.c:769-896), SURVEY.md 8f row 1
# ------------------------------------------------------------------------------------------

def _netfmt_golden():
    z = np.load(GOLDEN_DIR / "netfmt.npz")
    return z["msgs"], {k: z[k].tobytes() for k in z.files if k != "msgs"}


@pytest.mark.parametrize("verbatim", [False, True])
def test_port_writers_match_reference_golden(verbatim):
    msgs, want = _netfmt_golden()
    assert len(msgs) > 500 and int(np.sum(msgs["msgbits"] == 40)) > 0  # a length that is not sent
    v = int(verbatim)
    beast = port.format_beast(msgs, verbatim)
    assert beast == want[f"beast_v{v}"] and beast.count(b"\x1a\x1a") > 20
    for mlat in (False, True):
        assert port.format_raw(msgs, verbatim, mlat) == want[f"raw_v{v}_m{int(mlat)}"]
    assert b"\n*" in want[f"raw_v{v}_m1"]  # zero timestamps print '*' even with --mlat


@pytest.mark.skipif(ref.netfmt_binary() is None, reason="needs /root/reference or a prebuilt oracle/_ref")
def test_port_writers_match_live_reference():
    cfg = synth.SynthConfig(seed=401, nsamples=600_000, frames_per_s=3000, frac_biterror=0.2, modeac_per_s=2000)
    iq, _ = synth.generate(cfg)
    res = ref.run(iq, "uc8", modeac=True)
    assert len(res.msgs) > 300
    for verbatim in (False, True):
        for mlat in (False, True):
            wb, wr = ref.format_outputs(res, net_verbatim=verbatim, mlat=mlat)
            assert port.format_beast(res.msgs, verbatim) == wb
            assert port.format_raw(res.msgs, verbatim, mlat) == wr


def test_dcfilter_and_table_converter_properties():
    """Size-independent properties of the two converters added for SURVEY 8f rows 3 and 4 (oracle side; the
    GPU tests require the library to reproduce the oracle bit for bit)."""
    rng = np.random.default_rng(5)
    # DC block: a constant input leaves a magnitude that decays monotonically (z creeps towards the input at
    # 1 - exp(-2 pi / 2.4e6) per sample) and the state really is carried from call to call
    n = 400_000
    iq = np.tile(np.array([200, 90], dtype=np.uint8), n)
    mag, means = port.convert_dc(iq, "uc8", calls=[n // 2, n // 2])
    assert mag[0] > mag[n // 2] > mag[-1] > 0 and np.all(np.diff(mag.astype(np.int64)) <= 0)
    assert means[0][0] > means[1][0] > 0
    one_call, _ = port.convert_dc(iq, "uc8")
    assert np.array_equal(mag, one_call)
    # with the DC block off the same converter equals the plain float path for sc16 (dc_a = 0, dc_b = 1 there)
    v = rng.integers(-32768, 32768, 2 * 5000).astype("<i2").view(np.uint8)
    plain, _, _ = port.convert(v, "sc16")
    filtered, _ = port.convert_dc(v, "sc16")
    assert np.mean(plain != filtered) > 0.01  # the filter is on: it does change magnitudes ...
    assert np.max(np.abs(plain.astype(np.int64) - filtered.astype(np.int64))) < 2000  # ... by the small DC estimate only
    # table converter: depends on |I|, |Q| only, through their top bits; symmetric in I and Q
    for bits in (8, 5):
        k = 11 - bits
        base = rng.integers(0, 2048, (3000, 2))
        def mags(a):
            return port.convert_sc16q11_table(a.astype("<i2").reshape(-1).view(np.uint8), bits)[0]
        m0 = mags(base)
        assert np.array_equal(m0, mags(-base)) and np.array_equal(m0, mags(base[:, ::-1]))
        assert np.array_equal(m0, mags((base >> k) << k)) and np.array_equal(m0, mags(base + 2048 * 3))
