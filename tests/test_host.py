"""CPU: the C-ABI library loads without a GPU, exports what include/readsb_b200.h declares, fails
loudly when asked to compute, and its host-side tables / filter equal the oracle's."""
import ctypes
import re
from pathlib import Path

import numpy as np
import pytest

from oracle import port
from readsb_protobuf_b200 import api, results

ROOT = Path(__file__).resolve().parent.parent


def test_library_exports_every_declared_symbol():
    header = (ROOT / "include" / "readsb_b200.h").read_text()
    declared = set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 20
    L = api.load()
    for name in sorted(declared):
        assert hasattr(L, name), f"{name} declared in include/readsb_b200.h but not exported"
    assert declared == set(api.EXPORTED_SYMBOLS)


def test_abi_struct_sizes_match_python_mirrors():
    L = api.load()
    assert L.b200_abi_sizeof(0) == results.MSG_DTYPE.itemsize == 68
    assert L.b200_abi_sizeof(1) == results.STATS_DTYPE.itemsize == 136
    assert L.b200_abi_sizeof(2) == results.BLOCK_DTYPE.itemsize == 16
    assert L.b200_abi_sizeof(3) == ctypes.sizeof(api.Timing)
    assert L.b200_abi_sizeof(4) == api.PHASE_RECORD_DTYPE.itemsize
    assert L.b200_abi_sizeof(5) == api.ERRORINFO_DTYPE.itemsize == 12  # struct errorinfo, crc.h:32-37
    assert L.b200_abi_sizeof(6) == ctypes.sizeof(api._Config)


def test_no_cpu_fallback():
    """Without a usable B200 the product must refuse, not compute on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(api.B200Error) as e:
        api.Demodulator()
    assert e.value.code == -2 and "no CPU fallback" in str(e.value)


def test_bad_configuration_is_rejected():
    L = api.load()
    h = ctypes.c_void_p()
    for bad in (api._Config(99, 0, 0, 1, 58, 0, 0, 0), api._Config(1, 0, 7, 1, 58, 0, 0, 0),
                api._Config(1, 0, 0, 5, 58, 0, 0, 0), api._Config(1, 0, 0, 1, 0, 0, 0, 0)):
        assert L.b200_demod_create(ctypes.byref(bad), ctypes.byref(h)) == -1
        assert not h.value
    assert L.b200_demod_create(None, ctypes.byref(h)) == -1


def test_host_crc_tables_equal_oracle():
    rng = np.random.default_rng(3)
    for bits in (56, 112):
        for _ in range(300):
            m = rng.integers(0, 256, bits // 8, dtype=np.uint8).tobytes()
            assert api.host_checksum(m) == port.checksum(m)
    for nfix in (0, 1, 2):
        for bits in (56, 112):
            a, b = api.host_error_table(nfix, bits), port.error_table(nfix, bits)
            assert len(a) == len(b)
            for f in ("syndrome", "errors", "bit"):
                assert np.array_equal(a[f], b[f]), (nfix, bits, f)


def test_host_uc8_table_equals_oracle():
    assert np.array_equal(api.host_uc8_table(), port.uc8_table())


def test_host_icao_filter_semantics():
    # icao_filter.c: membership lasts from the add until the second flip after it
    add, test, expire = 0, 1, 2
    ops = [test, add, test, expire, test, expire, test, expire, test,  # expire(0) flips at once (next_flip = 0)
           add, expire, test, expire, test]
    args = [0x4B9696, 0x4B9696, 0x4B9696, 0, 0x4B9696, 59_999, 0x4B9696, 60_000, 0x4B9696,
            0xABCDEF, 60_001, 0xABCDEF, 120_000, 0xABCDEF]
    res = api.host_filter_script(ops, args)
    assert list(res[[0, 2, 4, 6, 8, 11, 13]]) == [0, 1, 1, 1, 0, 1, 1]
    # the low-16-bit alias (icao_filter.c:87-96) is stored but never matches a full-address test
    res = api.host_filter_script([add, test, test], [0x123456, 0x003456, 0x123456])
    assert list(res[1:]) == [0, 1]
    # a full table drops further addresses instead of looping (icao_filter.c:78-81)
    n = 5000
    ops = [add] * n + [test] * n
    addrs = [0x100000 + 7 * i for i in range(n)]
    res = api.host_filter_script(ops, addrs + addrs)
    assert 4000 <= int(res[n:].sum()) < n


@pytest.mark.parametrize("verbatim", [False, True])
def test_output_writers_match_the_reference(verbatim):
    """b200_format_beast / b200_format_raw (host functions of the library, no device needed) against bytes
    the reference's own modesSendBeastOutput / modesSendRawOutput produced (tests/golden/netfmt.npz)."""
    from conftest import GOLDEN
    z = np.load(GOLDEN / "netfmt.npz")
    msgs, v = z["msgs"], int(verbatim)
    assert api.format_beast(msgs, verbatim) == z[f"beast_v{v}"].tobytes()
    for mlat in (False, True):
        assert api.format_raw(msgs, verbatim, mlat) == z[f"raw_v{v}_m{int(mlat)}"].tobytes()
    # capacity handling: the byte count is returned whatever fits, nothing is written past cap
    L = api.load()
    m = np.ascontiguousarray(msgs)
    need = L.b200_format_beast(m.ctypes.data, len(m), v, None, 0)
    buf = np.full(101, 0xEE, dtype=np.uint8)
    assert L.b200_format_beast(m.ctypes.data, len(m), v, buf.ctypes.data, 100) == need
    assert buf[100] == 0xEE and bytes(buf[:100]) == z[f"beast_v{v}"].tobytes()[:100]


def test_host_resolver_over_recorded_kernel_outputs():
    """The order-dependent tail of demodulate2400 (best-phase pick, decode-time rejects, ICAO filter,
    skip-ahead, statistics: resolver.cc) run on the CPU over what the kernels produced on a B200 for a small
    seeded stream (tests/golden/resolver_span_*.bin, recorded by scripts/dump_spans.py fixture in two process
    calls), against the oracle on the regenerated stream: messages, block means and every counter."""
    from readsb_protobuf_b200 import synth
    paths = sorted((ROOT / "tests" / "golden").glob("resolver_span_*.bin"), key=lambda p: int(p.stem.split("_")[-1]))
    assert paths, "tests/golden/resolver_span_*.bin are part of the repository"
    cfg = synth.resolver_fixture_config()
    iq, _ = synth.generate(cfg)
    want = port.run(iq, "uc8")
    got = api.host_resolve_dumps(paths, nfix=1)
    assert got.n_samples == cfg.nsamples and len(paths) == 2
    assert int(got.stats["convert_cpu_s"]) == 0  # kernel-vs-host CRC disagreements
    got.stats["convert_cpu_s"] = got.stats["demod_cpu_s"] = 0
    assert len(want.msgs) > 500
    assert results.compare_results(got, want, float_rtol=0.0, signal_atol=0.0) == []
    # cut differently, the files no longer describe the stream: the resolver must not silently agree
    with pytest.raises(api.B200Error):
        api.host_resolve_dumps([ROOT / "tests" / "golden" / "kat_frame.npz"])


@pytest.mark.parametrize("predict", ["always-optimistic", "prescan"])
@pytest.mark.parametrize("threads", [2, 3, 8])
def test_host_resolver_runs_side_by_side(threads, predict, monkeypatch, capfd):
    """The same replay with the walk forced into several runs of mag_bufs per span, each started from a predicted
    ICAO-filter state on its own thread and kept only if the prediction held (resolver.cc, 'speculation'): whatever
    the number of threads and wherever the runs are cut, the result is the sequential one, bit for bit."""
    from readsb_protobuf_b200 import synth
    paths = sorted((ROOT / "tests" / "golden").glob("resolver_span_*.bin"), key=lambda p: int(p.stem.split("_")[-1]))
    assert paths, "tests/golden/resolver_span_*.bin are part of the repository"
    monkeypatch.setenv("B200_RESOLVER_THREADS", str(threads))
    monkeypatch.setenv("B200_RESOLVER_MIN_LIVE", "0")
    monkeypatch.setenv("B200_RESOLVER_MIN_BLOCKS", "1")
    monkeypatch.setenv("B200_RESOLVER_MIN_LIVE_PER_RUN", "1")
    monkeypatch.setenv("B200_RESOLVER_TRACE", "1")
    monkeypatch.setenv("B200_RESOLVER_PREDICT", predict)
    cfg = synth.resolver_fixture_config()
    iq, _ = synth.generate(cfg)
    want = port.run(iq, "uc8")
    got = api.host_resolve_dumps(paths, nfix=1)
    assert int(got.stats["convert_cpu_s"]) == 0
    got.stats["convert_cpu_s"] = got.stats["demod_cpu_s"] = 0
    assert results.compare_results(got, want, float_rtol=0.0, signal_atol=0.0) == []
    # ... and the walk really was split
    import re
    m = re.search(r"resolver: (\d+) spans, (\d+) as several runs \((\d+) runs", capfd.readouterr().err)
    assert m, "B200_RESOLVER_TRACE prints the resolver's run counts when it is destroyed"
    if predict == "prescan" and threads < 3:
        assert int(m.group(2)) == 0  # predicting costs about one more walk: not worth it on two workers
    else:
        # ("always-optimistic": every run starts from the true state in front of its round although the fixture is the
        # beginning of a stream, where every aircraft is new -- rounds, re-speculation and the one-run fall-back all run)
        assert int(m.group(2)) >= 1 and int(m.group(3)) > int(m.group(2))


def test_filter_speculation_rule_is_sound(tmp_path):
    """tests/cpp/test_icao_filter.cc: whenever IcaoFilter::same_members / differs_only_unprobed keep a speculatively
    walked run, replaying the run on the true filter state gives the same answer to every test() -- checked on
    thousands of random filter histories and runs (inserts, flips on different clocks, probes inside and outside the
    difference).  The resolver's parallel walk (resolver.cc) is exact because of this rule."""
    import shutil
    import subprocess
    gxx = shutil.which("g++")
    assert gxx, "g++ is part of this image"
    exe = tmp_path / "test_icao_filter"
    csrc = ROOT / "readsb_protobuf_b200" / "csrc"
    subprocess.run([gxx, "-std=c++17", "-O1", "-I", str(ROOT / "include"), "-I", str(csrc), str(ROOT / "tests" / "cpp" / "test_icao_filter.cc"),
                    str(csrc / "resolver.cc"), str(csrc / "host_tables.cc"), "-lpthread", "-o", str(exe)], check=True)
    r = subprocess.run([str(exe), "4000"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "kept by the probe rule" in r.stdout


def _synthetic_span_dumps(tmp_path, n_aircraft, seed):
    """Resolver inputs written by hand in the layout B200_DUMP_SPAN produces (cabi.cu, layout 2): three spans of 400
    mag_bufs (65 s of stream: the start-up flip of the ICAO filter and the one a minute later), a live position every
    ~6000 samples carrying one frame: a clean DF17 of one of `n_aircraft` addresses (an insert into the filter), or a DF4
    reply whose address -- the CRC residue -- is one of them (accepted only while the filter knows it)."""
    rng = np.random.default_rng(seed)
    B, blocks_per_span = 131072, 400
    pool = rng.choice(np.arange(0x100000, 0xf00000), size=n_aircraft, replace=False).astype(np.uint32)
    live_dt = np.dtype([("pos", "<u4"), ("info", "<u4"), ("dead_rank", "<u4"), ("pad", "<u4")])
    rec_dt = np.dtype([("pos", "<u4"), ("w0", "<u4"), ("w1", "<u4"), ("errbits", "<u4"), ("power", "<u8"), ("msg", "u1", (14,)), ("pad", "u1", (2,))])
    assert live_dt.itemsize == 16 and rec_dt.itemsize == 40
    paths = []
    for span in range(3):
        n = B * blocks_per_span
        pos = np.arange(500, n - 700, 6000, dtype=np.int64) + rng.integers(0, 4000, size=len(np.arange(500, n - 700, 6000)))
        pos = np.unique(pos).astype(np.uint32)
        k = len(pos)
        live = np.zeros(k, dtype=live_dt)
        recs = np.zeros(k, dtype=rec_dt)
        live["pos"] = pos
        live["pad"] = np.arange(k, dtype=np.uint32)
        addr = pool[rng.integers(0, n_aircraft, size=k)]
        is_es = rng.random(k) < 0.7
        phase = rng.integers(4, 9, size=k).astype(np.uint32)
        live["info"] = (1 << (phase - 4)).astype(np.uint32) | (1 << 8)  # try mask: that phase; one record
        for i in range(k):
            a = int(addr[i])
            if is_es[i]:
                body = bytes([0x8D, a >> 16, (a >> 8) & 255, a & 255]) + rng.integers(0, 256, size=7, dtype=np.uint8).tobytes()
                crc = port.checksum(body + b"\0\0\0")  # parity that makes the frame's syndrome 0
                msg = body + bytes([crc >> 16, (crc >> 8) & 255, crc & 255])
                assert port.checksum(msg) == 0
                recs["w0"][i] = 0 | (4 << 24) | (1 << 31)  # crc 0, kKindES, no repair, key in S
                recs["w1"][i] = a | (int(phase[i]) << 24)
            else:
                body = bytes([0x20]) + rng.integers(0, 256, size=3, dtype=np.uint8).tobytes()  # DF4
                crc = port.checksum(body + b"\0\0\0")
                ap = crc ^ a  # Address/Parity: the residue is the address
                msg = body + bytes([ap >> 16, (ap >> 8) & 255, ap & 255])
                assert port.checksum(msg) == a
                recs["w0"][i] = a | (1 << 24) | (1 << 31)  # crc = the address, kKindAP, key in S
                recs["w1"][i] = a | (int(phase[i]) << 24)
            recs["msg"][i, : len(msg)] = np.frombuffer(msg, dtype=np.uint8)
        recs["pos"] = pos
        recs["errbits"] = 0xFFFF
        recs["power"] = rng.integers(1 << 20, 1 << 34, size=k).astype(np.uint64)
        nblocks = blocks_per_span + 2
        hdr = np.array([n, span * n, B, 1 if span == 2 else 0, 0, 0, 0, k, k, nblocks, 2, 0], dtype=np.uint64)
        path = tmp_path / f"span_{span}.bin"
        with open(path, "wb") as f:
            f.write(hdr.tobytes())
            f.write(live.tobytes())
            f.write(recs.tobytes())
            f.write(np.zeros(4 * k, dtype=np.uint64).tobytes())           # hidden-dead counts: none
            f.write(np.zeros(8 * nblocks, dtype=np.uint32).tobytes())     # block dead counters
            sums = rng.integers(1 << 30, 1 << 34, size=2 * nblocks).astype(np.uint64)
            sums[2 * blocks_per_span:] = 0  # the empty last mag_buf of a stream of whole blocks (0 / 0 = NaN, as in the reference)
            f.write(sums.tobytes())  # integer block sums
            f.write(np.zeros(2 * nblocks, dtype=np.float64).tobytes())
        paths.append(path)
    return paths


@pytest.mark.parametrize("n_aircraft", [300, 3000, 12000])
def test_host_resolver_side_by_side_with_large_populations(n_aircraft, tmp_path, monkeypatch):
    """The walk in several runs against the walk in one, on hand-made resolver inputs with more aircraft than the
    fixtures have: 300 (the filter stays small), 3000 (its insert lists outgrow what a copy reproduces, so speculation
    has to hand over to the one-run walk in the middle of a span) and 12000 (its 8192-entry tables fill up and drop
    inserts, icao_filter.c:78-81).  Same messages, same counters, whatever the mode and the number of threads."""
    paths = _synthetic_span_dumps(tmp_path, n_aircraft, seed=n_aircraft)
    monkeypatch.setenv("B200_RESOLVER_THREADS", "1")
    want = api.host_resolve_dumps(paths, nfix=1)
    assert int(want.stats["convert_cpu_s"]) == 0  # no kernel-vs-host CRC disagreement: the frames are well formed
    n_es = int(np.sum(want.msgs["msgtype"] == 17))
    n_ap = int(np.sum(want.msgs["msgtype"] == 4))
    assert n_es > 15000 and n_ap > 0 and int(want.stats["demod_rejected_unknown_icao"]) > 0, (n_es, n_ap)
    for predict in ("optimistic", "always-optimistic", "prescan"):
        for threads in (2, 5, 8):
            monkeypatch.setenv("B200_RESOLVER_THREADS", str(threads))
            monkeypatch.setenv("B200_RESOLVER_PREDICT", predict)
            monkeypatch.setenv("B200_RESOLVER_MIN_LIVE", "0")
            got = api.host_resolve_dumps(paths, nfix=1)
            got.stats["convert_cpu_s"] = want.stats["convert_cpu_s"]
            assert results.compare_results(got, want, float_rtol=0.0, signal_atol=0.0) == [], (predict, threads)
