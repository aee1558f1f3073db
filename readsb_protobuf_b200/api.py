"""ctypes front-end of the C ABI (include/readsb_b200.h -> libreadsb_b200.so).

The Python layer mirrors the reference's own vocabulary for this path: a `Demodulator` is one
receiver stream (what `struct _Modes` + the FIFO are to readsb); `process()` is the
`ifileRun` / `demodulate2400` loop over a span of IQ; results are `modesMessage`-shaped records
and the demodulator's share of `struct stats`.

There is no CPU fallback: loading the library without a built CUDA extension, or creating a
demodulator without a B200, raises.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass

import numpy as np

from . import build as _build
from .results import BLOCK_DTYPE, MSG_DTYPE, STATS_DTYPE, DemodResult

FORMATS = {"uc8": 0, "sc16": 1, "sc16q11": 2}
BYTES_PER_SAMPLE = {"uc8": 2, "sc16": 4, "sc16q11": 4}
OVERLAP_SAMPLES = 326
DEFAULT_BLOCK_SAMPLES = 131072
FLAG_FINAL = 1

KIND_AP, KIND_AP_COMMB, KIND_DF11, KIND_ES = 1, 2, 3, 4

PHASE_RECORD_DTYPE = np.dtype(
    [("position", "<u4"), ("crc", "<u4"), ("key", "<u4"), ("phase", "u1"), ("kind", "u1"), ("errors", "u1"),
     ("reserved", "u1")], align=False)
ERRORINFO_DTYPE = np.dtype([("syndrome", "<u4"), ("errors", "<i4"), ("bit", "i1", (2,)), ("padding", "<u2")])

EXPORTED_SYMBOLS = (
    "b200_demod_create", "b200_demod_destroy", "b200_demod_reset", "b200_last_error",
    "b200_demod_process", "b200_demod_process_device", "b200_demod_message_count", "b200_demod_messages",
    "b200_demod_block_count", "b200_demod_blocks", "b200_demod_get_stats", "b200_demod_get_timing",
    "b200_scan_device", "b200_convert", "b200_uc8_table", "b200_debug_scan", "b200_crc_batch",
    "b200_error_table", "b200_abi_sizeof", "b200_host_checksum", "b200_host_error_table", "b200_host_uc8_table",
    "b200_host_filter_script", "b200_host_resolve_dumps", "b200_host_alloc", "b200_host_free", "b200_demod_modeac_count", "b200_format_beast", "b200_format_raw",
)


class B200Error(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"readsb_b200 error {code}: {message}")
        self.code = code


ABI_VERSION = 3  # B200_ABI_VERSION


class _Config(ctypes.Structure):
    _fields_ = [
        ("abi_version", ctypes.c_int32), ("device", ctypes.c_int32), ("input_format", ctypes.c_int32),
        ("nfix_crc", ctypes.c_int32), ("preamble_threshold", ctypes.c_int32), ("block_samples", ctypes.c_uint32),
        ("startup_time_ms", ctypes.c_uint64), ("max_span_samples", ctypes.c_uint64),
        ("mode_ac", ctypes.c_int32), ("filter_dc", ctypes.c_int32),
        ("sc16q11_table_bits", ctypes.c_int32), ("reserved", ctypes.c_int32),
    ]


class Timing(ctypes.Structure):
    _fields_ = [
        ("h2d_ms", ctypes.c_float), ("scan_ms", ctypes.c_float), ("classify_ms", ctypes.c_float),
        ("d2h_ms", ctypes.c_float), ("resolve_ms", ctypes.c_float), ("total_ms", ctypes.c_float),
        ("n_candidates", ctypes.c_uint64), ("n_phase_records", ctypes.c_uint64), ("n_live", ctypes.c_uint64),
        ("scan_launches", ctypes.c_uint32), ("chunks", ctypes.c_uint32), ("d2h_bytes", ctypes.c_uint64),
        ("slice_ms", ctypes.c_float), ("reserved", ctypes.c_uint32),
    ]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_ if k != "reserved"}


_lib = None


def library_path():
    return _build.ensure_cuda()


def load():
    """Load libreadsb_b200.so (building it in-tree if needed).  Raises if it cannot be built."""
    global _lib
    if _lib is not None:
        return _lib
    L = ctypes.CDLL(str(library_path()))
    vp, u64, u32, i32 = ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint32, ctypes.c_int
    L.b200_last_error.restype = ctypes.c_char_p
    L.b200_demod_create.restype = i32
    L.b200_demod_create.argtypes = [ctypes.POINTER(_Config), ctypes.POINTER(vp)]
    L.b200_demod_destroy.restype = None
    L.b200_demod_destroy.argtypes = [vp]
    L.b200_format_beast.restype = u64
    L.b200_format_beast.argtypes = [vp, u64, i32, vp, u64]
    L.b200_format_raw.restype = u64
    L.b200_format_raw.argtypes = [vp, u64, i32, i32, vp, u64]
    L.b200_demod_modeac_count.restype = u64
    L.b200_demod_modeac_count.argtypes = [vp]
    L.b200_demod_reset.restype = i32
    L.b200_demod_reset.argtypes = [vp]
    L.b200_demod_process.restype = i32
    L.b200_demod_process.argtypes = [vp, vp, u64, u32]
    L.b200_demod_process_device.restype = i32
    L.b200_demod_process_device.argtypes = [vp, vp, u64, u32, vp]
    L.b200_demod_message_count.restype = u64
    L.b200_demod_message_count.argtypes = [vp]
    L.b200_demod_messages.restype = vp
    L.b200_demod_messages.argtypes = [vp]
    L.b200_demod_block_count.restype = u64
    L.b200_demod_block_count.argtypes = [vp]
    L.b200_demod_blocks.restype = vp
    L.b200_demod_blocks.argtypes = [vp]
    L.b200_demod_get_stats.restype = i32
    L.b200_demod_get_stats.argtypes = [vp, vp]
    L.b200_demod_get_timing.restype = i32
    L.b200_demod_get_timing.argtypes = [vp, ctypes.POINTER(Timing)]
    L.b200_scan_device.restype = i32
    L.b200_scan_device.argtypes = [vp, vp, u64, i32, vp, ctypes.POINTER(ctypes.c_float), ctypes.POINTER(u64)]
    L.b200_convert.restype = i32
    L.b200_convert.argtypes = [vp, vp, u32, vp, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double)]
    L.b200_uc8_table.restype = i32
    L.b200_uc8_table.argtypes = [vp, vp]
    L.b200_debug_scan.restype = i32
    L.b200_debug_scan.argtypes = [vp, vp, u64, vp, vp, u64, ctypes.POINTER(u64)]
    L.b200_crc_batch.restype = i32
    L.b200_crc_batch.argtypes = [vp, vp, u32, vp, vp, vp]
    L.b200_error_table.restype = i32
    L.b200_error_table.argtypes = [vp, i32, vp, i32]
    L.b200_abi_sizeof.restype = i32
    L.b200_abi_sizeof.argtypes = [i32]
    L.b200_host_checksum.restype = u32
    L.b200_host_checksum.argtypes = [vp, i32]
    L.b200_host_error_table.restype = i32
    L.b200_host_error_table.argtypes = [i32, i32, vp, i32]
    L.b200_host_uc8_table.restype = None
    L.b200_host_uc8_table.argtypes = [vp]
    L.b200_host_filter_script.restype = i32
    L.b200_host_filter_script.argtypes = [vp, vp, u32, vp]
    L.b200_host_resolve_dumps.restype = i32
    L.b200_host_resolve_dumps.argtypes = [vp, u32, i32, vp, u64, vp, vp, u64, vp, vp]
    _lib = L
    return L


def host_checksum(msg: bytes) -> int:
    buf = (ctypes.c_uint8 * len(msg)).from_buffer_copy(msg)
    return int(load().b200_host_checksum(buf, len(msg) * 8))


def host_error_table(nfix: int, bits: int) -> np.ndarray:
    L = load()
    n = L.b200_host_error_table(nfix, bits, None, 0)
    t = np.zeros(max(n, 1), dtype=ERRORINFO_DTYPE)
    L.b200_host_error_table(nfix, bits, t.ctypes.data, n)
    return t[:n]


def host_uc8_table() -> np.ndarray:
    t = np.empty(65536, dtype=np.uint16)
    load().b200_host_uc8_table(t.ctypes.data)
    return t


def host_filter_script(ops, args) -> np.ndarray:
    ops = np.ascontiguousarray(ops, dtype=np.uint8)
    args = np.ascontiguousarray(args, dtype=np.uint64)
    res = np.zeros(len(ops), dtype=np.uint8)
    _check(load().b200_host_filter_script(ops.ctypes.data, args.ctypes.data, len(ops), res.ctypes.data))
    return res


def host_resolve_dumps(paths, nfix: int = 1) -> DemodResult:
    """The host resolver alone over recorded kernel outputs (files written with B200_DUMP_SPAN set, one per
    pipeline chunk, in stream order).  No GPU involved: this is how the CPU suite tests the host logic."""
    L = load()
    arr = (ctypes.c_char_p * len(paths))(*[str(p).encode() for p in paths])
    nm, nb = ctypes.c_uint64(), ctypes.c_uint64()
    cap = 1 << 16
    while True:
        msgs = np.empty(cap, dtype=MSG_DTYPE)
        blocks = np.empty(cap, dtype=BLOCK_DTYPE)
        st = np.zeros(1, dtype=STATS_DTYPE)
        rc = L.b200_host_resolve_dumps(arr, len(paths), nfix, msgs.ctypes.data, cap, ctypes.byref(nm), blocks.ctypes.data, cap,
                                       ctypes.byref(nb), st.ctypes.data)
        if rc == -4 and cap < (1 << 26):  # B200_ERR_CAPACITY
            cap = max(int(nm.value), int(nb.value)) + 16
            continue
        _check(rc)
        break
    n = sum(int(np.fromfile(p, dtype=np.uint64, count=1)[0]) for p in paths)
    return DemodResult(msgs[: int(nm.value)].copy(), st[0], blocks[: int(nb.value)].copy(), n)


def _check(rc: int):
    if rc != 0:
        raise B200Error(rc, load().b200_last_error().decode("utf-8", "replace"))


def _as_bytes(iq) -> np.ndarray:
    a = np.ascontiguousarray(iq)
    return a.view(np.uint8).reshape(-1)


@dataclass
class SpanResult:
    msgs: np.ndarray
    blocks: np.ndarray
    timing: dict


def format_beast(msgs: np.ndarray, net_verbatim: bool = True) -> bytes:
    """Beast binary frames of a message array (MSG_DTYPE), as modesSendBeastOutput writes them."""
    L = load()
    msgs = np.ascontiguousarray(msgs, dtype=MSG_DTYPE)
    need = int(L.b200_format_beast(msgs.ctypes.data, len(msgs), 1 if net_verbatim else 0, None, 0))
    buf = np.empty(max(need, 1), dtype=np.uint8)
    got = int(L.b200_format_beast(msgs.ctypes.data, len(msgs), 1 if net_verbatim else 0, buf.ctypes.data, need))
    assert got == need
    return buf[:need].tobytes()


def format_raw(msgs: np.ndarray, net_verbatim: bool = True, mlat: bool = False) -> bytes:
    """Raw-service lines of a message array, as modesSendRawOutput writes them."""
    L = load()
    msgs = np.ascontiguousarray(msgs, dtype=MSG_DTYPE)
    need = int(L.b200_format_raw(msgs.ctypes.data, len(msgs), 1 if net_verbatim else 0, 1 if mlat else 0, None, 0))
    buf = np.empty(max(need, 1), dtype=np.uint8)
    got = int(L.b200_format_raw(msgs.ctypes.data, len(msgs), 1 if net_verbatim else 0, 1 if mlat else 0, buf.ctypes.data, need))
    assert got == need
    return buf[:need].tobytes()


class Demodulator:
    """One receiver stream: converter + demodulator + CRC tables + ICAO filter state.

    Parameters mirror the reference's flags: fmt = --iformat, nfix = --fix/--no-fix/--aggressive,
    threshold = --preamble-threshold, modeac = --modeac, dcfilter = --dcfilter; table_bits = the
    SC16Q11_TABLE_BITS a reference build was compiled with (sc16q11 only; 0 = the float converter).
    """

    def __init__(self, fmt: str = "uc8", nfix: int = 1, threshold: int = 58,
                 block_samples: int = DEFAULT_BLOCK_SAMPLES, device: int = 0,
                 max_span_samples: int = 0, startup_time_ms: int = 0, modeac: bool = False, dcfilter: bool = False,
                 table_bits: int = 0):
        self._L = load()
        self.fmt = fmt
        self.bytes_per_sample = BYTES_PER_SAMPLE[fmt]
        self.block_samples = block_samples
        cfg = _Config(ABI_VERSION, device, FORMATS[fmt], nfix, threshold, block_samples, startup_time_ms, max_span_samples,
                      1 if modeac else 0, 1 if dcfilter else 0, table_bits, 0)
        h = ctypes.c_void_p()
        _check(self._L.b200_demod_create(ctypes.byref(cfg), ctypes.byref(h)))
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            self._L.b200_demod_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def reset(self):
        _check(self._L.b200_demod_reset(self._h))

    # ---- the hot path ----
    def _collect(self, copy: bool = True) -> SpanResult:
        if not copy:
            # the messages stay where the library put them (host memory it owns, b200_demod_messages): a view, valid
            # until the next process call -- a benchmark loop should not time a 68-byte-per-message memmove into numpy
            n = int(self._L.b200_demod_message_count(self._h))
            msgs = np.ctypeslib.as_array(ctypes.cast(self._L.b200_demod_messages(self._h), ctypes.POINTER(ctypes.c_uint8)),
                                         shape=(n * MSG_DTYPE.itemsize,)).view(MSG_DTYPE) if n else np.empty(0, dtype=MSG_DTYPE)
            return SpanResult(msgs, np.empty(0, dtype=BLOCK_DTYPE), self.timing())
        n = int(self._L.b200_demod_message_count(self._h))
        msgs = np.empty(n, dtype=MSG_DTYPE)
        if n:
            ctypes.memmove(msgs.ctypes.data, self._L.b200_demod_messages(self._h), n * MSG_DTYPE.itemsize)
        nb = int(self._L.b200_demod_block_count(self._h))
        blocks = np.empty(nb, dtype=BLOCK_DTYPE)
        if nb:
            ctypes.memmove(blocks.ctypes.data, self._L.b200_demod_blocks(self._h), nb * BLOCK_DTYPE.itemsize)
        return SpanResult(msgs, blocks, self.timing())

    def process(self, iq, final: bool = True) -> SpanResult:
        """Demodulate a span of raw IQ held in host memory (numpy array or anything buffer-like)."""
        b = _as_bytes(iq)
        n = b.size // self.bytes_per_sample
        _check(self._L.b200_demod_process(self._h, b.ctypes.data, n, FLAG_FINAL if final else 0))
        return self._collect()

    def process_ptr(self, host_ptr: int, nsamples: int, final: bool = True, copy: bool = True) -> SpanResult:
        _check(self._L.b200_demod_process(self._h, host_ptr, nsamples, FLAG_FINAL if final else 0))
        return self._collect(copy)

    def process_device(self, dev_ptr: int, nsamples: int, final: bool = True, stream: int = 0, copy: bool = True) -> SpanResult:
        """Demodulate a span already resident in device memory (e.g. a torch uint8 tensor's data_ptr())."""
        _check(self._L.b200_demod_process_device(self._h, dev_ptr, nsamples, FLAG_FINAL if final else 0, stream))
        return self._collect(copy)

    def stats(self) -> np.ndarray:
        s = np.zeros(1, dtype=STATS_DTYPE)
        _check(self._L.b200_demod_get_stats(self._h, s.ctypes.data))
        return s[0]

    def modeac_count(self) -> int:
        """Modes.stats_current.demod_modeac: Mode A/C replies decoded since create/reset."""
        return int(self._L.b200_demod_modeac_count(self._h))

    def crc_mismatches(self) -> int:
        """Kernel-vs-host CRC disagreements seen by the resolver (must be 0)."""
        return int(self.stats()["convert_cpu_s"])

    def timing(self) -> dict:
        t = Timing()
        _check(self._L.b200_demod_get_timing(self._h, ctypes.byref(t)))
        return t.as_dict()

    def run(self, iq, span_samples: int | None = None) -> DemodResult:
        """Whole stream in one or more spans -> the same record the oracles produce."""
        b = _as_bytes(iq)
        n = b.size // self.bytes_per_sample
        self.reset()
        msgs, blocks = [], []
        if span_samples is None:
            spans = [(0, n)]
        else:
            assert span_samples % self.block_samples == 0
            spans = [(s, min(n, s + span_samples)) for s in range(0, max(n, 1), span_samples)]
            if n and n % span_samples == 0:
                spans.append((n, n))  # the empty final span closes the stream
        for i, (s0, s1) in enumerate(spans):
            r = self.process(b[s0 * self.bytes_per_sample: s1 * self.bytes_per_sample], final=(i == len(spans) - 1))
            msgs.append(r.msgs)
            blocks.append(r.blocks)
        st = self.stats().copy()
        st["convert_cpu_s"] = 0.0
        st["demod_cpu_s"] = 0.0
        return DemodResult(np.concatenate(msgs) if msgs else np.empty(0, MSG_DTYPE), st,
                           np.concatenate(blocks) if blocks else np.empty(0, BLOCK_DTYPE), n)

    # ---- kernel-level entry points ----
    def scan_device(self, dev_ptr: int, nsamples: int, mode: int = 0, stream: int = 0):
        """K1 alone over a device-resident span; returns (milliseconds, candidate count)."""
        ms = ctypes.c_float()
        nc = ctypes.c_uint64()
        _check(self._L.b200_scan_device(self._h, dev_ptr, nsamples, mode, stream, ctypes.byref(ms), ctypes.byref(nc)))
        return ms.value, int(nc.value)

    def convert(self, iq):
        """iq_convert_fn: (u16 magnitudes, mean_level, mean_power)."""
        b = _as_bytes(iq)
        n = b.size // self.bytes_per_sample
        mag = np.empty(n, dtype=np.uint16)
        ml, mp = ctypes.c_double(), ctypes.c_double()
        _check(self._L.b200_convert(self._h, b.ctypes.data, n, mag.ctypes.data, ctypes.byref(ml), ctypes.byref(mp)))
        return mag, ml.value, mp.value

    def uc8_table(self) -> np.ndarray:
        t = np.empty(65536, dtype=np.uint16)
        _check(self._L.b200_uc8_table(self._h, t.ctypes.data))
        return t

    def debug_scan(self, iq):
        """(try mask per scan position, class records) of K1 over a stream-start span."""
        b = _as_bytes(iq)
        n = b.size // self.bytes_per_sample
        masks = np.zeros(n, dtype=np.uint8)
        cap = max(n * 5, 16)
        recs = np.zeros(cap, dtype=PHASE_RECORD_DTYPE)
        nrec = ctypes.c_uint64()
        _check(self._L.b200_debug_scan(self._h, b.ctypes.data, n, masks.ctypes.data, recs.ctypes.data, cap,
                                       ctypes.byref(nrec)))
        return masks, recs[: int(nrec.value)]

    def crc_batch(self, frames14: np.ndarray):
        """modesChecksum + modesChecksumDiagnose for n frames (n x 14 uint8)."""
        f = np.ascontiguousarray(frames14, dtype=np.uint8).reshape(-1, 14)
        n = len(f)
        syn = np.zeros(n, dtype=np.uint32)
        err = np.zeros(n, dtype=np.int8)
        bits = np.zeros((n, 2), dtype=np.int8)
        _check(self._L.b200_crc_batch(self._h, f.ctypes.data, n, syn.ctypes.data, err.ctypes.data, bits.ctypes.data))
        return syn, err, bits

    def error_table(self, bits: int) -> np.ndarray:
        n = self._L.b200_error_table(self._h, bits, None, 0)
        if n < 0:
            _check(n)
        t = np.zeros(max(n, 1), dtype=ERRORINFO_DTYPE)
        self._L.b200_error_table(self._h, bits, t.ctypes.data, n)
        return t[:n]
