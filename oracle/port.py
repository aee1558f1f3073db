"""TEST INFRASTRUCTURE ONLY: ctypes front-end of oracle/modes_oracle.c (the CPU restatement)."""
from __future__ import annotations

import ctypes
import subprocess
from pathlib import Path

import numpy as np

from readsb_protobuf_b200.results import BLOCK_DTYPE, MSG_DTYPE, STATS_DTYPE, DemodResult

HERE = Path(__file__).resolve().parent
FORMATS = {"uc8": 0, "sc16": 1, "sc16q11": 2}
OVERLAP = 326
BLOCK_SAMPLES = 131072

ERRORINFO_DTYPE = np.dtype([("syndrome", "<u4"), ("errors", "<i4"), ("bit", "i1", (2,)), ("padding", "<u2")])


class _Config(ctypes.Structure):
    _fields_ = [("format", ctypes.c_int32), ("nfix", ctypes.c_int32), ("threshold", ctypes.c_int32),
                ("block_samples", ctypes.c_uint32), ("modeac", ctypes.c_int32), ("dcfilter", ctypes.c_int32), ("sc16q11_table_bits", ctypes.c_int32)]


class _Result(ctypes.Structure):
    _fields_ = [("msgs", ctypes.c_void_p), ("n_msgs", ctypes.c_uint64), ("blocks", ctypes.c_void_p),
                ("n_blocks", ctypes.c_uint64), ("stats", ctypes.c_uint8 * STATS_DTYPE.itemsize),
                ("n_samples", ctypes.c_uint64)]


_lib = None


def build() -> Path:
    out = HERE / "_build" / "libmodes_oracle.so"
    srcs = [HERE / "modes_oracle.c", HERE / "modes_oracle.h"]
    if not out.exists() or any(s.stat().st_mtime > out.stat().st_mtime for s in srcs):
        subprocess.run(["make", "-C", str(HERE), "port", "CC=gcc"], check=True, capture_output=True)
    return out


def lib():
    global _lib
    if _lib is None:
        L = ctypes.CDLL(str(build()))
        L.mo_run.restype = ctypes.c_int
        L.mo_run.argtypes = [ctypes.POINTER(_Config), ctypes.c_void_p, ctypes.c_uint64, ctypes.POINTER(_Result)]
        L.mo_result_free.argtypes = [ctypes.POINTER(_Result)]
        L.mo_uc8_table.argtypes = [ctypes.c_void_p]
        L.mo_convert.restype = ctypes.c_int
        L.mo_convert.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_uint32, ctypes.c_void_p,
                                 ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double)]
        L.mo_format_beast.restype = ctypes.c_size_t
        L.mo_format_beast.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t]
        L.mo_format_raw.restype = ctypes.c_size_t
        L.mo_format_raw.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t]
        L.mo_checksum.restype = ctypes.c_uint32
        L.mo_checksum.argtypes = [ctypes.c_void_p, ctypes.c_int]
        L.mo_single_bit_syndrome.restype = ctypes.c_uint32
        L.mo_single_bit_syndrome.argtypes = [ctypes.c_int]
        L.mo_error_table.restype = ctypes.c_int
        L.mo_error_table.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int]
        L.mo_try_mask.restype = ctypes.c_int
        L.mo_try_mask.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_int]
        L.mo_slice.restype = None
        L.mo_slice.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
        _lib = L
    return _lib


def run(iq: np.ndarray, fmt: str = "uc8", nfix: int = 1, threshold: int = 58,
        block_samples: int = BLOCK_SAMPLES, modeac: bool = False, dcfilter: bool = False,
        table_bits: int = 0) -> DemodResult:
    """Demodulate a whole stream of raw IQ bytes with the CPU restatement."""
    iq = np.ascontiguousarray(iq).view(np.uint8).reshape(-1)
    bps = 2 if fmt == "uc8" else 4
    nsamples = iq.size // bps
    cfg = _Config(FORMATS[fmt], nfix, threshold, block_samples, 1 if modeac else 0, 1 if dcfilter else 0, table_bits)
    res = _Result()
    rc = lib().mo_run(ctypes.byref(cfg), iq.ctypes.data, nsamples, ctypes.byref(res))
    if rc != 0:
        raise ValueError("mo_run failed: %d" % rc)
    try:
        n = int(res.n_msgs)
        msgs = np.empty(n, dtype=MSG_DTYPE)
        if n:
            ctypes.memmove(msgs.ctypes.data, res.msgs, n * MSG_DTYPE.itemsize)
        nb = int(res.n_blocks)
        blocks = np.empty(nb, dtype=BLOCK_DTYPE)
        if nb:
            ctypes.memmove(blocks.ctypes.data, res.blocks, nb * BLOCK_DTYPE.itemsize)
        stats = np.frombuffer(bytes(res.stats), dtype=STATS_DTYPE)[0].copy()
        return DemodResult(msgs, stats, blocks, int(res.n_samples))
    finally:
        lib().mo_result_free(ctypes.byref(res))


def uc8_table() -> np.ndarray:
    t = np.empty(65536, dtype=np.uint16)
    lib().mo_uc8_table(t.ctypes.data)
    return t


def convert(iq: np.ndarray, fmt: str):
    """(mag u16 array, mean_level, mean_power) for one converter call over the whole input."""
    iq = np.ascontiguousarray(iq).view(np.uint8).reshape(-1)
    bps = 2 if fmt == "uc8" else 4
    n = iq.size // bps
    mag = np.empty(n, dtype=np.uint16)
    ml, mp = ctypes.c_double(), ctypes.c_double()
    rc = lib().mo_convert(FORMATS[fmt], iq.ctypes.data, n, mag.ctypes.data, ctypes.byref(ml), ctypes.byref(mp))
    assert rc == 0
    return mag, ml.value, mp.value


def sc16q11_table(bits: int = 8) -> np.ndarray:
    """init_sc16q11_lookup (convert.c:270-294) of a reference built with -DSC16Q11_TABLE_BITS=bits."""
    t = np.empty(1 << (2 * bits), dtype=np.uint16)
    lib().mo_sc16q11_table.argtypes = [ctypes.c_int, ctypes.c_void_p]
    lib().mo_sc16q11_table(bits, t.ctypes.data)
    return t


def convert_sc16q11_table(iq: np.ndarray, bits: int = 8):
    """convert_sc16q11_table (convert.c:296-328): (mag, mean_level, mean_power) of one converter call."""
    iq = np.ascontiguousarray(iq).view(np.uint8).reshape(-1)
    n = iq.size // 4
    mag = np.empty(n, dtype=np.uint16)
    ml, mp = ctypes.c_double(), ctypes.c_double()
    L = lib()
    L.mo_convert_sc16q11_table.restype = ctypes.c_int
    L.mo_convert_sc16q11_table.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_uint32, ctypes.c_void_p,
                                           ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double)]
    assert L.mo_convert_sc16q11_table(bits, iq.ctypes.data, n, mag.ctypes.data, ctypes.byref(ml), ctypes.byref(mp)) == 0
    return mag, ml.value, mp.value


class _DcState(ctypes.Structure):
    _fields_ = [("dc_a", ctypes.c_float), ("dc_b", ctypes.c_float), ("z1_I", ctypes.c_float), ("z1_Q", ctypes.c_float)]


def convert_dc(iq: np.ndarray, fmt: str, calls=None):
    """--dcfilter converters (convert.c:113-213, 374-423): magnitudes of the whole input, converted in
    consecutive calls of the given sample counts (the filter state runs on across calls), and the
    (mean_level, mean_power) of every call."""
    iq = np.ascontiguousarray(iq).view(np.uint8).reshape(-1)
    bps = 2 if fmt == "uc8" else 4
    n = iq.size // bps
    calls = [n] if calls is None else list(calls)
    assert sum(calls) == n
    L = lib()
    L.mo_dc_init.argtypes = [ctypes.POINTER(_DcState), ctypes.c_double]
    L.mo_convert_dc.restype = ctypes.c_int
    L.mo_convert_dc.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_uint32, ctypes.POINTER(_DcState), ctypes.c_void_p,
                                ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double)]
    st = _DcState()
    L.mo_dc_init(ctypes.byref(st), 2400000.0)
    mag = np.empty(n, dtype=np.uint16)
    means, at = [], 0
    for c in calls:
        ml, mp = ctypes.c_double(), ctypes.c_double()
        rc = L.mo_convert_dc(FORMATS[fmt], iq.ctypes.data + at * bps, c, ctypes.byref(st), mag.ctypes.data + 2 * at,
                             ctypes.byref(ml), ctypes.byref(mp))
        assert rc == 0
        means.append((ml.value, mp.value))
        at += c
    return mag, means


def checksum(msg: bytes) -> int:
    buf = (ctypes.c_uint8 * len(msg)).from_buffer_copy(msg)
    return int(lib().mo_checksum(buf, len(msg) * 8))


def error_table(nfix: int, bits: int) -> np.ndarray:
    cap = 8192
    t = np.zeros(cap, dtype=ERRORINFO_DTYPE)
    n = lib().mo_error_table(nfix, bits, t.ctypes.data, cap)
    assert n <= cap
    return t[:n]


def try_masks(mag: np.ndarray, threshold: int = 58) -> np.ndarray:
    """5-bit try mask for every scan position j with a full 19-sample window."""
    mag = np.ascontiguousarray(mag, dtype=np.uint16)
    n = max(len(mag) - 18, 0)
    out = np.zeros(n, dtype=np.uint8)
    L = lib()
    for j in range(n):
        out[j] = L.mo_try_mask(mag.ctypes.data, j, threshold)
    return out


def slice_bytes(mag: np.ndarray, j: int, try_phase: int, nbytes: int) -> bytes:
    mag = np.ascontiguousarray(mag, dtype=np.uint16)
    out = (ctypes.c_uint8 * nbytes)()
    lib().mo_slice(mag.ctypes.data, j, try_phase, nbytes, out)
    return bytes(out)


def format_beast(msgs: np.ndarray, net_verbatim: bool = True) -> bytes:
    """net_io.c:769-835 restated: Beast binary frames of a message array (MSG_DTYPE)."""
    msgs = np.ascontiguousarray(msgs, dtype=MSG_DTYPE)
    need = lib().mo_format_beast(msgs.ctypes.data, len(msgs), int(net_verbatim), None, 0)
    buf = np.empty(max(need, 1), dtype=np.uint8)
    lib().mo_format_beast(msgs.ctypes.data, len(msgs), int(net_verbatim), buf.ctypes.data, need)
    return buf[:need].tobytes()


def format_raw(msgs: np.ndarray, net_verbatim: bool = True, mlat: bool = False) -> bytes:
    """net_io.c:870-896 restated: raw-service lines of a message array."""
    msgs = np.ascontiguousarray(msgs, dtype=MSG_DTYPE)
    need = lib().mo_format_raw(msgs.ctypes.data, len(msgs), int(net_verbatim), int(mlat), None, 0)
    buf = np.empty(max(need, 1), dtype=np.uint8)
    lib().mo_format_raw(msgs.ctypes.data, len(msgs), int(net_verbatim), int(mlat), buf.ctypes.data, need)
    return buf[:need].tobytes()
