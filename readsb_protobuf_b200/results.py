"""Result records shared by the product, the oracles and the tests.

The binary layout is the one oracle/ref_harness.c writes ("MDSR" files) and the one the
C-ABI returns (include/readsb_b200.h: b200_message, b200_demod_stats).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

MSG_DTYPE = np.dtype(
    [
        ("timestampMsg", "<u8"),
        ("sysTimestampMsg", "<u8"),
        ("signalLevel", "<f8"),
        ("crc", "<u4"),
        ("addr", "<u4"),
        ("score", "<i4"),
        ("msgbits", "u1"),
        ("msgtype", "u1"),
        ("correctedbits", "u1"),
        ("reserved", "u1"),
        ("msg", "u1", (14,)),
        ("verbatim", "u1", (14,)),
    ],
    align=False,
)
assert MSG_DTYPE.itemsize == 68

STATS_DTYPE = np.dtype(
    [
        ("demod_preambles", "<u4"),
        ("demod_rejected_bad", "<u4"),
        ("demod_rejected_unknown_icao", "<u4"),
        ("demod_accepted", "<u4", (3,)),
        ("demod_preamblePhase", "<u4", (5,)),
        ("demod_bestPhase", "<u4", (5,)),
        ("strong_signal_count", "<u4"),
        ("messages_total", "<u4"),
        ("samples_processed", "<u8"),
        ("noise_power_count", "<u8"),
        ("signal_power_count", "<u8"),
        ("noise_power_sum", "<f8"),
        ("signal_power_sum", "<f8"),
        ("peak_signal_power", "<f8"),
        ("convert_cpu_s", "<f8"),
        ("demod_cpu_s", "<f8"),
    ],
    align=False,
)
assert STATS_DTYPE.itemsize == 136

BLOCK_DTYPE = np.dtype([("mean_level", "<f8"), ("mean_power", "<f8")])

HEADER_DTYPE = np.dtype([("magic", "S4"), ("version", "<u4"), ("n_msgs", "<u8"), ("n_blocks", "<u8"),
                         ("n_samples", "<u8")])

# integer counters compared bit-exactly between implementations
EXACT_STATS = (
    "demod_preambles", "demod_rejected_bad", "demod_rejected_unknown_icao", "demod_accepted",
    "demod_preamblePhase", "demod_bestPhase", "strong_signal_count", "messages_total",
    "samples_processed", "noise_power_count", "signal_power_count",
)
FLOAT_STATS = ("noise_power_sum", "signal_power_sum", "peak_signal_power")
# message fields compared bit-exactly (signalLevel is compared to 1e-5 per north_star, and is
# in fact identical: it is an integer sum divided twice)
EXACT_MSG_FIELDS = ("timestampMsg", "sysTimestampMsg", "crc", "addr", "score", "msgbits", "msgtype",
                    "correctedbits", "msg", "verbatim")


@dataclass
class DemodResult:
    msgs: np.ndarray  # MSG_DTYPE
    stats: np.ndarray  # STATS_DTYPE scalar (0-d structured)
    blocks: np.ndarray  # BLOCK_DTYPE
    n_samples: int


def read_result_file(path) -> DemodResult:
    raw = np.fromfile(path, dtype=np.uint8)
    hdr = raw[: HEADER_DTYPE.itemsize].view(HEADER_DTYPE)[0]
    assert hdr["magic"] == b"MDSR" and hdr["version"] == 1, hdr
    off = HEADER_DTYPE.itemsize
    stats = raw[off: off + STATS_DTYPE.itemsize].view(STATS_DTYPE)[0].copy()
    off += STATS_DTYPE.itemsize
    n = int(hdr["n_msgs"])
    msgs = raw[off: off + n * MSG_DTYPE.itemsize].view(MSG_DTYPE).copy()
    off += n * MSG_DTYPE.itemsize
    nb = int(hdr["n_blocks"])
    blocks = raw[off: off + nb * BLOCK_DTYPE.itemsize].view(BLOCK_DTYPE).copy()
    return DemodResult(msgs, stats, blocks, int(hdr["n_samples"]))


def write_result_file(path, res: DemodResult) -> None:
    """The inverse of read_result_file (same layout as oracle/ref_harness.c writes)."""
    hdr = np.zeros(1, dtype=HEADER_DTYPE)
    hdr["magic"] = b"MDSR"
    hdr["version"] = 1
    hdr["n_msgs"] = len(res.msgs)
    hdr["n_blocks"] = len(res.blocks)
    hdr["n_samples"] = res.n_samples
    with open(path, "wb") as f:
        f.write(hdr.tobytes())
        f.write(np.asarray(res.stats, dtype=STATS_DTYPE).reshape(1).tobytes())
        f.write(np.ascontiguousarray(res.msgs, dtype=MSG_DTYPE).tobytes())
        f.write(np.ascontiguousarray(res.blocks, dtype=BLOCK_DTYPE).tobytes())


def compare_results(got: DemodResult, want: DemodResult, float_rtol: float = 0.0,
                    signal_atol: float = 1e-5, check_blocks: bool = True) -> list[str]:
    """Differences between two results as human-readable strings (empty list = parity)."""
    diffs: list[str] = []
    if len(got.msgs) != len(want.msgs):
        diffs.append(f"message count {len(got.msgs)} != {len(want.msgs)}")
    n = min(len(got.msgs), len(want.msgs))
    for f in EXACT_MSG_FIELDS:
        a, b = got.msgs[f][:n], want.msgs[f][:n]
        bad = np.nonzero((a != b) if a.ndim == 1 else (a != b).any(axis=1))[0]
        if len(bad):
            i = int(bad[0])
            diffs.append(f"msg field {f}: {len(bad)} mismatches, first at #{i}: {a[i]!r} != {b[i]!r}")
    if n:
        d = np.abs(got.msgs["signalLevel"][:n] - want.msgs["signalLevel"][:n])
        if float(d.max()) > signal_atol:
            diffs.append(f"signalLevel max abs diff {float(d.max())} > {signal_atol}")
    for f in EXACT_STATS:
        if not np.array_equal(got.stats[f], want.stats[f]):
            diffs.append(f"stats.{f}: {got.stats[f]!r} != {want.stats[f]!r}")
    for f in FLOAT_STATS:
        a, b = float(got.stats[f]), float(want.stats[f])
        if np.isnan(a) and np.isnan(b):
            continue
        if not (abs(a - b) <= float_rtol * max(abs(a), abs(b))):
            diffs.append(f"stats.{f}: {a!r} != {b!r} (rtol {float_rtol})")
    if check_blocks:
        if len(got.blocks) != len(want.blocks):
            diffs.append(f"block count {len(got.blocks)} != {len(want.blocks)}")
        else:
            for f in ("mean_level", "mean_power"):
                a, b = got.blocks[f], want.blocks[f]
                both_nan = np.isnan(a) & np.isnan(b)
                ok = both_nan | (np.abs(a - b) <= float_rtol * np.maximum(np.abs(a), np.abs(b)))
                if not ok.all():
                    i = int(np.nonzero(~ok)[0][0])
                    diffs.append(f"block {f}: {int((~ok).sum())} mismatches, first at #{i}: {a[i]!r} != {b[i]!r}")
    return diffs
