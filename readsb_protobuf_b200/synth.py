"""Seeded synthetic Mode S IQ streams (ctypes front-end of csrc/synth_iq.c).

Workload shapes follow SURVEY.md section 8(d) / BASELINE.json `configs`; the generator is
not on the demodulation path.
"""
from __future__ import annotations

import ctypes
import hashlib
from dataclasses import dataclass

import numpy as np

from . import build as _build

FORMATS = {"uc8": 0, "sc16": 1, "sc16q11": 2}
BYTES_PER_SAMPLE = {"uc8": 2, "sc16": 4, "sc16q11": 4}
SAMPLE_RATE = 2_400_000


class _Cfg(ctypes.Structure):
    _fields_ = [
        ("seed", ctypes.c_uint64),
        ("nsamples", ctypes.c_uint64),
        ("format", ctypes.c_int32),
        ("n_icao", ctypes.c_int32),
        ("frames_per_s", ctypes.c_double),
        ("noise_sigma", ctypes.c_double),
        ("amp_min", ctypes.c_double),
        ("amp_max", ctypes.c_double),
        ("frac_biterror", ctypes.c_double),
        ("frac_df17", ctypes.c_double),
        ("frac_df11", ctypes.c_double),
        ("modeac_per_s", ctypes.c_double),
    ]


FRAME_DTYPE = np.dtype(
    [
        ("start_tick", "<u8"),
        ("amp", "<f4"),
        ("phase0", "<f4"),
        ("dphase", "<f4"),
        ("errbit", "<i2"),
        ("nbytes", "u1"),
        ("df", "u1"),
        ("msg", "u1", (14,)),
        ("pad", "u1", (2,)),
    ],
    align=False,
)


@dataclass(frozen=True)
class SynthConfig:
    seed: int
    nsamples: int
    fmt: str = "uc8"
    frames_per_s: float = 200.0
    noise_sigma: float = 0.02
    amp_min: float = 0.05
    amp_max: float = 0.9
    frac_biterror: float = 0.0
    n_icao: int = 200
    frac_df17: float = 0.6
    frac_df11: float = 0.2
    modeac_per_s: float = 0.0  # Mode A/C replies per second (frames with df == 32 in the plan)

    def _c(self) -> _Cfg:
        return _Cfg(
            self.seed, self.nsamples, FORMATS[self.fmt], self.n_icao, self.frames_per_s,
            self.noise_sigma, self.amp_min, self.amp_max, self.frac_biterror,
            self.frac_df17, self.frac_df11, self.modeac_per_s,
        )


_lib = None


def _load():
    global _lib
    if _lib is None:
        lib = ctypes.CDLL(str(_build.ensure_synth()))
        lib.synth_plan.restype = ctypes.c_int64
        lib.synth_plan.argtypes = [ctypes.POINTER(_Cfg), ctypes.c_void_p, ctypes.c_int64]
        lib.synth_render.restype = None
        lib.synth_render.argtypes = [
            ctypes.POINTER(_Cfg), ctypes.c_void_p, ctypes.c_int64,
            ctypes.c_uint64, ctypes.c_uint64, ctypes.c_void_p,
        ]
        lib.synth_frame_size.restype = ctypes.c_int
        assert lib.synth_frame_size() == FRAME_DTYPE.itemsize, (lib.synth_frame_size(), FRAME_DTYPE.itemsize)
        _lib = lib
    return _lib


def plan(cfg: SynthConfig) -> np.ndarray:
    """Ground-truth frame list (structured array, FRAME_DTYPE), sorted by start tick."""
    lib = _load()
    c = cfg._c()
    n = lib.synth_plan(ctypes.byref(c), None, 0)
    frames = np.zeros(max(int(n), 1), dtype=FRAME_DTYPE)
    got = lib.synth_plan(ctypes.byref(c), frames.ctypes.data, int(n))
    assert got == n
    return frames[: int(n)]


def render(cfg: SynthConfig, frames: np.ndarray | None = None, first: int = 0,
           count: int | None = None, out: np.ndarray | None = None) -> np.ndarray:
    """Render samples [first, first+count) as raw little-endian IQ bytes (uint8 array)."""
    lib = _load()
    if frames is None:
        frames = plan(cfg)
    if count is None:
        count = cfg.nsamples - first
    nbytes = count * BYTES_PER_SAMPLE[cfg.fmt]
    if out is None:
        out = np.empty(nbytes, dtype=np.uint8)
    assert out.dtype == np.uint8 and out.size >= nbytes and out.flags["C_CONTIGUOUS"]
    c = cfg._c()
    frames = np.ascontiguousarray(frames)
    lib.synth_render(ctypes.byref(c), frames.ctypes.data, len(frames), first, count, out.ctypes.data)
    return out[:nbytes]


def generate(cfg: SynthConfig):
    """(iq_bytes, frames) for the whole stream."""
    frames = plan(cfg)
    return render(cfg, frames), frames


def sha256(iq: np.ndarray) -> str:
    return hashlib.sha256(memoryview(np.ascontiguousarray(iq))).hexdigest()


def resolver_fixture_config() -> SynthConfig:
    """The small dense stream whose recorded kernel outputs (tests/golden/resolver_span_*.bin, written by
    scripts/dump_spans.py fixture on a B200) let the CPU suite run the host resolver against the oracle."""
    return SynthConfig(seed=77, nsamples=1_000_000, fmt="uc8", frames_per_s=3000.0, frac_biterror=0.3, n_icao=40)


# BASELINE.json `configs`, by index
def baseline_config(index: int, seed: int | None = None, seconds: float | None = None) -> SynthConfig:
    if index == 0:  # configs[0]: 1 s uc8 plumbing case
        return SynthConfig(seed=1 if seed is None else seed,
                           nsamples=int((seconds or 1.0) * SAMPLE_RATE), fmt="uc8", frames_per_s=200.0)
    if index == 1:  # configs[1]: 60 s uc8, ~200 frames/s, no bit errors
        return SynthConfig(seed=2 if seed is None else seed,
                           nsamples=int((seconds or 60.0) * SAMPLE_RATE), fmt="uc8", frames_per_s=200.0)
    if index == 2:  # configs[2]: 60 s sc16 (sc16q11: use .fmt override)
        return SynthConfig(seed=3 if seed is None else seed,
                           nsamples=int((seconds or 60.0) * SAMPLE_RATE), fmt="sc16", frames_per_s=200.0)
    if index in (3, 4):  # configs[3]/[4]: 10 min dense uc8 with 20 % one-bit errors
        return SynthConfig(seed=(4 if index == 3 else 10) if seed is None else seed,
                           nsamples=int((seconds or 600.0) * SAMPLE_RATE), fmt="uc8",
                           frames_per_s=5000.0, frac_biterror=0.2)
    raise ValueError(index)
