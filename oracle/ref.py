"""TEST INFRASTRUCTURE ONLY: runs the unmodified reference (oracle/_ref/ref_demod).

One subprocess per stream: the reference keeps function-static filter state
(icao_filter.c:151) that cannot be reset in-process.
"""
from __future__ import annotations

import os
import subprocess
import tempfile
from pathlib import Path

import numpy as np

from readsb_protobuf_b200 import build as _build
from readsb_protobuf_b200.results import DemodResult, read_result_file

HERE = Path(__file__).resolve().parent


def binary() -> Path | None:
    """Path of ref_demod; built from /root/reference when that exists, else a prebuilt copy."""
    return _build.ensure_ref()


def available() -> bool:
    return binary() is not None


def run_file(path, fmt: str = "uc8", nfix: int = 1, threshold: int = 58, block_samples: int | None = None,
             max_samples: int | None = None, repeat: int = 1, mag_out=None, modeac: bool = False,
             dcfilter: bool = False, table_bits: int = 0) -> DemodResult:
    exe = binary()
    if exe is None:
        raise RuntimeError("oracle/_ref/ref_demod is not built and /root/reference is absent")
    if table_bits:
        # the reference as its armhf package builds it (-DSC16Q11_TABLE_BITS=8, debian/rules:19), or with the larger
        # tables of its oneoff/convert_benchmark.c (9, 10, 11 bits)
        assert table_bits in (8, 9, 10, 11), "oracle/Makefile builds the table variants for 8..11 bits"
        exe = exe.with_name(f"ref_demod_tb{table_bits}")
    with tempfile.TemporaryDirectory() as td:
        out = os.path.join(td, "ref.res")
        cmd = [str(exe), "--in", str(path), "--out", out, "--format", fmt, "--nfix", str(nfix),
               "--threshold", str(threshold), "--repeat", str(repeat)]
        if block_samples:
            cmd += ["--block", str(block_samples)]
        if max_samples is not None:
            cmd += ["--max-samples", str(max_samples)]
        if mag_out:
            cmd += ["--mag-out", str(mag_out)]
        if modeac:
            cmd += ["--modeac"]
        if dcfilter:
            cmd += ["--dcfilter"]
        subprocess.run(cmd, check=True, capture_output=True)
        return read_result_file(out)


def run(iq: np.ndarray, fmt: str = "uc8", **kw) -> DemodResult:
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "in.bin")
        np.ascontiguousarray(iq).view(np.uint8).tofile(path)
        return run_file(path, fmt, **kw)


def magnitudes(iq: np.ndarray, fmt: str = "uc8", dcfilter: bool = False, table_bits: int = 0) -> np.ndarray:
    """The reference converter's u16 magnitudes for the whole stream."""
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "in.bin")
        mag = os.path.join(td, "mag.bin")
        np.ascontiguousarray(iq).view(np.uint8).tofile(path)
        run_file(path, fmt, mag_out=mag, dcfilter=dcfilter, table_bits=table_bits)
        return np.fromfile(mag, dtype=np.uint16)


def netfmt_binary() -> Path | None:
    p = HERE / "_ref" / "ref_netfmt"
    return p if p.exists() else None


def format_outputs(res: DemodResult, net_verbatim: bool = True, mlat: bool = False):
    """(beast bytes, raw bytes): the reference's own modesSendBeastOutput / modesSendRawOutput
    (net_io.c:769-896) over the messages of a result (oracle/ref_netfmt.c)."""
    from readsb_protobuf_b200.results import write_result_file
    exe = netfmt_binary()
    if exe is None:
        raise RuntimeError("oracle/_ref/ref_netfmt is not built and /root/reference is absent")
    with tempfile.TemporaryDirectory() as td:
        rp, bp, tp = os.path.join(td, "r.res"), os.path.join(td, "beast.bin"), os.path.join(td, "raw.txt")
        write_result_file(rp, res)
        cmd = [str(exe), "--in", rp, "--beast-out", bp, "--raw-out", tp]
        if mlat:
            cmd.append("--mlat")
        if not net_verbatim:
            cmd.append("--no-verbatim")
        subprocess.run(cmd, check=True, capture_output=True)
        return open(bp, "rb").read(), open(tp, "rb").read()
