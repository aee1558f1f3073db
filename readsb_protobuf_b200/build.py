"""In-tree builds: the CUDA/C-ABI library, the workload generator and (test-only) the oracles.

Everything lands in git-ignored directories inside the repo so that the built files travel
to the GPU box with the gpurun snapshot:
    readsb_protobuf_b200/_build/libreadsb_b200.so    product: CUDA kernels + C-ABI + host resolver
    readsb_protobuf_b200/_build/libmodes_synth.so    workload generator
    oracle/_build/libmodes_oracle.so                 test infrastructure: C restatement
    oracle/_ref/{ref_demod,readsb_ref,crctests}      test infrastructure: the reference itself
"""
from __future__ import annotations

import os
import shutil
import subprocess
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
CSRC = PKG / "csrc"
BUILD = PKG / "_build"
ORACLE = ROOT / "oracle"
REFERENCE = Path(os.environ.get("READSB_REFERENCE", "/root/reference"))

NVCC_ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _newer(target: Path, sources) -> bool:
    if not target.exists():
        return False
    t = target.stat().st_mtime
    return all(Path(s).stat().st_mtime <= t for s in sources)


def _run(cmd, **kw):
    proc = subprocess.run([str(c) for c in cmd], capture_output=True, text=True, **kw)
    if proc.returncode != 0:
        raise RuntimeError("build failed: %s\n%s\n%s" % (" ".join(map(str, cmd)), proc.stdout, proc.stderr))
    return proc


def ensure_synth() -> Path:
    out = BUILD / "libmodes_synth.so"
    src = CSRC / "synth_iq.c"
    if not _newer(out, [src]):
        BUILD.mkdir(exist_ok=True)
        _run(["gcc", "-std=c11", "-O2", "-fopenmp", "-fPIC", "-shared", "-Wall", "-Wextra",
              "-o", out, src, "-lm"])
    return out


def cuda_sources():
    return sorted(CSRC.glob("*.cu")) + sorted(CSRC.glob("*.cc"))


def cuda_headers():
    return sorted(CSRC.glob("*.h")) + sorted(CSRC.glob("*.cuh")) + sorted(CSRC.glob("*.inl")) + sorted((ROOT / "include").glob("*.h"))


def find_nvcc() -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not Path(nvcc).exists():
        raise RuntimeError("nvcc not found; the CUDA library cannot be built (there is no CPU fallback)")
    return nvcc


def ensure_cuda(force: bool = False, verbose: bool = False) -> Path:
    """Compile the product library for sm_100a (nvcc cross-compiles without a GPU)."""
    out = BUILD / "libreadsb_b200.so"
    srcs = cuda_sources()
    if not force and _newer(out, list(srcs) + list(cuda_headers())):
        return out
    BUILD.mkdir(exist_ok=True)
    cmd = [find_nvcc(), *NVCC_ARCH, "-O3", "-lineinfo", "-std=c++17", "-shared",
           "-Xcompiler", "-fPIC,-Wall,-Wextra,-ffp-contract=off", "-fmad=false",
           "-I", ROOT / "include", "-I", CSRC, "-o", out, *srcs]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    proc = _run(cmd)
    if verbose:
        print(proc.stderr)
    return out


def ensure_oracle() -> Path:
    """TEST INFRASTRUCTURE: the C restatement of the reference path."""
    out = ORACLE / "_build" / "libmodes_oracle.so"
    if not _newer(out, [ORACLE / "modes_oracle.c", ORACLE / "modes_oracle.h"]):
        _run(["make", "-C", ORACLE, "port", "CC=gcc"])
    return out


def have_reference_sources() -> bool:
    return (REFERENCE / "demod_2400.c").exists()


def ensure_ref() -> Path | None:
    """TEST INFRASTRUCTURE: the unmodified reference compiled from /root/reference.

    Returns the ref_demod binary, building it when the reference sources are present (this
    container); on the GPU box only a prebuilt oracle/_ref/ can exist.
    """
    binary = ORACLE / "_ref" / "ref_demod"
    if have_reference_sources():
        netfmt = ORACLE / "_ref" / "ref_netfmt"
        deps = [ORACLE / "ref_harness.c", ORACLE / "ref_netfmt.c", ORACLE / "ref_shim" / "stubs.c", ORACLE / "Makefile"]
        if not _newer(binary, deps) or not _newer(netfmt, deps) or not _newer(ORACLE / "_ref" / "ref_demod_tb11", deps):
            _run(["make", "-C", ORACLE, "ref", "CC=gcc", f"REF={REFERENCE}"])
    return binary if binary.exists() else None


def ensure_dropin() -> Path | None:
    """TEST INFRASTRUCTURE: the reference's own program linked with the shim + libreadsb_b200.so
    in place of convert.o / demod_2400.o / sdr_ifile.o (oracle/_ref/readsb_b200)."""
    binary = ORACLE / "_ref" / "readsb_b200"
    if have_reference_sources():
        lib = ensure_cuda()
        shim = PKG / "shim" / "readsb_b200_shim.c"
        if not _newer(binary, [shim, lib, ORACLE / "Makefile", ROOT / "include" / "readsb_b200.h"]):
            _run(["make", "-C", ORACLE, "ref", "dropin", "CC=gcc", f"REF={REFERENCE}"])
    return binary if binary.exists() else None
