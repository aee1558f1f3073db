"""Regenerates tests/golden/*: run in the build container, where /root/reference exists.

The reference's own tests hold no vector for this path, so the golden
results here are OUTPUTS OF THE UNMODIFIED REFERENCE (oracle/_ref/ref_demod, built by
oracle/Makefile from /root/reference) on small seeded IQ streams.  Each fixture is
    <name>.npz : iq (raw little-endian IQ bytes), msgs / stats / blocks (reference results),
                 meta (JSON: format, flags, generator config, sha256 of the IQ)
Small enough to commit; they travel to the GPU box, where /root/reference does not exist.

    python tests/golden/make_golden.py
"""
from __future__ import annotations

import json
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, str(ROOT))

from oracle import ref  # noqa: E402
from readsb_protobuf_b200 import synth  # noqa: E402

# name -> (generator config, demodulator flags)
FIXTURES = {
    "uc8_fix1": (synth.SynthConfig(seed=101, nsamples=200_000, fmt="uc8", frames_per_s=3000, frac_biterror=0.2),
                 dict(nfix=1, threshold=58, block_samples=131072)),
    "uc8_fix2_aggressive": (synth.SynthConfig(seed=102, nsamples=150_000, fmt="uc8", frames_per_s=4000, frac_biterror=0.4),
                            dict(nfix=2, threshold=58, block_samples=131072)),
    "uc8_nofix_thr75": (synth.SynthConfig(seed=103, nsamples=150_000, fmt="uc8", frames_per_s=3000, frac_biterror=0.2),
                        dict(nfix=0, threshold=75, block_samples=131072)),
    # stream length an exact multiple of the block: the reference then demodulates an empty
    # final block whose mean power is 0/0 (convert.c:108-110, demod_2400.c:425)
    "uc8_whole_blocks": (synth.SynthConfig(seed=104, nsamples=3 * 32768, fmt="uc8", frames_per_s=3000),
                         dict(nfix=1, threshold=58, block_samples=32768)),
    "sc16": (synth.SynthConfig(seed=105, nsamples=120_000, fmt="sc16", frames_per_s=3000, frac_biterror=0.2),
             dict(nfix=1, threshold=58, block_samples=131072)),
    # Mode A/C replies next to Mode S traffic, --modeac: a block's replies follow its Mode S messages
    "uc8_modeac": (synth.SynthConfig(seed=107, nsamples=300_000, fmt="uc8", frames_per_s=1500, frac_biterror=0.2,
                                     modeac_per_s=4000),
                   dict(nfix=1, threshold=58, block_samples=131072, modeac=True)),
    # the reference as its armhf package is built (-DSC16Q11_TABLE_BITS=8, debian/rules:19): convert_sc16q11_table
    "sc16q11_table8": (synth.SynthConfig(seed=110, nsamples=100_000, fmt="sc16q11", frames_per_s=3000, frac_biterror=0.2,
                                         modeac_per_s=2000, amp_max=1.3),
                       dict(nfix=1, threshold=58, block_samples=32768, table_bits=8, modeac=True)),
    "sc16q11": (synth.SynthConfig(seed=106, nsamples=120_000, fmt="sc16q11", frames_per_s=3000, frac_biterror=0.2),
                dict(nfix=1, threshold=58, block_samples=131072)),
}

DC_FIXTURES = {
    "uc8_dcfilter": (synth.SynthConfig(seed=108, nsamples=120_000, fmt="uc8", frames_per_s=3000, frac_biterror=0.2,
                                       modeac_per_s=3000), (9, -6)),
    "sc16_dcfilter": (synth.SynthConfig(seed=109, nsamples=80_000, fmt="sc16", frames_per_s=3000, frac_biterror=0.2,
                                        modeac_per_s=3000), (1500, -900)),
}


def add_dc_offset(iq: np.ndarray, fmt: str, di: int, dq: int) -> np.ndarray:
    if fmt == "uc8":
        v = iq.astype(np.int32).reshape(-1, 2) + [di, dq]
        return np.clip(v, 0, 255).astype(np.uint8).reshape(-1)
    full = 32767 if fmt == "sc16" else 2047
    v = iq.view("<i2").astype(np.int32).reshape(-1, 2) + [di, dq]
    return np.clip(v, -full - 1, full).astype("<i2").reshape(-1).view(np.uint8)


# the one known-answer frame in the reference tree (comment at net_io.c:1645)
KAT_FRAME_HEX = "8D4B969699155600E87406F5B69F"


def render_single_frame(frame: bytes, start_tick: int, nsamples: int, amp: float = 0.5) -> np.ndarray:
    """Noise-free uc8 rendering of one frame on the 12 MHz grid (same envelope as synth_iq.c)."""
    ticks = np.zeros(96 + len(frame) * 8 * 12 + 16, dtype=np.float64)
    for p in (0, 12, 42, 54):
        ticks[p:p + 6] = 1
    for i in range(len(frame) * 8):
        bit = (frame[i >> 3] >> (7 - (i & 7))) & 1
        o = 96 + 12 * i + (0 if bit else 6)
        ticks[o:o + 6] = 1
    env = np.zeros(nsamples * 5 + len(ticks) + 8)
    env[start_tick:start_tick + len(ticks)] = ticks
    mag = env[: nsamples * 5].reshape(nsamples, 5).mean(axis=1) * amp
    i = np.clip(np.rint(mag * 127.5 + 127.5), 0, 255).astype(np.uint8)
    q = np.full(nsamples, 128, dtype=np.uint8)
    return np.stack([i, q], axis=1).reshape(-1)


def df18_frames():
    """DF18 frames for every CF value and the ME types whose IMF bit decodeExtendedSquitter looks at
    (mode_s.c:1373-1470): the reference flags mm->addr with MODES_NON_ICAO_ADDRESS for some of them."""
    rng = np.random.default_rng(18)
    frames = []
    metypes = [0, 4, 5, 8, 9, 18, 19, 20, 22, 23, 28, 29, 31]
    for cf in range(8):
        for metype in metypes:
            for variant in range(3):
                msg = bytearray(14)
                msg[0] = (18 << 3) | cf
                msg[1:4] = bytes(rng.integers(1, 255, 3, dtype=np.uint8))
                me = bytearray(rng.integers(0, 256, 7, dtype=np.uint8))
                sub = [1, 0, 5][variant] if metype in (19, 28) else int(rng.integers(0, 8))
                me[0] = (metype << 3) | sub
                msg[4:11] = me
                pi = synth._load().synth_crc24(bytes(msg), 14)
                msg[11:14] = bytes([(pi >> 16) & 0xff, (pi >> 8) & 0xff, pi & 0xff])
                frames.append(bytes(msg))
    return frames


def all_df_frames():
    """Every downlink format demodulate2400 accepts (DF 0, 4, 5, 11, 16, 17, 18, 20, 21, 24-31) from a pool
    of six aircraft: squitters first so that Address/Parity replies find their address in the ICAO
    filter, all-call replies with IID != 0, and one-bit errors in each family."""
    crc24 = synth._load().synth_crc24
    rng = np.random.default_rng(77)
    icaos = [int(x) for x in rng.integers(1, 0xffffff, 6)]

    def es(df, icao):
        m = bytearray(14)
        m[0] = (df << 3) | int(rng.integers(0, 8))
        m[1:4] = icao.to_bytes(3, "big")
        m[4:11] = bytes(rng.integers(0, 256, 7, dtype=np.uint8))
        m[11:14] = crc24(bytes(m), 14).to_bytes(3, "big")
        return bytes(m)

    def ap(df, icao, nbytes):
        m = bytearray(nbytes)
        m[0] = (df << 3) | int(rng.integers(0, 8))
        m[1:nbytes - 3] = bytes(rng.integers(0, 256, nbytes - 4, dtype=np.uint8))
        m[nbytes - 3:] = (crc24(bytes(m), nbytes) ^ icao).to_bytes(3, "big")
        return bytes(m)

    def df11(icao, iid=0):
        m = bytearray(7)
        m[0] = (11 << 3) | 5
        m[1:4] = icao.to_bytes(3, "big")
        m[4:7] = (crc24(bytes(m), 7) ^ iid).to_bytes(3, "big")
        return bytes(m)

    def flip(frame, lo, hi):
        f = bytearray(frame)
        b = int(rng.integers(lo, hi))
        f[b >> 3] ^= 0x80 >> (b & 7)
        return bytes(f)

    frames = []
    for ic in icaos:
        frames += [es(17, ic), df11(ic)]
    for _ in range(12):
        for ic in icaos:
            for df, nb in ((0, 7), (4, 7), (5, 7), (16, 14), (20, 14), (21, 14), (24, 14), (25, 14), (27, 14), (31, 14)):
                frames.append(ap(df, ic, nb))
            frames += [es(17, ic), es(18, ic), df11(ic, int(rng.integers(0, 64))),
                       flip(es(17, ic), 5, 112), flip(df11(ic), 5, 56), flip(ap(20, ic, 14), 5, 112)]
    return frames


def render_frames(frames, gap=600, amp=0.5):
    """Noise-free uc8 rendering of frames one after another, start ticks cycling through the five phases."""
    parts = []
    for i, f in enumerate(frames):
        parts.append(render_single_frame(f, 100 * 5 + (i % 5), gap, amp))
    return np.concatenate(parts)


def netfmt_messages():
    """Messages for the output writers (net_io.c:769-896): the uc8_modeac reference messages plus crafted
    ones that hit every escape and clamp: 0x1a in the timestamp, the signal byte and the payload, signal
    levels 0 / tiny / 1.0 / above full scale, a zero timestamp (raw falls back to '*'), a length that is
    not sent at all."""
    from readsb_protobuf_b200.results import MSG_DTYPE
    z = np.load(HERE / "uc8_modeac.npz")
    base = z["msgs"]
    rng = np.random.default_rng(26)
    n = 400
    m = np.zeros(n, dtype=MSG_DTYPE)
    m["msgbits"] = rng.choice([16, 56, 112, 112, 56, 40], n)
    m["timestampMsg"] = rng.integers(0, 1 << 48, n, dtype=np.uint64)
    m["msg"] = rng.integers(0, 256, (n, 14), dtype=np.uint8)
    m["verbatim"] = rng.integers(0, 256, (n, 14), dtype=np.uint8)
    m["signalLevel"] = rng.random(n) ** 4
    for i in range(0, n, 7):      # 0x1a everywhere an escape can be needed
        m["timestampMsg"][i] = (int(m["timestampMsg"][i]) & ~(0xff << (8 * (i % 6)))) | (0x1a << (8 * (i % 6)))
        m["msg"][i, i % 14] = 0x1a
        m["verbatim"][i, (i + 3) % 14] = 0x1a
    m["signalLevel"][::11] = (26.0 / 255.0) ** 2   # signal byte 0x1a
    m["signalLevel"][1::13] = 0.0
    m["signalLevel"][2::17] = 1e-12                # rounds to 0 -> forced to 1
    m["signalLevel"][3::19] = 1.0
    m["signalLevel"][4::23] = 1.7                  # above full scale -> 255
    m["timestampMsg"][5::29] = 0                   # raw: '*' even with mlat
    return np.concatenate([base, m])


def main(only=None):
    assert ref.available(), "needs /root/reference (or a prebuilt oracle/_ref)"
    for name, (cfg, flags) in FIXTURES.items():
        if only and name not in only:
            continue
        iq, frames = synth.generate(cfg)
        res = ref.run(iq, cfg.fmt, **flags)
        meta = dict(fmt=cfg.fmt, flags=flags, generator=cfg.__dict__, sha256=synth.sha256(iq), n_frames=len(frames),
                    source="oracle/_ref/ref_demod (unmodified reference objects)")
        np.savez_compressed(HERE / f"{name}.npz", iq=iq, msgs=res.msgs, stats=np.array([res.stats]), blocks=res.blocks,
                            meta=np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8))
        print(f"{name}: {len(frames)} frames, {len(res.msgs)} reference messages, {iq.nbytes} IQ bytes")

    # --dcfilter (convert_*_generic, convert.c:113-213, 374-423): the same generator streams seen through a
    # receiver with a DC offset on both rails, demodulated by the reference with its DC block on
    for name, (cfg, offs) in DC_FIXTURES.items():
        if only and name not in only:
            continue
        iq, frames = synth.generate(cfg)
        iq = add_dc_offset(iq, cfg.fmt, *offs)
        flags = dict(nfix=1, threshold=58, block_samples=32768, dcfilter=True, modeac=True)
        res = ref.run(iq, cfg.fmt, **flags)
        meta = dict(fmt=cfg.fmt, flags=flags, generator=cfg.__dict__, dc_offset=offs, sha256=synth.sha256(iq), n_frames=len(frames),
                    source="oracle/_ref/ref_demod --dcfilter --modeac (unmodified reference objects)")
        np.savez_compressed(HERE / f"{name}.npz", iq=iq, msgs=res.msgs, stats=np.array([res.stats]), blocks=res.blocks,
                            meta=np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8))
        print(f"{name}: {len(frames)} frames, {len(res.msgs)} reference messages, {iq.nbytes} IQ bytes")

    if not only or "uc8_df18" in only:
        frames = df18_frames()
        iq = render_frames(frames)
        res = ref.run(iq, "uc8")
        meta = dict(fmt="uc8", flags=dict(nfix=1, threshold=58, block_samples=131072), sha256=synth.sha256(iq), n_frames=len(frames),
                    source="oracle/_ref/ref_demod (unmodified reference objects)")
        np.savez_compressed(HERE / "uc8_df18.npz", iq=iq, msgs=res.msgs, stats=np.array([res.stats]), blocks=res.blocks,
                            meta=np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8))
        flagged = int(np.sum((res.msgs["addr"] >> 24) & 1))
        print(f"uc8_df18: {len(frames)} frames, {len(res.msgs)} reference messages, {flagged} with a non-ICAO address")
    for name, nfix in (("uc8_all_df", 1), ("uc8_all_df_nofix", 0)):
        if only and name not in only:
            continue
        frames = all_df_frames()
        iq = render_frames(frames)
        flags = dict(nfix=nfix, threshold=58, block_samples=131072)
        res = ref.run(iq, "uc8", **flags)
        meta = dict(fmt="uc8", flags=flags, sha256=synth.sha256(iq), n_frames=len(frames),
                    source="oracle/_ref/ref_demod (unmodified reference objects)")
        np.savez_compressed(HERE / f"{name}.npz", iq=iq, msgs=res.msgs, stats=np.array([res.stats]), blocks=res.blocks,
                            meta=np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8))
        print(f"{name}: {len(frames)} frames, {len(res.msgs)} reference messages, DFs",
              sorted(set(int(x) for x in res.msgs["msgtype"])))
    if not only or "netfmt" in only:
        from readsb_protobuf_b200.results import DemodResult
        msgs = netfmt_messages()
        res = DemodResult(msgs, np.zeros((), dtype=np.load(HERE / "uc8_modeac.npz")["stats"].dtype), np.zeros(0, dtype=np.load(HERE / "uc8_modeac.npz")["blocks"].dtype), 0)
        out = {"msgs": msgs}
        for verb in (0, 1):
            for mlat in (0, 1):
                beast, raw = ref.format_outputs(res, net_verbatim=bool(verb), mlat=bool(mlat))
                out[f"beast_v{verb}"] = np.frombuffer(beast, dtype=np.uint8)
                out[f"raw_v{verb}_m{mlat}"] = np.frombuffer(raw, dtype=np.uint8)
        np.savez_compressed(HERE / "netfmt.npz", **out)
        print(f"netfmt: {len(msgs)} messages, beast {len(out['beast_v1'])} bytes, raw {len(out['raw_v1_m1'])} bytes "
              "(oracle/_ref/ref_netfmt: the reference's own writers)")
    if only and "kat_frame" not in only:
        return
    frame = bytes.fromhex(KAT_FRAME_HEX)
    iq = np.concatenate([render_single_frame(frame, 100003, 24000), render_single_frame(frame, 20001, 8000)])
    res = ref.run(iq, "uc8")
    meta = dict(fmt="uc8", flags=dict(nfix=1, threshold=58, block_samples=131072), frame=KAT_FRAME_HEX,
                sha256=synth.sha256(iq), source="oracle/_ref/ref_demod; frame from the comment at net_io.c:1645")
    np.savez_compressed(HERE / "kat_frame.npz", iq=iq, msgs=res.msgs, stats=np.array([res.stats]), blocks=res.blocks,
                        meta=np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8))
    print("kat_frame:", [(int(m["timestampMsg"]), bytes(m["msg"]).hex(), int(m["score"])) for m in res.msgs])


if __name__ == "__main__":
    main(set(sys.argv[1:]) or None)
