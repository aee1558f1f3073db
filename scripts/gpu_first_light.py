"""Development aid: run the CUDA path against both oracles on a few seeded streams (needs a B200)."""
import sys, time, traceback
import numpy as np
sys.path.insert(0, ".")
from readsb_protobuf_b200 import api, synth, results
from oracle import port, ref

def main():
    cases = [
        ("c0-1s", synth.baseline_config(0), {}),
        ("tiny", synth.SynthConfig(seed=9, nsamples=5000, frames_per_s=4000), {}),
        ("dense2s", synth.baseline_config(3, seconds=2.0), {}),
        ("sc16", synth.baseline_config(2, seconds=1.0), {}),
        ("sc16q11", synth.SynthConfig(seed=33, nsamples=1_500_000, fmt="sc16q11", frames_per_s=1000, frac_biterror=0.2), {}),
        ("nfix2", synth.baseline_config(3, seconds=1.0, seed=77), {"nfix": 2}),
        ("nfix0", synth.baseline_config(3, seconds=1.0, seed=78), {"nfix": 0}),
        ("multiple", synth.SynthConfig(seed=5, nsamples=131072 * 3, frames_per_s=2000), {}),
        ("ragged", synth.SynthConfig(seed=6, nsamples=1_000_003, frames_per_s=3000, frac_biterror=0.3), {"block_samples": 50000}),
    ]
    bad = 0
    for name, cfg, kw in cases:
        iq, frames = synth.generate(cfg)
        want = port.run(iq, cfg.fmt, **kw)
        try:
            with api.Demodulator(fmt=cfg.fmt, **kw) as d:
                t = time.time()
                got = d.run(iq)
                dt = time.time() - t
                tim = d.timing()
                mism = d.crc_mismatches()
            rtol = 0.0 if cfg.fmt == "uc8" else 1e-5
            diffs = results.compare_results(got, want, float_rtol=rtol)
            print(f"{name}: frames={len(frames)} msgs gpu={len(got.msgs)} oracle={len(want.msgs)} "
                  f"{'OK' if not diffs else 'DIFF'} mism={mism} wall={dt*1e3:.1f}ms timing={tim}")
            for x in diffs[:8]:
                print("   ", x)
            bad += bool(diffs)
        except Exception:
            traceback.print_exc()
            bad += 1
    # span-split run must equal the one-shot run
    cfg = synth.baseline_config(3, seconds=3.0, seed=99)
    iq, _ = synth.generate(cfg)
    want = port.run(iq, cfg.fmt)
    with api.Demodulator() as d:
        got = d.run(iq, span_samples=131072 * 4)
    diffs = results.compare_results(got, want)
    print("spans:", "OK" if not diffs else diffs[:8])
    bad += bool(diffs)
    print("FAILED" if bad else "ALL OK")
    return bad

if __name__ == "__main__":
    sys.exit(main())
