"""Development aid: launches every kernel of the library a few times, for one `ncu --set full` capture.

    ncu --set full --clock-control none --import-source on -o gpurun_out/r02_all python scripts/profile_all_kernels.py

One pipeline chunk each of: sparse uc8 (configs[1] traffic), dense uc8 (configs[3] traffic), sc16 (float converter +
float_block_sums_kernel), uc8 with --modeac, uc8 with --dcfilter (dc_prepare / dc_chain / dc_magnitude), sc16q11 through
the 8-bit table; then the boundary kernels (convert_kernel, crc_batch_kernel).  Spans are sized to one chunk
(about one wave of K1a tiles) so that the capture stays short.
"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from readsb_protobuf_b200 import api, synth  # noqa: E402

SECONDS = float(sys.argv[1]) if len(sys.argv) > 1 else 8.0  # 19.2 M samples: one chunk


def run(label, cfg, **flags):
    iq, _ = synth.generate(cfg)
    dev = torch.from_numpy(iq).cuda()
    with api.Demodulator(fmt=cfg.fmt, max_span_samples=cfg.nsamples + (1 << 20), **flags) as d:
        r = d.process_device(dev.data_ptr(), cfg.nsamples, final=True, stream=torch.cuda.current_stream().cuda_stream)
        print(label, len(r.msgs), {k: round(v, 3) if isinstance(v, float) else v for k, v in r.timing.items()})
    del dev


n = int(SECONDS * synth.SAMPLE_RATE)
run("sparse uc8", synth.SynthConfig(seed=2, nsamples=n, frames_per_s=200.0))
run("dense uc8", synth.SynthConfig(seed=4, nsamples=n, frames_per_s=5000.0, frac_biterror=0.2))
run("sc16", synth.SynthConfig(seed=3, nsamples=n, fmt="sc16", frames_per_s=200.0))
run("uc8 modeac", synth.SynthConfig(seed=5, nsamples=n // 2, frames_per_s=500.0, modeac_per_s=500.0), modeac=True)
run("uc8 dcfilter", synth.SynthConfig(seed=6, nsamples=n // 4, frames_per_s=500.0), dcfilter=True)
run("sc16q11 table8", synth.SynthConfig(seed=7, nsamples=n // 2, fmt="sc16q11", frames_per_s=500.0), table_bits=8)

with api.Demodulator(fmt="uc8") as d:
    rng = np.random.default_rng(1)
    mag, ml, mp = d.convert(rng.integers(0, 256, size=2 * 131072, dtype=np.uint8))
    syn, err, bits = d.crc_batch(rng.integers(0, 256, size=(65536, 14), dtype=np.uint8))
    print("convert / crc_batch", float(ml), int(np.count_nonzero(err >= 0)))
