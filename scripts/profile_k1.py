"""Development aid: a few device-resident passes of the workload, for `ncu` captures of K1/K2."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from readsb_protobuf_b200 import api, synth

seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
fmt = sys.argv[3] if len(sys.argv) > 3 else "uc8"
cfg = synth.baseline_config(1, seconds=seconds)
if fmt != "uc8":
    cfg = synth.SynthConfig(seed=3, nsamples=cfg.nsamples, fmt=fmt, frames_per_s=200.0)
iq, _ = synth.generate(cfg)
dev = torch.from_numpy(iq).cuda()
d = api.Demodulator(fmt=fmt, max_span_samples=cfg.nsamples + (1 << 20))
for i in range(reps):
    d.reset()
    r = d.process_device(dev.data_ptr(), cfg.nsamples, final=True, stream=torch.cuda.current_stream().cuda_stream)
    print(i, len(r.msgs), r.timing)
for mode in (0, 1, 2):
    for i in range(reps):
        ms, nc = d.scan_device(dev.data_ptr(), cfg.nsamples, mode=mode, stream=torch.cuda.current_stream().cuda_stream)
        print("scan mode", mode, "ms", ms, "cands", nc, "Gsamples/s", cfg.nsamples / ms / 1e6, "GB/s", cfg.nsamples * synth.BYTES_PER_SAMPLE[fmt] / ms / 1e6)
