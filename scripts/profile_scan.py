"""Development aid: two device-resident K1 launches (scan only, then scan+slice) for `ncu`."""
import sys
import torch
sys.path.insert(0, ".")
from readsb_protobuf_b200 import api, synth

seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 20.0
fmt = sys.argv[2] if len(sys.argv) > 2 else "uc8"
cfg = synth.baseline_config(1, seconds=seconds)
if fmt != "uc8":
    cfg = synth.SynthConfig(seed=3, nsamples=cfg.nsamples, fmt=fmt, frames_per_s=200.0)
iq, _ = synth.generate(cfg)
dev = torch.from_numpy(iq).cuda()
d = api.Demodulator(fmt=fmt, max_span_samples=cfg.nsamples + (1 << 20))
for mode in (0, 1, 2, 0, 1, 2):
    ms, nc = d.scan_device(dev.data_ptr(), cfg.nsamples, mode=mode, stream=torch.cuda.current_stream().cuda_stream)
    print("scan mode", mode, "ms", ms, "cands", nc)
