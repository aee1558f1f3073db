"""Development aid: two device-resident passes of the workload through the whole pipeline (K1a, K1b, K2
per chunk), for `ncu` captures."""
import sys
import torch
sys.path.insert(0, ".")
from readsb_protobuf_b200 import api, synth

seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
fmt = sys.argv[2] if len(sys.argv) > 2 else "uc8"
dense = len(sys.argv) > 3 and sys.argv[3] == "dense"  # the configs[3] traffic (5000 frames/s, 20 % one-bit errors)
cfg = synth.baseline_config(3 if dense else 1, seconds=seconds)
if fmt != "uc8":
    cfg = synth.SynthConfig(seed=3, nsamples=cfg.nsamples, fmt=fmt, frames_per_s=200.0)
iq, _ = synth.generate(cfg)
dev = torch.from_numpy(iq).cuda()
d = api.Demodulator(fmt=fmt, max_span_samples=cfg.nsamples + (1 << 20))
for i in range(2):
    d.reset()
    r = d.process_device(dev.data_ptr(), cfg.nsamples, final=True, stream=torch.cuda.current_stream().cuda_stream)
    print(i, len(r.msgs), r.timing)
