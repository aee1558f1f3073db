import sys; sys.path.insert(0, ".")
import torch, numpy as np
from readsb_protobuf_b200 import api, synth
cfg = synth.baseline_config(1, seconds=60.0)
host = torch.empty(cfg.nsamples*2, dtype=torch.uint8, pin_memory=True)
frames = synth.plan(cfg); synth.render(cfg, frames, out=host.numpy())
d = api.Demodulator(max_span_samples=cfg.nsamples + (1 << 20))
for i in range(4):
    d.reset()
    r = d.process_ptr(host.data_ptr(), cfg.nsamples, final=True)
    print(i, len(r.msgs), {k: round(v,3) if isinstance(v,float) else v for k,v in r.timing.items()})
# plain H2D timing
dev = torch.empty_like(host, device="cuda")
for i in range(3):
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record(); dev.copy_(host, non_blocking=True); e1.record(); torch.cuda.synchronize()
    print("plain H2D ms", e0.elapsed_time(e1))
