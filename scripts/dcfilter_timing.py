"""--dcfilter on the full-size BASELINE configs[1] stream (60 s of uc8): time of the DC front end + pipeline,
against the plain path.  The DC block is one dependent float chain per rail over the whole stream
(convert.c:136-138), so this path is bound by 8 cycles per sample on one warp, not by bandwidth."""
import sys, time
sys.path.insert(0, ".")
import torch
from readsb_protobuf_b200 import api, synth

cfg = synth.baseline_config(1, seconds=float(sys.argv[1]) if len(sys.argv) > 1 else 60.0)
iq, _ = synth.generate(cfg)
dev = torch.from_numpy(iq).cuda()
s = torch.cuda.current_stream().cuda_stream
for dc in (False, True):
    with api.Demodulator(fmt="uc8", dcfilter=dc, max_span_samples=cfg.nsamples + (1 << 20)) as d:
        best = 1e9
        for _ in range(3):
            d.reset()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            r = d.process_device(dev.data_ptr(), cfg.nsamples, final=True, stream=s)
            torch.cuda.synchronize()
            best = min(best, time.perf_counter() - t0)
        print(f"dcfilter={dc}: {best * 1e3:.2f} ms for {cfg.nsamples} samples = {cfg.nsamples / best / 1e6:.0f} Msamples/s, "
              f"{len(r.msgs)} messages")
