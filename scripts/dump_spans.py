import sys, os
sys.path.insert(0, ".")
import torch
from readsb_protobuf_b200 import api, synth
os.makedirs("gpurun_out/spans", exist_ok=True)
cfg = synth.baseline_config(1, seconds=60.0)
iq, _ = synth.generate(cfg)
dev = torch.from_numpy(iq).cuda()
d = api.Demodulator(fmt="uc8", max_span_samples=cfg.nsamples + (1 << 20))
os.environ["B200_DUMP_SPAN"] = "gpurun_out/spans"
r = d.process_device(dev.data_ptr(), cfg.nsamples, final=True, stream=torch.cuda.current_stream().cuda_stream)
print(len(r.msgs), r.timing)
