"""Development aid / fixture generator (on a GPU box): records what the kernels hand to the host resolver.

    python scripts/dump_spans.py            # the 60 s bench stream -> gpurun_out/spans/ (tools/resolver_bench.cc)
    python scripts/dump_spans.py fixture    # the small seeded stream of tests/test_host.py -> gpurun_out/resolver_fixture/

The fixture (one file per pipeline chunk) is committed under tests/golden/ so that the CPU suite can run the host
resolver over real kernel outputs and compare with the oracle on the regenerated stream.
"""
import os
import sys

sys.path.insert(0, ".")
import torch  # noqa: E402

from readsb_protobuf_b200 import api, synth  # noqa: E402

fixture = len(sys.argv) > 1 and sys.argv[1] == "fixture"
dense = len(sys.argv) > 1 and sys.argv[1] == "dense"   # 20 s of the configs[3] stream -> gpurun_out/spans_dense/
out = "gpurun_out/resolver_fixture" if fixture else ("gpurun_out/spans_dense" if dense else "gpurun_out/spans")
os.makedirs(out, exist_ok=True)
cfg = synth.resolver_fixture_config() if fixture else (synth.baseline_config(3, seconds=float(os.environ.get("DUMP_SECONDS", "20"))) if dense else synth.baseline_config(1, seconds=60.0))
iq, _ = synth.generate(cfg)
dev = torch.from_numpy(iq).cuda()
d = api.Demodulator(fmt="uc8", max_span_samples=cfg.nsamples + (1 << 20))
os.environ["B200_DUMP_SPAN"] = out
s = torch.cuda.current_stream().cuda_stream
if fixture:
    # two process calls, so that the second file starts from the filter / statistics state the first one left
    cut = 4 * 131072
    r0 = d.process_device(dev.data_ptr(), cut, final=False, stream=s)
    r = d.process_device(dev.data_ptr() + 2 * cut, cfg.nsamples - cut, final=True, stream=s)
    print(len(r0.msgs) + len(r.msgs), sorted(os.listdir(out)))
else:
    r = d.process_device(dev.data_ptr(), cfg.nsamples, final=True, stream=s)
    print(len(r.msgs), r.timing, sorted(os.listdir(out)))
