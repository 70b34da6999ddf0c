"""Kernel timeline of ONE warm training step replayed from the CUDA graph (C2, batch 32): name, stream,
start and duration of every kernel (CUPTI through torch.profiler), to show which branches of the step
run concurrently.  Output: gpurun_out/step_timeline.txt"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from torch.profiler import profile, ProfilerActivity  # noqa: E402
from fvta_memexqa_b200.config import fvta_config  # noqa: E402
from fvta_memexqa_b200.data import SyntheticDataset  # noqa: E402
from fvta_memexqa_b200.model import get_model  # noqa: E402
from fvta_memexqa_b200.trainer import Trainer  # noqa: E402

cfg = fvta_config(batch_size=32, seed=1234)
model = get_model(cfg)
trainer = Trainer(model, cfg).enable_graph()
data = SyntheticDataset(cfg, num_examples=32, seed=1234, pin=True)
batch = next(iter(data.get_batches(32, num_steps=1, shuffle=False)))
for _ in range(5):
    trainer.step(None, batch)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    trainer.step(None, batch)
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type.name == "CUDA" or "cuda" in str(e.device_type).lower()]
ev = sorted(ev, key=lambda e: e.time_range.start)
t0 = ev[0].time_range.start
streams = {}
lines = []
busy_until = 0.0
overlap = 0.0
for e in ev:
    s, d = e.time_range.start - t0, e.time_range.end - e.time_range.start
    sid = streams.setdefault(getattr(e, "device_resource_id", 0), len(streams))
    lines.append("%9.1f us  +%7.1f  stream %d  %s" % (s, d, sid, e.name[:90]))
    if s < busy_until:
        overlap += min(busy_until, s + d) - s
    busy_until = max(busy_until, s + d)
total = max(e.time_range.end for e in ev) - t0
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "step_timeline.txt"), "w") as f:
    f.write("one replay of the captured C2 step: %d kernels/copies on %d streams, %.1f us from first start to last "
            "end, %.1f us of kernel time overlapped with an earlier kernel\n" % (len(ev), len(streams), total, overlap))
    f.write("\n".join(lines) + "\n")
print(open(os.path.join(ROOT, "gpurun_out", "step_timeline.txt")).read()[:3000])
