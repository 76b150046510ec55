"""Kernel timeline of one CUDA-graph replay of the cfg2 training step (CUPTI through torch.profiler):
how much of the replay is kernel time and how much is the gap between dependent kernels."""
import os, sys, collections
import torch
from torch import nn
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mclstexp_b200 import model as mm
from mclstexp_b200.graphs import GraphedTrainStep
G = int(os.environ.get("G", 1000))
torch.manual_seed(0)
net = mm.mclSTExp_Attention("none", 1.0, 1024, G, 256, 8, 64, 2, targets="soft")
net.image_encoder = nn.Identity()
net = net.cuda()
g = torch.Generator(device="cuda"); g.manual_seed(7)
batch = {"image": torch.randn(1024, 1024, generator=g, device="cuda"),
         "expression": torch.rand(1024, G, generator=g, device="cuda"),
         "position": torch.randint(0, 64, (1024, 2), generator=g, device="cuda").float()}
step = GraphedTrainStep(net, batch, defer_weight_grads=os.environ.get("DEFER", "1") != "0")
for _ in range(5):
    step(batch)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step.graph.replay()
    torch.cuda.synchronize()
ev = sorted([e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range is not None],
            key=lambda e: e.time_range.start)
ev = [e for e in ev if "memcpy" not in e.name.lower() and "memset" not in e.name.lower()] or ev
t0, t1 = ev[0].time_range.start, max(e.time_range.end for e in ev)
busy = sum(e.time_range.end - e.time_range.start for e in ev)
gaps = [ev[i + 1].time_range.start - ev[i].time_range.end for i in range(len(ev) - 1)]
print(f"kernels {len(ev)}  span {(t1 - t0):.1f} us  sum of kernel durations {busy:.1f} us  "
      f"sum of positive gaps {sum(x for x in gaps if x > 0):.1f} us  median gap {sorted(gaps)[len(gaps)//2]:.2f} us")
if os.environ.get("DUMP"):
    with open(os.environ["DUMP"], "w") as f:
        for e in ev:
            f.write(f"{e.time_range.start - t0:9.1f} {e.time_range.end - e.time_range.start:8.2f}  "
                    f"{e.name.split('(')[0].replace('void ', '').replace('mclst::', '')[:70]}\n")
agg = collections.OrderedDict()
for e in ev:
    k = e.name.split("(")[0].replace("void ", "").replace("mclst::", "")[:60]
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += e.time_range.end - e.time_range.start
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]:
    print(f"{k:62s} {n:4d} {t:9.1f} us  {t / n:7.2f} us each")
