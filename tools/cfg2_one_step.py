"""One eager cfg2 training step (B=1024, G=1000, soft loss) for a kernel launch list under ncu."""
import os, sys
import torch
from torch import nn
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mclstexp_b200 import model as mm
G = int(os.environ.get("G", 1000))
torch.manual_seed(0)
net = mm.mclSTExp_Attention("none", 1.0, 1024, G, 256, 8, 64, 2, targets="soft")
net.image_encoder = nn.Identity()
net = net.cuda()
g = torch.Generator(device="cuda"); g.manual_seed(7)
batch = {"image": torch.randn(1024, 1024, generator=g, device="cuda"),
         "expression": torch.rand(1024, G, generator=g, device="cuda"),
         "position": torch.randint(0, 64, (1024, 2), generator=g, device="cuda").float()}
for _ in range(int(os.environ.get("STEPS", 3))):
    net.zero_grad(set_to_none=True)
    net(batch).backward()
torch.cuda.synchronize()
