"""Two training steps (batch 8 x 320^2) for ncu launch lists."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from microbeseg_b200.unets import build_unet
from microbeseg_b200.training import TrainEngine, train_step
torch.manual_seed(0)
dev = torch.device("cuda:0")
net = build_unet("DU", "relu", "conv", "bn", dev, 1, filters=[64, 1024]).train()
eng = TrainEngine(net, use_graph=False)
opt = torch.optim.Adam(net.parameters(), lr=8e-4, amsgrad=True)
B, S = 8, 320
g = torch.Generator(device="cpu").manual_seed(0)
img = (torch.rand(B, 1, S, S, generator=g) * 2 - 1).to(dev)
bl, cl = torch.rand(B, 1, S, S, generator=g).to(dev), torch.rand(B, 1, S, S, generator=g).to(dev)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    loss = train_step(eng, opt, img, bl, cl)
torch.cuda.synchronize()
print("done", float(loss))
