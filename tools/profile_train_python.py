"""cProfile of the eager training loop (host side): where the ~30 us per launch go."""
import cProfile
import os
import pstats
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import licos_b200 as L  # noqa: E402
from licos_b200 import synth  # noqa: E402

dev = torch.device("cuda", 0)
torch.manual_seed(100)
net = L.image_models["bmshj2018-factorized"](quality=1, pretrained=False).to(dev).train()
crit = L.RateDistortionLoss(lmbda=1e-2)
opt = L.net_aux_optimizer(net, {"net": {"type": "Adam", "lr": 1e-4}, "aux": {"type": "Adam", "lr": 1e-3}})
x = synth.make_input("rgb256", 32, seed=0, device=dev)


def train_step():
    opt["net"].zero_grad(); opt["aux"].zero_grad()
    out = net(x)
    loss = crit(out, x)
    loss["loss"].backward()
    torch.nn.utils.clip_grad_norm_(net.parameters(), 1.0)
    opt["net"].step()
    aux = net.aux_loss()
    aux.backward()
    opt["aux"].step()


for _ in range(5):
    train_step()
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(50):
    train_step()
torch.cuda.synchronize()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(45)
