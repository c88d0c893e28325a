import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import licos_b200 as L
from licos_b200 import synth
from torch.profiler import ProfilerActivity, profile
dev = torch.device("cuda", 0)
torch.manual_seed(42)
net = L.image_models["bmshj2018-hyperprior"](quality=6, pretrained=False)
synth.condition_weights(net)
net = net.to(dev).eval()
x = synth.make_input("rgb1024", 1, seed=77, device=dev)
with torch.no_grad():
    for _ in range(3): net(x)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(5): net(x)
        torch.cuda.synchronize()
rows = [(e.key, e.device_time_total / 5 / 1e3, e.count // 5) for e in prof.key_averages()]
rows.sort(key=lambda r: -r[1])
tot = sum(r[1] for r in rows)
for k, t, n in rows[:22]: print(f"{t:8.4f} ms x{n:2d}  {k[:100]}")
print("total", tot)
