"""Does capturing the codec step in a CUDA graph pay?  Eager launches vs graph replay of the same step."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import licos_b200 as L
from licos_b200 import synth
dev = torch.device("cuda:0")
torch.manual_seed(42)
net = L.image_models["bmshj2018-factorized"](quality=1, pretrained=False)
synth.condition_weights(net)
net = net.to(dev).eval()
eb = net.entropy_bottleneck
x = synth.make_input("rgb256", 256, device=dev)

def step():
    y = net.g_a(x)
    y_hat, lik, sym, y_nhwc = eb.forward_fused(y, want_symbols=True, want_nhwc=True)
    return net.g_s(y_hat, nhwc=y_nhwc), lik, sym

def timeit(fn, n=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

with torch.no_grad():
    print("eager ms/step", timeit(step))
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            step()
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        out = step()
    print("graph ms/step", timeit(g.replay))
    ref = step()
    g.replay()
    torch.cuda.synchronize()
    print("graph result equals eager:", all(torch.equal(a, b) for a, b in zip(out, ref)))
