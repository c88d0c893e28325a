"""Device time (CUDA events) and host time of every stage of one bench step."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import licos_b200 as L
from licos_b200 import ops, synth

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dev = torch.device("cuda:0")
torch.manual_seed(42)
net = L.image_models["bmshj2018-factorized"](quality=1, pretrained=False)
synth.condition_weights(net)
net = net.to(dev).eval()
eb = net.entropy_bottleneck
x = synth.make_input("rgb256", B, device=dev)

def timeit(name, fn, n=5):
    for _ in range(2):
        out = fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(n):
        out = fn()
    e1.record()
    host = (time.perf_counter() - t0) / n * 1e3
    torch.cuda.synchronize()
    print(f"{name:28s} device {e0.elapsed_time(e1) / n:8.3f} ms   host-issue {host:8.3f} ms")
    return out

with torch.no_grad():
    y = timeit("g_a(x)", lambda: net.g_a(x))
    yh, lik = timeit("eb(y) eval", lambda: eb(y))
    sym = timeit("eb.symbols(y)", lambda: eb.symbols(y))
    timeit("nchw_to_nhwc_bf16(y_hat)", lambda: ops.nchw_to_nhwc_bf16(yh))
    xh = timeit("g_s(y_hat)", lambda: net.g_s(yh))
    timeit("sum_log(lik)", lambda: ops.sum_log(lik))
    timeit("sum_sq_err(x_hat, x)", lambda: ops.sum_sq_err(xh, x))
    timeit("eb(y) noise(philox)", lambda: eb(y, training=True, seed=1))
    timeit("full step", lambda: (net.g_a(x), eb(y), eb.symbols(y), net.g_s(yh)))
