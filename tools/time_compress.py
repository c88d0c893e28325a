"""compress() end to end (g_a + symbols + rANS) with the device coder vs the host coder, 256 tiles."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import licos_b200 as L
from licos_b200 import synth
dev = torch.device("cuda:0")
torch.manual_seed(42)
net = L.image_models["bmshj2018-factorized"](quality=1, pretrained=False)
synth.condition_weights(net)
net.update()
net = net.to(dev).eval()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
x = synth.make_input("rgb256", B, device=dev)
for mode in (True, False, True):
    net.entropy_bottleneck.device_coder = mode
    for _ in range(2):
        out = net.compress(x)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 5
    for _ in range(n):
        out = net.compress(x)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / n
    nbytes = sum(len(s) for s in out["strings"][0])
    print(f"device_coder={mode}: {dt * 1e3:.2f} ms per {B} tiles = {B * 65536 / dt / 1e6:.0f} MPix/s, {nbytes / B:.0f} bytes/tile, "
          f"{nbytes * 8 / (B * 65536):.3f} bpp")

# breakdown of the device path
from licos_b200 import ops
eb = net.entropy_bottleneck
with torch.no_grad():
    y = net.g_a(x)
    sym = eb.symbols(y)
torch.cuda.synchronize()
def ev():
    return torch.cuda.Event(enable_timing=True)
for _ in range(2):
    e = [ev() for _ in range(3)]
    t0 = time.perf_counter()
    e[0].record()
    out = ops.rans_encode_device(sym.reshape(B, -1), None, 256, eb._quantized_cdf, eb._cdf_length, eb._offset)
    e[1].record()
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    print(f"rans_encode_device: wall {1e3 * (t1 - t0):.2f} ms, device span {e[0].elapsed_time(e[1]):.2f} ms")
import ctypes
from licos_b200._lib import lib
cap = sym[0].numel() + 1024
work = torch.empty((B, cap), dtype=torch.int32, device=dev)
rcp = torch.empty(eb._quantized_cdf.numel(), dtype=torch.int64, device=dev)
lengths = torch.empty(B, dtype=torch.int32, device=dev)
s2 = sym.reshape(B, -1).contiguous()
for _ in range(3):
    e0, e1 = ev(), ev()
    e0.record()
    lib.licos_rans_encode_device(s2.data_ptr(), None, 0, B, s2.shape[1], 256, eb._quantized_cdf.data_ptr(), eb._quantized_cdf.shape[0],
                                 eb._quantized_cdf.shape[1], eb._cdf_length.data_ptr(), eb._offset.data_ptr(), rcp.data_ptr(),
                                 work.data_ptr(), cap, lengths.data_ptr(), torch.cuda.current_stream().cuda_stream)
    e1.record()
    torch.cuda.synchronize()
    print(f"rans kernels only: {e0.elapsed_time(e1):.2f} ms")

# decompress(): device decoder + dequantise + g_s
comp = net.compress(x)
for _ in range(2):
    dec = net.decompress(comp["strings"], comp["shape"])
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    dec = net.decompress(comp["strings"], comp["shape"])
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 5
print(f"decompress: {dt * 1e3:.2f} ms per {B} tiles = {B * 65536 / dt / 1e6:.0f} MPix/s")
strings = list(comp["strings"][0])
for _ in range(3):
    e0, e1 = ev(), ev()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0.record()
    symd = ops.rans_decode_device(strings, None, sym[0].numel(), 256, eb._quantized_cdf, eb._cdf_length, eb._offset)
    e1.record()
    torch.cuda.synchronize()
    print(f"rans_decode_device: wall {1e3 * (time.perf_counter() - t0):.2f} ms, device span {e0.elapsed_time(e1):.2f} ms, "
          f"equal to the encoder's input: {bool(torch.equal(symd.reshape(sym.shape), sym))}")

# decoder kernel alone (strings already packed on the device)
import numpy as np
sizes = [len(s) for s in strings]
words = np.frombuffer(b"".join(strings), dtype=np.uint32)
n_words = np.asarray([sz // 4 for sz in sizes], dtype=np.int32)
offs = np.zeros(B, dtype=np.int64)
np.cumsum(n_words[:-1], out=offs[1:])
packed_d = torch.from_numpy(words.view(np.int32).copy()).to(dev)
offs_d, nw_d = torch.from_numpy(offs).to(dev), torch.from_numpy(n_words).to(dev)
outd = torch.empty((B, sym[0].numel()), dtype=torch.int32, device=dev)
status = torch.empty(B, dtype=torch.int32, device=dev)
for _ in range(3):
    e0, e1 = ev(), ev()
    e0.record()
    lib.licos_rans_decode_device(packed_d.data_ptr(), offs_d.data_ptr(), nw_d.data_ptr(), None, 0, B, outd.shape[1], 256,
                                 eb._quantized_cdf.data_ptr(), eb._quantized_cdf.shape[0], eb._quantized_cdf.shape[1],
                                 eb._cdf_length.data_ptr(), eb._offset.data_ptr(), outd.data_ptr(), status.data_ptr(),
                                 torch.cuda.current_stream().cuda_stream)
    e1.record()
    torch.cuda.synchronize()
    print(f"rans decode kernel only: {e0.elapsed_time(e1):.2f} ms (cdf stride {eb._quantized_cdf.shape[1]}, ok {int(status.min()) == 0})")
