"""PCIe probe: contiguous vs row-pitched (2-D) pinned copies, both directions, and both at once."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import wrf_model_cuda_sample_b200 as wrf

def t(fn, n=3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n

h = torch.empty(1 << 28, dtype=torch.float32, pin_memory=True)      # 1 GiB
d = torch.empty_like(h, device="cuda")
gb = h.numel() * 4 / 1e9
print("contiguous H2D %.1f GB/s" % (gb / t(lambda: d.copy_(h, non_blocking=True))))
print("contiguous D2H %.1f GB/s" % (gb / t(lambda: h.copy_(d, non_blocking=True))))
h2 = torch.empty_like(h).pin_memory(); d2 = torch.empty_like(d)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def both():
    with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
print("duplex: %.1f GB/s each way" % (gb / t(both)))
g = wrf.Grid.from_shape(1800, 1060, 50, halo=5)
f = wrf.synth_fields(g, names=("u", "t"), pinned=True)
with wrf.Patch(g) as p:
    up = lambda: (p.upload(f, names=("u",)), p.sync())
    print("pitched 2-D H2D (row 7240 B -> pitch 7296 B) %.1f GB/s" % (f["u"].nbytes / 1e9 / t(up)))
    dn = lambda: p.download(f, names=("t",))
    print("pitched 2-D D2H %.1f GB/s" % (f["t"].nbytes / 1e9 / t(dn)))
