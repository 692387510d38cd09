"""Time RQAE.forward_host (the e2e path) for the library named by RQAE_B200_LIB.  usage: python tools/e2e_time.py [tokens]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rqae_b200 import RQAE
T = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 19
torch.manual_seed(0)
m = RQAE().eval().cuda(); m.freeze_packed()
xh = torch.randn(T, 2304).pin_memory()
qh = torch.empty(T, 2304, pin_memory=True); ch = torch.empty(T, 1024, dtype=torch.int64, pin_memory=True)
m.forward_host(xh[: 1 << 15], out=(qh[: 1 << 15], ch[: 1 << 15]))
best = 0.0
for _ in range(3):
    t0 = time.perf_counter(); m.forward_host(xh, out=(qh, ch)); dt = time.perf_counter() - t0
    best = max(best, T / dt)
print("e2e tok/s best of 3:", round(best), "checksum", int(ch[:4096].sum()))
