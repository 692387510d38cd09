"""Per-role clock breakdown of the intensity GEMM (RQAE_INT_DBG bit 1024): where the issuer, a builder thread, an
epilogue warp and the feature-tile producer of every CTA spend their cycles.  usage: python tools/int_profile.py [dbg]"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
dbg = int(sys.argv[1]) if len(sys.argv) > 1 else 0
os.environ["RQAE_INT_DBG"] = str(dbg | 1024)
import numpy as np
import torch
from rqae_b200 import _lib
from rqae_b200.feature import intensity_many
from tools.bench_intensity import Stub, CUTS

dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(7)
T, F = 131072, 1024
codes = torch.randint(0, 625, (T, 1024), generator=g, device=dev, dtype=torch.int16)
centers = torch.randint(0, 625, (F, 1024), generator=g, device=dev, dtype=torch.int32)
lw = (torch.rand(1024, generator=g, device=dev) * 2 + 13).half()
m = Stub(dev)
out = torch.empty(F, len(CUTS), T, dtype=torch.float16, device=dev)
for _ in range(3):
    intensity_many(m, codes, centers, CUTS, layer_weights=lw, out=out)
torch.cuda.synchronize()
buf = np.zeros((148, 16), dtype=np.uint64)
_lib.check(_lib.load().rqae_intensity_profile(buf.ctypes.data, 148), "profile")
a = buf.astype(np.float64).mean(0)
kb = 70 * 2048 / 148
names = ["issuer wait V", "issuer wait U", "issuer wait acc", "issuer total", "builder wait free", "builder build", "builder total",
         "epilogue wait cut", "epilogue hold acc", "epilogue wait store-read", "epilogue total", "producer wait free", "producer total",
         "epi: cut -> stores 1 issued", "epi: kept-half roundings", "epi: wait + stage 2 + stores 2"]
print(f"dbg={dbg}: mean clocks per CTA, and per K-block ({kb:.0f} K-blocks per CTA)")
for n, v in zip(names, a):
    print(f"  {n:26s} {v:12.0f}  {v / kb:8.1f}")
