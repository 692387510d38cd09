"""Small driver for ncu captures: one forward (and optionally decode) launch of the 2B-shape model."""
import argparse
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rqae_b200 import RQAE

ap = argparse.ArgumentParser()
ap.add_argument("--tokens", type=int, default=148 * 16 * 2)
ap.add_argument("--dim", type=int, default=2304)
ap.add_argument("--nq", type=int, default=1024)
ap.add_argument("--decode", action="store_true")
ap.add_argument("--reps", type=int, default=1)
a = ap.parse_args()
torch.manual_seed(0)
m = RQAE(dim=a.dim, num_quantizers=a.nq).eval().cuda()
m.freeze_packed()   # no per-call walk over the 2 * nq parameter versions: the timing below is the kernel's
x = torch.randn(1, a.tokens, a.dim, device="cuda")
for _ in range(a.reps):
    q, idx = m(x)
    if a.decode:
        d = m.decode(idx)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n_t = 5 if a.tokens <= 65536 else 1
e0.record()
for _ in range(n_t):
    q, idx = m(x)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n_t
print("forward ms", ms, "tokens", a.tokens, "tok/s", a.tokens / ms * 1e3)
if a.decode:
    e0.record(); d = m.decode(idx); e1.record(); torch.cuda.synchronize()
    print("decode ms", e0.elapsed_time(e1), "tok/s", a.tokens / e0.elapsed_time(e1) * 1e3)
    for prec in ("f16", "f16x3"):
        d2 = m.decode(idx, precision=prec); torch.cuda.synchronize()
        e0.record(); d2 = m.decode(idx, precision=prec); e1.record(); torch.cuda.synchronize()
        err = ((d2 - d).abs().max() / d.abs().max()).item()
        print("decode", prec, "ms", e0.elapsed_time(e1), "tok/s", a.tokens / e0.elapsed_time(e1) * 1e3, "rel err", err)
