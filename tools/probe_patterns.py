"""FMA-pipe rate of the forward kernel's operand patterns (modes 2/3 of rqae_fp32_peak_probe) vs the plain FFMA2 peak."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rqae_b200 import _lib
lib = _lib.load()
sink = torch.rand(256, device="cuda") + 0.5
st = torch.cuda.current_stream().cuda_stream
for mode, name in [(1, "FFMA2 dense (pair,pair,pair; reuse)"), (0, "FFMA scalar"), (2, "sweep-2 pattern: acc += w*r"), (3, "sweep-1 pattern: o-chain, r -= o")]:
    fl = ctypes.c_double(0)
    lib.rqae_fp32_peak_probe(mode, 2000, ctypes.byref(fl), sink.data_ptr(), st); torch.cuda.synchronize()
    best = 0
    for _ in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); lib.rqae_fp32_peak_probe(mode, 200000 if mode >= 2 else 20000, ctypes.byref(fl), sink.data_ptr(), st); b.record(); torch.cuda.synchronize()
        best = max(best, fl.value / (a.elapsed_time(b) * 1e-3) / 1e12)
    print(f"mode {mode} {name}: {best:.2f} TFLOP/s")
