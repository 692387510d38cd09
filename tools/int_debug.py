"""Progressive diagnostics of the intensity GEMM on the GPU box (prints error statistics per configuration so
that one visit tells which part of the tcgen05 pipeline is off).   usage: python tools/int_debug.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from oracle import feature_oracle as fo
from rqae_b200 import RQAE
from rqae_b200.feature import intensity_many

dev = torch.device("cuda:0")
torch.manual_seed(0)
nq = int(os.environ.get("NQ", 160))
m = RQAE(dim=64, num_quantizers=nq).eval()
lw = fo.layer_weights(torch.stack([l[1].weight.data for l in m.layers]))
sims = fo.codebook_sims(m.codebook.data[0])
md = m.to(dev)
g = torch.Generator().manual_seed(11)


def run(T, F, layers, tag):
    codes = torch.randint(0, 625, (T, nq), generator=g)
    centers = torch.randint(0, 625, (F, nq), generator=g)
    out = intensity_many(md, codes.to(dev), centers, layers, layer_weights=lw)
    torch.cuda.synchronize()
    out = out.cpu()
    worst = 0.0
    rows = []
    for f in sorted(set([0, min(1, F - 1), F // 2, F - 1])):
        exact = fo.intensity_f64(sims, centers[f], codes, lw, layers).T
        err = (out[f].double() - exact).abs()
        worst = max(worst, float(err.max()))
        rows.append((f, [round(float(e), 5) for e in err.max(dim=1).values]))
    print(f"[{tag}] T={T} F={F} cuts={layers} worst={worst:.5f} {'OK' if worst < 2e-3 else 'BAD'}")
    if worst >= 2e-3:
        for f, r in rows:
            print("   f", f, "max err per cut", r)
        f = 0
        exact = fo.intensity_f64(sims, centers[f], codes, lw, layers).T
        print("   out[0,0,:8]  ", [round(float(v), 4) for v in out[0, 0, :8]])
        print("   exact[0,:8]  ", [round(float(v), 4) for v in exact[0, :8]])
        e = (out[0, 0].double() - exact[0]).abs()
        bad = (e > 1.5e-3).nonzero().flatten()
        print("   bad tokens (cut 0, f 0):", bad[:16].tolist(), "count", len(bad))
    sys.stdout.flush()


run(8, 1, [0], "one layer")
run(8, 1, [3], "4 layers = 1 MMA k-step")
run(8, 1, [15], "one full K-block")
run(8, 1, [31], "two K-blocks")
run(256, 128, [63], "full token tile, full feature tile")
run(300, 130, [63], "2 token tiles, 2 feature tiles")
run(300, 130, [2, 4, 6, 8, 12, 16, 24, 32, 48, 64, 128, nq - 1], "script-3 cuts")
run(70000, 300, [5, nq - 1], "more units than SMs")
