"""Time the feature-intensity GEMM (BASELINE configs[4] pattern: F = 1024 feature centers, the 14 cuts of
scripts/3_make_rqae_features.py:178, nq = 1024) on synthetic codes resident in HBM; CUDA events.
usage: python tools/bench_intensity.py [--tokens N] [--features F] [--reps R]"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rqae_b200.feature import intensity_many, select_top_middle_bottom

CUTS = [2, 4, 6, 8, 12, 16, 24, 32, 48, 64, 128, 256, 512, 1023]


class Stub:
    quantization_method = "round_fsq"

    def __init__(self, dev):
        from rqae_b200.model import _fsq_grid
        self.codebook = torch.nn.Parameter(_fsq_grid(5, 4, True)[None].to(dev), requires_grad=False)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tokens", type=int, default=262144)
    ap.add_argument("--features", type=int, default=1024)
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(7)
    codes = torch.randint(0, 625, (a.tokens, 1024), generator=g, device=dev, dtype=torch.int16)
    centers = torch.randint(0, 625, (a.features, 1024), generator=g, device=dev, dtype=torch.int32)
    lw = (torch.rand(1024, generator=g, device=dev) * 2 + 13).half()
    m = Stub(dev)
    T_pad = (a.tokens + 255) // 256 * 256
    out = torch.empty(a.features, len(CUTS), T_pad, dtype=torch.float16, device=dev)
    for _ in range(2):
        intensity_many(m, codes, centers, CUTS, layer_weights=lw, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.reps):
        intensity_many(m, codes, centers, CUTS, layer_weights=lw, out=out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.reps
    flop = 2.0 * a.tokens * a.features * 4096
    byts = a.tokens * 2048 + a.features * len(CUTS) * a.tokens * 2
    res = intensity_many(m, codes, centers, CUTS, layer_weights=lw, out=out)
    select_top_middle_bottom(res, 100)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(a.reps):
        idx, val = select_top_middle_bottom(res, 100)
    e1.record()
    torch.cuda.synchronize()
    ms_sel = e0.elapsed_time(e1) / a.reps
    print(json.dumps({"select_ms": ms_sel, "select_gbs": a.features * len(CUTS) * a.tokens * 2 / ms_sel / 1e6,
                      "rows": a.features * len(CUTS)}))
    print(json.dumps({"tokens": a.tokens, "features": a.features, "ms": ms, "tokens_per_s": a.tokens / ms * 1e3,
                      "tflops": flop / ms / 1e9, "out_gbs": byts / ms / 1e6,
                      "note": "whole call: schedule + code transpose + feature operand + GEMM"}))


if __name__ == "__main__":
    main()
