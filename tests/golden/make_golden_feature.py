"""Golden vectors of the feature-intensity path from the UNMODIFIED reference (rqae/feature.py).

Run in the build container only (needs /root/reference):  python tests/golden/make_golden_feature.py

Writes tests/golden/kat_feature.npz with two cases, each holding everything the implementation needs
(layer-0 codebook, fp16 layer weights as the reference computes them, feature centers, token codes, the
layer cuts) and what ``RQAEFeature.from_quantizer(ref_model, center=..., layers=...).intensity(codes)``
returned for every feature:

  small/  RQAE(dim=64, num_quantizers=64), random codes (5,7,64), 3 centers, cuts incl. an unsorted list
  2b/     RQAE() at the 2B shape (torch.manual_seed(0)); codes = the reference's own codes of the 2B KAT
          (tests/golden/kat_2b.npz, first 512 tokens); 8 centers = codes of tokens 600..607; the cuts of
          scripts/3_make_rqae_features.py:178
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from rqae.model import RQAE as RefRQAE          # noqa: E402
from rqae.feature import RQAEFeature as RefFeature  # noqa: E402

SCRIPT3_LAYERS = [2, 4, 6, 8, 12, 16, 24, 32, 48, 64, 128, 256, 512, 1023]


def run_case(model, codes, centers, layers):
    outs = []
    lw = None
    for c in centers:
        f = RefFeature.from_quantizer(model, center=c.numpy(), layers=list(layers))
        lw = f.layer_weights
        with torch.inference_mode():
            outs.append(f.intensity(codes).clone())
    return torch.stack(outs), lw   # (F, ..., C) fp16


def main():
    out = {}
    # ---- small
    torch.manual_seed(0)
    m = RefRQAE(dim=64, num_quantizers=64).eval()
    g = torch.Generator().manual_seed(3)
    codes = torch.randint(0, 625, (5, 7, 64), generator=g)
    codes[0, 0, :8] = 312           # the all-zero codeword: sims row/column of zeros
    centers = torch.randint(0, 625, (3, 64), generator=g)
    centers[1, 3] = 312
    layers = [2, 4, 6, 8, 12, 16, 24, 32, 48, 63]
    o, lw = run_case(m, codes, centers, layers)
    layers_u = [63, 0, 17, 5]
    o_u, _ = run_case(m, codes, centers, layers_u)
    out.update({"small/cb0": m.codebook.data[0].numpy(), "small/lw": lw.numpy(), "small/codes": codes.numpy().astype(np.int16),
                "small/centers": centers.numpy().astype(np.int32), "small/layers": np.array(layers, np.int32),
                "small/out": o.numpy(), "small/layers_unsorted": np.array(layers_u, np.int32), "small/out_unsorted": o_u.numpy()})
    # ---- 2B
    torch.manual_seed(0)
    m = RefRQAE().eval()
    kat = np.load(os.path.join(HERE, "kat_2b.npz"))
    c_all = torch.from_numpy(kat["codes1024"].astype(np.int64))      # (1024 tokens, 1024 layers)
    codes = c_all[:512]
    centers = c_all[600:608]
    o, lw = run_case(m, codes, centers, SCRIPT3_LAYERS)
    out.update({"2b/cb0": m.codebook.data[0].numpy(), "2b/lw": lw.numpy(), "2b/codes": codes.numpy().astype(np.int16),
                "2b/centers": centers.numpy().astype(np.int32), "2b/layers": np.array(SCRIPT3_LAYERS, np.int32),
                "2b/out": o.numpy()})
    path = os.path.join(HERE, "kat_feature.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes;", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
