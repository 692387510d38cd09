"""Golden vectors of the nearest-example search from the UNMODIFIED reference (demo/server/server.py).

Run in the build container only (needs /root/reference):  python tests/golden/make_golden_search.py

server.py imports ``modal`` at module level and that package is not in this image, so a stub module whose
decorators return the decorated object is put into ``sys.modules`` first (tests/golden/_ref_loader.py); nothing of
the reference is edited.
``IntensityEngine.find_examples`` (server.py:159-325) calls ``.cuda()`` on its tensors; there is no GPU in the
build container, so ``torch.Tensor.cuda`` is an identity for the duration of the script and the reference's
torch code runs on the CPU.  The engine object is created without ``setup()`` (which reads a Modal volume)
and given exactly the two attributes ``setup()`` would leave behind (server.py:104-115,139):
``sims = subfeature_sims * layer_norms`` of the reference ``RQAE`` and ``activations`` = int32 code shards.

Writes tests/golden/kat_search.npz with two cases:

  k81/   RQAE(dim=48, codebook_size=3 -> K=81, num_quantizers=160); 3 shards of 8 sequences x 7 positions;
         layers [4, 6, 8, 12, 16, 24, 32, 48, 64, 150] (the last range is 86 layers: the chunks-of-64 branch,
         server.py:216-234); query = dataset sequence 5 (idx=) and an external activation (activation=)
  k625/  RQAE(dim=32, num_quantizers=4) with the 625-row round_fsq table; 2 shards of 8 x 5; layers [2, 3]

Each case holds the scaled fp16 table, the code shards, the query and, per yielded layer, the reference's
top / middle / bottom ``indices`` and ``intensities``.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
sys.path.insert(0, REF)


sys.path.insert(0, HERE)
from _ref_loader import load_reference_file  # noqa: E402


def load_reference_server():
    return load_reference_file("demo/server/server.py", "ref_server")


def run(engine, out, prefix, tag, **kw):
    layers = kw["layers"]
    for res, layer in engine.find_examples(**kw):
        for part in ("top", "middle", "bottom"):
            out[f"{prefix}/{tag}/{layer}/{part}/indices"] = res[part]["indices"].numpy()
            out[f"{prefix}/{tag}/{layer}/{part}/intensities"] = res[part]["intensities"].numpy()
    out[f"{prefix}/{tag}/layers"] = np.array(layers, np.int32)


def make_engine(srv, model, shards):
    sims = model.subfeature_sims
    sims *= model.layer_norms.unsqueeze(-1).unsqueeze(-1)        # server.py:111
    eng = srv.IntensityEngine.__new__(srv.IntensityEngine)
    eng.sims = sims.clone().detach()
    eng.activations = shards
    return eng


def main():
    srv = load_reference_server()
    torch.Tensor.cuda = lambda self, *a, **k: self
    from rqae.model import RQAE as RefRQAE

    out = {}
    # ---- K = 81, deep enough for the chunked branch
    torch.manual_seed(0)
    m = RefRQAE(dim=48, codebook_size=3, num_quantizers=160).eval()
    K = m.codebook.shape[1]
    g = torch.Generator().manual_seed(11)
    shards = [torch.randint(0, K, (8, 7, 160), generator=g, dtype=torch.int32) for _ in range(3)]
    shards[1][2, 3] = shards[0][5, 1]              # an exact repeat of a query position elsewhere in the dataset
    eng = make_engine(srv, m, shards)
    layers = [4, 6, 8, 12, 16, 24, 32, 48, 64, 150]
    kw = dict(top_examples=6, middle_examples=4, bottom_examples=4, layers=layers)
    run(eng, out, "k81", "idx5", idx=5, **kw)
    ext = torch.randint(0, K, (7, 160), generator=g, dtype=torch.int32)
    run(eng, out, "k81", "ext", activation=ext, **kw)
    out["k81/sims"] = eng.sims.numpy()
    out["k81/shards"] = torch.stack(shards).numpy().astype(np.int16)
    out["k81/ext"] = ext.numpy().astype(np.int16)
    out["k81/args"] = np.array([6, 4, 4], np.int32)
    # ---- K = 625
    torch.manual_seed(1)
    m = RefRQAE(dim=32, num_quantizers=4).eval()
    g = torch.Generator().manual_seed(12)
    shards = [torch.randint(0, 625, (8, 5, 4), generator=g, dtype=torch.int32) for _ in range(2)]
    shards[0][1, 0, :2] = 312                       # the zero codeword
    eng = make_engine(srv, m, shards)
    kw = dict(top_examples=5, middle_examples=2, bottom_examples=3, layers=[2, 3])
    run(eng, out, "k625", "idx3", idx=3, **kw)
    out["k625/sims"] = eng.sims.numpy()
    out["k625/shards"] = torch.stack(shards).numpy().astype(np.int16)
    out["k625/args"] = np.array([5, 2, 3], np.int32)

    path = os.path.join(HERE, "kat_search.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", len(out), "arrays")


if __name__ == "__main__":
    main()
