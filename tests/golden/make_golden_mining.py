"""Golden vectors of the mining step from the UNMODIFIED reference (scripts/3_make_rqae_features.py:98-149).

Run in the build container only (needs /root/reference):  python tests/golden/make_golden_mining.py

``FeatureHelper.get_activations`` is called on an object created without ``setup()`` (which reads a Modal volume)
and given the three attributes ``setup()`` leaves behind (scripts/3:37-51): ``tokens`` (only its shape is used),
``texts`` (here: the sequence number, so the order of the returned sequences can be read back) and ``indices``
(the code store).  The feature is ``RQAEFeature.from_quantizer(ref_model, center=..., layers=...)`` of the reference.

Writes tests/golden/kat_mining.npz:  RQAE(dim=64, num_quantizers=64) (the model of kat_feature.npz "small"),
a store of 2100 sequences x 3 positions (so that the 1024-sequence batching of scripts/3:104-108 takes three
rounds), two features, top_k = 7 and top_k = 100; per feature, top_k and layer the sequence numbers in the order
the reference returns them and their per-position activations."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from _ref_loader import load_reference_file  # noqa: E402


def main():
    s3 = load_reference_file("scripts/3_make_rqae_features.py", "ref_scripts3")
    from rqae.model import RQAE as RefRQAE
    from rqae.feature import RQAEFeature as RefFeature

    torch.manual_seed(0)
    m = RefRQAE(dim=64, num_quantizers=64).eval()
    g = torch.Generator().manual_seed(21)
    N, S, nq = 2100, 3, 64
    codes = torch.randint(0, 625, (N, S, nq), generator=g)
    centers = [codes[17, 1].clone(), torch.randint(0, 625, (nq,), generator=g)]
    layers = [2, 4, 6, 8, 12, 16, 24, 32, 48, 63]
    helper = s3.FeatureHelper.__new__(s3.FeatureHelper)
    helper.tokens = torch.zeros(N, S, dtype=torch.int64)
    helper.texts = list(range(N))
    helper.indices = codes
    out = {"cb0": m.codebook.data[0].numpy(), "codes": codes.numpy().astype(np.int16), "layers": np.array(layers, np.int32),
           "centers": torch.stack(centers).numpy().astype(np.int32)}
    for f, c in enumerate(centers):
        feat = RefFeature.from_quantizer(m, center=c.numpy(), layers=list(layers))
        out["lw"] = feat.layer_weights.numpy()
        for top_k in (7, 100):
            with torch.inference_mode():
                acts = helper.get_activations(feat, top_k=top_k)
            for l in layers:
                out[f"f{f}/k{top_k}/{l}/sequences"] = np.array([a["text"] for a in acts[l]], np.int32)
                out[f"f{f}/k{top_k}/{l}/activations"] = np.stack([a["activations"] for a in acts[l]])
    # get_unique_token_indices (scripts/3:53-82) under a fixed seed: a small vocabulary with repeats
    vocab_tokens = torch.randint(0, 50, (40, 9), generator=g)
    helper.tokens = vocab_tokens
    torch.manual_seed(1234)
    out["unique/tokens"] = vocab_tokens.numpy()
    out["unique/seed"] = np.array([1234])
    out["unique/indices"] = helper.get_unique_token_indices().numpy()
    path = os.path.join(HERE, "kat_mining.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", len(out), "arrays")


if __name__ == "__main__":
    main()
