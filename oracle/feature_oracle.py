"""Torch-CPU restatement of the feature-intensity path (oracle, kind = "port").

TEST INFRASTRUCTURE -- see ``oracle/__init__.py``.  Follows the reference
(harish-kamath/rqae) op for op:

    rqae/model.py:133-143      codebook_sims = (normalize(cb0) @ normalize(cb0).T).half()
    rqae/feature.py:95-100     layer_weights[l] = mean column norm of layers[l][1].weight, cast to fp16
    rqae/feature.py:102-129    intensity(): gather, *= weights, cumsum, /= cumsum(weights), select cuts
    scripts/3_make_rqae_features.py:116-128   argsort(descending) -> top-k / middle-k / bottom-k

``intensity_steps`` spells the fp16 arithmetic out (what torch's CPU kernels do, probed against the
reference: products rounded to fp16; the running sums are kept in fp32 and EVERY prefix is rounded to
fp16; the quotient is an fp32 divide rounded to fp16).  ``intensity`` issues the same ATen calls as the
reference.  Both are checked against outputs of the unmodified reference (tests/golden/kat_feature.npz).
``intensity_f64`` is the exact value of the same formula (fp16 table, fp16 weights, float64 arithmetic),
used to state how far any fp32-accumulating implementation may sit from the reference.
"""
from __future__ import annotations

from typing import Sequence

import torch
import torch.nn.functional as F


def codebook_sims(codebook0: torch.Tensor) -> torch.Tensor:
    """rqae/model.py:140-142.  (K, K) fp16."""
    n = F.normalize(codebook0.detach().clone().float(), dim=-1)
    return (n @ n.T).to(torch.float16)


def normalized_codebook(codebook0: torch.Tensor) -> torch.Tensor:
    """F.normalize(codebook[0]) as in rqae/model.py:141 -- the rank-4 factor of codebook_sims."""
    return F.normalize(codebook0.detach().clone().float(), dim=-1)


def layer_weights(w_out: torch.Tensor) -> torch.Tensor:
    """rqae/feature.py:97-99.  w_out: (nq, D, cd) stacked layers.{l}.1.weight -> (nq,) fp16."""
    return torch.tensor([w.norm(dim=0).mean().item() for w in w_out]).to(torch.float16)


def intensity(sims_h: torch.Tensor, center: torch.Tensor, codes: torch.Tensor, w_h: torch.Tensor,
              layers: Sequence[int]) -> torch.Tensor:
    """rqae/feature.py:115-129, same ATen calls.  codes (..., nq) -> (..., len(layers)) fp16."""
    max_layer = max(layers) + 1
    sims = sims_h[center[:max_layer].int(), codes[..., :max_layer].int()]
    sims *= w_h[:max_layer]
    sims = sims.cumsum(dim=-1)
    sims /= w_h[:max_layer].cumsum(dim=-1)
    return sims[..., list(layers)]


def intensity_steps(sims_h: torch.Tensor, center: torch.Tensor, codes: torch.Tensor, w_h: torch.Tensor,
                    layers: Sequence[int]) -> torch.Tensor:
    """The same computation with every rounding written out."""
    L = max(layers) + 1
    s = sims_h[center[:L].long(), codes[..., :L].long()]                   # fp16 gather
    p = (s.float() * w_h[:L].float()).to(torch.float16)                    # fp16 product, one rounding
    run = torch.zeros(p.shape[:-1], dtype=torch.float32)
    wrun = torch.zeros((), dtype=torch.float32)
    out = []
    cuts = set(int(l) for l in layers)
    res = {}
    for l in range(L):
        run = run + p[..., l].float()                                      # fp32 running sum
        wrun = wrun + w_h[l].float()
        if l in cuts:
            ph = run.to(torch.float16)                                     # prefix rounded to fp16
            wh = wrun.to(torch.float16)
            res[l] = (ph.float() / wh.float()).to(torch.float16)           # fp32 divide, rounded to fp16
    out = [res[int(l)] for l in layers]
    return torch.stack(out, dim=-1)


def intensity_f64(sims_h: torch.Tensor, center: torch.Tensor, codes: torch.Tensor, w_h: torch.Tensor,
                  layers: Sequence[int]) -> torch.Tensor:
    L = max(layers) + 1
    s = sims_h[center[:L].long(), codes[..., :L].long()].double()
    w = w_h[:L].double()
    num = (s * w).cumsum(-1)
    den = w.cumsum(-1)
    return (num / den)[..., list(layers)]


def select_top_middle_bottom(values: torch.Tensor, top_k: int = 100):
    """scripts/3_make_rqae_features.py:120-128 for one (feature, cut): indices into `values` (1-D) of the
    top_k largest, the top_k around the median rank and the top_k smallest, in the argsort order."""
    order = torch.argsort(values, descending=True)
    n = len(order)
    top = order[:top_k]
    bottom = order[-top_k:]
    middle = order[n // 2 - top_k // 2: n // 2 + top_k // 2]
    return top, middle, bottom


def get_activations(all_intensities: torch.Tensor, layers: Sequence[int], top_k: int, seq_len: int, stable: bool = False):
    """scripts/3_make_rqae_features.py:115-149 after the intensities are known.  all_intensities (T, len(layers)),
    tokens sequence-major.  Per layer: the sequence numbers of the selected tokens (top, then middle, then bottom),
    de-duplicated in order of first appearance (:133-139), and each such sequence's activations at all of its
    positions (:143-147).  ``stable=True`` orders equal values by index, as the CUDA selection does; the reference's
    ``torch.argsort`` (stable=False) leaves that order unspecified.  Returns {layer: (sequences list, (n, seq_len) array)}."""
    out = {}
    for j, l in enumerate(layers):
        col = all_intensities[:, j]
        if stable:
            order = torch.sort(col.float(), descending=True, stable=True).indices
            n = len(order)
            picked = torch.cat([order[:top_k], order[n // 2 - top_k // 2: n // 2 + top_k // 2], order[-top_k:]])
        else:
            picked = torch.cat(select_top_middle_bottom(col, top_k))
        seqs = []
        for k in picked.tolist():
            if k // seq_len not in seqs:
                seqs.append(k // seq_len)
        acts = torch.stack([col[s * seq_len:(s + 1) * seq_len] for s in seqs]).numpy()
        out[l] = (seqs, acts)
    return out
