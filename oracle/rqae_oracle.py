"""Torch-CPU restatement of the RQAE hot path (oracle, kind = "port").

TEST INFRASTRUCTURE -- see ``oracle/__init__.py``.  Follows the reference
``rqae/model.py`` (harish-kamath/rqae) op for op, with the per-layer
``nn.Linear`` pairs stacked into four tensors so that no reference class is
needed at run time:

    w_in  (nq, cd, D)   = layers.{l}.0.weight      rqae/model.py:32
    b_in  (nq, cd)      = layers.{l}.0.bias
    w_out (nq, D, cd)   = layers.{l}.1.weight      rqae/model.py:33
    b_out (nq, D)       = layers.{l}.1.bias
    codebook (nq, K, cd)                           rqae/model.py:37-45

It issues the same ATen calls as the reference (``F.linear``, ``norm``,
``matmul``, ``argmax``, index, add/sub) in the same order, so in fp32 it is
bit-identical to the reference on the same machine (checked against the golden
files) and costs what the reference costs -- which is why ``bench.py`` uses it as
the CPU baseline on boxes where /root/reference does not exist.  In fp64 it also
returns, for every (token, layer), the cos-sim margin between the winning code
and the best *different-valued* competitor, which the parity protocol
 uses to classify disagreements as near-ties.
"""
from __future__ import annotations

from dataclasses import dataclass
from itertools import product
from typing import Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as F


@dataclass
class StackedWeights:
    w_in: torch.Tensor      # (nq, cd, D)
    b_in: torch.Tensor      # (nq, cd)
    w_out: torch.Tensor     # (nq, D, cd)
    b_out: torch.Tensor     # (nq, D)
    codebook: torch.Tensor  # (nq, K, cd)
    quantization_method: str = "round_fsq"

    @property
    def nq(self) -> int:
        return self.w_in.shape[0]

    @property
    def dim(self) -> int:
        return self.w_in.shape[2]

    @property
    def cd(self) -> int:
        return self.w_in.shape[1]

    def to(self, dtype) -> "StackedWeights":
        return StackedWeights(self.w_in.to(dtype), self.b_in.to(dtype), self.w_out.to(dtype),
                              self.b_out.to(dtype), self.codebook.to(dtype), self.quantization_method)

    @classmethod
    def from_state_dict(cls, sd, quantization_method: str = "round_fsq") -> "StackedWeights":
        nq = sd["codebook"].shape[0]
        g = lambda k: sd[k].detach().cpu()
        return cls(
            torch.stack([g(f"layers.{l}.0.weight") for l in range(nq)]),
            torch.stack([g(f"layers.{l}.0.bias") for l in range(nq)]),
            torch.stack([g(f"layers.{l}.1.weight") for l in range(nq)]),
            torch.stack([g(f"layers.{l}.1.bias") for l in range(nq)]),
            g("codebook"),
            quantization_method,
        )


def fsq_codebook(codebook_size: int, codebook_dim: int, normalise: bool) -> torch.Tensor:
    """The (K, cd) grid of rqae/model.py:63-72: linspace(-1,1,cbs)^cd in
    itertools.product order (first coordinate slowest); for round_fsq each row
    is divided by its float64 norm (the all-zero row is left zero), then the
    whole table is cast to fp32 by the ``copy_`` into the parameter."""
    axis = np.linspace(-1, 1, codebook_size)
    grid = np.array(list(product(axis, repeat=codebook_dim)))
    if normalise:
        n = np.linalg.norm(grid, axis=-1, keepdims=True)
        n = np.where(n == 0, 1.0, n)
        grid = grid / n
    return torch.from_numpy(grid).to(torch.float32)


def random_init(dim=2304, codebook_dim=4, codebook_size=5, num_quantizers=1024,
                quantization_method="round_fsq", seed=0) -> StackedWeights:
    """Same RNG consumption order as ``torch.manual_seed(seed); RQAE(...)`` in the
    reference constructor (rqae/model.py:29-55): per layer Linear(D,cd) then
    Linear(cd,D) (each: kaiming-uniform weight, then uniform bias), then one
    ``randn`` for the codebook parameter."""
    torch.manual_seed(seed)
    w_in, b_in, w_out, b_out = [], [], [], []
    for _ in range(num_quantizers):
        a = torch.nn.Linear(dim, codebook_dim)
        b = torch.nn.Linear(codebook_dim, dim)
        w_in.append(a.weight.detach()); b_in.append(a.bias.detach())
        w_out.append(b.weight.detach()); b_out.append(b.bias.detach())
    if quantization_method in ("fsq", "round_fsq"):
        torch.randn(num_quantizers, codebook_size ** codebook_dim, codebook_dim)  # consumed, then overwritten
        cb = fsq_codebook(codebook_size, codebook_dim, quantization_method == "round_fsq")
        cb = cb.unsqueeze(0).repeat(num_quantizers, 1, 1)
    else:
        cb = torch.randn(num_quantizers, codebook_size, codebook_dim)
        cb = cb / cb.norm(dim=-1, keepdim=True)  # normalize_codebooks(), rqae/model.py:126-131
    return StackedWeights(torch.stack(w_in), torch.stack(b_in), torch.stack(w_out), torch.stack(b_out),
                          cb.contiguous(), quantization_method)


def _distinct_margin(cos: torch.Tensor, cb_l: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    """cos[..., idx] minus the best cos among rows whose codeword differs from
    codebook[idx] (duplicate rows are not competitors: they decode identically)."""
    win = cb_l[idx]                                              # (..., cd)
    same = (cb_l.unsqueeze(0) == win.reshape(-1, 1, cb_l.shape[-1])).all(-1)  # (T, K)
    c = cos.reshape(-1, cos.shape[-1]).masked_fill(same, float("-inf"))
    top = cos.reshape(-1, cos.shape[-1]).gather(1, idx.reshape(-1, 1)).squeeze(1)
    return (top - c.max(dim=1).values).reshape(idx.shape)


@torch.inference_mode()
def forward(w: StackedWeights, x: torch.Tensor, max_layers: float = float("inf"),
            dtype: torch.dtype = torch.float32, want_margins: bool = False,
            teacher_codes: Optional[torch.Tensor] = None):
    """rqae/model.py:199-230 (eval mode / temperature 0: plain argmax).

    Returns (quantized_out (B,S,D), indices (B,S,nq') int64[, margins (B,S,nq')]).
    ``teacher_codes`` forces the residual recurrence to follow the given codes
    while still reporting this implementation's own argmax per layer."""
    w = w.to(dtype)
    residual = x.to(dtype)
    quantized_out = 0
    all_idx, all_m = [], []
    learned = w.quantization_method not in ("fsq", "round_fsq")
    cb_all = w.codebook
    for l in range(w.nq):
        if l >= max_layers:
            break
        z = F.linear(residual, w.w_in[l], w.b_in[l])                 # model.py:211
        if learned:
            # model.py:126-131,196: quantize() re-normalises the WHOLE codebook parameter in
            # place before every layer, so layer l sees a table normalised l+1 times (in the
            # parameter's own dtype, fp32) since the call started.
            cb32 = cb_all.to(torch.float32)
            cb_all = (cb32 / cb32.norm(dim=-1, keepdim=True)).to(dtype)
        cb_l = cb_all[l]
        zn = z / z.norm(dim=-1, keepdim=True)                          # model.py:188
        cos = torch.matmul(zn, cb_l.T)                                 # model.py:190
        idx = cos.argmax(dim=-1)                                       # model.py:182
        all_idx.append(idx)
        if want_margins:
            all_m.append(_distinct_margin(cos, cb_l, idx))
        use = idx if teacher_codes is None else teacher_codes[..., l].to(torch.int64)
        c = cb_l[use]                                                  # model.py:192
        c = z + (c - z)                                                # model.py:218-220 (STE)
        o = F.linear(c, w.w_out[l], w.b_out[l])                        # model.py:221
        residual = residual - o                                        # model.py:223
        quantized_out = quantized_out + o                              # model.py:224
    idx = torch.stack(all_idx, dim=-1)                                 # model.py:226 (contiguous here)
    if want_margins:
        return quantized_out, idx, torch.stack(all_m, dim=-1)
    return quantized_out, idx


@torch.inference_mode()
def decode(w: StackedWeights, indices: torch.Tensor, layers: Optional[Sequence[int]] = None,
           dtype: torch.dtype = torch.float32) -> torch.Tensor:
    """rqae/model.py:232-252: codewords always come from codebook[0]; the
    out-projections are summed in ascending layer order, the first selected
    layer initialising the accumulator."""
    w = w.to(dtype)
    cv = w.codebook[0][indices.to(torch.int64)]                        # model.py:234
    return decode_from_codebook_values(w, cv, layers, dtype)


@torch.inference_mode()
def decode_from_codebook_values(w: StackedWeights, cv: torch.Tensor,
                                layers: Optional[Sequence[int]] = None,
                                dtype: torch.dtype = torch.float32) -> torch.Tensor:
    w = w.to(dtype)
    cv = cv.to(dtype)
    q = None
    for l in range(w.nq):                                              # model.py:239-247
        if layers is not None and l not in layers:
            continue
        o = F.linear(cv[:, :, l], w.w_out[l], w.b_out[l])
        q = o if q is None else q + o
    return q


def codebook_sims(w: StackedWeights) -> torch.Tensor:
    """rqae/model.py:133-143: fp16 cos-sim table of the layer-0 codebook."""
    cb = w.codebook[0].float()
    n = F.normalize(cb, dim=-1)
    return (n @ n.T).to(torch.float16)


def layer_weights_fp16(w: StackedWeights) -> torch.Tensor:
    """rqae/feature.py:95-100: mean column norm of each W_out, cast to fp16."""
    return torch.tensor([w.w_out[l].norm(dim=0).mean().item() for l in range(w.nq)]).to(torch.float16)


@torch.inference_mode()
def intensity(w: StackedWeights, center: torch.Tensor, token_indices: torch.Tensor,
              layers: Sequence[int]) -> torch.Tensor:
    """rqae/feature.py:102-129, fp16 semantics included (gather from the fp16
    table, in-place fp16 multiply, fp16 cumsum, fp16 divide)."""
    max_layer = max(layers) + 1
    sims_t = codebook_sims(w)
    lw = layer_weights_fp16(w)
    sims = sims_t[center[:max_layer].int(), token_indices[..., :max_layer].int()]
    sims *= lw[:max_layer]
    sims = sims.cumsum(dim=-1)
    sims /= lw[:max_layer].cumsum(dim=-1)
    return sims[..., list(layers)]
