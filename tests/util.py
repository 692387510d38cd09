"""Helpers shared by the tests: golden-case access and weight plumbing."""
from __future__ import annotations

import numpy as np
import torch

from oracle import rqae_oracle as orc
from oracle import c_oracle

SMALL_CASES = ["round_fsq_d256", "round_fsq_d256_zero", "round_fsq_d512_ml16", "fsq_d384", "vq_d256",
               "round_fsq_d200_ragged", "round_fsq_d768_trained", "round_fsq_cbs3"]


def small_case(g, name):
    p = name + "/"
    keys = ["w_in", "b_in", "w_out", "b_out", "codebook", "x", "q", "codes", "codes_full", "dec", "margins_fp64",
            "dec_layers", "codebook_post"]
    d = {k: g[p + k] for k in keys}
    d["method"] = str(g[p + "method"])
    d["cbs"] = int(g[p + "cbs"])
    ml = int(g[p + "max_layers"])
    d["max_layers"] = None if ml < 0 else ml
    d["dec_layers"] = list(d["dec_layers"]) if int(g[p + "has_dec_layers"]) else None
    return d


def stacked(d) -> orc.StackedWeights:
    t = torch.from_numpy
    return orc.StackedWeights(t(d["w_in"]), t(d["b_in"]), t(d["w_out"]), t(d["b_out"]), t(d["codebook"]), d["method"])


def cweights(d) -> c_oracle.CWeights:
    return c_oracle.CWeights(d["w_in"], d["b_in"], d["w_out"], d["b_out"], d["codebook"],
                             d["method"] in ("fsq", "round_fsq"))


def module_from_case(d, device="cpu"):
    """rqae_b200.RQAE carrying the golden case's weights (state-dict layout of the reference)."""
    from rqae_b200 import RQAE
    nq, cd, D = d["w_in"].shape
    m = RQAE(dim=D, codebook_dim=cd, codebook_size=d["cbs"], num_quantizers=nq, quantization_method=d["method"])
    sd = {}
    for l in range(nq):
        sd[f"layers.{l}.0.weight"] = torch.from_numpy(d["w_in"][l])
        sd[f"layers.{l}.0.bias"] = torch.from_numpy(d["b_in"][l])
        sd[f"layers.{l}.1.weight"] = torch.from_numpy(d["w_out"][l])
        sd[f"layers.{l}.1.bias"] = torch.from_numpy(d["b_out"][l])
    sd["codebook"] = torch.from_numpy(d["codebook"])
    sd["codebook_counts"] = torch.zeros_like(m.codebook_counts)
    m.load_state_dict(sd, strict=True)
    return m.eval().to(device)


def stacked_from_module(m) -> orc.StackedWeights:
    return orc.StackedWeights.from_state_dict({k: v.detach().cpu() for k, v in m.state_dict().items()},
                                              m.quantization_method)


def x_2b(rows=None):
    """The 2B KAT input: randn(32,128,2304) from CPU generator seed 1 (sha-checked by the tests)."""
    x = torch.randn(32, 128, 2304, generator=torch.Generator().manual_seed(1))
    return x if rows is None else x.view(-1, 2304)[:rows]
