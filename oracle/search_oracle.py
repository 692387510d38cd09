"""Torch-CPU restatement of the nearest-example search of the demo server (oracle, kind = "port").

TEST INFRASTRUCTURE -- see ``oracle/__init__.py``.  Follows the reference (harish-kamath/rqae)
``demo/server/server.py`` op for op:

    server.py:41-68     get_intensities(): gather sims[s2, l, code[b, s, l]]
    server.py:101-115   sims = subfeature_sims * layer_norms[:, None, None]      (fp16, in place)
    server.py:176-196   query table: query_sims[l, s2, :] = sims[l, q_code[s2][l], :]
    server.py:198-263   per layer range [a, b): sum of the gathered values over the range (fp16 result;
                        ranges longer than 64 layers in chunks of 64 whose fp16 sums are added in fp16),
                        accumulated over ranges in fp16
    server.py:265-324   max over the dataset sequence's positions, argsort(descending) over sequences,
                        top / middle / bottom slices, intensities of the selected sequences

``find_examples`` issues the same ATen calls as the reference and is checked against outputs of the
UNMODIFIED reference (tests/golden/kat_search.npz, made by tests/golden/make_golden_search.py, which
imports server.py with a stub ``modal`` module).  ``accumulate_steps`` writes every rounding out the way
the CUDA kernel does it (fp32 running sum in ascending layer order inside a chunk, one rounding to fp16
per chunk, fp16 adds between chunks and between ranges); it may differ from torch's own fp16 ``sum`` only
through the order of the fp32 additions, i.e. by one fp16 rounding step, rarely.
"""
from __future__ import annotations

from typing import Iterator, List, Sequence, Tuple

import torch

SERVER_LAYERS = [4, 6, 8, 12, 16, 24, 32, 48, 64, 128, 256, 512, 1023]   # server.py:167
CHUNK = 64                                                               # server.py:219-222


def scaled_sims(subfeature_sims_h: torch.Tensor, layer_norms: torch.Tensor) -> torch.Tensor:
    """server.py:104-115: the (nq, K, K) fp16 table times the per-layer norm, IN fp16 (in-place multiply)."""
    sims = subfeature_sims_h.clone()
    sims *= layer_norms.unsqueeze(-1).unsqueeze(-1)
    return sims


def layer_ranges(layers: Sequence[int]) -> List[Tuple[int, int]]:
    """server.py:199-202."""
    lr = [0] + list(layers)
    return [(lr[i], lr[i + 1]) for i in range(len(lr) - 1)]


def query_table(sims_h: torch.Tensor, query: torch.Tensor, n_layers: int) -> torch.Tensor:
    """server.py:176-196.  query (Sq, nq) int -> (n_layers, Sq, K) fp16."""
    qa = query[:, :n_layers].T
    qs = sims_h[:n_layers].to(torch.float16)
    return torch.stack([qs[l, qa[l].long()] for l in range(n_layers)])


def get_intensities(activations: torch.Tensor, sims: torch.Tensor) -> torch.Tensor:
    """server.py:41-68: out[b, s, s2, l] = sims[s2, l, activations[b, s, l]].  The reference spells this as
    an expand + torch.gather; a gather has no arithmetic, so plain advanced indexing gives the same values.
    The result is made contiguous in the reference's (B, S, S2, D) order because the caller's ``sum(dim=-1)``
    must see the memory layout the reference's sum sees."""
    n_layers = activations.shape[-1]
    lay = torch.arange(n_layers)
    picked = sims[:, lay, activations.long()]            # (S2, B, S, D)
    return picked.permute(1, 2, 0, 3).contiguous()


def range_intensities(shard: torch.Tensor, qtab: torch.Tensor, a: int, b: int) -> torch.Tensor:
    """server.py:214-249 for one shard (n, S, nq) and the layer range [a, b) -> (n, S, Sq) fp16."""
    if b - a > CHUNK:
        out = None
        for c0 in range(a, b, CHUNK):
            c1 = min(c0 + CHUNK, b)
            part = get_intensities(shard[..., c0:c1], qtab[c0:c1].transpose(0, 1)).sum(dim=-1)
            if out is None:
                out = part
            else:
                out += part
        return out
    return get_intensities(shard[..., a:b], qtab[a:b].transpose(0, 1)).sum(dim=-1)


def accumulate(activations: Sequence[torch.Tensor], sims_h: torch.Tensor, query: torch.Tensor,
               layers: Sequence[int]) -> Iterator[torch.Tensor]:
    """The running ``intensity_accumulation`` (N, S, Sq) fp16 after every layer range (server.py:204-263)."""
    qtab = query_table(sims_h, query, max(layers))
    acc = None
    for a, b in layer_ranges(layers):
        step = torch.cat([range_intensities(sh, qtab, a, b) for sh in activations], dim=0)
        if acc is None:
            acc = step
        else:
            acc += step
        yield acc


def select(acc: torch.Tensor, top: int, middle: int, bottom: int):
    """server.py:265-312: (indices, intensities) for top / middle / bottom, indices (Sq, k) int32,
    intensities (Sq, k, S) fp16."""
    max_values = acc.max(dim=1).values
    order = max_values.argsort(dim=0, descending=True)
    n = order.shape[0]
    lists = {"top": order[:top].T, "middle": order[n // 2 - middle // 2: n // 2 + middle // 2].T,
             "bottom": order[-bottom:].T}
    out = {}
    for k, lst in lists.items():
        inten = torch.stack([acc[lst[i], :, i] for i in range(acc.shape[2])])
        out[k] = {"indices": lst.int(), "intensities": inten.to(torch.float16)}
    return out


def find_examples(activations: Sequence[torch.Tensor], sims_h: torch.Tensor, query: torch.Tensor,
                  top_examples: int = 30, middle_examples: int = 10, bottom_examples: int = 10,
                  layers: Sequence[int] = SERVER_LAYERS):
    """server.py:159-325 as a generator of (result dict, layer)."""
    for layer, acc in zip(layers, accumulate(activations, sims_h, query, layers)):
        yield select(acc, top_examples, middle_examples, bottom_examples), layer


def accumulate_steps(codes: torch.Tensor, sims_h: torch.Tensor, query: torch.Tensor,
                     layers: Sequence[int]) -> Iterator[torch.Tensor]:
    """The same accumulation with every rounding written out, in the order the CUDA kernel uses.
    codes (T, nq) int (all dataset tokens, sequence-major) -> running (T, Sq) fp16 per range."""
    L = max(layers)
    Sq = query.shape[0]
    T = codes.shape[0]
    acc = None
    for a, b in layer_ranges(layers):
        rng = None
        for c0 in range(a, b, CHUNK):
            c1 = min(c0 + CHUNK, b)
            run = torch.zeros(T, Sq, dtype=torch.float32)
            for l in range(c0, c1):
                rows = sims_h[l][query[:, l].long()]                 # (Sq, K) fp16: table rows of the query codes
                run = run + rows[:, codes[:, l].long()].T.float()    # fp32 add, ascending layer order
            h = run.to(torch.float16)                                # one rounding per chunk
            rng = h if rng is None else (rng.float() + h.float()).to(torch.float16)
        acc = rng if acc is None else (acc.float() + rng.float()).to(torch.float16)
        assert L >= b
        yield acc


# ---------------------------------------------------------------------------------------------------------------
# the rank-5 form of the table (groundwork for a tensor-core variant of the search)
# ---------------------------------------------------------------------------------------------------------------
def rank5_factors(w_out: torch.Tensor, b_out: torch.Tensor, codebook: torch.Tensor):
    """subfeature_sims[l][a][b] (rqae/model.py:145-167) is the cosine of the affine images W_out[l] c + b_out[l] of two
    codewords, i.e. u_l(a) . u_l(b) with the unit 5-vectors u_l(c) = R_l [c; 1] / |R_l [c; 1]|, R_l^T R_l = [W|b]^T [W|b].
    w_out (nq, D, cd), b_out (nq, D), codebook (nq, K, cd) -> (nq, K, cd + 1) float64 (zero rows for a zero image)."""
    g = torch.cat([w_out.double(), b_out.double().unsqueeze(-1)], dim=-1)          # (nq, D, 5)
    m = g.transpose(1, 2) @ g                                                      # (nq, 5, 5) Gram
    evals, evecs = torch.linalg.eigh(m)
    r = evals.clamp_min(0).sqrt().unsqueeze(-1) * evecs.transpose(1, 2)            # R = sqrt(L) V^T, R^T R = M
    x = torch.cat([codebook.double(), torch.ones_like(codebook[..., :1]).double()], dim=-1)   # (nq, K, 5)
    u = x @ r.transpose(1, 2)                                                      # rows R [c; 1]
    n = u.norm(dim=-1, keepdim=True)
    return torch.where(n > 1e-12, u / n.clamp_min(1e-12), torch.zeros_like(u))


def accumulate_rank5(codes: torch.Tensor, factors: torch.Tensor, layer_norms: torch.Tensor, query: torch.Tensor,
                     layers: Sequence[int], passes: int = 1) -> Iterator[torch.Tensor]:
    """What a tensor-core form of the search would compute: per layer the 5-term dot product of the dataset token's
    factor (fp16) with the query position's factor times the layer norm (fp16), exact products accumulated in fp32,
    with the reference's chunk / range roundings (as in ``accumulate_steps``).  ``passes = 3`` adds the fp16 remainders
    of both operands (hi*hi + hi*lo + lo*hi, the split the tensor-core decode uses).  codes (T, nq) -> (T, Sq) fp16."""
    def split(v):
        hi = v.to(torch.float16)
        lo = (v - hi.double()).to(torch.float16)
        return hi.float(), lo.float()
    d_hi, d_lo = split(factors)                                                    # dataset side
    q_hi, q_lo = split(factors * layer_norms.double().reshape(-1, 1, 1))           # query side carries the layer norm
    T, Sq = codes.shape[0], query.shape[0]
    acc = None
    for a, b in layer_ranges(layers):
        rng = None
        for c0 in range(a, b, CHUNK):
            run = torch.zeros(T, Sq, dtype=torch.float32)
            for l in range(c0, min(c0 + CHUNK, b)):
                dc, qc = codes[:, l].long(), query[:, l].long()
                run = run + d_hi[l][dc] @ q_hi[l][qc].T
                if passes == 3:
                    run = run + d_hi[l][dc] @ q_lo[l][qc].T + d_lo[l][dc] @ q_hi[l][qc].T
            h = run.to(torch.float16)
            rng = h if rng is None else (rng.float() + h.float()).to(torch.float16)
        acc = rng if acc is None else (acc.float() + rng.float()).to(torch.float16)
        yield acc
