"""Flip statistics of a tensor-core in-projection (VERDICT r1 "missing" item 7) -- a CPU study, not a
test and not product code: how often would the argmax of rqae/model.py:182-197 change if z = W_in r + b_in were computed
from operands rounded to a tensor-core input type (products exact, fp32 accumulation), teacher-forced along the fp32
path so that every layer sees the reference's own residual?  Run:  python tests/study_tc_encode_flips.py [tokens] [layers]

Variants: tf32 (10-bit mantissa, both operands), bf16, fp16 (one pass); fp16x3 = r_hi W_hi + r_hi W_lo + r_lo W_hi with
hi/lo fp16 splits (the decode kernel's "f16x3"); tf32x3 the same with tf32 splits; bf16x6 = three-way bf16 splits, the six
largest cross terms.  A flip whose fp64 top-1 / top-2 margin is <= 1e-5 is a near-tie the parity protocol accepts anyway
(SURVEY 8c); the others are real disagreements with the reference."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F

from oracle import rqae_oracle as orc


def rnd_bits(x: torch.Tensor, keep: int) -> torch.Tensor:
    """round-to-nearest-even of fp32 to `keep` explicit mantissa bits (tf32: 10)"""
    i = x.contiguous().view(torch.int32)
    drop = 23 - keep
    half = (1 << (drop - 1)) - 1
    lsb = (i >> drop) & 1
    return ((i + half + lsb) & ~((1 << drop) - 1)).view(torch.float32)


def split(x, fn, n):
    out, rem = [], x
    for _ in range(n):
        p = fn(rem)
        out.append(p)
        rem = rem - p
    return out


TYPES = {
    "tf32": lambda v: rnd_bits(v, 10),
    "bf16": lambda v: v.to(torch.bfloat16).float(),
    "fp16": lambda v: v.to(torch.float16).float(),
}


def variants(r, w):
    """name -> z without bias, fp32 accumulation of exact products"""
    res = {}
    for name, fn in TYPES.items():
        res[name] = fn(r) @ fn(w).T
    for name, base, n in (("fp16x3", "fp16", 2), ("tf32x3", "tf32", 2)):
        rs, ws = split(r, TYPES[base], n), split(w, TYPES[base], n)
        res[name] = rs[0] @ ws[0].T + (rs[0] @ ws[1].T + rs[1] @ ws[0].T)
    rs, ws = split(r, TYPES["bf16"], 3), split(w, TYPES["bf16"], 3)
    res["bf16x6"] = rs[0] @ ws[0].T + (rs[0] @ ws[1].T + rs[1] @ ws[0].T) + (rs[1] @ ws[1].T + rs[0] @ ws[2].T + rs[2] @ ws[0].T)
    return res


@torch.inference_mode()
def main():
    T = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    L = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
    w = orc.random_init(num_quantizers=L)
    x = torch.randn(T, w.dim, generator=torch.Generator().manual_seed(1))
    r = x.clone()
    r64 = x.double()
    w64 = w.to(torch.float64)
    names = list(TYPES) + ["fp16x3", "tf32x3", "bf16x6"]
    flips = {n: 0 for n in names}
    real = {n: 0 for n in names}
    tok = {n: torch.zeros(T, dtype=torch.bool) for n in names}
    tok_real = {n: torch.zeros(T, dtype=torch.bool) for n in names}
    zerr = {n: 0.0 for n in names}
    for l in range(L):
        cb = w.codebook[l]
        z = F.linear(r, w.w_in[l], w.b_in[l])
        cos = (z / z.norm(dim=-1, keepdim=True)) @ cb.T
        idx = cos.argmax(-1)
        # fp64 margin of the reference's choice on the fp64 image of the same residual
        z64 = F.linear(r.double(), w64.w_in[l], w64.b_in[l])
        cos64 = (z64 / z64.norm(dim=-1, keepdim=True)) @ w64.codebook[l].T
        m64 = orc._distinct_margin(cos64, w64.codebook[l], idx)
        for n, zz in variants(r, w.w_in[l]).items():
            zv = zz + w.b_in[l]
            iv = ((zv / zv.norm(dim=-1, keepdim=True)) @ cb.T).argmax(-1)
            diff = (cb[iv] != cb[idx]).any(-1)                      # duplicate rows decode identically: not a flip
            flips[n] += int(diff.sum())
            hard = diff & (m64 > 1e-5)
            real[n] += int(hard.sum())
            tok[n] |= diff
            tok_real[n] |= hard
            zerr[n] = max(zerr[n], float(((zv - z).abs().max(-1).values / z.abs().max(-1).values).max()))
        c = cb[idx]
        c = z + (c - z)
        o = F.linear(c, w.w_out[l], w.b_out[l])
        r = r - o
    print(f"{T} tokens x {L} layers, random-init 2B shape, teacher-forced along the fp32 path")
    print(f"{'variant':8s} {'flips/step':>12s} {'real flips/step':>16s} {'tokens with a flip':>19s} {'... a real flip':>16s} {'max rel err of z':>17s}")
    for n in names:
        print(f"{n:8s} {flips[n] / (T * L):12.2e} {real[n] / (T * L):16.2e} {float(tok[n].float().mean()):19.3f} "
              f"{float(tok_real[n].float().mean()):16.3f} {zerr[n]:17.1e}")


if __name__ == "__main__":
    main()
