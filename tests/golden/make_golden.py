"""Generate the golden vectors under tests/golden/ from the UNMODIFIED reference.

Run in the build container only (needs /root/reference, which does not exist on
the GPU box):  python tests/golden/make_golden.py

The reference ships no tests or fixtures, so parity is
pinned on outputs of the reference itself.  Two files are written:

  kat_2b.npz     2B-shape random-init model (torch.manual_seed(0); RQAE()),
                 x = randn(32,128,2304, seed 1): reference codes for the first
                 1024 tokens, x / q_out / decode for the first 128, fp64 margins
                 for the first 128 (from oracle.rqae_oracle, dtype=float64),
                 sha256 fingerprints of weights, x and the full 4096-token codes.
  kat_small.npz  small configs with the weights stored in full: round_fsq, fsq,
                 a learned-codebook method, max_layers, decode(layers=...),
                 zero rows (NaN rule), duplicate-codeword tie rule.
  kat_2b_codes4096.npz  the reference's int16 codes of ALL 4096 tokens of the 2B KAT
                 (BASELINE configs[0]); sha-checked against kat_2b.npz's fp_codes_i16.
  kat_9b.npz     Gemma-2-9B width with the deeper stack of BASELINE configs[3]
                 (torch.manual_seed(0); RQAE(dim=3584, num_quantizers=2048)), 64 tokens:
                 reference codes, reconstruction, decode, fp64 margins teacher-forced
                 along the reference's trajectory, weight fingerprints.

`--only NAME` (small | 2b | codes4096 | 9b) regenerates one file.
"""
import hashlib
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from rqae.model import RQAE as RefRQAE  # noqa: E402  (the unmodified reference)
from oracle import rqae_oracle as orc    # noqa: E402


def sha16(t: torch.Tensor) -> str:
    return hashlib.sha256(t.detach().contiguous().cpu().numpy().tobytes()).hexdigest()[:16]


def layers_fingerprint(sd) -> str:
    h = hashlib.sha256()
    for k, v in sd.items():
        if k.startswith("layers."):
            h.update(v.detach().contiguous().numpy().tobytes())
    return h.hexdigest()[:16]


def make_2b():
    torch.manual_seed(0)
    ref = RefRQAE().eval()
    sd = ref.state_dict()
    x = torch.randn(32, 128, 2304, generator=torch.Generator().manual_seed(1))
    t0 = time.time()
    with torch.inference_mode():
        q, idx = ref(x)
    t_fwd = time.time() - t0
    idx = idx.contiguous()
    t0 = time.time()
    with torch.inference_mode():
        dec = ref.decode(idx[:1])
        dec_sub = ref.decode(idx[:1], layers=list(range(0, 1024, 3)))
    t_dec = time.time() - t0
    w = orc.StackedWeights.from_state_dict(sd)
    _, idx64, margins = orc.forward(w, x[:1], dtype=torch.float64, want_margins=True)
    # teacher-forced fp64 margins along the REFERENCE's code path (so a margin exists at
    # every (token, layer) of the reference trajectory even after an fp32/fp64 flip)
    _, idx64_tf, margins_tf = orc.forward(w, x[:1], dtype=torch.float64, want_margins=True,
                                          teacher_codes=idx[:1])
    tok_sha = np.frombuffer(
        b"".join(hashlib.sha256(r.numpy().tobytes()).digest()[:8] for r in idx.view(-1, 1024).to(torch.int16)),
        dtype=np.uint64)
    out = dict(
        x128=x[0].numpy(),
        codes1024=idx.view(-1, 1024)[:1024].to(torch.int16).numpy(),
        q128=q[0].numpy(),
        dec128=dec[0].numpy(),
        dec128_every3=dec_sub[0].numpy(),
        margins128_fp64=margins_tf[0].numpy().astype(np.float32),
        codes128_fp64=idx64_tf[0].to(torch.int16).numpy(),
        tok_sha4096=tok_sha,
        fp_layers=layers_fingerprint(sd),
        fp_codebook0=sha16(sd["codebook"][0]),
        fp_x=sha16(x),
        fp_codes_i16=sha16(idx.to(torch.int16)),
        codes_sum=np.int64(idx.sum().item()),
        q_norm_mean=np.float64(q.norm(dim=-1).mean().item()),
        ref_forward_seconds_4096tok=np.float64(t_fwd),
        ref_decode_seconds_128tok=np.float64(t_dec),
        ref_threads=np.int64(torch.get_num_threads()),
        torch_version=torch.__version__,
    )
    np.savez_compressed(os.path.join(HERE, "kat_2b.npz"), **out)
    print("2b:", {k: out[k] for k in ("fp_layers", "fp_codebook0", "fp_x", "fp_codes_i16", "codes_sum")},
          "fwd s", t_fwd)
    free_run = (idx64[0] == idx[0]).all(-1).float().mean().item()
    print("   fp64 free-run tokens identical to fp32 reference:", free_run)


def small_case(name, out, *, dim, nq, method, cbs=5, seed=0, tokens=(3, 17), max_layers=None,
               dec_layers=None, zero_rows=(), scale_out=None):
    torch.manual_seed(seed)
    ref = RefRQAE(dim=dim, num_quantizers=nq, quantization_method=method, codebook_size=cbs).eval()
    if scale_out is not None:  # "pseudo-trained": make the residual shrink so recon tolerance is meaningful
        with torch.no_grad():
            for l, layer in enumerate(ref.layers):
                w_in = layer[0].weight
                layer[1].weight.copy_(torch.linalg.pinv(w_in) * scale_out * (0.97 ** l))
                layer[1].bias.mul_(0.01)
    # clone: for learned codebooks the reference renormalises the parameter in place during forward
    sd = {k: v.detach().clone() for k, v in ref.state_dict().items()}
    x = torch.randn(*tokens, dim, generator=torch.Generator().manual_seed(seed + 100))
    for r in zero_rows:
        x.view(-1, dim)[r] = 0
    kw = {} if max_layers is None else dict(max_layers=max_layers)
    with torch.inference_mode():
        q, idx = ref(x, **kw)
        idx = idx.contiguous()
        if idx.shape[-1] == nq:
            full = idx
        else:
            ref.load_state_dict(sd)
            full = ref(x)[1].contiguous()
        cb_post = ref.codebook.detach().clone()   # the table decode() indexes (state after forward)
        dec = ref.decode(full, layers=dec_layers)
    w = orc.StackedWeights.from_state_dict(sd, method)
    _, i64, m = orc.forward(w, x, dtype=torch.float64, want_margins=True, teacher_codes=full,
                            max_layers=float("inf") if max_layers is None else max_layers)
    p = name + "/"
    out[p + "w_in"] = w.w_in.numpy(); out[p + "b_in"] = w.b_in.numpy()
    out[p + "w_out"] = w.w_out.numpy(); out[p + "b_out"] = w.b_out.numpy()
    out[p + "codebook"] = w.codebook.numpy()
    out[p + "codebook_post"] = cb_post.numpy()
    out[p + "x"] = x.numpy(); out[p + "q"] = q.numpy(); out[p + "codes"] = idx.to(torch.int16).numpy()
    out[p + "codes_full"] = full.to(torch.int16).numpy()
    out[p + "dec"] = dec.numpy(); out[p + "margins_fp64"] = m.numpy().astype(np.float32)
    out[p + "method"] = method; out[p + "cbs"] = np.int64(cbs)
    out[p + "max_layers"] = np.int64(-1 if max_layers is None else max_layers)
    out[p + "dec_layers"] = np.array([] if dec_layers is None else list(dec_layers), dtype=np.int64)
    out[p + "has_dec_layers"] = np.int64(dec_layers is not None)
    print(name, "codes", tuple(idx.shape), "q", float(q.norm(dim=-1).mean()))


def make_small():
    out = {}
    small_case("round_fsq_d256", out, dim=256, nq=8, method="round_fsq")
    small_case("round_fsq_d256_zero", out, dim=256, nq=8, method="round_fsq", zero_rows=(0, 5), seed=1)
    small_case("round_fsq_d512_ml16", out, dim=512, nq=32, method="round_fsq", max_layers=16,
               dec_layers=range(16), seed=2)
    small_case("fsq_d384", out, dim=384, nq=12, method="fsq", seed=3)
    small_case("vq_d256", out, dim=256, nq=10, method="vq", cbs=64, seed=4)
    small_case("round_fsq_d200_ragged", out, dim=200, nq=6, method="round_fsq", tokens=(1, 5), seed=5)
    small_case("round_fsq_d768_trained", out, dim=768, nq=48, method="round_fsq", tokens=(2, 33), seed=6,
               scale_out=1.0, dec_layers=[0, 1, 5, 47])
    small_case("round_fsq_cbs3", out, dim=128, nq=5, method="round_fsq", cbs=3, seed=7)
    np.savez_compressed(os.path.join(HERE, "kat_small.npz"), **out)


def make_codes4096():
    """All 4096 tokens of the 2B KAT through the unmodified reference (about two minutes on 8 cores)."""
    torch.manual_seed(0)
    ref = RefRQAE().eval()
    x = torch.randn(32, 128, 2304, generator=torch.Generator().manual_seed(1))
    t0 = time.time()
    with torch.inference_mode():
        _, idx = ref(x)
    idx16 = idx.contiguous().to(torch.int16)
    g = np.load(os.path.join(HERE, "kat_2b.npz"))
    assert sha16(idx16) == str(g["fp_codes_i16"]), "codes differ from the committed fingerprint"
    np.savez_compressed(os.path.join(HERE, "kat_2b_codes4096.npz"), codes4096=idx16.view(-1, 1024).numpy(),
                        fp_codes_i16=sha16(idx16), fp_x=sha16(x))
    print("codes4096:", sha16(idx16), "fwd s", time.time() - t0)


def make_9b(tokens=64):
    torch.manual_seed(0)
    ref = RefRQAE(dim=3584, num_quantizers=2048).eval()
    sd = ref.state_dict()
    x = torch.randn(1, tokens, 3584, generator=torch.Generator().manual_seed(3))
    t0 = time.time()
    with torch.inference_mode():
        q, idx = ref(x)
        idx = idx.contiguous()
        dec = ref.decode(idx[:, :8])
    t_fwd = time.time() - t0
    w = orc.StackedWeights.from_state_dict(sd)
    _, idx64_tf, margins_tf = orc.forward(w, x, dtype=torch.float64, want_margins=True, teacher_codes=idx)
    out = dict(
        x=x[0].numpy(), codes=idx[0].to(torch.int16).numpy(), q=q[0].numpy(), dec8=dec[0].numpy(),
        margins_fp64=margins_tf[0].numpy().astype(np.float32), codes_fp64=idx64_tf[0].to(torch.int16).numpy(),
        fp_layers=layers_fingerprint(sd), fp_codebook0=sha16(sd["codebook"][0]), fp_x=sha16(x),
        fp_codes_i16=sha16(idx.to(torch.int16)), codes_sum=np.int64(idx.sum().item()),
        ref_forward_seconds=np.float64(t_fwd), ref_threads=np.int64(torch.get_num_threads()),
        torch_version=torch.__version__,
    )
    np.savez_compressed(os.path.join(HERE, "kat_9b.npz"), **out)
    print("9b:", {k: out[k] for k in ("fp_layers", "fp_x", "fp_codes_i16", "codes_sum")}, "fwd s", t_fwd)
    print("   fp64 teacher-forced argmax identical to fp32 reference at",
          float((idx64_tf[0] == idx[0]).float().mean()), "of (token, layer) pairs")


if __name__ == "__main__":
    only = sys.argv[sys.argv.index("--only") + 1] if "--only" in sys.argv else None
    if only in (None, "small"):
        make_small()
    if only in (None, "2b") and "--small-only" not in sys.argv:
        make_2b()
    if only in (None, "codes4096") and "--small-only" not in sys.argv:
        make_codes4096()
    if only in (None, "9b") and "--small-only" not in sys.argv:
        make_9b()
