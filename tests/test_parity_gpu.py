"""GPU parity tests (run on the B200 box with -m gpu): the CUDA path, called through the C ABI via
rqae_b200.RQAE, against (a) the plain-C oracle evaluated in the kernel's documented summation order --
bit-exact, codes AND reconstruction -- and (b) the golden vectors of the unmodified reference under the
near-tie protocol of tests/parity.py."""
import hashlib

import numpy as np
import pytest
import torch

from oracle import c_oracle
from tests import parity, util

pytestmark = pytest.mark.gpu

KERNEL_ORDER = c_oracle.KERNEL_ORDER  # see rq_forward.cuh header


def _cuda():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def model_2b():
    from rqae_b200 import RQAE
    torch.manual_seed(0)
    m = RQAE().eval()
    w = util.stacked_from_module(m)
    return m.to(_cuda()), c_oracle.CWeights.from_stacked(w)


FSQ_CASES = [c for c in util.SMALL_CASES if not c.startswith("vq")]


@pytest.mark.parametrize("name", FSQ_CASES)
def test_small_forward_bit_exact_vs_c_oracle_and_reference(golden_small, name):
    d = util.small_case(golden_small, name)
    m = util.module_from_case(d, _cuda())
    x = torch.from_numpy(d["x"]).to(_cuda())
    kw = {} if d["max_layers"] is None else dict(max_layers=d["max_layers"])
    q, idx = m(x, **kw)
    assert idx.dtype == torch.int64 and idx.shape == d["codes"].shape and q.shape == d["q"].shape
    qo, co = c_oracle.forward_f32(util.cweights(d), d["x"], max_layers=d["max_layers"], **KERNEL_ORDER)
    assert np.array_equal(idx.cpu().numpy(), co.astype(np.int64))
    assert np.array_equal(q.cpu().numpy(), qo)
    rep = parity.compare_codes(idx.cpu().numpy(), d["codes"], d["margins_fp64"])
    assert rep.failures == 0, str(rep)
    ok = parity.exact_token_mask(idx.cpu().numpy(), d["codes"])
    qr = d["q"].reshape(-1, d["q"].shape[-1])
    qt = q.cpu().numpy().reshape(qr.shape)
    assert np.abs(qt[ok] - qr[ok]).max() <= 2e-5 * np.abs(qr).max()   # stated fp tolerance


def test_learned_codebook_matches_reference(golden_small):
    d = util.small_case(golden_small, "vq_d256")
    m = util.module_from_case(d, _cuda())
    q, idx = m(torch.from_numpy(d["x"]).to(_cuda()))
    rep = parity.compare_codes(idx.cpu().numpy(), d["codes"], d["margins_fp64"])
    assert rep.failures == 0 and rep.exact >= rep.tokens - 1, str(rep)
    # the reference leaves the parameter renormalised in place (model.py:126-131); the device's
    # x / x.norm() differs from the CPU's in the last bit, hence a tolerance rather than equality
    assert np.allclose(m.codebook.detach().cpu().numpy(), d["codebook_post"], rtol=0, atol=3e-7)
    dec = m.decode(torch.from_numpy(d["codes_full"].astype(np.int64)).to(_cuda()))
    assert np.abs(dec.cpu().numpy() - d["dec"]).max() <= 2e-5 * np.abs(d["dec"]).max()


@pytest.mark.parametrize("name", FSQ_CASES)
def test_small_decode_bit_exact_vs_reference(golden_small, name):
    d = util.small_case(golden_small, name)
    m = util.module_from_case(d, _cuda())
    codes = torch.from_numpy(d["codes_full"].astype(np.int64)).to(_cuda())
    dec = m.decode(codes, layers=d["dec_layers"])
    assert np.array_equal(dec.cpu().numpy(), d["dec"])
    cv = m.indices_to_codebook_values(codes)
    dec2 = m.decode_from_codebook_values(cv, layers=d["dec_layers"])
    assert np.array_equal(dec2.cpu().numpy(), d["dec"])
    for dt in (torch.int16, torch.int32):
        assert torch.equal(m.decode(codes.to(dt), layers=d["dec_layers"]), dec)
    assert m.decode(codes, layers=[]) is None


def test_2b_kat_128_tokens(model_2b, golden_2b):
    m, cw = model_2b
    g = golden_2b
    x = torch.from_numpy(g["x128"]).to(_cuda())
    q, idx = m(x.view(1, 128, 2304))
    codes = idx[0].cpu().numpy()
    # (a) bit-exact against the C oracle in the kernel's summation order
    qo, co = c_oracle.forward_f32(cw, g["x128"], **KERNEL_ORDER)
    assert np.array_equal(codes, co.astype(np.int64))
    assert np.array_equal(q[0].cpu().numpy(), qo)
    # (b) the reference's own codes: near-tie protocol with the committed fp64 margins
    rep = parity.compare_codes(codes, g["codes1024"][:128], g["margins128_fp64"])
    print("2B KAT, 128 tokens vs reference:", rep)
    assert rep.failures == 0, str(rep)
    assert rep.exact >= 120
    ok = parity.exact_token_mask(codes, g["codes1024"][:128])
    rel = np.abs(q[0].cpu().numpy()[ok] - g["q128"][ok]).max() / np.abs(g["q128"]).max()
    print("   reconstruction max rel err on exact tokens:", rel)
    assert rel <= 2e-5


def test_2b_1024_tokens_vs_reference_codes(model_2b, golden_2b):
    m, cw = model_2b
    g = golden_2b
    x = util.x_2b(1024)
    assert hashlib.sha256(util.x_2b().numpy().tobytes()).hexdigest()[:16] == str(g["fp_x"])
    codes = m.encode(x.to(_cuda()).view(8, 128, 2304), out_dtype=torch.int16).cpu().numpy().reshape(1024, 1024)
    ref = g["codes1024"]
    bad = np.flatnonzero((codes != ref).any(axis=1))
    margins = np.full(ref.shape, np.inf, np.float32)
    if len(bad):
        _, _, mb = c_oracle.forward_f64(cw, x.numpy()[bad], teacher=ref[bad])
        margins[bad] = mb
    rep = parity.compare_codes(codes, ref, margins)
    print("2B, 1024 tokens vs reference:", rep)
    assert rep.failures == 0, str(rep)
    assert rep.near_tie <= 32   # expected ~0.2-1 % of tokens at nq=1024


def test_2b_all_4096_tokens_of_config0_vs_reference_codes(model_2b, golden_2b):
    """BASELINE configs[0] in full: every one of the 4096 KAT tokens against the reference's codes
    (tests/golden/kat_2b_codes4096.npz, sha-checked against the per-token fingerprints in kat_2b.npz)."""
    import os
    m, cw = model_2b
    g = golden_2b
    g4 = np.load(os.path.join(os.path.dirname(__file__), "golden", "kat_2b_codes4096.npz"))
    ref = g4["codes4096"]
    assert ref.shape == (4096, 1024) and str(g4["fp_codes_i16"]) == str(g["fp_codes_i16"])
    assert hashlib.sha256(ref.tobytes()).hexdigest()[:16] == str(g["fp_codes_i16"])
    tok_sha = np.frombuffer(b"".join(hashlib.sha256(r.tobytes()).digest()[:8] for r in ref), dtype=np.uint64)
    assert np.array_equal(tok_sha, g["tok_sha4096"])
    x = util.x_2b()
    codes = m.encode(x.to(_cuda()), out_dtype=torch.int16).cpu().numpy().reshape(4096, 1024)
    got_sha = np.frombuffer(b"".join(hashlib.sha256(r.tobytes()).digest()[:8] for r in codes), dtype=np.uint64)
    bad = np.flatnonzero(got_sha != g["tok_sha4096"])
    assert np.array_equal(bad, np.flatnonzero((codes != ref).any(axis=1)))
    margins = np.full(ref.shape, np.inf, np.float32)
    if len(bad):
        _, _, mb = c_oracle.forward_f64(cw, x.view(-1, 2304).numpy()[bad], teacher=ref[bad])
        margins[bad] = mb
    rep = parity.compare_codes(codes, ref, margins)
    print("2B, all 4096 tokens of configs[0] vs reference:", rep)
    assert rep.failures == 0, str(rep)
    assert rep.near_tie <= 128   # expected ~0.2-1 % of tokens at nq=1024


def test_9b_kat_full_depth_vs_reference(golden_9b):
    """BASELINE configs[3] shape (d=3584, nq=2048) pinned on the UNMODIFIED reference: codes under the near-tie
    protocol with fp64 margins along the reference's trajectory, reconstruction within the stated tolerance,
    decode bit-exact; and bit-exact against the C oracle in kernel order at full depth."""
    from rqae_b200 import RQAE
    g = golden_9b
    torch.manual_seed(0)
    m = RQAE(dim=3584, num_quantizers=2048).eval()
    h = hashlib.sha256()
    for k, v in m.state_dict().items():
        if k.startswith("layers."):
            h.update(v.numpy().tobytes())
    assert h.hexdigest()[:16] == str(g["fp_layers"])
    cw = c_oracle.CWeights.from_stacked(util.stacked_from_module(m))
    m = m.to(_cuda())
    x = torch.from_numpy(g["x"])
    assert hashlib.sha256(x.numpy().tobytes()).hexdigest()[:16] == str(g["fp_x"])
    n = x.shape[0]
    q, idx = m(x.to(_cuda()).view(1, n, 3584))
    codes = idx[0].cpu().numpy()
    qo, co = c_oracle.forward_f32(cw, g["x"], **KERNEL_ORDER)
    assert np.array_equal(codes, co.astype(np.int64)) and np.array_equal(q[0].cpu().numpy(), qo)
    rep = parity.compare_codes(codes, g["codes"], g["margins_fp64"])
    print("9B-width KAT (d=3584, nq=2048), %d tokens vs reference:" % n, rep)
    assert rep.failures == 0, str(rep)
    assert rep.exact >= n - 8
    ok = parity.exact_token_mask(codes, g["codes"])
    rel = np.abs(q[0].cpu().numpy()[ok] - g["q"][ok]).max() / np.abs(g["q"]).max()
    print("   reconstruction max rel err on exact tokens:", rel)
    assert rel <= 2e-5
    # teacher-forced: every (token, layer) of the reference trajectory checked independently
    teacher = torch.from_numpy(g["codes"].astype(np.int32)).to(_cuda()).view(1, n, 2048)
    _, tf, _ = m._run_forward(x.to(_cuda()).view(1, n, 3584), float("inf"), 0.0, False, torch.int32, teacher=teacher)
    mism = tf[0].cpu().numpy() != g["codes"]
    assert (g["margins_fp64"][mism] < parity.EPS).all() and mism.mean() < 1e-4
    dec = m.decode(torch.from_numpy(g["codes"][:8].astype(np.int64)).to(_cuda()).view(1, 8, 2048))
    assert np.array_equal(dec[0].cpu().numpy(), g["dec8"])


def test_2b_decode_bit_exact(model_2b, golden_2b):
    m, _ = model_2b
    g = golden_2b
    codes = torch.from_numpy(g["codes1024"][:128].astype(np.int64)).to(_cuda()).view(1, 128, 1024)
    dec = m.decode(codes)
    assert np.array_equal(dec[0].cpu().numpy(), g["dec128"])
    dec3 = m.decode(codes, layers=list(range(0, 1024, 3)))
    assert np.array_equal(dec3[0].cpu().numpy(), g["dec128_every3"])
    # forward's reconstruction and decode(forward codes) agree to rounding (reference: 4.8e-7 rel)
    q, idx = m(torch.from_numpy(g["x128"]).to(_cuda()).view(1, 128, 2304))
    d2 = m.decode(idx)
    assert (q - d2).abs().max().item() <= 2e-5 * q.abs().max().item()


def test_position_tile_and_batch_invariance(model_2b):
    """Codes of a token do not depend on where it sits in the batch, on the batch size, or on the
    number of CTAs (the summation order is fixed by construction)."""
    m, _ = model_2b
    dev = _cuda()
    base = torch.randn(37, 2304, generator=torch.Generator().manual_seed(7)).to(dev)
    ref = m.encode(base.view(1, 37, 2304), max_layers=64)[0]
    big = base.repeat(150, 1)[: 148 * 16 + 5]                 # more than one wave of unit pairs, ragged tail
    out = m.encode(big.view(1, -1, 2304), max_layers=64)[0]
    for i in range(0, big.shape[0], 37):
        n = min(37, big.shape[0] - i)
        assert torch.equal(out[i:i + n], ref[:n])
    one = m.encode(base[5:6].view(1, 1, 2304), max_layers=64)[0]
    assert torch.equal(one[0], ref[5])


def test_max_layers_prefix_and_dtypes(model_2b):
    m, _ = model_2b
    x = torch.randn(2, 9, 2304, generator=torch.Generator().manual_seed(9)).to(_cuda())
    q_full, idx_full = m(x, max_layers=48)
    q16, idx16 = m(x, max_layers=16)
    assert idx16.shape == (2, 9, 16) and torch.equal(idx16, idx_full[..., :16])
    dec16 = m.decode(idx_full, layers=range(16))
    assert (q16 - dec16).abs().max().item() <= 2e-5 * q16.abs().max().item()
    for dt in (torch.int16, torch.int32, torch.int64):
        assert torch.equal(m.encode(x, max_layers=48, out_dtype=dt).to(torch.int64), idx_full)
    assert m.encode(x[:0], max_layers=8).shape == (0, 9, 8)


def test_teacher_forced_mode_checks_every_layer(model_2b, golden_2b):
    """Teacher forcing: the recurrence follows the reference's codes while the kernel reports its own
    argmax per layer, so every (token, layer) is checked independently."""
    m, cw = model_2b
    g = golden_2b
    x = torch.from_numpy(g["x128"][:64]).to(_cuda())
    teacher = torch.from_numpy(g["codes1024"][:64].astype(np.int32)).to(_cuda())
    _, codes, z = m._run_forward(x.view(1, 64, 2304), float("inf"), 0.0, False, torch.int32, teacher=teacher.view(1, 64, 1024), want_z=True)
    codes = codes[0].cpu().numpy()
    mism = codes != g["codes1024"][:64]
    assert (g["margins128_fp64"][:64][mism] < parity.EPS).all()
    assert mism.mean() < 1e-4
    assert torch.isfinite(z).all()


def test_nan_and_inf_rows(model_2b):
    m, _ = model_2b
    x = torch.randn(1, 8, 2304, generator=torch.Generator().manual_seed(11)).to(_cuda())
    x[0, 3, 17] = float("inf")
    x[0, 5, 100] = float("nan")
    _, idx = m(x, max_layers=8)
    clean = m.encode(torch.randn(1, 8, 2304, generator=torch.Generator().manual_seed(11)).to(_cuda()), max_layers=8)
    assert (idx[0, 3] == 0).all() and (idx[0, 5] == 0).all()     # NaN row -> first NaN index = 0
    keep = [0, 1, 2, 4, 6, 7]
    assert torch.equal(idx[0, keep], clean[0, keep])              # neighbours in the same unit unaffected


def test_zero_and_tiny_coordinates_fall_back_to_full_scan():
    """The sign-orthant search is only valid when no coordinate of the normalised in-projection is tiny;
    rows that differ in the sign of a coordinate multiplied by (almost) zero tie exactly and the LOWEST index
    must win (torch.argmax).  Layer 0 has z1 == 0 exactly, layer 1 a denormal-scale z2, layer 2 two zeros."""
    from rqae_b200 import RQAE
    torch.manual_seed(21)
    m = RQAE(dim=256, num_quantizers=6).eval()
    with torch.no_grad():
        m.layers[0][0].weight[1].zero_(); m.layers[0][0].bias[1] = 0.0
        m.layers[1][0].weight[2].mul_(1e-12); m.layers[1][0].bias[2] = 0.0
        m.layers[2][0].weight[0].zero_(); m.layers[2][0].bias[0] = 0.0
        m.layers[2][0].weight[3].zero_(); m.layers[2][0].bias[3] = 0.0
    cw = c_oracle.CWeights.from_stacked(util.stacked_from_module(m))
    m = m.to(_cuda())
    x = torch.randn(5, 23, 256, generator=torch.Generator().manual_seed(22))
    q, idx = m(x.to(_cuda()))
    qo, co = c_oracle.forward_f32(cw, x.numpy(), **KERNEL_ORDER)
    assert np.array_equal(idx.cpu().numpy(), co.astype(np.int64))
    assert np.array_equal(q.cpu().numpy(), qo)
    # the tie really is resolved towards the negative twin (lower index) at layer 0: coordinate 1 of the
    # chosen codeword is never positive
    cb = m.codebook[0].cpu().numpy()
    assert (cb[idx[..., 0].cpu().numpy()][..., 1] <= 0).all()


def _designed_z_model(method="round_fsq", cbs=5):
    """Layer 0 copies x[..., :4] into z exactly (unit in-projection rows, zero bias), so a test can place z on
    the decision boundaries of the codebook search."""
    from rqae_b200 import RQAE
    torch.manual_seed(31)
    m = RQAE(dim=256, num_quantizers=3, quantization_method=method, codebook_size=cbs).eval()
    with torch.no_grad():
        m.layers[0][0].weight.zero_()
        m.layers[0][0].bias.zero_()
        for k in range(4):
            m.layers[0][0].weight[k, k] = 1.0
    return m


@pytest.mark.parametrize("method,cbs", [("round_fsq", 5), ("fsq", 5), ("round_fsq", 3), ("round_fsq", 4)])
def test_search_shortcut_on_decision_boundaries(method, cbs):
    """Adversarial inputs for the canonical-row search: exact and near ties between codewords (bisectors of
    row pairs), equal and nearly equal coordinate magnitudes, tiny and zero coordinates, huge and tiny scales.
    Codes and reconstruction must equal the exhaustive fp32 argmax of the C oracle bit for bit."""
    m = _designed_z_model(method, cbs)
    cw = c_oracle.CWeights.from_stacked(util.stacked_from_module(m))
    cb = m.codebook[0].double()
    K = cb.shape[0]
    g = torch.Generator().manual_seed(32)
    zs = []
    ia, ib = torch.randint(0, K, (6000,), generator=g), torch.randint(0, K, (6000,), generator=g)
    bis = cb[ia] + cb[ib]                                   # bisectors: exact ties in exact arithmetic
    zs.append(bis)
    zs.append(bis * (1 + 3e-7 * torch.randn(6000, 4, generator=g, dtype=torch.float64)))   # near ties
    base = torch.randn(4000, 4, generator=g, dtype=torch.float64)
    eq = base.clone(); eq[:, 1] = eq[:, 0] * torch.where(torch.rand(4000, generator=g) < 0.5, 1.0, -1.0)
    zs.append(eq)                                           # |z0| == |z1|
    zs.append(eq * (1 + 2e-7 * torch.randn(4000, 4, generator=g, dtype=torch.float64)))
    tiny = base.clone(); tiny[:, 2] *= 1e-7; tiny[:2000, 3] = 0.0
    zs.append(tiny)                                         # tiny / zero coordinates
    zs.append(cb[torch.randint(0, K, (2000,), generator=g)] * 3.0)          # exactly on codewords
    zs.append(torch.randn(4000, 4, generator=g, dtype=torch.float64))       # generic
    z = torch.cat(zs).float()
    z = z * torch.logspace(-6, 6, z.shape[0]).view(-1, 1)[torch.randperm(z.shape[0], generator=g)]
    x = torch.randn(z.shape[0], 256, generator=g)
    x[:, :4] = z
    m = m.to(_cuda())
    q, idx = m(x.to(_cuda()).view(1, -1, 256))
    qo, co = c_oracle.forward_f32(cw, x.numpy(), **KERNEL_ORDER)
    assert np.array_equal(idx[0].cpu().numpy(), co.astype(np.int64))
    assert np.array_equal(q[0].cpu().numpy(), qo)


def test_search_shortcut_random_stress():
    """1.3 M (token, layer) decisions of a small random model against the exhaustive C oracle."""
    from rqae_b200 import RQAE
    torch.manual_seed(41)
    m = RQAE(dim=256, num_quantizers=64).eval()
    cw = c_oracle.CWeights.from_stacked(util.stacked_from_module(m))
    x = torch.randn(20000, 256, generator=torch.Generator().manual_seed(42))
    codes = m.to(_cuda()).encode(x.to(_cuda()).view(1, -1, 256), out_dtype=torch.int32)[0].cpu().numpy()
    _, co = c_oracle.forward_f32(cw, x.numpy(), want_q=False, **KERNEL_ORDER)
    assert np.array_equal(codes, co)


def test_hook_with_stub_llm(model_2b):
    m, _ = model_2b

    class Stub:
        w = torch.linspace(-0.1, 0.1, 2304, device="cuda")

        def norm(self, hs):
            return hs * torch.rsqrt(hs.pow(2).mean(-1, keepdim=True) + 1e-6) * (1.0 + self.w)

        def denorm(self, hs, orig):
            return hs / (1.0 + self.w) / torch.rsqrt(orig.float().pow(2).mean(-1, keepdim=True) + 1e-6)

    stash = {}
    hook = m.hook(llm=Stub(), store=lambda k, v: stash.__setitem__(k, v))
    hs = (torch.randn(2, 6, 2304, generator=torch.Generator().manual_seed(13)) * 3).to(_cuda()).half()
    out = (hs.clone(),)
    old = m.num_quantizers
    hook(None, None, out)
    assert set(stash) == {"original", "normed", "quantized", "indices", "new"}
    assert stash["indices"].shape == (2, 6, old) and stash["indices"].dtype == torch.int64
    assert torch.equal(out[0][:, 0], hs[:, 0])                    # BOS passthrough
    assert torch.equal(out[0], stash["new"].half())
    q, idx = m(stash["normed"])
    assert torch.equal(idx, stash["indices"]) and torch.equal(q, stash["quantized"])


class _GemmaStub:
    """rqae/llm.py:60-73 with a synthetic RMSNorm weight: norm = Gemma2RMSNorm (x * rsqrt(mean(x^2) + eps) * (1 + w)),
    denorm = its inverse with the reference's hard-coded 1e-6."""

    def __init__(self, dim, eps=1e-6, seed=5):
        self.eps = eps
        self.weight = (0.1 * torch.randn(dim, generator=torch.Generator().manual_seed(seed))).cuda()

    def norm(self, hs):
        x = hs.float()
        return (x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + self.eps) * (1.0 + self.weight.float())).type_as(hs)

    def denorm(self, hs, orig):
        hs = hs / (1.0 + self.weight.float())
        return hs.float() / torch.rsqrt(orig.float().pow(2).mean(-1, keepdim=True) + 1e-6)


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16, torch.float32])
def test_fused_hook_equals_generic_hook(model_2b, dtype):
    """hook_fn (model.py:276-289) as one launch (rqae_hook_rmsnorm) against the generic path (torch
    .float / norm / denorm / BOS / copy_ around the forward kernel) with the arithmetic of llm.py:65-73.  The two sum
    the squares in a different order, so the normalised input differs in the last fp32 bit: codes are compared under
    the near-tie protocol (margins from the fp64 C oracle along the generic path's codes), replaced hidden states on
    tokens with identical codes within one rounding step of the output dtype."""
    m, cw = model_2b
    stub = _GemmaStub(2304)
    B, S = 4, 37
    hs = (3.0 * torch.randn(B, S, 2304, generator=torch.Generator().manual_seed(14))).to(_cuda()).to(dtype)
    stash = {}
    ref_out = (hs.clone(),)
    m.hook(norm=stub.norm, denorm=stub.denorm, store=lambda k, v: stash.__setitem__(k, v))(None, None, ref_out)
    fused = hs.clone()
    codes = m.hook_rmsnorm_(fused, stub.weight, stub.eps, return_codes=True)
    torch.cuda.synchronize()
    assert torch.equal(fused[:, 0], hs[:, 0])                                      # BOS passthrough, bit for bit
    ref_codes = stash["indices"].cpu().numpy().reshape(B * S, -1)
    got = codes.cpu().numpy().reshape(B * S, -1)
    bad = np.flatnonzero((got != ref_codes).any(axis=1))
    margins = np.full(ref_codes.shape, np.inf, np.float32)
    if len(bad):
        xn = stash["normed"].cpu().numpy().reshape(B * S, -1)
        _, _, mb = c_oracle.forward_f64(cw, xn[bad], teacher=ref_codes[bad])
        margins[bad] = mb
    rep = parity.compare_codes(got, ref_codes, margins)
    print(f"fused hook ({dtype}) vs generic hook:", rep)
    assert rep.failures == 0, str(rep)
    same = torch.from_numpy(parity.exact_token_mask(got, ref_codes)).view(B, S).to(_cuda())
    a, b = fused.float()[same], ref_out[0].float()[same]
    step = {torch.float16: 2.0 ** -10, torch.bfloat16: 2.0 ** -7, torch.float32: 2e-5}[dtype]
    assert ((a - b).abs() <= step * b.abs() + 2e-5 * b.abs().max()).all()   # one output rounding step + the fp32 tolerance
    # the hook object takes the fused path on its own when it is told the norm weight, and leaves the input alone
    # with replace=False
    via_hook = hs.clone()
    m.hook(norm=stub.norm, denorm=stub.denorm, rms_weight=stub.weight, rms_eps=stub.eps)(None, None, (via_hook,))
    assert torch.equal(via_hook, fused)
    untouched = hs.clone()
    m.hook(norm=stub.norm, denorm=stub.denorm, rms_weight=stub.weight, replace=False)(None, None, (untouched,))
    assert torch.equal(untouched, hs)
    no_bos = hs.clone()
    m.hook_rmsnorm_(no_bos, stub.weight, stub.eps, skip_bos=False)
    assert not torch.equal(no_bos[:, 0], hs[:, 0]) and torch.equal(no_bos[:, 1:], fused[:, 1:])


def test_forward_host_matches_device_path(model_2b):
    m, _ = model_2b
    x = torch.randn(3000, 2304, generator=torch.Generator().manual_seed(17))
    q_d, idx_d = m(x.to(_cuda()).view(1, -1, 2304), max_layers=32)
    q_h, idx_h = m.forward_host(x.pin_memory(), max_layers=32, chunk_tokens=1024)
    assert torch.equal(idx_h, idx_d[0].cpu()) and torch.equal(q_h, q_d[0].cpu())
    # int32 / int16 codes, caller-owned result tensors reused across calls, codes-only mode, ragged last chunk
    for dt in (torch.int32, torch.int16):
        qo = torch.empty(3000, 2304).pin_memory()
        co = torch.empty(3000, 32, dtype=dt).pin_memory()
        for chunk in (700, 4096):
            co.zero_()
            q2, c2 = m.forward_host(x.pin_memory(), max_layers=32, chunk_tokens=chunk, out_dtype=dt, out=(qo, co))
            assert c2 is co and q2 is qo
            assert torch.equal(co.to(torch.int64), idx_d[0].cpu()) and torch.equal(qo, q_d[0].cpu())
    _, c3 = m.forward_host(x.pin_memory(), max_layers=32, want_q=False, chunk_tokens=999)
    assert torch.equal(c3, idx_d[0].cpu())
    # both code-transfer modes of the host pipeline, every widening thread count
    for mode, thr in (("narrow", 1), ("narrow", 3), ("direct", 0), ("auto", 0)):
        for dt in (torch.int64, torch.int32):
            q4, c4 = m.forward_host(x.pin_memory(), max_layers=32, chunk_tokens=500, out_dtype=dt, code_transfer=mode,
                                    widen_threads=thr)
            assert torch.equal(c4.to(torch.int64), idx_d[0].cpu()) and torch.equal(q4, q_d[0].cpu()), (mode, thr, dt)


def test_gemma9b_width_bit_exact_vs_c_oracle():
    from rqae_b200 import RQAE
    torch.manual_seed(5)
    m = RQAE(dim=3584, num_quantizers=24).eval()
    cw = c_oracle.CWeights.from_stacked(util.stacked_from_module(m))
    m = m.to(_cuda())
    x = torch.randn(1, 29, 3584, generator=torch.Generator().manual_seed(6))
    q, idx = m(x.to(_cuda()))
    qo, co = c_oracle.forward_f32(cw, x.numpy(), **KERNEL_ORDER)
    assert np.array_equal(idx.cpu().numpy(), co.astype(np.int64)) and np.array_equal(q.cpu().numpy(), qo)
    dec = m.decode(idx)
    assert np.array_equal(dec.cpu().numpy(), c_oracle.decode_f32(cw, codes=co))


def test_gemma9b_width_cluster_variant_bit_exact_in_its_own_order(golden_9b):
    """Opt-in D-split cluster variant (rqae_forward_variant(1)): two CTAs per unit, each owning half of the hidden
    dimension, partials exchanged through distributed shared memory.  Bit-exact against the C oracle evaluated in
    THAT summation order (KERNEL_ORDER_9B), the reference's golden codes under the near-tie protocol, ragged and
    multi-wave token counts, teacher forcing; and the default variant is restored."""
    from rqae_b200 import RQAE, _lib
    lib = _lib.load()
    g = golden_9b
    torch.manual_seed(0)
    m = RQAE(dim=3584, num_quantizers=2048).eval()
    cw = c_oracle.CWeights.from_stacked(util.stacked_from_module(m))
    m = m.to(_cuda())
    x = torch.from_numpy(g["x"])
    n = x.shape[0]
    assert lib.rqae_forward_variant(1) == 0
    try:
        q, idx = m(x.to(_cuda()).view(1, n, 3584))
        codes = idx[0].cpu().numpy()
        qo, co = c_oracle.forward_f32(cw, g["x"], **c_oracle.KERNEL_ORDER_9B)
        assert np.array_equal(codes, co.astype(np.int64)) and np.array_equal(q[0].cpu().numpy(), qo)
        rep = parity.compare_codes(codes, g["codes"], g["margins_fp64"])
        print("9B-width KAT, cluster variant vs reference:", rep)
        assert rep.failures == 0, str(rep)
        # more units than clusters (lock-step path), ragged tail, shallow prefix
        xb = torch.randn(1, 74 * 16 * 2 + 21, 3584, generator=torch.Generator().manual_seed(8))
        ib = m.encode(xb.to(_cuda()), max_layers=12, out_dtype=torch.int32)[0].cpu().numpy()
        _, cb = c_oracle.forward_f32(cw, xb[0].numpy(), max_layers=12, want_q=False, **c_oracle.KERNEL_ORDER_9B)
        assert np.array_equal(ib, cb)
        teacher = torch.from_numpy(g["codes"].astype(np.int32)).to(_cuda()).view(1, n, 2048)
        _, tf, _ = m._run_forward(x.to(_cuda()).view(1, n, 3584), float("inf"), 0.0, False, torch.int32, teacher=teacher)
        mism = tf[0].cpu().numpy() != g["codes"]
        assert (g["margins_fp64"][mism] < parity.EPS).all()
    finally:
        assert lib.rqae_forward_variant(0) == 1
    _, idx0 = m(x.to(_cuda()).view(1, n, 3584))
    _, c0 = c_oracle.forward_f32(cw, g["x"], want_q=False, **KERNEL_ORDER)
    assert np.array_equal(idx0[0].cpu().numpy(), c0.astype(np.int64))


def test_derived_tables_on_gpu_match_the_reference_formulas():
    """SURVEY 8f-4 / model.py:133-178 on the device: ``subfeature_sims`` from the per-layer 5x5 Gram of [W_out | b_out]
    (no (nq, K, D) intermediate) against the reference's own construction written out -- F.normalize(lin_out[l](codebook[l]))
    and its Gram matrix in fp16 -- and ``layer_norms`` / ``codebook_sims`` against their formulas.  Load-time code on
    cuBLAS, not a kernel of this library; the test pins its values on the GPU as the CPU tests do on the host."""
    import torch.nn.functional as F
    from rqae_b200 import RQAE
    torch.manual_seed(3)
    m = RQAE(dim=320, num_quantizers=9).eval().to(_cuda())
    with torch.no_grad():
        sims = m.subfeature_sims
        assert sims.shape == (9, 625, 625) and sims.dtype == torch.float16 and sims.is_cuda
        worst, differ = 0.0, 0.0
        for l in range(9):
            sub = F.normalize(m.layers[l][1](m.codebook[l]), dim=-1)                       # (K, D): model.py:145-167
            ref = (sub @ sub.T).to(torch.float16)
            worst = max(worst, float((ref.float() - sims[l].float()).abs().max()))
            differ = max(differ, float((ref != sims[l]).float().mean()))
        assert worst <= 2.0 ** -10 and differ < 0.01, (worst, differ)                    # one fp16 step, rare
        ln = torch.stack([m.layers[l][1].weight.norm(dim=0).mean() for l in range(9)])
        assert torch.allclose(m.layer_norms.to(ln.device), ln, rtol=1e-6, atol=0)
        cb = F.normalize(m.codebook[0], dim=-1)
        assert torch.equal(m.codebook_sims, (cb @ cb.T).to(torch.float16))


def test_cpu_tensors_are_rejected(model_2b):
    m, _ = model_2b
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 1, 2304))


def test_many_units_per_cta_lockstep_path_bit_exact():
    """More units than SMs: every CTA loops over several units and the grid lock-step (cooperative launch,
    rq_forward.cuh) is active; a ragged last unit leaves some CTAs with one unit fewer."""
    from rqae_b200 import RQAE
    torch.manual_seed(11)
    m = RQAE(dim=256, num_quantizers=6).eval()
    cw = c_oracle.CWeights.from_stacked(util.stacked_from_module(m))
    m = m.to(_cuda())
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    n = sms * 16 * 3 + 16 * 7 + 5
    x = torch.randn(1, n, 256, generator=torch.Generator().manual_seed(12))
    q, idx = m(x.to(_cuda()))
    qo, co = c_oracle.forward_f32(cw, x.numpy(), **KERNEL_ORDER)
    assert np.array_equal(idx.cpu().numpy(), co.astype(np.int64)) and np.array_equal(q.cpu().numpy(), qo)
    # two launches in flight on different streams use different lock-step counters
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    xd = x.to(_cuda())
    torch.cuda.synchronize()
    with torch.cuda.stream(s1):
        i1 = m.encode(xd, out_dtype=torch.int32)
    with torch.cuda.stream(s2):
        i2 = m.encode(xd, out_dtype=torch.int32)
    torch.cuda.synchronize()
    assert np.array_equal(i1.cpu().numpy(), co) and np.array_equal(i2.cpu().numpy(), co)


# ---------------------------------------------------------------------------------------------------
# opt-in tensor-core decode (tcgen05 GEMM over codes): tolerance against the bit-exact decode
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("precision,rtol", [("f16", 1e-3), ("f16x3", 5e-5)])
def test_tensor_core_decode_within_stated_tolerance(model_2b, golden_2b, precision, rtol):
    """a tensor-core decode is judged by relative error (<= 1e-3 for a single reduced-precision pass);
    the three-pass split removes the operand rounding and is left with the tensor core's fp32 accumulation
    (measured 1.8e-5 over 262 144 tokens; bound 5e-5)."""
    m, _ = model_2b
    codes = torch.from_numpy(golden_2b["codes1024"].astype(np.int64))[:300].to(_cuda())     # ragged: 2 token tiles
    exact = m.decode(codes)
    fast = m.decode(codes, precision=precision)
    torch.cuda.synchronize()
    assert fast.shape == exact.shape and fast.dtype == torch.float32
    err = (fast - exact).abs().max().item() / exact.abs().max().item()
    assert err <= rtol, f"{precision}: max abs error / max |q| = {err}"
    # the layers= filter and narrow code dtypes go through the same path
    sub = list(range(0, 1024, 3))
    a = m.decode(codes.to(torch.int16), layers=sub, precision=precision)
    b = m.decode(codes, layers=sub)
    assert (a - b).abs().max().item() / b.abs().max().item() <= rtol
    with pytest.raises(ValueError):
        m.decode(codes, precision="bf16")


def test_tensor_core_decode_small_shapes(golden_small):
    d = util.small_case(golden_small, "round_fsq_d200_ragged")      # D = 200: one partial feature tile
    m = util.module_from_case(d, _cuda())
    codes = torch.from_numpy(d["codes"].astype(np.int64)).to(_cuda())
    exact = m.decode(codes)
    fast = m.decode(codes, precision="f16x3")
    assert (fast - exact).abs().max().item() <= 5e-5 * exact.abs().max().item()
