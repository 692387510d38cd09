"""Host-side logic of rqae_b200.RQAE that needs no GPU: construction parity with the reference,
state-dict contract, error behaviour, hook wiring, learned-codebook normalisation schedule."""
import hashlib

import numpy as np
import pytest
import torch

from rqae_b200 import RQAE
from tests import util


def test_constructor_reproduces_reference_parameters(golden_2b):
    torch.manual_seed(0)
    m = RQAE()
    h = hashlib.sha256()
    sd = m.state_dict()
    for k, v in sd.items():
        if k.startswith("layers."):
            h.update(v.numpy().tobytes())
    assert h.hexdigest()[:16] == str(golden_2b["fp_layers"])
    assert hashlib.sha256(sd["codebook"][0].numpy().tobytes()).hexdigest()[:16] == str(golden_2b["fp_codebook0"])
    assert len(sd) == 4098 and sum(p.numel() for p in m.parameters()) == 23_797_760 + 1024 * 625 * 4 - 1024 * 625 * 4
    assert sd["layers.0.0.weight"].shape == (4, 2304) and sd["layers.0.1.weight"].shape == (2304, 4)
    assert sd["codebook"].shape == (1024, 625, 4) and sd["codebook_counts"].shape == (1024, 625)
    assert not m.codebook.requires_grad
    assert (m.codebook[0] == m.codebook[-1]).all()


def test_state_dict_round_trip_strict(golden_small):
    d = util.small_case(golden_small, "round_fsq_d256")
    m = util.module_from_case(d)
    m2 = RQAE(dim=256, num_quantizers=8)
    m2.load_state_dict(m.state_dict(), strict=True)
    for (k, a), (_, b) in zip(m.state_dict().items(), m2.state_dict().items()):
        assert torch.equal(a, b), k
    with pytest.raises(RuntimeError):
        m2.load_state_dict({"bogus": torch.zeros(1)}, strict=True)


def test_learned_codebook_shapes_and_normalisation():
    torch.manual_seed(4)
    m = RQAE(dim=256, num_quantizers=10, quantization_method="vq", codebook_size=64)
    assert m.codebook.shape == (10, 64, 4) and m.codebook.requires_grad
    assert torch.allclose(m.codebook.detach().norm(dim=-1), torch.ones(10, 64), atol=1e-6)


def test_no_cpu_fallback():
    m = RQAE(dim=256, num_quantizers=4).eval()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.zeros(1, 2, 256))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m.decode(torch.zeros(1, 2, 4, dtype=torch.int64))


def test_training_gumbel_path_not_supported():
    m = RQAE(dim=256, num_quantizers=4).train()
    with pytest.raises(NotImplementedError):
        m(torch.zeros(1, 2, 256), temperature=0.5)


def test_hook_argument_contract_and_wiring(monkeypatch):
    m = RQAE(dim=64, num_quantizers=3).eval()
    with pytest.raises(ValueError):
        m.hook(denorm=lambda a, b: a)
    with pytest.raises(ValueError):
        m.hook(norm=lambda a: a)
    with pytest.raises(AssertionError):
        m.hook(llm=object())
    calls = {}

    def fake_forward(self, x, max_layers=float("inf"), temperature=0.0):
        calls["x"] = x
        return x * 0.5, torch.zeros(*x.shape[:-1], 3, dtype=torch.int64)

    monkeypatch.setattr(RQAE, "forward", fake_forward)
    stash = {}
    hook = m.hook(norm=lambda h: h * 2, denorm=lambda q, h: q + 1, store=lambda k, v: stash.__setitem__(k, v))
    hs = torch.arange(2 * 3 * 64, dtype=torch.float16).reshape(2, 3, 64) / 100
    out = (hs.clone(),)
    hook(None, None, out)
    assert list(stash) == ["original", "normed", "quantized", "indices", "new"]
    assert torch.equal(calls["x"], hs.float() * 2)
    expect = hs.float() * 2 * 0.5 + 1
    expect[:, 0] = hs.float()[:, 0]
    assert torch.equal(out[0], expect.half())
    # stored tensors are clones taken at the time of the call (model.py:278-288): the later denorm / BOS overwrite
    # of q_out must not show in "quantized"
    assert torch.equal(stash["quantized"], hs.float() * 2 * 0.5)
    assert stash["quantized"].data_ptr() != stash["new"].data_ptr()
    # without a store the five clones are skipped
    clones = {"n": 0}
    real_clone = torch.Tensor.clone

    def counting_clone(self, *a, **k):
        clones["n"] += 1
        return real_clone(self, *a, **k)
    monkeypatch.setattr(torch.Tensor, "clone", counting_clone)
    m.hook(norm=lambda h: h, denorm=lambda q, h: q)(None, None, (real_clone(hs),))
    assert clones["n"] == 0
    m.hook(norm=lambda h: h, denorm=lambda q, h: q, store=lambda k, v: None)(None, None, (real_clone(hs),))
    assert clones["n"] == 5
    monkeypatch.setattr(torch.Tensor, "clone", real_clone)
    out2 = (hs.clone(),)
    m.hook(norm=lambda h: h, denorm=lambda q, h: q, replace=False, skip_bos=False)(None, None, out2)
    assert torch.equal(out2[0], hs)


def test_derived_tables_match_reference_formulas():
    torch.manual_seed(1)
    m = RQAE(dim=64, num_quantizers=3)
    cs = m.codebook_sims
    assert cs.shape == (625, 625) and cs.dtype == torch.float16
    assert cs[312].abs().max() == 0               # the zero row normalises to zero (F.normalize eps)
    assert m.layer_norms.shape == (3,)
    assert m.subfeature_sims.shape == (3, 625, 625) and m.subfeature_sims.dtype == torch.float16
    ln = torch.tensor([l[1].weight.data.norm(dim=0).mean().item() for l in m.layers])
    assert torch.equal(m.layer_norms, ln)
    # subfeature_sims comes from the 5x5 Gram of [W_out | b_out], not from the (nq, K, D) subfeatures
    # tensor the reference materialises (model.py:145-167): same table within one fp16 ulp
    import torch.nn.functional as F
    with torch.no_grad():
        n = F.normalize(m.subfeatures, dim=-1)
        ref = (n @ n.transpose(-1, -2)).to(torch.float16)
    assert float((m.subfeature_sims.float() - ref.float()).abs().max()) <= 2 ** -10
    assert float((m.subfeature_sims == ref).float().mean()) > 0.99


def test_make_feature_driver_follows_scripts3(tmp_path, monkeypatch):
    """scripts/3_make_rqae_features.py:164-200 as rqae_b200.feature.make_feature: which tokens become features, their
    centers, batching over features and the saved files -- with the GPU step (get_activations_many) stubbed out."""
    from rqae_b200 import feature as ft
    torch.manual_seed(0)
    m = RQAE(dim=64, num_quantizers=6).eval()
    g = torch.Generator().manual_seed(1)
    tokens = torch.randint(0, 450, (300, 7), generator=g)
    codes = torch.randint(0, 625, (300, 7, 6), generator=g, dtype=torch.int32)
    helper = ft.FeatureHelper.__new__(ft.FeatureHelper)          # the constructor insists on a CUDA store
    helper.tokens, helper.texts, helper.indices, helper.feature_folder = tokens, list(range(300)), codes, str(tmp_path / "features")
    launches = []

    def fake_many(features, layers=None, top_k=100):
        launches.append(len(features))
        return [{l: [{"text": 0, "activations": np.zeros(7, np.float16)}] for l in layers} for _ in features]
    helper.get_activations_many = fake_many
    torch.manual_seed(7)
    feats = ft.make_feature(m, helper, num_tokens=10, layers=[1, 3, 5], top_k=4, features_per_launch=4)
    # the same two draws by hand: unique picks, [200:-200], shuffle, first 10
    torch.manual_seed(7)
    picks = ft.unique_token_indices(tokens)[200:-200]
    picks = picks[torch.randperm(picks.shape[0])][:10]
    assert len(feats) == 10 and launches == [4, 4, 2]
    for i, f in enumerate(feats):
        assert torch.equal(f.center, codes[picks[i, 0], picks[i, 1]]) and f.layers == [1, 3, 5]
        assert f.layer_weights.dtype == torch.float16 and f.rqae is m
        back = ft.RQAEFeature.load(str(tmp_path / "features" / f"{i:06d}.npz"))
        assert torch.equal(back.center, f.center) and list(back.layers) == [1, 3, 5]


@pytest.mark.parametrize("method", ["round_fsq", "fsq"])
def test_fsq_grid_equals_reference_construction_for_every_size(method):
    """model.py:63-72 (numpy linspace ^ codebook_dim in itertools.product order, float64 row norms, zero row kept):
    the oracle restates it with numpy, the module builds it with torch -- equal bit for bit for sizes 2..9 (the
    axis comes from numpy's linspace in both, so any size whose middle value is not exactly 0 there agrees too)."""
    from oracle import rqae_oracle as orc
    from rqae_b200.model import _fsq_grid
    for cbs in range(2, 10):
        want = orc.fsq_codebook(cbs, 4, method == "round_fsq")
        got = _fsq_grid(cbs, 4, method == "round_fsq")
        assert got.dtype == torch.float32 and torch.equal(got, want), cbs


def test_int16_codes_are_range_checked():
    m = RQAE(dim=64, num_quantizers=2, quantization_method="vq", codebook_size=40000).eval()
    with pytest.raises(ValueError, match="int16"):
        m.encode(torch.zeros(1, 1, 64), out_dtype=torch.int16)
