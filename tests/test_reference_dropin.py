"""The UNMODIFIED reference callers running on top of ``rqae_b200.RQAE`` ("rqae/feature.py and
scripts/3_make_rqae_features.py work unchanged").

Build container only: needs /root/reference (skipped elsewhere).  Nothing here touches a GPU -- the callers read
``codebook_sims``, ``layers[l][1].weight``, ``num_quantizers`` and ``codebook_dim`` of the model object, which are
plain torch on the new class -- and nothing of the reference is edited: ``rqae.feature`` is imported as is,
``scripts/3`` through the stub-``modal`` loader of tests/golden/_ref_loader.py.

What is compared: the reference caller on the NEW class against the same caller on the REFERENCE class, same
weights (same RNG order), same codes.  Equality is exact."""
import os
import sys

import numpy as np
import pytest
import torch

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "rqae")), reason="needs /root/reference")

HERE = os.path.dirname(os.path.abspath(__file__))


def _ref_modules():
    if REF not in sys.path:
        sys.path.insert(0, REF)
    from rqae.model import RQAE as RefRQAE
    from rqae.feature import RQAEFeature as RefFeature
    return RefRQAE, RefFeature


def _pair(dim=64, nq=64, seed=0):
    """(reference model, rqae_b200 model) with identical parameters: same constructor RNG order."""
    from rqae_b200 import RQAE
    RefRQAE, _ = _ref_modules()
    torch.manual_seed(seed)
    ref = RefRQAE(dim=dim, num_quantizers=nq).eval()
    torch.manual_seed(seed)
    new = RQAE(dim=dim, num_quantizers=nq).eval()
    for (k, a), (_, b) in zip(ref.state_dict().items(), new.state_dict().items()):
        assert torch.equal(a, b), k
    return ref, new


def test_reference_state_dict_loads_strict_both_ways():
    ref, new = _pair()
    new.load_state_dict(ref.state_dict(), strict=True)
    ref.load_state_dict(new.state_dict(), strict=True)
    assert list(ref.state_dict().keys()) == list(new.state_dict().keys())


def test_reference_rqaefeature_runs_on_new_class_and_equals_reference_class():
    """rqae/feature.py:95-136: from_quantizer -> load_model (layer weights from layers[l][1].weight) ->
    intensity (gather from codebook_sims)."""
    _, RefFeature = _ref_modules()
    ref, new = _pair()
    g = torch.Generator().manual_seed(5)
    codes = torch.randint(0, 625, (7, 11, 64), generator=g)
    center = codes[3, 4].numpy()
    layers = [2, 4, 6, 8, 12, 16, 24, 32, 48, 63]
    f_ref = RefFeature.from_quantizer(ref, center=center, layers=list(layers))
    f_new = RefFeature.from_quantizer(new, center=center, layers=list(layers))
    assert f_new.num_quantizers == 64 and f_new.dim == 4
    assert torch.equal(f_ref.layer_weights, f_new.layer_weights)
    assert torch.equal(ref.codebook_sims, new.codebook_sims)
    with torch.inference_mode():
        a = f_ref.intensity(codes)
        b = f_new.intensity(codes)
        a2 = f_ref.intensity(codes, layers=[5, 63])
        b2 = f_new.intensity(codes, layers=[5, 63])
    assert a.dtype == torch.float16 and a.shape == (7, 11, len(layers))
    assert torch.equal(a, b) and torch.equal(a2, b2)


def test_reference_scripts3_get_activations_runs_on_new_class():
    """scripts/3_make_rqae_features.py:98-149 (FeatureHelper.get_activations, imported unmodified with a stub
    ``modal``) driving the reference's RQAEFeature built from the NEW model class: same selected sequences in the
    same order and the same activation rows as with the reference model class."""
    sys.path.insert(0, os.path.join(HERE, "golden"))
    from _ref_loader import load_reference_file
    s3 = load_reference_file("scripts/3_make_rqae_features.py", "ref_scripts3_dropin")
    _, RefFeature = _ref_modules()
    ref, new = _pair()
    g = torch.Generator().manual_seed(21)
    N, S, nq = 1100, 3, 64                      # two rounds of the 1024-sequence batching (scripts/3:104-108)
    codes = torch.randint(0, 625, (N, S, nq), generator=g)
    layers = [2, 8, 16, 63]

    def run(model):
        helper = s3.FeatureHelper.__new__(s3.FeatureHelper)
        helper.tokens = torch.zeros(N, S, dtype=torch.int64)
        helper.texts = list(range(N))
        helper.indices = codes
        feat = RefFeature.from_quantizer(model, center=codes[17, 1].numpy(), layers=list(layers))
        with torch.inference_mode():
            return helper.get_activations(feat, top_k=7)

    a, b = run(ref), run(new)
    assert list(a.keys()) == list(b.keys()) == layers
    for l in layers:
        assert [r["text"] for r in a[l]] == [r["text"] for r in b[l]]
        for ra, rb in zip(a[l], b[l]):
            assert np.array_equal(ra["activations"], rb["activations"])


def test_reference_server_tables_from_new_class():
    """demo/server/server.py:101-115 builds its engine table from subfeature_sims and layer_norms of the model:
    the new class's tables (5x5 Gram form, no (nq, K, D) intermediate) against the reference's own properties."""
    ref, new = _pair(dim=96, nq=12, seed=3)
    with torch.inference_mode():
        s_ref = ref.subfeature_sims
        s_new = new.subfeature_sims
        n_ref, n_new = ref.layer_norms, new.layer_norms
    assert s_new.shape == s_ref.shape == (12, 625, 625) and s_new.dtype == torch.float16
    assert torch.equal(n_ref, n_new)
    # one fp16 ulp at magnitude <= 1 is 2^-11 (the two forms sum in a different order)
    assert float((s_ref.float() - s_new.float()).abs().max()) <= 2.0 ** -10
    assert float((s_ref != s_new).float().mean()) < 0.01
