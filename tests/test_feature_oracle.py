"""CPU tests of the feature-intensity oracle against outputs of the unmodified reference
(tests/golden/kat_feature.npz, written by tests/golden/make_golden_feature.py) and of the host-side
mirror of rqae/feature.py."""
import os

import numpy as np
import pytest
import torch

from oracle import feature_oracle as fo

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def kat():
    return np.load(os.path.join(ROOT, "tests", "golden", "kat_feature.npz"))


def _case(kat, name, layers_key="layers", out_key="out"):
    t = torch.from_numpy
    return (t(kat[f"{name}/cb0"]), t(kat[f"{name}/lw"]), t(kat[f"{name}/codes"].astype(np.int64)),
            t(kat[f"{name}/centers"]), [int(l) for l in kat[f"{name}/{layers_key}"]], t(kat[f"{name}/{out_key}"]))


@pytest.mark.parametrize("name,lk,ok", [("small", "layers", "out"), ("small", "layers_unsorted", "out_unsorted"),
                                        ("2b", "layers", "out")])
def test_oracle_bit_identical_to_reference(kat, name, lk, ok):
    cb0, lw, codes, centers, layers, ref = _case(kat, name, lk, ok)
    sims = fo.codebook_sims(cb0)
    for f in range(centers.shape[0]):
        assert torch.equal(fo.intensity(sims, centers[f], codes, lw, layers), ref[f])
        assert torch.equal(fo.intensity_steps(sims, centers[f], codes, lw, layers), ref[f])


def test_reference_sits_within_1e3_of_exact_arithmetic(kat):
    """The reference's own fp16 roundings move a value by < 1e-3 from the exact evaluation of the same formula:
    the yardstick for the tolerance of the tensor-core path (tests/test_feature_gpu.py)."""
    cb0, lw, codes, centers, layers, ref = _case(kat, "2b")
    sims = fo.codebook_sims(cb0)
    for f in range(centers.shape[0]):
        exact = fo.intensity_f64(sims, centers[f], codes, lw, layers)
        assert float((exact - ref[f].double()).abs().max()) < 1e-3


def test_sims_table_is_rank4_gram_of_normalised_codebook(kat):
    cb0 = torch.from_numpy(kat["small/cb0"])
    n = fo.normalized_codebook(cb0)
    sims = fo.codebook_sims(cb0)
    assert torch.equal((n @ n.T).half(), sims)
    assert torch.count_nonzero(sims[312]) == 0 and torch.count_nonzero(sims[:, 312]) == 0   # the zero codeword


def test_selection_matches_argsort_definition():
    g = torch.Generator().manual_seed(5)
    v = torch.randn(5000, generator=g).half()
    top, mid, bot = fo.select_top_middle_bottom(v, 100)
    s = torch.sort(v.float(), descending=True).values
    assert torch.equal(v[top].float(), s[:100]) and torch.equal(v[bot].float(), s[-100:])
    assert torch.equal(v[mid].float(), s[2500 - 50:2500 + 50])


def test_host_mirror_constructor_and_errors():
    from rqae_b200 import RQAE, RQAEFeature
    torch.manual_seed(0)
    m = RQAE(dim=64, num_quantizers=16).eval()
    f = RQAEFeature.from_quantizer(m, center=np.arange(16), layers=[3, 15])
    assert f.layer_weights.dtype == torch.float16 and f.layer_weights.shape == (16,)
    assert f.center.dtype == torch.int32 and f.num_quantizers == 16 and f.dim == 4 and f.layers == [3, 15]
    w_out = torch.stack([l[1].weight.data for l in m.layers])
    assert torch.equal(f.layer_weights, fo.layer_weights(w_out))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        f.intensity(torch.zeros(2, 3, 16, dtype=torch.int64))
    g = RQAEFeature(num_quantizers=16)
    with pytest.raises(ValueError, match="Model not loaded"):
        g.intensity(torch.zeros(2, 16, dtype=torch.int64))
    assert g.layers == [15] and g.to_feature(0).id == ""


def test_feature_save_load_roundtrip(tmp_path):
    from rqae_b200 import RQAEFeature
    f = RQAEFeature(id="7", num_quantizers=8, layers=[1, 7], center=np.arange(8), layer_weights=np.ones(8))
    path = str(tmp_path / "f.npz")
    f.save(path)
    g = RQAEFeature.load(path)
    assert g.id == "7" and list(g.layers) == [1, 7] and torch.equal(g.center, f.center)


@pytest.mark.parametrize("f", [0, 1])
@pytest.mark.parametrize("top_k", [7, 100])
def test_mining_selection_bit_identical_to_reference_get_activations(f, top_k):
    """scripts/3_make_rqae_features.py:98-149 run unmodified (tests/golden/make_golden_mining.py): intensities, argsort,
    the three slices, de-duplication by sequence and the per-sequence activation rows."""
    kat = np.load(os.path.join(os.path.dirname(__file__), "golden", "kat_mining.npz"))
    codes = torch.from_numpy(kat["codes"].astype(np.int64))                  # (N, S, nq)
    N, S, _ = codes.shape
    layers = [int(l) for l in kat["layers"]]
    sims = fo.codebook_sims(torch.from_numpy(kat["cb0"]))
    inten = fo.intensity(sims, torch.from_numpy(kat["centers"][f]), codes, torch.from_numpy(kat["lw"]), layers)
    got = fo.get_activations(inten.flatten(0, 1), layers, top_k, S)
    for l in layers:
        seqs, acts = got[l]
        assert seqs == [int(s) for s in kat[f"f{f}/k{top_k}/{l}/sequences"]], l
        want = kat[f"f{f}/k{top_k}/{l}/activations"]
        assert acts.dtype == want.dtype and np.array_equal(acts.view(np.uint16), want.view(np.uint16)), l


def test_unique_token_indices_equal_the_reference_under_the_same_seed():
    """FeatureHelper.get_unique_token_indices (scripts/3:53-82, run unmodified by make_golden_mining.py): one random
    occurrence per distinct token; the mirror draws the same single randperm, so the same seed gives the same picks."""
    from rqae_b200.feature import unique_token_indices
    kat = np.load(os.path.join(os.path.dirname(__file__), "golden", "kat_mining.npz"))
    tokens = torch.from_numpy(kat["unique/tokens"])
    torch.manual_seed(int(kat["unique/seed"][0]))
    got = unique_token_indices(tokens)
    assert got.dtype == torch.int32 and np.array_equal(got.numpy(), kat["unique/indices"])
    # every pick is an occurrence of its token, tokens ascending
    picked = tokens[got[:, 0].long(), got[:, 1].long()]
    assert torch.equal(picked, torch.unique(tokens))
