"""The search oracle (oracle/search_oracle.py) against outputs of the unmodified reference server code
(tests/golden/kat_search.npz; demo/server/server.py:159-325 run by tests/golden/make_golden_search.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import search_oracle as so

GOLD = os.path.join(os.path.dirname(__file__), "golden", "kat_search.npz")
CASES = [("k81", "idx5"), ("k81", "ext"), ("k625", "idx3")]


def load_case(kat, case, tag):
    sims = torch.from_numpy(kat[f"{case}/sims"])
    shards = [torch.from_numpy(s.astype(np.int32)) for s in kat[f"{case}/shards"]]
    if tag.startswith("idx"):
        i = int(tag[3:])
        n_per = shards[0].shape[0]
        query = shards[i // n_per][i % n_per]
    else:
        query = torch.from_numpy(kat[f"{case}/{tag}"].astype(np.int32))
    layers = [int(v) for v in kat[f"{case}/{tag}/layers"]]
    top, mid, bot = (int(v) for v in kat[f"{case}/args"])
    return sims, shards, query, layers, (top, mid, bot)


@pytest.fixture(scope="module")
def kat():
    return np.load(GOLD)


@pytest.mark.parametrize("case,tag", CASES)
def test_find_examples_bit_identical_to_reference(kat, case, tag):
    sims, shards, query, layers, (top, mid, bot) = load_case(kat, case, tag)
    n = 0
    for res, layer in so.find_examples(shards, sims, query, top, mid, bot, layers):
        for part in ("top", "middle", "bottom"):
            assert np.array_equal(res[part]["indices"].numpy(), kat[f"{case}/{tag}/{layer}/{part}/indices"]), (layer, part)
            a = res[part]["intensities"].numpy()
            b = kat[f"{case}/{tag}/{layer}/{part}/intensities"]
            assert a.dtype == np.float16 and np.array_equal(a.view(np.uint16), b.view(np.uint16)), (layer, part)
        n += 1
    assert n == len(layers)


@pytest.mark.parametrize("case,tag", CASES)
def test_written_out_roundings_match_the_aten_form(kat, case, tag):
    """accumulate_steps (the kernel's order of operations) against the ATen form: identical except where the
    order of the fp32 additions inside a chunk moves a sum across an fp16 rounding boundary."""
    sims, shards, query, layers, _ = load_case(kat, case, tag)
    codes = torch.cat(shards).reshape(-1, shards[0].shape[-1])
    worst_ulps, differing, total = 0, 0, 0
    for ref, mine in zip(so.accumulate(shards, sims, query, layers), so.accumulate_steps(codes, sims, query, layers)):
        r = ref.reshape(-1, ref.shape[-1])
        d = (r.view(torch.int16).int() - mine.view(torch.int16).int()).abs()
        same_sign = (r.float() * mine.float()) >= 0
        assert bool(same_sign.all())
        worst_ulps = max(worst_ulps, int(d.max()))
        differing += int((d != 0).sum())
        total += d.numel()
    assert worst_ulps <= 2, worst_ulps                 # fp16 steps, accumulated over the ranges
    assert differing <= 0.01 * total, (differing, total)


def test_layer_ranges_and_chunks():
    assert so.layer_ranges([4, 6, 150]) == [(0, 4), (4, 6), (6, 150)]
    # the chunked branch (server.py:216-234): a range of 144 layers is 64 + 64 + 16
    sims = torch.randn(150, 9, 9).half()
    q = torch.randint(0, 9, (3, 150), dtype=torch.int32)
    sh = torch.randint(0, 9, (2, 4, 150), dtype=torch.int32)
    qt = so.query_table(sims, q, 150)
    whole = so.range_intensities(sh, qt, 6, 150)
    parts = [so.get_intensities(sh[..., a:b], qt[a:b].transpose(0, 1)).sum(-1) for a, b in [(6, 70), (70, 134), (134, 150)]]
    manual = parts[0].clone()
    manual += parts[1]
    manual += parts[2]
    assert torch.equal(whole, manual)


def test_rank5_form_of_the_table_and_its_distance_from_the_reference():
    """Groundwork for the tensor-core variant: the engine table is a rank-5 Gram matrix per layer, and
    a search computed from fp16 factors lands within a few fp16 steps of the reference's table-gather form -- it cannot
    be bit-exact, because the reference sums fp16-ROUNDED table entries."""
    from rqae_b200 import RQAE
    from rqae_b200.search import engine_sims
    torch.manual_seed(3)
    m = RQAE(dim=96, num_quantizers=160).eval()                       # host-side tables only, no kernel
    sims = engine_sims(m)                                             # (nq, K, K) fp16 = subfeature_sims * layer_norms
    w = torch.stack([l[1].weight.detach() for l in m.layers])
    b = torch.stack([l[1].bias.detach() for l in m.layers])
    fac = so.rank5_factors(w, b, m.codebook.detach())
    # 1. the factorisation reproduces the table up to its own fp16 roundings
    exact = (fac @ fac.transpose(1, 2)) * m.layer_norms.double().reshape(-1, 1, 1)
    assert float((exact - sims.double()).abs().max()) <= 2 ** -9 * float(m.layer_norms.max())
    # 2. accumulated over the server's ranges: distance in fp16 steps of the running value
    g = torch.Generator().manual_seed(4)
    codes = torch.randint(0, 625, (6 * 9, 160), generator=g, dtype=torch.int32)
    query = codes[:9].clone()
    layers = [4, 6, 8, 12, 16, 24, 32, 48, 64, 150]

    def steps(t):
        i = t.contiguous().view(torch.int16).int()
        return torch.where(i < 0, -(i & 0x7FFF), i)
    for passes, bar in ((1, 4), (3, 3)):
        worst = 0
        for ref, alt in zip(so.accumulate_steps(codes, sims, query, layers),
                            so.accumulate_rank5(codes, fac, m.layer_norms, query, layers, passes=passes)):
            worst = max(worst, int((steps(ref) - steps(alt)).abs().max()))
        assert worst <= bar, (passes, worst)


def test_engine_table_modes_follow_the_server_setup():
    """server.py:103-115: 'projected' = subfeature_sims * layer_norms, 'original' = codebook_sims repeated per layer
    * layer_norms, both multiplied IN fp16; anything else is the reference's ValueError."""
    from rqae_b200 import RQAE
    from rqae_b200.search import engine_sims
    torch.manual_seed(5)
    m = RQAE(dim=48, num_quantizers=5).eval()
    ln = m.layer_norms
    proj = engine_sims(m)
    assert proj.dtype == torch.float16 and proj.shape == (5, 625, 625)
    assert torch.equal(proj, so.scaled_sims(m.subfeature_sims, ln))
    assert torch.equal(m.subfeature_sims, so.scaled_sims(m.subfeature_sims, torch.ones(5)))   # the cached table is not scaled in place
    orig = engine_sims(m, "original")
    want = m.codebook_sims.unsqueeze(0).repeat(5, 1, 1)
    want *= ln.unsqueeze(-1).unsqueeze(-1)
    assert torch.equal(orig, want)
    with pytest.raises(ValueError, match="Invalid mode"):
        engine_sims(m, "other")


@pytest.mark.parametrize("mode", ["projected", "original"])
def test_product_rank_factors_reproduce_the_engine_table(mode):
    """rqae_b200.search.rank_factors (what the tensor-core ranking builds its fp16 tables from) against the engine
    table of the same mode (server.py:103-115) and, for 'projected', against the oracle's factorisation: the Gram
    matrices agree (the factors themselves are only defined up to a rotation)."""
    from rqae_b200 import RQAE
    from rqae_b200.search import engine_sims, rank_factors
    torch.manual_seed(7)
    m = RQAE(dim=80, num_quantizers=12).eval()
    f = rank_factors(m, mode)                                         # (nq, K, R) float64
    assert f.dtype == torch.float64 and f.shape[:2] == (12, 625) and f.shape[2] == (5 if mode == "projected" else 4)
    gram = f @ f.transpose(1, 2)
    table = engine_sims(m, mode).double()
    scaled = gram * m.layer_norms.double().reshape(-1, 1, 1)
    assert float((scaled - table).abs().max()) <= 2 ** -9 * float(m.layer_norms.max())      # the table's own fp16 roundings
    if mode == "projected":
        w = torch.stack([l[1].weight.detach() for l in m.layers])
        b = torch.stack([l[1].bias.detach() for l in m.layers])
        fo_ = so.rank5_factors(w, b, m.codebook.detach())
        assert float((gram - fo_ @ fo_.transpose(1, 2)).abs().max()) <= 1e-12
    with pytest.raises(ValueError, match="Invalid mode"):
        rank_factors(m, "other")
