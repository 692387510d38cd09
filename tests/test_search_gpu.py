"""GPU parity tests of the nearest-example search (rq_search.cuh through the C ABI via rqae_b200.search)
against outputs of the unmodified reference server code (tests/golden/kat_search.npz) and the oracle.

Bars: the running accumulation is BIT-EXACT against the oracle's written-out form (same fp16 table values, fp32
sum in ascending layer order inside a chunk of <= 64 layers, the reference's fp16 roundings) and within
ULPS fp16 steps of the reference's own ATen result (whose fp32 summation order inside `sum` is torch's; on the
goldens the two are identical).  The selected sequences equal a stable descending argsort of the per-position
maxima index for index; against the reference's (unstable) argsort they are compared through the selected VALUES,
which are order-independent."""
import os

import numpy as np
import pytest
import torch

from oracle import search_oracle as so
from tests.test_search_oracle import CASES, GOLD, load_case

pytestmark = pytest.mark.gpu
ULPS = 2


def _dev():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def kat():
    return np.load(GOLD)


def _ulps(a: torch.Tensor, b: torch.Tensor) -> int:
    """largest distance in fp16 steps between two fp16 tensors (+0 and -0 coincide)."""
    def key(t):
        i = t.contiguous().view(torch.int16).int()
        return torch.where(i < 0, -(i & 0x7FFF), i)
    return int((key(a) - key(b)).abs().max()) if a.numel() else 0


def _engine(sims, shards, dtype=torch.int32):
    from rqae_b200.search import IntensityEngine
    dev = _dev()
    return IntensityEngine(sims=sims.to(dev), activations=[s.to(dev).to(dtype) for s in shards])


@pytest.mark.parametrize("case,tag", CASES)
def test_accumulation_bit_exact_vs_oracle_and_reference_form(kat, case, tag):
    sims, shards, query, layers, _ = load_case(kat, case, tag)
    eng = _engine(sims, shards)
    q = eng._query(None, query, max(layers))
    codes = torch.cat(shards).reshape(-1, shards[0].shape[-1])
    N, S = eng.activations.shape[:2]
    steps = so.accumulate_steps(codes, sims, query, layers)
    aten = so.accumulate(shards, sims, query, layers)
    for (acc, maxv), want, ref in zip(eng.accumulate(q, layers), steps, aten):
        got = acc.cpu()
        assert torch.equal(got.reshape(N * S, -1).float(), want.float())
        assert _ulps(got, ref) <= ULPS
        assert torch.equal(maxv.cpu().float(), want.reshape(N, S, -1).max(dim=1).values.T.float())


@pytest.mark.parametrize("case,tag", CASES)
def test_find_examples_matches_reference_golden(kat, case, tag):
    sims, shards, query, layers, (top, mid, bot) = load_case(kat, case, tag)
    eng = _engine(sims, shards, torch.int16)
    kw = dict(idx=int(tag[3:])) if tag.startswith("idx") else dict(activation=query)
    accs = so.accumulate(shards, sims, query, layers)
    n = 0
    for (res, layer), acc in zip(eng.find_examples(top_examples=top, middle_examples=mid, bottom_examples=bot,
                                                   layers=layers, **kw), accs):
        maxv = acc.max(dim=1).values                                   # (N, Sq), the reference's max_values
        order = torch.sort(maxv.float(), dim=0, descending=True, stable=True).indices
        N = order.shape[0]
        want = {"top": order[:top].T, "middle": order[N // 2 - mid // 2: N // 2 + mid // 2].T, "bottom": order[-bot:].T}
        for part in ("top", "middle", "bottom"):
            idx = res[part]["indices"]
            gold_idx = torch.from_numpy(kat[f"{case}/{tag}/{layer}/{part}/indices"])
            gold_int = torch.from_numpy(kat[f"{case}/{tag}/{layer}/{part}/intensities"])
            assert idx.dtype == torch.int32 and idx.shape == gold_idx.shape
            assert res[part]["intensities"].dtype == torch.float16 and res[part]["intensities"].shape == gold_int.shape
            assert torch.equal(idx.long(), want[part]), (layer, part)                 # stable order, index for index
            # against the reference's own selection: the same values at every rank (ties may name other sequences)
            qpos = torch.arange(idx.shape[0]).unsqueeze(-1)
            assert torch.equal(maxv[idx.long(), qpos].float(), maxv[gold_idx.long(), qpos].float()), (layer, part)
            # intensities of the selected sequences = the reference's accumulation rows
            assert torch.equal(res[part]["intensities"].float(), acc[idx.long(), :, qpos].float()), (layer, part)
            same = (idx == gold_idx).all(dim=1)
            assert torch.equal(res[part]["intensities"][same].float(), gold_int[same].float())
        n += 1
    assert n == len(layers)


@pytest.mark.parametrize("dtype", [torch.int16, torch.int32, torch.int64])
def test_random_table_ragged_sizes_and_code_dtypes(dtype):
    """K = 625, 127 query positions, a token count that is no multiple of the 64-token tile, ranges of 1, 63, 64,
    65 and 130 layers (chunk boundaries, server.py:216-234), a few codes outside [0, K) (contribute 0)."""
    g = torch.Generator().manual_seed(5)
    nq, K, N, S, Sq = 330, 625, 13, 21, 127
    sims = (torch.randn(nq, K, K, generator=g) * 0.7).half()
    codes = torch.randint(0, K, (N, S, nq), generator=g, dtype=torch.int32)
    codes[3, 5, 7] = -1
    codes[4, 0, 100] = K
    query = torch.randint(0, K, (Sq, nq), generator=g, dtype=torch.int32)
    query[9, 2] = K + 3                                                # a zero row of the query table
    layers = [1, 64, 128, 193, 323]
    eng = _engine(sims, [codes], dtype)
    q = eng._query(None, query, max(layers))
    # oracle with the out-of-range codes mapped to an appended all-zero row / column
    sims_z = torch.zeros(nq, K + 1, K + 1, dtype=torch.float16)
    sims_z[:, :K, :K] = sims
    codes_z = codes.clone().reshape(-1, nq)
    codes_z[(codes_z < 0) | (codes_z >= K)] = K
    query_z = query.clone()
    query_z[(query_z < 0) | (query_z >= K)] = K
    for (acc, maxv), want in zip(eng.accumulate(q, layers), so.accumulate_steps(codes_z, sims_z, query_z, layers)):
        assert torch.equal(acc.cpu().reshape(N * S, Sq).float(), want.float())
        assert torch.equal(maxv.cpu().float(), want.reshape(N, S, Sq).max(dim=1).values.T.float())


def test_single_query_position_and_full_128():
    g = torch.Generator().manual_seed(6)
    nq, K, N, S = 40, 81, 70, 9
    sims = torch.randn(nq, K, K, generator=g).half()
    codes = torch.randint(0, K, (N, S, nq), generator=g, dtype=torch.int32)
    eng = _engine(sims, [codes], torch.int16)
    for Sq in (1, 128):
        query = torch.randint(0, K, (Sq, nq), generator=g, dtype=torch.int32)
        layers = [3, 40]
        got = list(eng.find_examples(activation=query, top_examples=7, middle_examples=5, bottom_examples=2, layers=layers))
        for (res, layer), want in zip(got, so.accumulate_steps(codes.reshape(-1, nq), sims, query, layers)):
            maxv = want.reshape(N, S, Sq).max(dim=1).values
            order = torch.sort(maxv.float(), dim=0, descending=True, stable=True).indices
            assert torch.equal(res["top"]["indices"].long(), order[:7].T)
            assert torch.equal(res["middle"]["indices"].long(), order[N // 2 - 2: N // 2 + 2].T)   # 2 * (5 // 2) entries
            assert torch.equal(res["bottom"]["indices"].long(), order[-2:].T)
            assert res["top"]["intensities"].shape == (Sq, 7, S)


def test_argument_errors_are_the_reference_s():
    g = torch.Generator().manual_seed(7)
    sims = torch.randn(8, 9, 9, generator=g).half()
    codes = torch.randint(0, 9, (6, 4, 8), generator=g, dtype=torch.int32)
    eng = _engine(sims, [codes])
    with pytest.raises(ValueError, match="Cannot specify both idx and activation"):      # server.py:174-175
        next(eng.find_examples(idx=1, activation=codes[0], layers=[2, 4]))
    with pytest.raises(ValueError, match="Must specify either idx or activation"):        # server.py:181-182
        next(eng.find_examples(layers=[2, 4]))
    with pytest.raises(ValueError):
        next(eng.find_examples(idx=1, layers=[4, 2]))
    with pytest.raises(ValueError):
        next(eng.find_examples(idx=1, layers=[2, 9]))                                     # deeper than the store
    from rqae_b200.search import IntensityEngine
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        IntensityEngine(sims=sims, activations=codes)


def test_engine_from_model_and_store(tmp_path):
    """End to end on a real (random-init) model: codes from RQAE.encode, the store's file format, the engine's table
    from the model's own derived tables (server.py:103-115), query by idx; checked against the oracle run on the same
    table."""
    from rqae_b200 import RQAE, store
    from rqae_b200.search import IntensityEngine
    dev = _dev()
    torch.manual_seed(0)
    m = RQAE(dim=256, num_quantizers=24, name="t").eval().to(dev)
    x = torch.randn(12, 8, 256, generator=torch.Generator().manual_seed(2)).to(dev)
    codes = m.encode(x, out_dtype=torch.int32)                         # (12, 8, 24), BOS position included
    store.save_code_shard(str(tmp_path), m.name, 0, codes[:6])
    store.save_code_shard(str(tmp_path), m.name, 1, codes[6:])
    eng = IntensityEngine.from_store(m, str(tmp_path))
    assert tuple(eng.activations.shape) == (12, 7, 24) and eng.activations.dtype == torch.int16
    layers = [4, 6, 8, 12, 23]
    sims_cpu, act_cpu = eng.sims.cpu(), eng.activations.cpu().int()
    want = so.find_examples([act_cpu], sims_cpu, act_cpu[5], 4, 2, 2, layers)
    for (res, layer), (ref, rlayer) in zip(eng.find_examples(idx=5, top_examples=4, middle_examples=2, bottom_examples=2,
                                                             layers=layers), want):
        assert layer == rlayer
        # the query sequence itself is its own best match at every position's own slot
        for part in ("top", "middle", "bottom"):
            assert res[part]["indices"].shape == ref[part]["indices"].shape
            same = (res[part]["indices"] == ref[part]["indices"]).all(dim=1)
            assert _ulps(res[part]["intensities"][same], ref[part]["intensities"][same]) <= ULPS


# ---------------------------------------------------------------------------------------------------------------
# opt-in tensor-core ranking (precision="tc"): stated tolerance against the exact mode, exact rows for the selected
# ---------------------------------------------------------------------------------------------------------------
def _fp16_step(x: float) -> float:
    import math
    return 2.0 ** (max(math.floor(math.log2(max(abs(x), 6.2e-5))), -14) - 10)


@pytest.mark.parametrize("mode", ["projected", "original"])
@pytest.mark.parametrize("n_seq,S,Sq", [(9, 21, 13), (32, 127, 127)])
def test_tensor_core_maxima_within_stated_tolerance_of_exact_mode(mode, n_seq, S, Sq):
    """The tensor-core maxima (rank-5 factors in fp16, fp32 running prefix) against the exact mode's (fp16 table entries,
    the reference's chunk / range roundings): at most TC_STEPS fp16 steps of the largest value of the cut -- the distance
    is the reference's own rounding of its running value, not the factor rounding (oracle: test_search_oracle rank-5 test)."""
    from rqae_b200 import RQAE
    from rqae_b200.search import IntensityEngine
    TC_STEPS = 4
    dev = _dev()
    torch.manual_seed(3)
    nq = 150
    model = RQAE(dim=256, num_quantizers=nq).eval().to(dev)
    K = model.codebook.shape[1]
    g = torch.Generator().manual_seed(11)
    codes = torch.randint(0, K, (n_seq, S, nq), generator=g, dtype=torch.int32)
    codes[1, 2, 3] = -1                                                # contributes 0 in both modes
    query = torch.randint(0, K, (Sq, nq), generator=g, dtype=torch.int32)
    layers = [4, 6, 8, 12, 16, 24, 32, 48, 64, 128, 150]
    exact = IntensityEngine(model, codes.to(dev).to(torch.int16), mode=mode)
    tc = IntensityEngine(model, codes.to(dev).to(torch.int16), mode=mode, precision="tc")
    q = exact._query(None, query, max(layers))
    got = tc.maxima_tc(q, layers).float().cpu()                        # (cuts, Sq, N)
    worst = 0.0
    for ci, (acc, maxv) in enumerate(exact.accumulate(q, layers)):
        want = maxv.float().cpu()
        step = _fp16_step(float(want.abs().max()))
        d = float((got[ci] - want).abs().max())
        worst = max(worst, d / step)
        assert d <= TC_STEPS * step, (mode, layers[ci], d, step)
    print(f"tensor-core maxima vs exact ({mode}, {n_seq} x {S}, {Sq} query positions): worst {worst:.2f} fp16 steps of the cut's largest value")


@pytest.mark.parametrize("dtype", [torch.int16, torch.int32, torch.int64])
def test_tensor_core_find_examples_reports_exact_rows_and_near_equal_ranks(dtype):
    """precision="tc": the reported intensities are the exact accumulation rows of the reported sequences (bit for bit),
    and at every rank the reported sequence's exact maximum is within the stated tolerance of the exact mode's value."""
    from rqae_b200 import RQAE
    from rqae_b200.search import IntensityEngine
    dev = _dev()
    torch.manual_seed(4)
    nq, n_seq, S = 96, 41, 33
    Sq = S                                                             # the query is sequence 5 of the store
    model = RQAE(dim=256, num_quantizers=nq).eval().to(dev)
    K = model.codebook.shape[1]
    g = torch.Generator().manual_seed(12)
    codes = torch.randint(0, K, (n_seq, S, nq), generator=g, dtype=torch.int32).to(dev).to(dtype)   # int16: vector code loads in the rows kernel
    codes[3, 2, 5] = -1                                                # contributes 0 in both modes and in the exact rows
    layers = [4, 8, 16, 64, 96]
    exact = IntensityEngine(model, codes)
    tc = IntensityEngine(model, codes, precision="tc")
    res_e = list(exact.find_examples(idx=5, top_examples=6, middle_examples=4, bottom_examples=3, layers=layers))
    res_t = list(tc.find_examples(idx=5, top_examples=6, middle_examples=4, bottom_examples=3, layers=layers))
    q = exact._query(5, None, max(layers))
    accs = [(a.clone().cpu(), m.clone().cpu()) for a, m in exact.accumulate(q, layers)]
    assert [l for _, l in res_t] == layers
    qpos = torch.arange(Sq).unsqueeze(-1)
    for (rt, _), (re_, _), (acc, maxv) in zip(res_t, res_e, accs):
        step = _fp16_step(float(maxv.float().abs().max()))
        for part in ("top", "middle", "bottom"):
            it, ie = rt[part]["indices"].long(), re_[part]["indices"].long()
            assert it.shape == ie.shape and rt[part]["intensities"].shape == re_[part]["intensities"].shape
            assert torch.equal(rt[part]["intensities"].float(), acc[it, :, qpos].float()), part      # exact rows of ITS sequences
            vt, ve = maxv.T[it, qpos].float(), maxv.T[ie, qpos].float()                              # exact maxima at every rank
            assert float((vt - ve).abs().max()) <= 8 * step, part
    # the query sequence itself tops every position at the last cut in both modes
    assert bool((res_t[-1][0]["top"]["indices"][:, 0] == 5).all()) and bool((res_e[-1][0]["top"]["indices"][:, 0] == 5).all())


@pytest.mark.parametrize("n_seq,S,Sq,layers", [(1, 5, 3, [3]), (2, 128, 1, [9, 10]), (5, 1, 128, [1, 2, 40])])
def test_tensor_core_ranking_edge_shapes(n_seq, S, Sq, layers):
    """One sequence / one position / 128 positions per sequence (no padding column) / 128 query positions / ranges that
    start and end inside an 8-layer block: maxima within the stated tolerance of the exact mode, rows bit-equal."""
    from rqae_b200 import RQAE
    from rqae_b200.search import IntensityEngine
    dev = _dev()
    torch.manual_seed(9)
    nq = 48
    model = RQAE(dim=128, num_quantizers=nq).eval().to(dev)
    K = model.codebook.shape[1]
    g = torch.Generator().manual_seed(13)
    codes = torch.randint(0, K, (n_seq, S, nq), generator=g, dtype=torch.int32).to(dev).to(torch.int16)
    query = torch.randint(0, K, (Sq, nq), generator=g, dtype=torch.int32)
    exact = IntensityEngine(model, codes)
    tc = IntensityEngine(model, codes, precision="tc")
    q = exact._query(None, query, max(layers))
    got = tc.maxima_tc(q, layers).float().cpu()
    accs = [(a.clone(), m.clone()) for a, m in exact.accumulate(q, layers)]
    for ci, (acc, maxv) in enumerate(accs):
        want = maxv.float().cpu()
        assert got[ci].shape == want.shape
        assert float((got[ci] - want).abs().max()) <= 4 * _fp16_step(float(want.abs().max()))
    sel = torch.arange(n_seq, dtype=torch.int32, device=dev).repeat(Sq, 1)               # every sequence for every position
    rows = tc.rows_exact(tc._build_qrows(q, max(layers)), sel, layers).cpu()                # (Sq, n_seq, S) after the last range
    acc = accs[-1][0].cpu()                                                                  # (N, S, Sq)
    assert torch.equal(rows.float(), acc.permute(2, 0, 1).float())
