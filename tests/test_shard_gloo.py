"""N>1 plumbing on CPU: two gloo ranks shard a token batch with rqae_b200.shard, each runs the ORACLE on its
shard (standing in for the kernel -- there is no CPU compute path in the product), the codes are gathered
and must equal the single-process result bit for bit."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from rqae_b200 import shard  # noqa: E402


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_tokens, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import c_oracle
        from oracle import rqae_oracle as orc
        w = orc.random_init(dim=256, num_quantizers=12, seed=3)
        cw = c_oracle.CWeights.from_stacked(w)
        x = torch.randn(n_tokens, 256, generator=torch.Generator().manual_seed(5))
        mine = shard.shard_tokens(x, rank, world)
        a, b = shard.token_range(n_tokens, rank, world)
        assert mine.shape[0] == b - a
        _, codes = c_oracle.forward_f32(cw, mine.numpy(), want_q=False, **c_oracle.KERNEL_ORDER)
        local = torch.from_numpy(codes.astype(np.int16))
        full = shard.gather_codes(local, n_tokens)                 # all ranks
        only0 = shard.gather_codes(local, n_tokens, dst=0)         # rank 0 only
        assert (only0 is None) == (rank != 0)
        if rank == 0:
            assert torch.equal(full, only0)
        np.save(os.path.join(out_dir, f"codes_{rank}.npy"), full.numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_tokens", [37, 64])
def test_two_rank_shard_and_gather_equals_single_process(tmp_path, n_tokens):
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n_tokens, str(tmp_path)), nprocs=world, join=True)
    from oracle import c_oracle
    from oracle import rqae_oracle as orc
    w = orc.random_init(dim=256, num_quantizers=12, seed=3)
    cw = c_oracle.CWeights.from_stacked(w)
    x = torch.randn(n_tokens, 256, generator=torch.Generator().manual_seed(5))
    _, ref = c_oracle.forward_f32(cw, x.numpy(), want_q=False, **c_oracle.KERNEL_ORDER)
    for r in range(world):
        got = np.load(os.path.join(str(tmp_path), f"codes_{r}.npy"))
        assert got.shape == ref.shape and np.array_equal(got, ref.astype(np.int16))


def test_token_range_partitions_exactly():
    for n in [0, 1, 7, 16, 1000, 1 << 20]:
        for world in [1, 2, 3, 4, 8]:
            spans = [shard.token_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard.token_range(10, 2, 2)


# ---------------------------------------------------------------------------------------------------
# mining exchange: token-sharded intensities -> feature-sharded rows -> selection over the whole dataset
# ---------------------------------------------------------------------------------------------------
def _oracle_select(rows, k):
    """Stand-in for the CUDA radix select on CPU ranks: stable argsort windows (value desc, index asc)."""
    F, C, n = rows.shape
    idx = torch.full((F, C, 3, k), -1, dtype=torch.int32)
    for f in range(F):
        for c in range(C):
            order = torch.sort(rows[f, c].float(), descending=True, stable=True).indices
            idx[f, c, 0] = order[:k]
            mid = order[n // 2 - k // 2: n // 2 + k // 2]
            idx[f, c, 1, :len(mid)] = mid
            idx[f, c, 2] = order[n - k:]
    return idx, None


def _mine_worker(rank, world, port, n_tokens, n_feat, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        full = (torch.randn(n_feat, 3, n_tokens, generator=torch.Generator().manual_seed(9)) * 0.2).half()
        a, b = shard.token_range(n_tokens, rank, world)
        idx, _, (fa, fb) = shard.mine_sharded(full[:, :, a:b].contiguous(), n_tokens, top_k=5, select_fn=_oracle_select)
        np.save(os.path.join(out_dir, f"mine_{rank}.npy"), idx.numpy())
        np.save(os.path.join(out_dir, f"range_{rank}.npy"), np.array([fa, fb]))
    finally:
        dist.destroy_process_group()


def test_two_rank_mining_exchange_equals_single_process(tmp_path):
    world, n_tokens, n_feat = 2, 101, 7
    port = _free_port()
    mp.spawn(_mine_worker, args=(world, port, n_tokens, n_feat, str(tmp_path)), nprocs=world, join=True)
    full = (torch.randn(n_feat, 3, n_tokens, generator=torch.Generator().manual_seed(9)) * 0.2).half()
    ref, _ = _oracle_select(full, 5)
    got = torch.empty_like(ref)
    seen = 0
    for r in range(world):
        fa, fb = np.load(os.path.join(str(tmp_path), f"range_{r}.npy"))
        got[fa:fb] = torch.from_numpy(np.load(os.path.join(str(tmp_path), f"mine_{r}.npy")))
        seen += fb - fa
    assert seen == n_feat and torch.equal(got, ref)


# ---------------------------------------------------------------------------------------------------
# example search: sequence-sharded store, global ranking (shard.find_examples_sharded)
# ---------------------------------------------------------------------------------------------------
class _OracleEngine:
    """Stand-in for rqae_b200.search.IntensityEngine on CPU ranks: the search ORACLE produces the accumulation and
    the per-position maxima the kernels would (there is no CPU compute path in the product)."""

    def __init__(self, sims, activations):
        self.sims, self.activations = sims, activations

    def _query(self, idx, activation, n_layers):
        q = self.activations[int(idx)] if idx is not None else torch.as_tensor(activation)
        return q[:, :n_layers].to(torch.int32).contiguous()

    def accumulate(self, query, layers):
        from oracle import search_oracle as so
        N, S, nq = self.activations.shape
        for acc in so.accumulate_steps(self.activations.reshape(N * S, nq), self.sims, query, layers):
            a3 = acc.reshape(N, S, -1)
            yield a3, a3.max(dim=1).values.T.contiguous()


def _select_rows(rows, k, n=None):
    idx, _ = _oracle_select(rows.unsqueeze(0), k)
    return idx[0], None


SEARCH_CASE = dict(nq=40, K=27, N=13, S=5, layers=[3, 8, 40], top=4, mid=3, bot=2)


def _search_inputs():
    c = SEARCH_CASE
    g = torch.Generator().manual_seed(13)
    sims = torch.randn(c["nq"], c["K"], c["K"], generator=g).half()
    codes = torch.randint(0, c["K"], (c["N"], c["S"], c["nq"]), generator=g, dtype=torch.int32)
    ext = torch.randint(0, c["K"], (4, c["nq"]), generator=g, dtype=torch.int32)
    return sims, codes, ext


def _search_worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        c = SEARCH_CASE
        sims, codes, ext = _search_inputs()
        a, b = shard.token_range(c["N"], rank, world)
        eng = _OracleEngine(sims, codes[a:b])
        res = {}
        for tag, kw in (("idx", dict(idx=9)), ("ext", dict(activation=ext))):      # sequence 9 lives on rank 1
            for out, layer in shard.find_examples_sharded(eng, c["N"], top_examples=c["top"], middle_examples=c["mid"],
                                                          bottom_examples=c["bot"], layers=c["layers"],
                                                          select_fn=_select_rows, **kw):
                for part in ("top", "middle", "bottom"):
                    res[f"{tag}/{layer}/{part}/indices"] = out[part]["indices"].numpy()
                    res[f"{tag}/{layer}/{part}/intensities"] = out[part]["intensities"].numpy()
        with pytest.raises(ValueError, match="Cannot specify both"):
            next(shard.find_examples_sharded(eng, c["N"], idx=1, activation=ext, layers=c["layers"], select_fn=_select_rows))
        np.savez(os.path.join(out_dir, f"search_{rank}.npz"), **res)
    finally:
        dist.destroy_process_group()


def test_two_rank_example_search_equals_single_process(tmp_path):
    from rqae_b200.search import window_k, window_lists
    world = 2
    port = _free_port()
    mp.spawn(_search_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    c = SEARCH_CASE
    sims, codes, ext = _search_inputs()
    eng = _OracleEngine(sims, codes)
    k = window_k(c["top"], c["mid"], c["bot"], c["N"])
    got = [np.load(os.path.join(str(tmp_path), f"search_{r}.npz")) for r in range(world)]
    for tag, query in (("idx", eng._query(9, None, 40)), ("ext", eng._query(None, ext, 40))):
        qpos = torch.arange(query.shape[0]).unsqueeze(-1)
        for layer, (acc, maxv) in zip(c["layers"], eng.accumulate(query, c["layers"])):
            sel, _ = _select_rows(maxv, k)
            for part, lst in window_lists(sel, c["top"], c["mid"], c["bot"]).items():
                want_i = lst.int().numpy()
                want_v = acc[lst.long(), :, qpos].numpy()
                for r in range(world):
                    assert np.array_equal(got[r][f"{tag}/{layer}/{part}/indices"], want_i), (tag, layer, part, r)
                    assert np.array_equal(got[r][f"{tag}/{layer}/{part}/intensities"], want_v), (tag, layer, part, r)
