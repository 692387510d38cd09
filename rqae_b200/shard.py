"""Token sharding of the RQ hot path across the GPUs of one box (one process per GPU).

The reference has no multi-GPU code: it fans shards of 1024 sequences out to separate containers and
meets again on a network volume (scripts/1_create_activations.py:283-305).  Tokens are independent in
``RQAE.forward`` (rqae/model.py:199-224 has no cross-token op), so the B200 build shards the flattened
token axis into contiguous ranges, replicates the 85 MB of weights on every GPU, and needs NO collective on
the data path.  The only exchange is the optional gather of the code tensors when a caller wants them in
one place; it runs over ``torch.distributed`` (NCCL on the GPUs, gloo in the CPU tests).

Because the kernel's arithmetic does not depend on where in a tile, chunk or grid a token lands
(rq_forward.cuh: fixed summation order), the concatenation of the per-rank results is bit-identical to
the single-GPU result (tests/test_shard_gloo.py checks the plumbing on CPU with the oracle standing in
for the kernel; tests/test_parity_gpu.py checks the position invariance of the kernel itself).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist


def token_range(n_tokens: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced range [start, stop) of the flattened token axis owned by ``rank``."""
    if world <= 0 or not (0 <= rank < world) or n_tokens < 0:
        raise ValueError(f"bad shard request: n_tokens={n_tokens} rank={rank} world={world}")
    return (n_tokens * rank) // world, (n_tokens * (rank + 1)) // world


def shard_tokens(x: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    """The rows of ``x`` (any leading shape, last dim = model dim) that ``rank`` processes, as a (n, D) view."""
    flat = x.reshape(-1, x.shape[-1])
    a, b = token_range(flat.shape[0], rank, world)
    return flat[a:b]


def gather_codes(codes_local: torch.Tensor, n_tokens: int, group: Optional[dist.ProcessGroup] = None,
                 dst: Optional[int] = None) -> Optional[torch.Tensor]:
    """Assemble the (n_tokens, nq) code tensor from the per-rank shards produced with ``token_range``.

    ``codes_local`` is this rank's (n_local, nq) tensor (int16/int32/int64, on the device the process group
    communicates on).  With ``dst=None`` every rank receives the full tensor (all_gather); otherwise only
    rank ``dst`` does and the others get None.  Shards may differ by one row; they are padded to the largest
    for the collective and trimmed afterwards."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    a, b = token_range(n_tokens, rank, world)
    if codes_local.dim() != 2 or codes_local.shape[0] != b - a:
        raise ValueError(f"rank {rank} must pass its ({b - a}, nq) shard, got {tuple(codes_local.shape)}")
    nq = codes_local.shape[1]
    sizes = [token_range(n_tokens, r, world)[1] - token_range(n_tokens, r, world)[0] for r in range(world)]
    m = max(sizes) if sizes else 0
    send = codes_local
    if codes_local.shape[0] != m:
        send = torch.zeros(m, nq, dtype=codes_local.dtype, device=codes_local.device)
        send[: codes_local.shape[0]] = codes_local
    # codes travel as raw bytes: int16 is not a collective dtype of every backend (gloo rejects it)
    send = send.contiguous().view(torch.uint8)
    if dst is None:
        bufs = [torch.empty_like(send) for _ in range(world)]
        dist.all_gather(bufs, send, group=group)
    else:
        # `dst` is a rank of `group`; dist.gather wants the global rank
        bufs = [torch.empty_like(send) for _ in range(world)] if rank == dst else None
        dist.gather(send, bufs, dst=dist.get_global_rank(group, dst) if group is not None else dst, group=group)
        if rank != dst:
            return None
    return torch.cat([buf.view(codes_local.dtype)[:n] for buf, n in zip(bufs, sizes)], dim=0)


def encode_sharded(model, x: torch.Tensor, max_layers=float("inf"), out_dtype: torch.dtype = torch.int16,
                   gather: bool = False, group: Optional[dist.ProcessGroup] = None):
    """Encode this rank's contiguous share of ``x`` (identical on every rank, or at least identically shaped)
    on the current CUDA device; optionally all_gather the codes.  Returns (codes, (start, stop))."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    flat = x.reshape(-1, x.shape[-1])
    a, b = token_range(flat.shape[0], rank, world)
    codes = model.encode(flat[a:b].unsqueeze(0), max_layers=max_layers, out_dtype=out_dtype)[0]
    if gather and world > 1:
        codes = gather_codes(codes, flat.shape[0], group=group)
        return codes, (0, flat.shape[0])
    return codes, (a, b)


# ---------------------------------------------------------------------------------------------------
# feature mining across GPUs (BASELINE configs[4]: encode + per-feature intensities on 8 GPUs)
# ---------------------------------------------------------------------------------------------------
def exchange_to_feature_shards(intens_local: torch.Tensor, n_tokens: int,
                               group: Optional[dist.ProcessGroup] = None) -> Tuple[torch.Tensor, Tuple[int, int]]:
    """Token-sharded intensities -> feature-sharded intensities, the one exchange step of the mining path.

    Every rank holds ``intens_local`` (F, C, T_r) fp16 for its contiguous token range (``token_range``) and
    ALL F features -- what ``rqae_b200.feature.intensity_many`` returns for the rank's codes.  The selection
    of scripts/3_make_rqae_features.py:116-128 ranks each (feature, cut) row over the WHOLE dataset, and the
    median window does not compose from per-shard results, so rows are made whole instead: rank d receives
    features ``token_range(F, d, world)`` from everybody (one all_to_all, NVLink/NVSwitch under NCCL) and
    concatenates the pieces in rank order = global token order.  Returns ((F_d, C, n_tokens) with rows padded
    to a multiple of 8 tokens, (f_start, f_stop))."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    F, C, T_r = intens_local.shape
    a, b = token_range(n_tokens, rank, world)
    if T_r != b - a:
        raise ValueError(f"rank {rank} must pass its {b - a} tokens, got {T_r}")
    fa, fb = token_range(F, rank, world)
    sizes = [token_range(n_tokens, s, world)[1] - token_range(n_tokens, s, world)[0] for s in range(world)]
    esz = intens_local.element_size()
    in_splits = [(token_range(F, d, world)[1] - token_range(F, d, world)[0]) * C * T_r * esz for d in range(world)]
    if intens_local.is_contiguous():   # the feature ranges are consecutive slabs of the buffer: send it as it is
        send_flat = intens_local.view(torch.uint8).reshape(-1)
    else:
        send_flat = torch.cat([intens_local[token_range(F, d, world)[0]:token_range(F, d, world)[1]].contiguous()
                               .view(torch.uint8).reshape(-1) for d in range(world)])
    out_splits = [(fb - fa) * C * sizes[s] * esz for s in range(world)]
    recv_flat = torch.empty(sum(out_splits), dtype=torch.uint8, device=intens_local.device)
    dist.all_to_all_single(recv_flat, send_flat, output_split_sizes=out_splits, input_split_sizes=in_splits, group=group)
    recv = list(torch.split(recv_flat, out_splits))
    stride = (n_tokens + 7) // 8 * 8
    full = torch.zeros(fb - fa, C, stride, dtype=intens_local.dtype, device=intens_local.device)
    t0 = 0
    for s in range(world):
        full[:, :, t0:t0 + sizes[s]] = recv[s].view(intens_local.dtype).reshape(fb - fa, C, sizes[s])
        t0 += sizes[s]
    return full[:, :, :n_tokens], (fa, fb)


def mine_sharded(intens_local: torch.Tensor, n_tokens: int, top_k: int = 100, select_fn=None,
                 group: Optional[dist.ProcessGroup] = None):
    """Top / middle / bottom-k over the whole dataset for this rank's share of the features.
    ``select_fn(rows (F_d, C, n_tokens) fp16, top_k) -> (idx, val)`` defaults to the CUDA radix select;
    returned indices are global token positions.  Returns (idx, val, (f_start, f_stop))."""
    if select_fn is None:
        from .feature import select_top_middle_bottom as select_fn
    rows, frange = exchange_to_feature_shards(intens_local, n_tokens, group=group)
    idx, val = select_fn(rows, top_k)
    return idx, val, frange


# ---------------------------------------------------------------------------------------------------
# nearest-example search across GPUs: sequences shard, the ranking is global
# ---------------------------------------------------------------------------------------------------
def find_examples_sharded(engine, n_sequences: int, idx: Optional[int] = None, activation: Optional[torch.Tensor] = None,
                          top_examples: int = 30, middle_examples: int = 10, bottom_examples: int = 10,
                          layers=None, group: Optional[dist.ProcessGroup] = None, select_fn=None):
    """``IntensityEngine.find_examples`` (demo/server/server.py:159-325) over a code store whose SEQUENCES are split
    across ranks: ``engine`` is this rank's ``rqae_b200.search.IntensityEngine`` over the contiguous range
    ``token_range(n_sequences, rank, world)``.  The accumulation has no cross-sequence term, so each rank runs the
    kernels on its shard; per layer cut the only exchange is an all_gather of the per-position maxima (Sq x N fp16:
    9.4 MB for the reference's 36 864 sequences) -- every rank then ranks all sequences with the radix select -- and
    an all_reduce that collects the (Sq, k, S) intensity rows of the selected sequences from their owners.  Yields
    the reference's ``(result dict, layer)`` on every rank, equal to the single-GPU result.  ``idx`` is a GLOBAL
    sequence number; its owner broadcasts the query codes.  ``select_fn(rows (Sq, N) fp16, k) -> (idx (Sq, 3, k), val)``
    defaults to the CUDA radix select."""
    from .search import SERVER_LAYERS, window_k, window_lists
    if select_fn is None:
        from .feature import select_top_middle_bottom as select_fn
    layers = list(SERVER_LAYERS if layers is None else layers)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    lo, hi = token_range(n_sequences, rank, world)
    n_local, seq_len = engine.activations.shape[:2]
    if n_local != hi - lo:
        raise ValueError(f"rank {rank} must hold sequences [{lo}, {hi}), its engine holds {n_local}")
    if activation is not None and idx is not None:
        raise ValueError("Cannot specify both idx and activation")                 # server.py:174-175
    if activation is None and idx is None:
        raise ValueError("Must specify either idx or activation")                  # server.py:181-182
    L = max(layers)
    dev = engine.sims.device
    if idx is not None:
        if not 0 <= int(idx) < n_sequences:
            raise IndexError(f"idx {idx} outside the {n_sequences} sequences of the store")
        owner = next(r for r in range(world) if token_range(n_sequences, r, world)[0] <= int(idx) < token_range(n_sequences, r, world)[1])
        query = engine._query(int(idx) - lo, None, L) if rank == owner else torch.empty(seq_len, L, dtype=torch.int32, device=dev)
        dist.broadcast(query, src=dist.get_global_rank(group, owner) if group is not None else owner, group=group)
    else:
        query = engine._query(None, activation, L)
    Sq = query.shape[0]
    k = window_k(top_examples, middle_examples, bottom_examples, n_sequences)
    sizes = [token_range(n_sequences, r, world)[1] - token_range(n_sequences, r, world)[0] for r in range(world)]
    widest = max(sizes)
    qpos = torch.arange(Sq, device=dev).unsqueeze(-1)
    n_pad = (n_sequences + 7) // 8 * 8
    for layer, (acc, maxv) in zip(layers, engine.accumulate(query, layers)):
        mine = torch.zeros(Sq, widest, dtype=torch.float16, device=dev)
        mine[:, :n_local] = maxv
        parts = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(parts, mine, group=group)
        rows = torch.zeros(Sq, n_pad, dtype=torch.float16, device=dev)             # row layout of the radix select
        rows[:, :n_sequences] = torch.cat([p[:, :sizes[r]] for r, p in enumerate(parts)], dim=1)
        sel, _ = select_fn(rows[:, :n_sequences], k)
        out = {}
        for name, lst in window_lists(sel, top_examples, middle_examples, bottom_examples).items():
            g = lst.long()
            own = (g >= lo) & (g < hi)
            inten = torch.zeros(Sq, g.shape[1], seq_len, dtype=torch.float32, device=dev)
            if bool(own.any()):
                qq = qpos.expand_as(g)
                inten[own] = acc[g[own] - lo, :, qq[own]].float()
            dist.all_reduce(inten, group=group)                                    # one owner per row: the sum is exact
            out[name] = {"indices": lst.cpu().int(), "intensities": inten.cpu().to(torch.float16)}
        yield out, layer
