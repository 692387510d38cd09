"""The reference's on-disk code store, read and written in its own format.

``scripts/1_create_activations.py:251-261`` saves, per shard of 1024 sequences, the codes returned by the hook
as ``torch.save`` of a contiguous **int32** tensor ``(sequences, seq_len, num_quantizers)`` (BOS position
included) under ``<folder>/<rqae.name>/{shard:06d}.pt`` next to ``{shard:06d}_ce.pt`` (the CE loss).  The
mining script reads every ``{i:06d}.pt`` back in order and concatenates (``scripts/3_make_rqae_features.py:
44-51``); the demo server lists the folder, skips ``*_ce*`` files, sorts numerically and drops the BOS
position (``demo/server/server.py:118-128``).  This module is that contract as functions, so that shards written
by either implementation are read by the other.  Plain file I/O -- host logic, no kernels.

New here: ``RQAE.encode(..., out_dtype=torch.int32)`` emits the store's dtype straight from the kernel, which
removes the ``.cpu().to(torch.int32)`` conversion of ``scripts/1:184-186``."""
from __future__ import annotations

import os
from typing import Iterable, List, Optional

import torch

SHARD_SEQUENCES = 1024   # scripts/1_create_activations.py:291 ("Set for the rest of all scripts")


def shard_path(folder: str, model_name: str, shard: int, ce: bool = False) -> str:
    """``<folder>/<rqae.name>/{shard:06d}.pt`` (scripts/1:252-261)."""
    return os.path.join(folder, model_name, f"{shard:06d}{'_ce' if ce else ''}.pt")


def save_code_shard(folder: str, model_name: str, shard: int, codes: torch.Tensor, ce: Optional[float] = None) -> str:
    """Write one shard the way scripts/1:225-261 does: int32, contiguous, on the CPU."""
    if codes.dim() != 3:
        raise ValueError(f"codes must be (sequences, seq_len, num_quantizers), got {tuple(codes.shape)}")
    path = shard_path(folder, model_name, shard)
    os.makedirs(os.path.dirname(path), exist_ok=True)
    torch.save(codes.detach().to(device="cpu", dtype=torch.int32).contiguous(), path)
    if ce is not None:
        torch.save(ce, shard_path(folder, model_name, shard, ce=True))
    return path


def list_code_shards(folder: str, model_name: str) -> List[str]:
    """Shard files in numeric order, CE files excluded (demo/server/server.py:120-121)."""
    d = os.path.join(folder, model_name)
    files = [f for f in os.listdir(d) if "_ce" not in f and f.endswith(".pt")]
    return [os.path.join(d, f) for f in sorted(files, key=lambda x: int(x.split(".")[0]))]


def load_code_shards(folder: str, model_name: str, shards: Optional[Iterable[int]] = None, skip_bos: bool = False,
                     dtype: torch.dtype = torch.int32, device=None) -> torch.Tensor:
    """Concatenate shards along the sequence axis (scripts/3:44-51).  ``skip_bos=True`` drops position 0 as the
    demo server does (server.py:123-126).  ``dtype=torch.int16`` halves the footprint for the GPU kernels
    (every code is < 32768); ``device`` moves the result (e.g. the mining GPU)."""
    paths = list_code_shards(folder, model_name) if shards is None else [shard_path(folder, model_name, i) for i in shards]
    parts = []
    for p in paths:
        t = torch.load(p)
        if skip_bos:
            t = t[:, 1:]
        if dtype == torch.int16 and t.numel() and int(t.max()) > 32767:
            raise ValueError(f"{p}: codes up to {int(t.max())} do not fit int16; load the store as int32")
        parts.append(t.to(dtype))
    out = torch.cat(parts, dim=0) if parts else torch.empty(0, 0, 0, dtype=dtype)
    return out if device is None else out.to(device)
