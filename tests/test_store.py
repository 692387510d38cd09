"""The code store contract of the reference (scripts/1_create_activations.py:251-261 writes,
scripts/3_make_rqae_features.py:44-51 and demo/server/server.py:118-128 read) as implemented by rqae_b200.store."""
import os

import torch

from rqae_b200 import store

NAME = "rqae-rqae-round_fsq-cbd4-cbs5-nq1024"


def test_shard_files_are_what_the_reference_reads(tmp_path):
    g = torch.Generator().manual_seed(0)
    shards = [torch.randint(0, 625, (4, 8, 16), generator=g) for _ in range(3)]
    for i, c in enumerate(shards):
        p = store.save_code_shard(str(tmp_path), NAME, i, c, ce=1.5 + i)
        assert p.endswith(f"{NAME}/{i:06d}.pt")
    # the reference's reader (scripts/3:44-51): torch.load of {i:06d}.pt, cat on dim 0
    ref = torch.cat([torch.load(os.path.join(str(tmp_path), NAME, f"{i:06d}.pt")) for i in range(3)], dim=0)
    assert ref.dtype == torch.int32 and ref.is_contiguous() and torch.equal(ref.long(), torch.cat(shards))
    assert torch.load(store.shard_path(str(tmp_path), NAME, 2, ce=True)) == 3.5
    # the server's listing (server.py:120-121): no CE files, numeric order
    assert [os.path.basename(p) for p in store.list_code_shards(str(tmp_path), NAME)] == ["000000.pt", "000001.pt", "000002.pt"]
    assert torch.equal(store.load_code_shards(str(tmp_path), NAME), ref)
    nb = store.load_code_shards(str(tmp_path), NAME, skip_bos=True, dtype=torch.int16)
    assert nb.dtype == torch.int16 and torch.equal(nb.long(), torch.cat(shards)[:, 1:])
    assert torch.equal(store.load_code_shards(str(tmp_path), NAME, shards=[1]).long(), shards[1])
