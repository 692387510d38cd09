"""N-GPU check of the mining path (torchrun, one process per GPU, NCCL): every rank encodes its token shard,
computes the intensities of all features over it, the all_to_all makes (feature, cut) rows whole on the rank
that owns the feature, the radix select runs there.  Rank 0 recomputes everything on one GPU and compares.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/mine_2gpu.py"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from rqae_b200 import RQAE, shard
from rqae_b200.feature import intensity_many, select_top_middle_bottom, layer_weights_f16

CUTS = [2, 4, 6, 8, 12, 16, 24, 32, 48, 64, 127]


def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(0)
    m = RQAE(dim=2304, num_quantizers=128).eval().to(dev)
    T, Fn = 20000, 300
    x = torch.randn(T, 2304, generator=torch.Generator().manual_seed(1)).to(dev)
    codes, (a, b) = shard.encode_sharded(m, x)                      # this rank's tokens
    all_codes = shard.gather_codes(codes, T)                          # feature centers are tokens of the whole set
    centers = all_codes[torch.arange(0, T, T // Fn)[:Fn].to(dev)].to(torch.int32)
    lw = layer_weights_f16(m).to(dev)
    local = intensity_many(m, codes, centers, CUTS, layer_weights=lw)  # (F, C, T_r)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    idx, val, (fa, fb) = shard.mine_sharded(local.contiguous(), T, top_k=100)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    # single-GPU recomputation on every rank for its own features
    full = intensity_many(m, all_codes, centers[fa:fb], CUTS, layer_weights=lw)
    idx1, val1 = select_top_middle_bottom(full, 100)
    torch.cuda.synchronize()
    ok = bool(torch.equal(idx, idx1) and torch.equal(val, val1))
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(json.dumps({"world": world, "tokens": T, "features": Fn, "cuts": len(CUTS), "sharded_equals_single_gpu": bool(flag.item()),
                          "exchange_plus_select_s": dt}))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
