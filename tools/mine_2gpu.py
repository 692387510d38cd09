"""N-GPU check and timing of the mining path (torchrun, one process per GPU, NCCL): every rank encodes its token
shard, computes the intensities of ALL features over it, one all_to_all makes (feature, cut) rows whole on the rank
that owns the feature, the radix select runs there.  Every rank then recomputes its features on the gathered codes
alone and compares index for index.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/mine_2gpu.py
env: NQ (128), TOKENS (20000, whole job), FEATURES (300)"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from rqae_b200 import RQAE, shard
from rqae_b200.feature import intensity_many, select_top_middle_bottom, layer_weights_f16


def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    nq, T, Fn = int(os.environ.get("NQ", 128)), int(os.environ.get("TOKENS", 20000)), int(os.environ.get("FEATURES", 300))
    cuts = [c for c in [2, 4, 6, 8, 12, 16, 24, 32, 48, 64, 128, 256, 512] if c < nq - 1] + [nq - 1]
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(0)
    m = RQAE(dim=2304, num_quantizers=nq).eval().to(dev)
    m.freeze_packed()
    a, b = shard.token_range(T, rank, world)
    x = torch.randn(b - a, 2304, device=dev, generator=torch.Generator(device=dev).manual_seed(100 + rank))
    lw = layer_weights_f16(m).to(dev)

    def sync():
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()

    m.encode(x[:1024].unsqueeze(0), out_dtype=torch.int16)       # warm-up (packs the weights)
    sync(); t0 = time.perf_counter()
    codes = m.encode(x.unsqueeze(0), out_dtype=torch.int16)[0]     # this rank's tokens
    sync(); t_enc = time.perf_counter() - t0
    all_codes = shard.gather_codes(codes, T)                       # feature centers are tokens of the whole set
    centers = all_codes[torch.arange(0, T, max(1, T // Fn))[:Fn].to(dev)].to(torch.int32)
    intensity_many(m, codes[:256], centers, cuts, layer_weights=lw)
    sync(); t0 = time.perf_counter()
    local = intensity_many(m, codes, centers, cuts, layer_weights=lw)   # (F, C, T_r)
    sync(); t_int = time.perf_counter() - t0
    local = local.contiguous()
    shard.mine_sharded(local[:, :, : 64].contiguous(), 64 * world, top_k=8)   # warm-up of the collective
    sync(); t0 = time.perf_counter()
    idx, val, (fa, fb) = shard.mine_sharded(local, T, top_k=100)
    sync(); t_mine = time.perf_counter() - t0
    full = intensity_many(m, all_codes, centers[fa:fb], cuts, layer_weights=lw)
    idx1, val1 = select_top_middle_bottom(full, 100)
    torch.cuda.synchronize()
    ok = bool(torch.equal(idx, idx1) and torch.equal(val, val1))
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(json.dumps({"world": world, "tokens": T, "features": Fn, "cuts": len(cuts), "num_quantizers": nq,
                          "sharded_equals_single_gpu": bool(flag.item()),
                          "encode_s": t_enc, "encode_tokens_per_s": T / t_enc,
                          "intensity_s": t_int, "intensity_tokens_per_s": T / t_int,
                          "exchange_plus_select_s": t_mine,
                          "note": "wall clock between barriers, max over ranks by construction; whole-job tokens"}))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
