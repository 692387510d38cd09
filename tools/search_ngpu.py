"""N-GPU check and timing of the sequence-sharded example search (torchrun, one process per GPU, NCCL): every rank
holds its contiguous range of the code store, shard.find_examples_sharded ranks all sequences per layer cut
(all_gather of the per-position maxima, radix select on every rank, all_reduce of the selected rows); rank 0 also runs
the single-GPU IntensityEngine over the whole store and compares index for index and value for value.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/search_ngpu.py
env: SEQUENCES (4096, whole job), NQ (1024), POSITIONS (127)"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from rqae_b200 import shard
from rqae_b200.search import IntensityEngine, SERVER_LAYERS


def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    N, nq, S, K = int(os.environ.get("SEQUENCES", 4096)), int(os.environ.get("NQ", 1024)), int(os.environ.get("POSITIONS", 127)), 625
    layers = [l for l in SERVER_LAYERS if l < nq] + ([nq - 1] if nq - 1 not in SERVER_LAYERS else [])
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    dist.init_process_group("nccl", device_id=dev)
    g = torch.Generator(device=dev).manual_seed(5)                 # the same table and store on every rank, then sliced
    sims = torch.randn(nq, K, K, generator=g, device=dev, dtype=torch.float16) * 0.25
    sims.diagonal(dim1=1, dim2=2).fill_(1.0)
    store = torch.randint(0, K, (N, S, nq), generator=g, device=dev, dtype=torch.int16)
    a, b = shard.token_range(N, rank, world)
    eng = IntensityEngine(sims=sims, activations=store[a:b].clone())
    idx = (2 * N) // 3

    def run():
        return [(r, l) for r, l in shard.find_examples_sharded(eng, N, idx=idx, layers=layers)]

    def sync():
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()

    run()
    sync(); t0 = time.perf_counter()
    got = run()
    sync(); t_sharded = time.perf_counter() - t0
    ok = True
    t_single = None
    if rank == 0:
        full = IntensityEngine(sims=sims, activations=store)
        list(full.find_examples(idx=idx, layers=layers))
        torch.cuda.synchronize(); t0 = time.perf_counter()
        want = list(full.find_examples(idx=idx, layers=layers))
        torch.cuda.synchronize(); t_single = time.perf_counter() - t0
        for (r, l), (w, wl) in zip(got, want):
            for part in ("top", "middle", "bottom"):
                ok = ok and l == wl and torch.equal(r[part]["indices"], w[part]["indices"]) \
                    and torch.equal(r[part]["intensities"].float(), w[part]["intensities"].float())
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(json.dumps({"world": world, "sequences": N, "positions": S, "num_quantizers": nq, "cuts": len(layers),
                          "sharded_equals_single_gpu": bool(flag.item()), "sharded_s": t_sharded, "single_gpu_s": t_single,
                          "sharded_sequences_per_s": N / t_sharded,
                          "note": "wall clock between barriers (max over ranks by construction); whole-job sequences"}))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
