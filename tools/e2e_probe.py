#!/usr/bin/env python
"""Where does the end-to-end (host-buffer) path lose its scaling?  One process per GPU (torchrun) or one process.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29611 \
        tools/e2e_probe.py [--tokens T] [--out gpurun_out/e2e_probe_nN.json]

Every stage is started behind a barrier on all ranks at once and timed on the host around a stream synchronise;
the numbers reported are the SLOWEST rank's (max over ranks), per rank:

  h2d / d2h / bidir     plain cudaMemcpyAsync of pinned buffers (2 GiB), GB/s per rank
  host_copy             one torch CPU copy of 1 GiB per rank (host memory bandwidth under N-fold contention)
  widen                 the narrow mode's int16 -> int64 step alone, 1 / 2 / 8 threads, Gcodes/s per rank
  kernel                RQAE.forward on resident data, tokens/s per rank
  forward_host          the e2e call in its modes: direct | narrow x threads, int64 | int32 | int16 result, chunk size

Rank 0 prints one JSON line and writes it to --out."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tokens", type=int, default=1 << 19)
    ap.add_argument("--out", default="")
    ap.add_argument("--quick", action="store_true")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    from rqae_b200 import RQAE, _lib

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def maxr(v):
        if world == 1:
            return v
        t = torch.tensor([v], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed(fn, reps=1):
        barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize(dev)
        dt = (time.perf_counter() - t0) / reps
        return maxr(dt)

    import faulthandler
    faulthandler.enable()

    def note(msg):
        if rank == 0:
            print("[probe]", msg, file=sys.stderr, flush=True)

    lib = _lib.load()
    res = {"world": world, "cores": os.cpu_count(), "tokens_per_rank": args.tokens}
    try:
        res["numa_nodes"] = len([d for d in os.listdir("/sys/devices/system/node") if d.startswith("node")])
    except Exception:
        res["numa_nodes"] = None
    try:
        import psutil
        res["host_mem_gb"] = psutil.virtual_memory().total / 1e9
        res["affinity"] = len(os.sched_getaffinity(0))
    except Exception:
        pass

    note("raw copies")
    # ---- raw copies ----
    nb = 1 << 31
    hp = torch.empty(nb, dtype=torch.uint8, pin_memory=True)
    hp2 = torch.empty(nb, dtype=torch.uint8, pin_memory=True)
    hp.fill_(1); hp2.fill_(2)
    dv = torch.empty(nb, dtype=torch.uint8, device=dev)
    dv2 = torch.empty(nb, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    def h2d():
        with torch.cuda.stream(s1):
            dv.copy_(hp, non_blocking=True)

    def d2h():
        with torch.cuda.stream(s2):
            hp2.copy_(dv2, non_blocking=True)

    def both():
        h2d(); d2h()
    h2d(); d2h(); torch.cuda.synchronize(dev)
    res["h2d_gbs_per_rank"] = nb / timed(h2d, 3) / 1e9
    res["d2h_gbs_per_rank"] = nb / timed(d2h, 3) / 1e9
    t = timed(both, 3)
    res["bidir_gbs_per_rank_each_way"] = nb / t / 1e9
    del dv, dv2

    note("host memory")
    # ---- host memory ----
    a = torch.empty(1 << 30, dtype=torch.uint8); b = torch.empty(1 << 30, dtype=torch.uint8)
    a.fill_(3); b.copy_(a)
    torch.set_num_threads(1)
    res["host_copy_gbs_per_rank_1thread"] = 2 * (1 << 30) / timed(lambda: b.copy_(a), 3) / 1e9
    del a, b
    import numpy as np
    n = 1 << 27
    src = np.random.default_rng(0).integers(0, 625, size=n, dtype=np.int16)
    dst = np.empty(n, np.int64)
    lib.rqae_widen_codes_host(src.ctypes.data, dst.ctypes.data, n, 2, 1)
    for th in (1, 2, 8):
        res[f"widen_gcodes_per_rank_{th}thr"] = n / timed(lambda: lib.rqae_widen_codes_host(src.ctypes.data, dst.ctypes.data, n, 2, th), 2) / 1e9
    del src, dst, hp, hp2

    note("model")
    # ---- the model ----
    torch.manual_seed(0)
    model = RQAE().eval().to(dev)
    model.freeze_packed()
    T = args.tokens
    D, NQ = 2304, 1024
    x = torch.randn(T, D, device=dev, generator=torch.Generator(device=dev).manual_seed(1234 + rank))
    model(x[None, : 1 << 14]); torch.cuda.synchronize(dev)
    res["kernel_tokens_per_s_per_rank"] = T / timed(lambda: model(x[None]), 2)
    xh = torch.empty(T, D, dtype=torch.float32, pin_memory=True)
    xh.copy_(x)
    del x
    qh = torch.empty(T, D, dtype=torch.float32, pin_memory=True)
    outs = {torch.int64: torch.empty(T, NQ, dtype=torch.int64, pin_memory=True),
            torch.int32: torch.empty(T, NQ, dtype=torch.int32, pin_memory=True),
            torch.int16: torch.empty(T, NQ, dtype=torch.int16, pin_memory=True)}
    cases = [("direct", 0, torch.int64, 9472), ("narrow", 0, torch.int64, 9472), ("narrow", 1, torch.int64, 9472),
             ("narrow", 8, torch.int64, 9472), ("direct", 0, torch.int32, 9472), ("direct", 0, torch.int16, 9472),
             ("direct", 0, torch.int64, 33152), ("direct", 0, torch.int64, 4736), ("narrow", 0, torch.int64, 33152)]
    if args.quick:
        cases = cases[:2]
    fh = {}
    ref_sum = None
    for mode, thr, dt, chunk in cases:
        note(f"forward_host {mode} {thr} {dt} {chunk}")
        ch = outs[dt]
        ch.zero_()

        def run():
            model.forward_host(xh, out=(qh, ch), out_dtype=dt, chunk_tokens=chunk, code_transfer=mode, widen_threads=thr)
        model.forward_host(xh[: 1 << 15], out=(qh[: 1 << 15], ch[: 1 << 15]), out_dtype=dt, chunk_tokens=chunk,
                           code_transfer=mode, widen_threads=thr)
        dtm = timed(run, 2)
        key = f"{mode}{'' if mode == 'direct' else '_thr' + str(thr or 'auto')}_{str(dt).split('.')[-1]}_chunk{chunk}"
        fh[key] = T / dtm
        sm = int(ch[:4096].long().sum().item())
        ref_sum = sm if ref_sum is None else ref_sum
        assert sm == ref_sum, "modes disagree on the codes"
    res["forward_host_tokens_per_s_per_rank"] = fh
    lib.rqae_forward_host_config(0, 0)
    if rank == 0:
        line = json.dumps(res)
        print(line, flush=True)
        if args.out:
            os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
            open(args.out, "w").write(line + "\n")
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
