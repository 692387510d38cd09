"""Workload for the ncu capture of the search kernels: one find_examples pass over N sequences (default 4096) at the
reference's shape (127 positions, 1024 layers, K = 625, server.py layer list).  Used by tools/gpu_search_prof.sh."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from rqae_b200.search import IntensityEngine, SERVER_LAYERS  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(1)
sims = torch.randn(1024, 625, 625, generator=g, device=dev, dtype=torch.float16) * 0.25
sims.diagonal(dim1=1, dim2=2).fill_(1.0)
codes = torch.randint(0, 625, (N, 127, 1024), generator=g, device=dev, dtype=torch.int16)
eng = IntensityEngine(sims=sims, activations=codes)
q = eng._query(N // 3, None, max(SERVER_LAYERS))
for _ in eng.accumulate(q, SERVER_LAYERS):
    pass
torch.cuda.synchronize()
print("ok")
