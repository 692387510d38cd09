"""Component times of one tensor-core-ranked query over a store of the reference's shape."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rqae_b200 import RQAE
from rqae_b200.search import IntensityEngine, SERVER_LAYERS, window_k, window_lists
from rqae_b200.feature import select_top_middle_bottom

dev = torch.device("cuda:0")
torch.manual_seed(0)
model = RQAE(dim=2304, num_quantizers=1024).eval().to(dev)
N, S = int(sys.argv[1]) if len(sys.argv) > 1 else 36864, 127
K = model.codebook.shape[1]
g = torch.Generator(device=dev).manual_seed(77)
codes = torch.randint(0, K, (N, S, 1024), generator=g, device=dev, dtype=torch.int16)
eng = IntensityEngine(model, codes, precision="tc")
layers = SERVER_LAYERS
query = eng._query(N // 3, None, max(layers))
k = window_k(30, 10, 10, N)


def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        r = fn()
    e1.record(); torch.cuda.synchronize()
    return r, e0.elapsed_time(e1) / reps


maxv, t_max = timed(lambda: eng.maxima_tc(query, layers, padded=True))
table, t_tab = timed(lambda: eng._build_qrows(query, max(layers)))
(sel, _), t_sel = timed(lambda: select_top_middle_bottom(maxv, k, n=N))
sel = sel[:, :127].contiguous()
lists = window_lists(sel.reshape(len(layers) * 127, 3, k), 30, 10, 10)
cat = torch.cat([lists[nm] for nm in lists], dim=1).reshape(len(layers), 127, -1)
rows, t_rows = timed(lambda: eng.rows_exact(table, cat, layers, first_range=0))
t0 = time.perf_counter(); rc = rows.cpu(); cc = cat.cpu(); t_d2h = (time.perf_counter() - t0) * 1e3
print(f"maxima GEMM {t_max:.2f} ms | table build {t_tab:.2f} | select {t_sel:.2f} | exact rows {t_rows:.2f} | D2H (pageable) {t_d2h:.2f} | rows bytes {rows.numel() * 2 / 1e6:.1f} MB")
