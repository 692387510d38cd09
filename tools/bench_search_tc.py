"""Time the example search over a store of the reference's shape in both modes (bench.py's search_extras, standalone).
usage: python tools/bench_search_tc.py [sequences]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from rqae_b200 import RQAE

dev = torch.device("cuda:0")
torch.manual_seed(0)
model = RQAE(dim=2304, num_quantizers=1024).eval().to(dev)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 36864
r = bench.search_extras(torch, model, dev, sequences=n)
r.pop("cpu_port", None)
print(json.dumps(r))
