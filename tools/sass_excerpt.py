"""SASS mnemonic counts per kernel of rqae_b200/librqae_b200.so -> profiles/sass_excerpt.txt
usage: python tools/sass_excerpt.py > profiles/sass_excerpt.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "rqae_b200", "librqae_b200.so")
KEYS = ["UTCHMMA", "LDTM", "UTCBAR", "UBLKCP", "UTMASTG", "UTMALDG", "UBLKPF", "FFMA2", "FFMA", "USETMAXREG", "SYNCS", "STAS", "UCGABAR", "HSET2", "IDP",
        "ATOMS", "REDUX", "HMMA"]
arch = sorted(set(re.findall(r"sm_\d+a?", subprocess.run(["cuobjdump", "-lelf", so], capture_output=True, text=True).stdout)))
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
dem = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
print("# SASS mnemonic counts per kernel of rqae_b200/librqae_b200.so (cuobjdump -sass; built by __graft_entry__.build(); tools/sass_excerpt.py).")
print("# UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit, UBLKCP = cp.async.bulk (1-D bulk TMA), UTMASTG = cp.async.bulk.tensor store")
print("# (tensor-map TMA), UBLKPF = bulk L2 prefetch, FFMA2 = fma.rn.f32x2, USETMAXREG = setmaxnreg, SYNCS = mbarrier ops, STAS = st.async (cluster variant),")
print("# UCGABAR = barrier.cluster, HSET2 = packed fp16 compare, IDP = dp4a (lane-private counter folds), ATOMS = shared-memory atomics, REDUX = warp reduce;")
print("# HMMA (legacy mma.sync) must not appear.")
print(f"architectures in the fatbin: {arch}\n")
cur, cnt, tot = None, None, 0
def flush():
    if cur is not None:
        print(f"{dem(cur)[:104]:104s} instr={tot:6d} " + " ".join(f"{k}={cnt[k]}" for k in KEYS if cnt[k]))
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        flush()
        cur, cnt, tot = m.group(1), collections.Counter(), 0
        continue
    m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        tot += 1
        op = m.group(1)
        for k in KEYS:
            if op == k or (k in ("FFMA",) and op == "FFMA"):
                cnt[k] += 1
flush()
