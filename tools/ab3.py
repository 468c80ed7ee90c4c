"""A/B builds of libb2kmeans.so on the cfg3 shape in one GPU job (alternating): python tools/ab3.py a.so b.so ..."""
import sys, os, subprocess
code = r'''
import sys, os, torch
sys.path.insert(0, os.getcwd())
from spark_rapids_ml_b200 import _native
n, d, k = 6_000_000, 256, 256
ctx = _native.Context(0)
g = torch.Generator(device="cuda").manual_seed(1)
ctr = torch.rand((k, d), generator=g, device="cuda") * 20 - 10
X = torch.empty((n, d), device="cuda")
for s in range(0, n, 1_000_000):
    e = min(n, s + 1_000_000)
    X[s:e] = ctr[torch.randint(0, k, (e - s,), generator=g, device="cuda")] + torch.randn((e - s, d), generator=g, device="cuda")
C0 = (ctr + 0.25 * torch.randn((k, d), generator=g, device="cuda")).contiguous()
C1 = X[:k].clone()
ctx.kmeans_lloyd(X, C0.clone(), 3, -1.0)
ctx.set_option("time_kernels", 1)
out = []
for C in (C0, C0, C1):
    Cc = C.clone(); ctx.kmeans_lloyd(X, Cc, 20, -1.0); out.append(round(ctx.stats()["last_fused_ms"] * 12.5e6 / n, 3))
print("near_true x2, first_k (ms scaled to 12.5M rows):", out)
'''
libs = sys.argv[1:]
for rnd in range(2):
    for lib in libs:
        subprocess.run(["cp", lib, "spark_rapids_ml_b200/libb2kmeans.so"], check=True)
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
        print(os.path.basename(lib), r.stdout.strip() or r.stderr[-300:], flush=True)
