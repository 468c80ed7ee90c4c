"""A/B two builds of libb2kmeans.so in one GPU job (alternating, to cancel thermal/power drift)."""
import sys, os, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
code = r'''
import sys, os, torch
sys.path.insert(0, os.getcwd())
from spark_rapids_ml_b200 import _native
n, d, k = 10_000_000, 128, 64
ctx = _native.Context(0)
for kv in os.environ.get("B2K_OPTS", "").split(","):
    if "=" in kv: ctx.set_option(kv.split("=")[0], int(kv.split("=")[1]))
g = torch.Generator(device="cuda").manual_seed(1)
ctr = torch.rand((k, d), generator=g, device="cuda") * 20 - 10
X = torch.empty((n, d), device="cuda")
for s in range(0, n, 1_000_000):
    e = min(n, s + 1_000_000)
    X[s:e] = ctr[torch.randint(0, k, (e - s,), generator=g, device="cuda")] + torch.randn((e - s, d), generator=g, device="cuda")
C = X[:k].clone(); ctx.kmeans_lloyd(X, C, 5, -1.0)
ctx.set_option("time_kernels", 1)
out = []
for rep in range(3):
    C = X[:k].clone(); ctx.kmeans_lloyd(X, C, 60, -1.0); out.append(round(ctx.stats()["last_fused_ms"], 4))
print(out)
'''
libs = sys.argv[1:]   # "path.so" or "path.so:opt=val,opt=val"
for rnd in range(2):
    for spec in libs:
        lib, _, opts = spec.partition(":")
        subprocess.run(["cp", lib, "spark_rapids_ml_b200/libb2kmeans.so"], check=True)
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=dict(os.environ, B2K_OPTS=opts))
        print(os.path.basename(lib), opts, r.stdout.strip() or r.stderr[-300:], flush=True)
