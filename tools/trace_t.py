"""Per-step event trace of the large-shape kernel (probe build + option profile_fused): cycles between pipeline events."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, ".")
from spark_rapids_ml_b200 import _native

n, d, k = 2_000_000, 256, 256
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(42)
ctr = torch.rand((k, d), generator=g, device=dev) * 20 - 10
z = torch.randint(0, k, (n,), generator=g, device=dev)
X = (ctr[z] + torch.randn((n, d), generator=g, device=dev)).contiguous()
C = (ctr + 0.25 * torch.randn((k, d), generator=g, device=dev)).contiguous()
if len(sys.argv) > 1 and sys.argv[1] == "first_k":
    C = X[:k].clone()
ctx = _native.Context(0)
ctx.set_option("kernel_path", 2)
import os
if os.environ.get("B2K_PROBE_N"):
    ctx.set_option("probe", int(os.environ["B2K_PROBE_N"]))
Cw = C.clone()
if not (len(sys.argv) > 2 and sys.argv[2] == "raw"):   # "first_k raw": trace the very first iteration of the bad start
    ctx.kmeans_lloyd(X, Cw, 10 if len(sys.argv) > 1 else 2, -1.0)   # settle the centres before tracing a bad start
ctx.set_option("profile_fused", 1)
ctx.set_option("collect_recheck", 1)
ctx.kmeans_lloyd(X, Cw, 1, -1.0)
print("recheck rows / candidates in the traced iteration:", ctx.stats()["recheck_rows"], ctx.stats()["recheck_candidates"], "of", n)
tr = ctx.fused_profile().reshape(-1)[:256].reshape(2, 8, 16)
names = {0: "tma:sfree_ok", 1: "upd:keys_ok", 2: "mma:dempty_ok", 3: "mma:xfull0_ok", 14: "mma:xfull7_ok", 4: "mma:issued", 5: "epi:dfull_ok",
         6: "epi:Dloop_done", 7: "epi:exch_done", 8: "epi:flags", 9: "epi:enum_done", 13: "epi:eval_done", 15: "epi:rx_ok", 10: "epi:lfull_sent", 11: "upd:lfull_ok", 12: "upd:rows_done"}
order = [0, 2, 3, 14, 4, 5, 6, 7, 8, 9, 13, 15, 10, 11, 1, 12]
for cta in range(2):
    print(f"== CTA {cta}: cycles relative to epi:dfull_ok of step 20")
    t0 = tr[cta, 0, 5]
    for st in range(8):
        row = " ".join(f"{names[e].split(':')[1]}={int(tr[cta, st, e] - t0):>6d}" for e in order if tr[cta, st, e] != 0)
        print(f"step {20+st}: {row}")
    print("period (epi:dfull_ok):", np.diff(tr[cta, :, 5]).tolist())
