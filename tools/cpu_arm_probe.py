"""How the CPU arm's throughput depends on thread count / binding on this host (run on the GPU box; CPU only)."""
import os, subprocess, sys, json
code = '''
import os, sys, time
sys.path.insert(0, ".")
import numpy as np
from oracle import c_oracle
t = int(sys.argv[1])
rng = np.random.default_rng(0)
X = rng.standard_normal((1_000_000, 128), dtype=np.float32)
c_oracle.set_threads(t)
X = c_oracle.first_touch_copy(X)
C0 = X[:64].copy()
c_oracle.lloyd(X[:2000], C0, 1, -1.0, want_labels=False)
best = 1e9
for _ in range(3):
    t0 = time.perf_counter(); c_oracle.lloyd(X, C0, 5, -1.0, want_labels=False); best = min(best, time.perf_counter() - t0)
print(f"{5e6/best/1e6:.2f}")
'''
print("affinity cpus:", len(os.sched_getaffinity(0)), "cpu_count:", os.cpu_count())
try:
    print("cgroup cpu.max:", open("/sys/fs/cgroup/cpu.max").read().strip())
except Exception as e:
    print("cgroup cpu.max: n/a")
for bind, places in ((None, None), ("close", "cores"), ("spread", "cores"), ("false", None)):
    for t in (16, 32, 64, 128):
        env = dict(os.environ)
        env.pop("OMP_PROC_BIND", None); env.pop("OMP_PLACES", None); env.pop("OMP_NUM_THREADS", None)
        if bind: env["OMP_PROC_BIND"] = bind
        if places: env["OMP_PLACES"] = places
        r = subprocess.run([sys.executable, "-c", code, str(t)], env=env, capture_output=True, text=True)
        print(f"bind={bind} places={places} threads={t}: {r.stdout.strip()} M samples/s {r.stderr.strip()[-100:]}", flush=True)
