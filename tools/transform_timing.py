"""KMeansModel.transform throughput over 10 000-row Arrow batches (python tools/transform_timing.py [rows])."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from spark_rapids_ml_b200.clustering import KMeansModel
from spark_rapids_ml_b200.sparkshim import LocalSession
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
d, k = 128, 64
rng = np.random.default_rng(0)
X = rng.standard_normal((n, d), dtype=np.float32)
sess = LocalSession()
df = sess.from_numpy(X, col="features", num_partitions=1)
model = KMeansModel(cluster_centers_=X[:k].astype(np.float64).tolist(), n_cols=d, dtype="float32")
model._set(featuresCol="features")
for arrow_backed in (True, False):
    df.arrow_backed_pandas = arrow_backed
    model.transform(df).count()
    t0 = time.perf_counter()
    out = model.transform(df)
    m = out.count()
    dt = time.perf_counter() - t0
    print(f"arrow_backed={arrow_backed}: {m} rows in {dt:.3f} s = {m / dt / 1e6:.1f} M rows/s ({dt / (n / 10000) * 1e3:.3f} ms per 10 000-row batch)")
