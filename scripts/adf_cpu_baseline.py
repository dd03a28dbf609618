"""CPU oracle rate for the ADF (reference algorithm restated): small system, same density / cutoff."""
import sys, time, json, numpy as np
sys.path.insert(0, ".")
from oracle import adf as oadf
n = int(sys.argv[1]) if len(sys.argv) > 1 else 600
L = (n / 0.05) ** (1 / 3)
rng = np.random.default_rng(1)
data = {"A": (rng.random((n // 2, 1, 3)) * L).astype(np.float32), "B": (rng.random((n - n // 2, 1, 3)) * L).astype(np.float32)}
t0 = time.perf_counter()
_, raw = oadf.adf_histograms(data, ["A", "B"], [L] * 3, np.array([0]), 6.0, 500, 4, 1, return_counts=True)
dt = time.perf_counter() - t0
tri = sum(int(v[0][1].sum()) for v in raw.values())
print(json.dumps({"oracle_adf_atoms": n, "s": dt, "triples": tri, "triples_per_s": tri / dt}))
