"""Micro-benchmark of the registration kernel alone (GPU box): maps are built once, then the batched match is timed
with the library's own CUDA events.  NDTB_LIB selects the library variant.  usage: bench_match.py [pairs] [reps] [cov]"""
import os
import pickle
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ndt_feature_graph_b200 as N  # noqa: E402
from ndt_feature_graph_b200 import synth  # noqa: E402

pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 148
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
cache = f"/tmp/ndtb_workload_v2_{pairs}.pkl"
if os.path.exists(cache):
    tg, sr, T0s, Ds = pickle.load(open(cache, "rb"))
else:
    tg, sr, T0s, Ds = synth.velodyne_batch(pairs, n_base=8, seed=0)
    pickle.dump((tg, sr, T0s, Ds), open(cache, "wb"))
e = N.Engine(0)
mt = [N.NDTMap(e, 0.5) for _ in range(pairs)]
ms = [N.NDTMap(e, 0.5) for _ in range(pairs)]
t0 = time.perf_counter()
e.build_maps(mt + ms, tg + sr)
t_build = time.perf_counter() - t0
e.enable_timing(True)
res = None
times = []
prm = e.default_params(ctas_per_match=int(os.environ.get("CTAS", "0")), pass_budget=int(os.environ.get("BUDGET", "0")))
for r in range(reps + 1):
    res, _ = e.match_batch(mt, ms, T0s, prm)
    t, n = e.match_time()
    if r:
        times.append(t)
passes = (res["n_hess_passes"] + res["n_grad_passes"]).sum()
print(f"lib={os.path.basename(N.lib_path())} pairs={pairs} build(host->maps)={1e3 * t_build:.1f} ms match={np.mean(times):.2f} ms "
      f"(min {min(times):.2f}) passes={passes} hess={res['n_hess_passes'].sum()} iters_max={res['iterations'].max()} "
      f"chk={float(np.abs(res['T']).sum()):.12f} sm_ms: sum={res['kernel_ms'].sum():.1f} mean={res['kernel_ms'].mean():.2f} "
      f"max={res['kernel_ms'].max():.2f} | passes pctl 50/90/99/max={np.percentile(res['n_hess_passes'] + res['n_grad_passes'], [50, 90, 99, 100])}")
