import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import fuser_common as FC
import fuser_oracle as F
import ndt_feature_graph_b200 as N
from ndt_feature_graph_b200 import fuser as GF, laser as Ls
e = N.Engine(0)
d = FC.bag()
tr = Ls.TfTrack(d["odom_stamp"], d["odom"])
rng = np.random.default_rng(3)
idx = list(range(300, 300 + 13 * 9, 9))
clouds = [FC.cloud_of(d, i, rng) for i in idx]
fo = F.FuserOracle(FC.oracle_fuser_params(F), FC.SENSOR, F.MotionParams(**FC.MOTION))
fg = GF.NDTFeatureFuserHMT(e, FC.gpu_fuser_params(e))
fo.initialize(np.eye(4), clouds[0]); fg.initialize(np.eye(4), clouds[0])
last = tr.lookup(d["stamp"][idx[0]])
for i, c in zip(idx[1:], clouds[1:]):
    P = tr.lookup(d["stamp"][i]); Tm = F.pmul(F.pinv(last), P); last = P
    center, cell, size = fo.map.grid()
    fg.map.from_cells(center, cell, size, fo.map.export_cells(False), use_idx=True)
    fg.Tnow = fo.Tnow
    po = fo.map.export_cells(False); pg = fg.map.export_cells(False)
    for f in ("idx", "n", "has_gaussian", "occ", "mean", "cov"):
        if not np.array_equal(po[f], pg[f]): print('  SYNC DIFF', f, int((po[f] != pg[f]).sum()), np.abs(po[f].astype(float)-pg[f]).max())
    To = fo.update(Tm, c); Tg = fg.update(Tm, c)
    a, b = fo.map.export_cells(False), fg.map.export_cells(False)
    print(i, 'pose diff', np.abs(To - Tg).max(), 'cells', a.shape, b.shape)
    if a.shape != b.shape: continue
    for f in ("idx", "n", "has_gaussian", "occ", "mean", "cov"):
        if not np.array_equal(a[f], b[f]):
            bad = np.nonzero((a[f] != b[f]).reshape(len(a), -1).any(1))[0]
            print('  DIFF', f, len(bad), 'max abs', np.abs(a[f].astype(float)-b[f]).max(), 'first', a[bad[0]], b[bad[0]])
