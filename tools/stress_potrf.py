import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import g3py_b200 as g3
from g3py_b200 import workloads
N = int(sys.argv[1]); B = int(sys.argv[2]); iters = int(sys.argv[3]); w = int(sys.argv[4]) if len(sys.argv) > 4 else 1 << 20
X, y, Theta = workloads.c2_inputs(N, B)
gp = g3.GP(X, g3.Bias(), g3.SE(X) + g3.MAT52(X)); gp.observed(X, y)
ctx = gp.ctx
ctx.set_potrf_block(w)
thk = gp._kernel_theta(gp.natural(Theta))
out, tiles = ctx.debug_potrf_stress(gp.desc, thk, iters)
print("G3_DBG", os.environ.get("G3_DBG"), "N", N, "B", B, "w", w, "bad iters:", [(i, out[i].tolist()) for i in range(iters) if out[i, 0]], flush=True)

if out[:, 0].any():
    bad, ref = tiles
    m = bad != ref
    rows = np.nonzero(m.any(axis=1))[0]; cols = np.nonzero(m.any(axis=0))[0]
    print("first bad tile: n mismatching elems", int(m.sum()), "rows", rows.min(), "..", rows.max(), "(%d distinct)" % len(rows), "cols", cols.min(), "..", cols.max(), "(%d distinct)" % len(cols))
    print("rows list", rows.tolist()[:70]); print("cols list", cols.tolist()[:70])
    rel = np.abs(bad - ref)[m] / np.maximum(np.abs(ref[m]), 1e-300)
    print("rel err min/median/max", rel.min(), np.median(rel), rel.max(), "any nan", np.isnan(bad).any())
