import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import g3py_b200 as g3
rows, B, launches, inplace = (int(a) for a in sys.argv[1:5])
kd = int(sys.argv[5]) if len(sys.argv) > 5 else 128
ctx = g3.Context(0)
out = ctx.debug_gemm_stress(rows, B, launches, inplace, kd)
print("G3_DBG", os.environ.get("G3_DBG"), "rows", rows, "B", B, "launches", launches, "inplace", inplace, "kd", kd,
      "-> bad launches %d, bad elems %d, first idx %d (row %d col %d) at launch %d" % (out[0], out[1], out[2], (out[2] // 128) if out[2] >= 0 else -1, out[2] % 128 if out[2] >= 0 else -1, out[3]), flush=True)
