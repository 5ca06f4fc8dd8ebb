"""Latency of the 128x128 factor+inverse kernel alone (csrc/diag.cu vs the first kernel in potrf.cu): microseconds per launch for
1 .. 148 CTAs, host-checked residuals, and the phase clocks of the new kernel."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from g3py_b200._cabi import Context

PHASES = ["load00"]
for kb in range(4):
    PHASES += [f"U{kb}", f"F{kb}(warp0)", f"F{kb}(fillers)", f"inv{kb}", f"below{kb}"]
PHASES += ["store"]

def main():
    ctx = Context(0)
    for shift in (0.05, -0.05):
        for variant in (1, 2):
            for B in (1, 8, 64, 148):
                us, st, err = ctx.debug_diag_time(variant, B, 20, shift)
                print(f"variant {variant}  B={B:4d}  shift={shift:g}  {us:8.2f} us/launch   |LL^T-A|/|A|={err[0]:.2e}  |XL-I|={err[1]:.2e}  "
                      f"dlogdet={err[2]:.2e}  info={int(err[3])}", flush=True)
                if variant == 2 and B == 1:
                    d = np.diff(st[: len(PHASES) + 1])
                    print("   phase cycles: " + "  ".join(f"{n}={int(c)}" for n, c in zip(PHASES, d)), flush=True)
                    print(f"   total cycles {int(st[len(PHASES)] - st[0])}", flush=True)
                    fstart = [st[2 + 5 * kb] for kb in range(4)]      # stamp after U_kb's barrier
                    for kb in range(4):
                        q = st[32 + 4 * kb: 32 + 4 * kb + 3]
                        fend = st[3 + 5 * kb]
                        print(f"   F{kb}: first half {int(q[0] - fstart[kb])}  dmma {int(q[1] - q[0])}  second half {int(q[2] - q[1])}  "
                              f"normalise {int(fend - q[2])}", flush=True)

if __name__ == "__main__":
    main()
