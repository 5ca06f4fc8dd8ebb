import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np
import g3py_b200 as g3
from g3py_b200 import workloads
X3, y, _ = workloads.c2_inputs(4096, 1)
gp = g3.GP(X3, g3.Bias(), g3.SE(X3)); gp.observed(X3, y)
th = gp.dict_to_array(gp.params_default)
for _ in range(3): gp.dlogp(th, array=True)
