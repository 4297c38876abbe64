"""Kernel timeline of the window path at C2 (BSR_WIN_TRACE=1 prints per-window stage times); run under gpurun."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from mcmc_symreg_b200 import capi
w = bench.WORKLOADS["c2"]
X, y = bench.make_data(w)
eng = capi.Engine(3, 4096, list(range(1, 11)), [0.1] * 10, beta=-1.0, val=0, plateau_rule=False)
eng.set_data(X, y); eng.init_chains(w["seed"])
eng.run(3000)
os.environ["BSR_WIN_TRACE"] = "1"
eng.run(int(sys.argv[1]) if len(sys.argv) > 1 else 256)
