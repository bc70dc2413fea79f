"""Debug helper: one small nearest_neighbor case against brute force, printing what differs."""
import importlib, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from oracle import ref_port as O
icp = importlib.import_module("2dto3dreconstruction_b200.utils.icp")
n, m = int(sys.argv[1]), int(sys.argv[2])
rng = np.random.default_rng(n * 7 + m)
src, dst = rng.normal(size=(n, 3)) * 100, rng.normal(size=(m, 3)) * 100
for rep in range(3):
    d, i = icp.nearest_neighbor(src, dst)
    db, ib = O.nearest_neighbor_brute(src, dst)
    print("idx", i, "want", ib, "dist", d, "want", db)
