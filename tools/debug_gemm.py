import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from scregclust_b200 import api
rng = np.random.default_rng(0)
for (n, px, py) in [(512, 2048, 4096), (1000, 1600, 3000), (300, 700, 5000), (96, 3000, 3000)]:
    x = rng.standard_normal((n, px)); y = rng.standard_normal((n, py))
    ref = x.T @ y
    for rep in range(2):
        got = api.crossprod(x, y)
        e = np.abs(got - ref).max() / np.abs(ref).max()
        print(n, px, py, "rep", rep, "rel err", e, "n bad", int((np.abs(got - ref) > 1e-9).sum()), flush=True)
    g = api.crossprod(x)
    print(n, px, "sym rel err", np.abs(g - x.T @ x).max() / np.abs(g).max(), flush=True)
