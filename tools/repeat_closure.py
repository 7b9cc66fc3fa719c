"""Determinism check: the all-separator closure case (dense reduced system, n = 1188) solved many times."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["STBA_PG_CLOSURE_BAND"] = "2"
import numpy as np
import stba
from oracle import pg_oracle as pg
G = pg.make_graph(200, offsets=(1, 2, 3, 4))
res = set()
for r in range(int(sys.argv[1]) if len(sys.argv) > 1 else 40):
    with stba.posegraph.PoseGraph(G["q0"], G["t0"], G["ei"], G["ej"], G["zq"], G["zt"]) as p:
        s = p.solve()
        q, t = p.get_state()
    res.add((len(s.iterations), s.final_cost.hex(), float(np.abs(q).sum()).hex()))
print(len(res), "distinct outcomes:", sorted(res)[:4])
