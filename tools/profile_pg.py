"""ncu driver: the 10 000-pose / 39 990-edge pose graph (BASELINE.json configs[4]), one LM solve."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import stba
G = stba.synth.pose_graph(10000)
with stba.posegraph.PoseGraph(G["q0"], G["t0"], G["ei"], G["ej"], G["zq"], G["zt"]) as p:
    s = p.solve()
    print(s.BriefReport(), "%.1f ms" % s.total_time_ms, s.gpu_launches, "launches")
