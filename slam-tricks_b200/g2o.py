"""g2o-shaped front door over the same engine: the vertex / edge / optimizer names of
`SolveWithG2O`, st20-g2o/src/include/test_g2o.h:94-147, so that the reference's second solve path can be replayed.

    VertexCamera     6-dof vertex, estimate = (SO3 quaternion xyzw, POS); oplus: SO3 <- SO3 * exp(v[0:3]), POS += v[3:6]
                     (test_g2o.h:36-39 — the same tangent convention as the Ceres path)
    VertexLandmark   3-dof vertex, `setMarginalized(True)` (test_g2o.h:121) = eliminated by the Schur complement
    EdgeProject      binary edge, error = Project(landmark) - measurement (test_g2o.h:73-80), information = I
    SparseOptimizer  addVertex / addEdge / initializeOptimization / optimize(iterations) / activeChi2

The reference uses g2o only as a comparator (no vertex is fixed, results are not written back, test_g2o.h:142-145),
so this is not a parity target: `optimize(n)` runs the engine's Ceres-style trust-region LM for at most n iterations
(g2o's own Levenberg damping schedule is not restated).  Gauge freedom is allowed: the LM diagonal keeps the reduced
system positive definite.
"""
import numpy as np

from . import capi, engine


class _Vertex:
    def __init__(self):
        self._id, self._fixed, self._estimate, self._marginalized = -1, False, None, False

    def setId(self, i):
        self._id = int(i)

    def id(self):
        return self._id

    def setFixed(self, fixed):
        self._fixed = bool(fixed)

    def fixed(self):
        return self._fixed

    def setMarginalized(self, m):
        self._marginalized = bool(m)

    def estimate(self):
        return self._estimate


class VertexCamera(_Vertex):
    def setEstimate(self, so3_xyzw, pos=None):
        """`setEstimate(OptPose)`: quaternion xyzw of the camera->world rotation and the position."""
        if pos is None:
            so3_xyzw, pos = so3_xyzw
        self._estimate = (np.array(so3_xyzw, dtype=np.float64), np.array(pos, dtype=np.float64))

    def Project(self, landmark):
        """test_g2o.h:28-33."""
        q, t = self._estimate
        x, y, z, w = q
        R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                      [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                      [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
        p = R.T @ (np.asarray(landmark, dtype=np.float64) - t)
        return p[:2] / p[2]


class VertexLandmark(_Vertex):
    def setEstimate(self, p):
        self._estimate = np.array(p, dtype=np.float64)


class EdgeProject:
    def __init__(self):
        self._v, self._z = [None, None], None

    def setVertex(self, i, v):
        self._v[i] = v

    def setMeasurement(self, z):
        self._z = np.array(z, dtype=np.float64)

    def setInformation(self, info):
        if not np.allclose(np.asarray(info), np.eye(2)):
            raise NotImplementedError("only the identity information matrix of test_g2o.h:128 is supported")

    def computeError(self):
        return self._v[0].Project(self._v[1].estimate()) - self._z


class SparseOptimizer:
    def __init__(self, device=0):
        self._cams, self._lms, self._edges, self._ready, self.summary, self.device, self.verbose = [], [], [], False, None, device, False

    def setVerbose(self, v):
        self.verbose = bool(v)

    def setAlgorithm(self, _solver=None):
        """`OptimizationAlgorithmLevenberg(BlockSolver<6,3>(LinearSolverCSparse))` — accepted; the engine's own solver runs."""

    def addVertex(self, v):
        (self._cams if isinstance(v, VertexCamera) else self._lms).append(v)
        return True

    def addEdge(self, e):
        self._edges.append(e)
        return True

    def initializeOptimization(self):
        self._ready = True
        return True

    def activeChi2(self):
        return float(sum(float(e.computeError() @ e.computeError()) for e in self._edges))

    def optimize(self, iterations):
        if not self._ready:
            raise RuntimeError("initializeOptimization() first (test_g2o.h:134)")
        cams = sorted(self._cams, key=lambda v: v.id()); lms = sorted(self._lms, key=lambda v: v.id())
        ci = {id(v): k for k, v in enumerate(cams)}; li = {id(v): k for k, v in enumerate(lms)}
        oc = np.array([ci[id(e._v[0])] for e in self._edges], dtype=np.int32)
        ol = np.array([li[id(e._v[1])] for e in self._edges], dtype=np.int32)
        uv = np.array([e._z for e in self._edges], dtype=np.float64).reshape(-1, 2)
        order = np.lexsort((oc, ol))                                  # landmark-major, camera-ascending
        q = np.array([v.estimate()[0] for v in cams]); t = np.array([v.estimate()[1] for v in cams]); p = np.array([v.estimate() for v in lms])
        cc = np.array([v.fixed() for v in cams], dtype=np.uint8); lc = np.array([v.fixed() for v in lms], dtype=np.uint8)
        with engine.BAEngine(q, t, p, oc[order], ol[order], uv[order], cc, lc if lc.any() else None, device=self.device) as e:
            self.summary = e.solve(capi.Options(max_num_iterations=int(iterations)))
            q, t, p = e.get_state()
        for k, v in enumerate(cams):
            v._estimate = (q[k].copy(), t[k].copy())
        for k, v in enumerate(lms):
            v._estimate = p[k].copy()
        if self.verbose:
            for it in self.summary.iterations:
                print("iteration= %d\t chi2= %.6f\t lambda= %.6g" % (it["iteration"], 2 * it["cost"], 1.0 / it["trust_region_radius"]))
        return len(self.summary.iterations) - 1
