"""An INDEPENDENT restatement of the time step in D = 2 or 3 dimensions, used only to cross-check the C++
oracle where the reference itself stops (tests/test_oracle_independent_3d.py): cajitafluids implements 2-D
only, the benchmark sizes are 3-D, and the 3-D oracle is "the natural extension" of SURVEY §8a —
a FaceK entity / w component, 7-point operator, 64-point splines — with nothing else to hold it against.

Written from the cajitafluids sources (file:line below) and SURVEY §8a's extension column, not from
oracle/cfo_oracle.cpp and not from tests/ref2d_numpy.py, dimension-generic on purpose (one code path for
D = 2 and 3: in 2-D it must agree with ref2d_numpy.py, which is how this file is itself pinned): whole-array
numpy over the ghosted index space with the axes kept in a tuple, the pressure matrix assembled with scipy.sparse
Kronecker products (the operator is a sum of 1-D second differences with modified end rows), a textbook PCG loop.
Cajita pieces (splines, LocalMesh coordinates, the CG loop) follow the published algorithm [Cajita-mem].

Arrays are indexed [i, j(, k)] (x first) in LOCAL GHOSTED indices, H = 3 (src/Solver.hpp:78).
Boundary types: btype[d] = low wall of dimension d, btype[D + d] = high wall (src/BoundaryConditions.hpp:131).
"""
import itertools

import numpy as np
import scipy.sparse as sp

H = 3
SOLID, FREE = 0, 1


class RefND:
    def __init__(self, cells, *, box=1.0, dt=0.005, density=0.1, inflow_loc=None, inflow_size=None, inflow_vel=None,
                 inflow_q=3.0, force=None, btype=None, tol=1e-6, max_iter=2000, quirk_q1=True, quirk_q2=True):
        self.n = tuple(cells)
        D = self.D = len(self.n)
        self.h = box / self.n[0]                                  # src/Mesh.hpp:50-51 (box = extent of dimension 0)
        self.low = -H * self.h
        self.rho = density
        lo = tuple(inflow_loc) if inflow_loc is not None else (0.2, 0.45, 0.45)[:D]   # examples/advection.cpp:179-182
        sz = tuple(inflow_size) if inflow_size is not None else (0.02, 0.1, 0.1)[:D]
        self.src_lo, self.src_hi = lo, tuple(a + b for a, b in zip(lo, sz))
        self.src_vel = tuple(inflow_vel) if inflow_vel is not None else (1.0, 0.0, 0.0)[:D]
        self.src_q = inflow_q
        self.force = tuple(force) if force is not None else (0.0,) * D
        self.btype = tuple(btype) if btype is not None else (SOLID,) * (2 * D)
        self.tol, self.max_iter = tol, max_iter                   # src/VelocityCorrector.hpp:103-104
        self.q1, self.q2 = quirk_q1, quirk_q2
        # dt clamp  src/Solver.hpp:96-106
        umax = max(abs(v) for v in self.src_vel) + np.sqrt(np.sqrt(sum(f * f for f in self.force)) * self.h)
        self.dt = dt
        if umax > 0 and self.dt > self.h / umax:
            self.dt = self.h / umax
        # fields: cells, then one face entity per dimension (+1 along its normal)   src/ProblemManager.hpp:149-165
        self.q = np.zeros(self.shape(None))
        self.vel = [np.zeros(self.shape(d)) for d in range(D)]
        self.p = np.zeros(self.shape(None))
        self.time, self.cg_iterations, self.last_iters = 0.0, 0, 0
        self._assemble()

    # ---- index spaces and coordinates ---------------------------------------------------------------------
    def shape(self, ent):
        """ghosted extents of the cell entity (ent None) or of the faces normal to dimension ent"""
        return tuple(n + 2 * H + (1 if ent == d else 0) for d, n in enumerate(self.n))

    def own(self, ent):
        return tuple(slice(H, H + n + (1 if ent == d else 0)) for d, n in enumerate(self.n))

    def own_grid(self, ent):
        """local indices of the owned entities, one broadcastable array per dimension"""
        return np.meshgrid(*[np.arange(s.start, s.stop) for s in self.own(ent)], indexing="ij", sparse=True)

    def coords(self, ent, idx):
        """Cajita LocalMesh::coordinates: faces sit on the low side of their cell along the normal"""
        return [self.low + (idx[d] + (0.0 if ent == d else 0.5)) * self.h for d in range(self.D)]

    # ---- pressure matrix: src/VelocityCorrector.hpp:116-144, src/BoundaryConditions.hpp:56-97 -----------------
    def _assemble(self):
        D, scale = self.D, self.dt / (self.rho * self.h * self.h)
        self.scale = scale
        # off-diagonal part: sum_d I x ... x N_d x ... x I with N_d = -scale on the two off-diagonals (couplings to
        # ghosts do not exist: the reference zeroes those coefficients)
        A = None
        for d in range(D):
            term = None
            for e in range(D):
                n = self.n[e]
                f = sp.diags([np.full(n - 1, -scale), np.full(n - 1, -scale)], [-1, 1], format="csr") if e == d \
                    else sp.identity(n, format="csr")
                term = f if term is None else sp.kron(term, f, format="csr")   # x slowest: C order of [i, j, k]
            A = term if A is None else A + term
        # diagonal in the reference's own sequence (:137, BoundaryConditions.hpp:56-97): 2 D scale, then one
        # `scale` off for every SOLID wall the cell touches
        diag = np.full(self.n, 2.0 * D * scale)
        for d in range(D):
            sl_lo = tuple(0 if e == d else slice(None) for e in range(D))
            sl_hi = tuple(self.n[d] - 1 if e == d else slice(None) for e in range(D))
            if self.btype[d] == SOLID:
                diag[sl_lo] -= scale
            if self.btype[D + d] == SOLID:
                diag[sl_hi] -= scale
        A = A + sp.diags(diag.ravel(), 0, format="csr")
        self.A = A.tocsr()
        self.minv = 1.0 / self.A.diagonal()                       # src/VelocityCorrector.hpp:178

    # ---- Cajita ReferenceConjugateGradient ------------------------------------------------------
    def pcg(self, b):
        A, Mi = self.A, self.minv
        x = np.zeros_like(b)
        r = b.copy()
        res = np.sqrt(r @ r)
        if res <= self.tol:
            return x, 0, res
        z = Mi * r
        p = z.copy()
        zr = z @ r
        for it in range(1, self.max_iter + 1):
            q = A @ p
            alpha = zr / (p @ q)
            x += alpha * p
            r -= alpha * q
            res = np.sqrt(r @ r)
            if res <= self.tol:
                return x, it, res
            z = Mi * r
            zr, zr_old = z @ r, zr
            p = z + (zr / zr_old) * p
        raise RuntimeError("CG solver did not converge")

    # ---- Solver::_addInputs  src/Solver.hpp:181-263 -----------------------------------------------------------
    def _in_source(self, x):
        m = True
        for d in range(self.D):
            m = m & (x[d] >= self.src_lo[d]) & (x[d] < self.src_hi[d])        # src/InflowSource.hpp:40-41
        return m

    def _wall_mask(self, d, idx):
        """BoundaryCondition::operator()(Face...) src/BoundaryConditions.hpp:102-129: entities of index space idx whose
        global index along d is <= min (low SOLID wall) or > max (high SOLID wall); min = 0, max = n - 1 (Solver.hpp:109-110)"""
        g = idx[d] - H
        m = np.zeros(np.broadcast(*idx).shape, dtype=bool)
        if self.btype[d] == SOLID:
            m |= np.broadcast_to(g <= 0, m.shape)
        if self.btype[self.D + d] == SOLID:
            m |= np.broadcast_to(g > self.n[d] - 1, m.shape)
        return m

    def add_inputs(self):
        idx = self.own_grid(None)
        qv = self.q[self.own(None)]
        qv[self._in_source(self.coords(None, idx)) & (qv < self.src_q)] = self.src_q
        for d in range(self.D):
            idx = self.own_grid(d)
            a = self.vel[d][self.own(d)]
            m = self._in_source(self.coords(d, idx)) & (np.abs(a) < abs(self.src_vel[d]))
            a[m] = self.src_vel[d]
            a += self.force[d] * self.dt                                       # src/BodyForce.hpp:50,59
            a[self._wall_mask(d, idx)] = 0.0

    # ---- VelocityCorrector  src/VelocityCorrector.hpp:182-282 -------------------------------------------------
    def _shifted(self, sl, d, by):
        return tuple(slice(s.start + by, s.stop + by) if e == d else s for e, s in enumerate(sl))

    def correct_velocity(self):
        D, h = self.D, self.h
        oc = self.own(None)
        div = 0.0
        for d in range(D):
            div = div + (self.vel[d][self._shifted(oc, d, 1)] - self.vel[d][oc])
        rhs = -(1.0 / h) * div                                                 # :199-211
        x, it, res = self.pcg(rhs.ravel())
        self.last_iters, self.last_resid = it, res
        self.cg_iterations += it
        self.p[...] = 0.0
        self.p[oc] = x.reshape(self.n)
        scale = self.dt / (self.rho * h)                                       # :217
        for d in range(D):
            of = self.own(d)
            idx = self.own_grid(d)
            self.vel[d][of] -= scale * (self.p[of] - self.p[self._shifted(of, d, -1)])
            if d == 1 and self.q1:
                # :260: the FaceJ kernel hands `u` to the boundary functor: the FaceJ wall test zeroes u(i, j, k)
                self.vel[0][of][self._wall_mask(1, idx)] = 0.0
            else:
                self.vel[d][of][self._wall_mask(d, idx)] = 0.0

    # ---- Interpolation (Cajita splines, [Cajita-mem]) ---------------------------------------------------------
    @staticmethod
    def _weights(xl, order):
        i0 = np.floor(xl).astype(np.int64)
        f = xl - i0
        if order == 1:
            return i0, [1.0 - f, f]
        return i0 - 1, [(1.0 - f) ** 3 / 6.0, (3.0 * f ** 3 - 6.0 * f ** 2 + 4.0) / 6.0,
                        (-3.0 * f ** 3 + 3.0 * f ** 2 + 3.0 * f + 1.0) / 6.0, f ** 3 / 6.0]

    def sample(self, ent, arr, x, order):
        origin = self.coords(ent, [0] * self.D)
        base, w = zip(*[self._weights((x[d] - origin[d]) / self.h, order) for d in range(self.D)])
        out = 0.0
        for offs in itertools.product(range(order + 1), repeat=self.D):
            ii = tuple(np.clip(base[d] + offs[d], 0, arr.shape[d] - 1) for d in range(self.D))
            wt = w[0][offs[0]]
            for d in range(1, self.D):
                wt = wt * w[d][offs[d]]
            out = out + arr[ii] * wt
        return out

    # ---- TimeIntegrator  src/TimeIntegrator.hpp:36-177 --------------------------------------------------------
    def _velocity_at(self, x):
        return [self.sample(d, self.vel[d], x, 1) for d in range(self.D)]

    def _backtrace(self, x0):
        D, dt = self.D, self.dt
        v0 = self._velocity_at(x0)
        v1 = self._velocity_at([x0[d] - 0.5 * dt * v0[d] for d in range(D)])
        vs = v0 if self.q2 else v1                                             # :57-58
        v2 = self._velocity_at([x0[d] - 0.75 * dt * vs[d] for d in range(D)])
        return [x0[d] - dt * ((2.0 / 9.0) * v0[d] + (3.0 / 9.0) * v1[d] + (4.0 / 9.0) * v2[d]) for d in range(D)]

    def advect_all(self, order=3):
        new = []
        for ent, arr in [(None, self.q)] + [(d, self.vel[d]) for d in range(self.D)]:
            x0 = [np.broadcast_to(c, tuple(s.stop - s.start for s in self.own(ent)))
                  for c in self.coords(ent, self.own_grid(ent))]
            nxt = np.zeros_like(arr)                                           # the Next array's ghosts stay zero
            nxt[self.own(ent)] = self.sample(ent, arr, self._backtrace(x0), order)
            new.append(nxt)
        self.q, self.vel = new[0], new[1:]                                     # pm.advance

    # ---- Solver  src/Solver.hpp:125-147 -------------------------------------------------------------------------
    def setup(self):
        self.add_inputs()
        self.correct_velocity()

    def step(self, order=3):
        self.advect_all(order)
        self.add_inputs()
        self.correct_velocity()
        self.time += self.dt

    def output(self):
        """what SiloWriter hands to Silo (src/SiloWriter.hpp:136-197): the owned quantity and the velocity interpolated
        (order 1) to the cell centres, as [(k,) j, i] arrays"""
        x = [np.broadcast_to(c, self.n) for c in self.coords(None, self.own_grid(None))]
        return self.owned("q"), [np.ascontiguousarray(self.sample(d, self.vel[d], x, 1).T) for d in range(self.D)]

    def owned(self, name):
        """owned entities as [(k,) j, i] like the C ABI's dense host arrays"""
        ent = {"q": None, "u": 0, "v": 1, "w": 2, "p": None}[name]
        arr = {"q": self.q, "p": self.p}.get(name)
        if arr is None:
            arr = self.vel[ent]
        return np.ascontiguousarray(arr[self.own(ent)].T)
