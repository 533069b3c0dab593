"""An INDEPENDENT second restatement of the reference's 2-D time step, used only to cross-check the
C++ oracle (tests/test_oracle_independent.py).

Written from the cajitafluids sources (file:line below), not from oracle/cfo_oracle.cpp, and in a
different style on purpose: whole-array numpy over the reference's own ghosted (i, j) index space, an
explicitly assembled scipy.sparse pressure matrix (the reference stores its matrix too), and a
textbook PCG loop with numpy dot products.  Single rank, 2-D (the only dimension the reference
implements).  Cajita pieces (splines, LocalMesh::coordinates, the CG loop) follow the
published algorithm [Cajita-mem] exactly like the oracle has to — this file cannot pin those, it pins
the oracle's transcription of everything around them and gives a second, differently-ordered
floating-point evaluation of the whole path.

Arrays are indexed [i, j] (x first) in LOCAL GHOSTED indices like the reference's Kokkos views:
cells (n+2H, n+2H), I-faces (n+2H+1, n+2H), J-faces (n+2H, n+2H+1), H = 3 (src/Solver.hpp:78,
tests/tstMesh.cpp:61-68).
"""
import numpy as np
import scipy.sparse as sp

H = 3
SOLID, FREE = 0, 1


class Ref2D:
    def __init__(self, n, *, box=1.0, dt=0.005, density=0.1, inflow_loc=(0.2, 0.45), inflow_size=(0.02, 0.1),
                 inflow_vel=(1.0, 0.0), inflow_q=3.0, force=(0.0, 0.0), btype=(SOLID,) * 4, tol=1e-6, max_iter=2000,
                 quirk_q1=True, quirk_q2=True):
        self.n = n
        self.h = box / n                                   # src/Mesh.hpp:50-51
        self.low = -H * self.h                             # ghosted low corner of the local mesh
        self.rho = density
        self.force = np.array(force, dtype=float)
        self.src_box = (inflow_loc[0], inflow_loc[1], inflow_loc[0] + inflow_size[0], inflow_loc[1] + inflow_size[1])
        self.src_vel = inflow_vel                          # src/InflowSource.hpp:80-90
        self.src_q = inflow_q
        self.btype = btype                                 # [-x, -y, +x, +y]  src/BoundaryConditions.hpp:131
        self.bmin, self.bmax = (0, 0), (n - 1, n - 1)      # src/Solver.hpp:109-110
        self.tol, self.max_iter = tol, max_iter            # src/VelocityCorrector.hpp:103-104
        self.q1, self.q2 = quirk_q1, quirk_q2
        # dt clamp  src/Solver.hpp:96-106
        fmax_ = np.sqrt(force[0] ** 2 + force[1] ** 2)
        umax = max(abs(inflow_vel[0]), abs(inflow_vel[1])) + np.sqrt(fmax_ * self.h)
        self.dt = dt
        if umax > 0 and self.dt > self.h / umax:
            self.dt = self.h / umax
        m = n + 2 * H
        self.q = np.zeros((m, m))                          # src/ProblemManager.hpp:149-165: zero incl. ghosts
        self.u = np.zeros((m + 1, m))
        self.v = np.zeros((m, m + 1))
        self.p = np.zeros((m, m))
        self.time = 0.0
        self.cg_iterations = 0
        self.last_iters = 0
        self._build_matrix()

    # ---- geometry ------------------------------------------------------------------------------
    def own(self, ent):
        """owned local index ranges (Cajita::Own) of an entity: cells n x n, faces n+1 along the normal"""
        n = self.n
        ni = n + (1 if ent == "I" else 0)
        nj = n + (1 if ent == "J" else 0)
        return slice(H, H + ni), slice(H, H + nj)

    def coords(self, ent, i, j):
        """Cajita LocalMesh::coordinates of local entity (i, j)"""
        x = self.low + (i + (0.0 if ent == "I" else 0.5)) * self.h
        y = self.low + (j + (0.0 if ent == "J" else 0.5)) * self.h
        return x, y

    # ---- pressure matrix: src/VelocityCorrector.hpp:116-144 + src/BoundaryConditions.hpp:56-97 ----
    def _build_matrix(self):
        n = self.n
        scale = self.dt / (self.rho * self.h * self.h)
        gi, gj = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
        e = np.empty((5, n, n))
        e[0] = 4.0 * scale
        e[1:] = -1.0 * scale
        for cond, k, side in ((gi <= self.bmin[0], 1, 0), (gi > self.bmax[0] - 1, 2, 2),
                              (gj <= self.bmin[1], 3, 1), (gj > self.bmax[1] - 1, 4, 3)):
            e[k][cond] = 0.0
            if self.btype[side] == SOLID:
                e[0][cond] -= scale
        self.entries = e
        idx = (gi * n + gj)
        rows, cols, vals = [idx.ravel()], [idx.ravel()], [e[0].ravel()]
        for k, (di, dj) in zip((1, 2, 3, 4), ((-1, 0), (1, 0), (0, -1), (0, 1))):
            ok = (gi + di >= 0) & (gi + di < n) & (gj + dj >= 0) & (gj + dj < n)
            rows.append(idx[ok])
            cols.append(((gi + di) * n + (gj + dj))[ok])
            vals.append(e[k][ok])
            # off-domain couplings multiply ghost zeros in the reference; their coefficient is 0 too
            assert not e[k][~ok].any()
        self.A = sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(n * n, n * n))
        self.minv = 1.0 / e[0].ravel()                     # src/VelocityCorrector.hpp:178

    # ---- Cajita ReferenceConjugateGradient -------------------------------------------
    def pcg(self, b):
        A, Mi = self.A, self.minv
        x = np.zeros_like(b)
        r = b.copy()
        rr = float(r @ r)
        it = 0
        if np.sqrt(rr) <= self.tol:
            return x, it, np.sqrt(rr)
        z = Mi * r
        p = z.copy()
        zr_old = float(z @ r)
        q = A @ p
        pAp = float(p @ q)
        conv = False
        while it < self.max_iter:
            alpha = zr_old / pAp
            x += alpha * p
            r -= alpha * q
            rr = float(r @ r)
            it += 1
            if np.sqrt(rr) <= self.tol:
                conv = True
                break
            z = Mi * r
            zr_new = float(z @ r)
            beta = zr_new / zr_old
            p = z + beta * p
            q = A @ p
            pAp = float(p @ q)
            zr_old = zr_new
        if not conv:
            raise RuntimeError("CG solver did not converge")
        return x, it, np.sqrt(rr)

    # ---- Solver::_addInputs  src/Solver.hpp:181-263 ---------------------------------------------------
    def _in_box(self, x, y):
        b = self.src_box
        return (x >= b[0]) & (x < b[2]) & (y >= b[1]) & (y < b[3])          # src/InflowSource.hpp:40-41

    def _bc_face(self, ent, arr, si, sj):
        """BoundaryCondition::operator()(Face, ...)  src/BoundaryConditions.hpp:102-129 on owned faces"""
        ii = np.arange(si.start, si.stop) - H      # global index (single rank)
        jj = np.arange(sj.start, sj.stop) - H
        if ent == "I":
            if self.btype[0] == SOLID:
                arr[si, sj][ii <= self.bmin[0], :] = 0.0
            if self.btype[2] == SOLID:
                arr[si, sj][ii > self.bmax[0], :] = 0.0
        else:
            if self.btype[1] == SOLID:
                arr[si, sj][:, jj <= self.bmin[1]] = 0.0
            if self.btype[3] == SOLID:
                arr[si, sj][:, jj > self.bmax[1]] = 0.0

    def add_inputs(self):
        si, sj = self.own("C")
        I, J = np.meshgrid(np.arange(si.start, si.stop), np.arange(sj.start, sj.stop), indexing="ij")
        x, y = self.coords("C", I, J)
        qv = self.q[si, sj]
        m = self._in_box(x, y) & (qv < self.src_q)
        qv[m] = self.src_q
        for ent, arr, d in (("I", self.u, 0), ("J", self.v, 1)):
            si, sj = self.own(ent)
            I, J = np.meshgrid(np.arange(si.start, si.stop), np.arange(sj.start, sj.stop), indexing="ij")
            x, y = self.coords(ent, I, J)
            a = arr[si, sj]
            m = self._in_box(x, y) & (np.abs(a) < abs(self.src_vel[d]))
            a[m] = self.src_vel[d]
            a += self.force[d] * self.dt                                   # src/BodyForce.hpp:50,59
            self._bc_face(ent, arr, si, sj)

    # ---- VelocityCorrector  src/VelocityCorrector.hpp:182-282 ------------------------------------------
    def correct_velocity(self):
        n, h = self.n, self.h
        si, sj = self.own("C")
        u, v = self.u, self.v
        rhs = -(1.0 / h) * (u[H + 1:H + n + 1, sj] - u[H:H + n, sj] + v[si, H + 1:H + n + 1] - v[si, H:H + n])
        x, it, res = self.pcg(rhs.ravel())
        self.last_iters, self.last_resid = it, res
        self.cg_iterations += it
        self.p[:] = 0.0
        self.p[si, sj] = x.reshape(n, n)
        scale = self.dt / (self.rho * h)
        p = self.p
        ui, uj = self.own("I")
        u[ui, uj] -= scale * (p[ui, uj] - p[ui.start - 1:ui.stop - 1, uj])
        self._bc_face("I", u, ui, uj)
        vi, vj = self.own("J")
        v[vi, vj] -= scale * (p[vi, vj] - p[vi, vj.start - 1:vj.stop - 1])
        if self.q1:
            # src/VelocityCorrector.hpp:260: bc( FaceJ(), u, ... ) over the J-face index space:
            # zeroes U (not v) at local (i, j) with gj <= 0 or gj > max
            jj = np.arange(vj.start, vj.stop) - H
            if self.btype[1] == SOLID:
                u[vi, vj][:, jj <= self.bmin[1]] = 0.0
            if self.btype[3] == SOLID:
                u[vi, vj][:, jj > self.bmax[1]] = 0.0
        else:
            self._bc_face("J", v, vi, vj)

    # ---- Interpolation (Cajita splines, [Cajita-mem]) ---------------------------------------------------
    def _logical(self, ent, x, y):
        x0, y0 = self.coords(ent, 0, 0)
        return (x - x0) / self.h, (y - y0) / self.h

    @staticmethod
    def _w1(xl):
        i0 = xl.astype(np.int64)
        f = xl - i0
        return i0, [1.0 - f, f]

    @staticmethod
    def _w3(xl):
        i0 = xl.astype(np.int64)
        f = xl - i0
        w = [(1.0 - f) ** 3 / 6.0, (3.0 * f ** 3 - 6.0 * f ** 2 + 4.0) / 6.0,
             (-3.0 * f ** 3 + 3.0 * f ** 2 + 3.0 * f + 1.0) / 6.0, f ** 3 / 6.0]
        return i0 - 1, w

    def sample(self, ent, arr, x, y, order):
        xl, yl = self._logical(ent, x, y)
        fw = self._w1 if order == 1 else self._w3
        si, wi = fw(xl)
        sj, wj = fw(yl)
        out = np.zeros_like(x)
        for a, wa in enumerate(wi):
            for b, wb in enumerate(wj):
                ia = np.clip(si + a, 0, arr.shape[0] - 1)
                jb = np.clip(sj + b, 0, arr.shape[1] - 1)
                out += arr[ia, jb] * wa * wb
        return out

    # ---- TimeIntegrator  src/TimeIntegrator.hpp:36-177 ---------------------------------------------------
    def _rk3(self, x0, y0):
        dt, u, v = self.dt, self.u, self.v
        vel = lambda x, y: (self.sample("I", u, x, y, 1), self.sample("J", v, x, y, 1))
        v0 = vel(x0, y0)
        v1 = vel(x0 - 0.5 * dt * v0[0], y0 - 0.5 * dt * v0[1])
        if self.q2:     # :57-58: the third stage is taken along v0, not v1
            v2 = vel(x0 - 0.75 * dt * v0[0], y0 - 0.75 * dt * v0[1])
        else:
            v2 = vel(x0 - 0.75 * dt * v1[0], y0 - 0.75 * dt * v1[1])
        tx = x0 - dt * ((2.0 / 9.0) * v0[0] + (3.0 / 9.0) * v1[0] + (4.0 / 9.0) * v2[0])
        ty = y0 - dt * ((2.0 / 9.0) * v0[1] + (3.0 / 9.0) * v1[1] + (4.0 / 9.0) * v2[1])
        return tx, ty

    def advect_all(self, order=3):
        new = {}
        for ent, name in (("C", "q"), ("I", "u"), ("J", "v")):
            arr = getattr(self, name)
            si, sj = self.own(ent)
            I, J = np.meshgrid(np.arange(si.start, si.stop), np.arange(sj.start, sj.stop), indexing="ij")
            x0, y0 = self.coords(ent, I, J)
            tx, ty = self._rk3(x0, y0)
            nxt = np.zeros_like(arr)                    # the Next array's ghosts stay zero
            nxt[si, sj] = self.sample(ent, arr, tx, ty, order)
            new[name] = nxt
        self.q, self.u, self.v = new["q"], new["u"], new["v"]   # pm.advance x3

    # ---- Solver  src/Solver.hpp:125-147 ----------------------------------------------------------------
    def setup(self):
        self.add_inputs()
        self.correct_velocity()

    def step(self, order=3):
        self.advect_all(order)
        self.add_inputs()
        self.correct_velocity()
        self.time += self.dt

    def owned(self, name):
        """owned entities as [j, i] (y, x) like the C ABI's dense host arrays"""
        ent = {"q": "C", "u": "I", "v": "J", "p": "C"}[name]
        si, sj = self.own(ent)
        return getattr(self, name)[si, sj].T.copy()
