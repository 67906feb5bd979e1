"""An INDEPENDENT numpy / scipy restatement of the two registrations on the hot path, used to pin the C++ oracle.

The C++ oracle (oracle/*.cpp) and the CUDA product were written by the same hand and share their restatements of the
Eigen / FLANN algorithms the reference reaches through un-vendored libraries (Jacobi SVD, tridiagonal-QR eigen solver, LDLT,
kd-tree).  A shared misreading of one of those would be invisible to every GPU-vs-oracle test.  This module states the same
published algorithms a second time with none of that code:

  * linear algebra   numpy.linalg.eigh / svd / solve / inv (LAPACK) instead of the hand-written Jacobi / QR / LDLT;
  * neighbours       scipy.spatial.cKDTree instead of the oracle's kd-tree;
  * rotations        scipy.spatial.transform.Rotation for so3_exp; the NDT pose derivatives are taken from products of
                     derivative rotation matrices (d/da of Rx Ry Rz), not from the 8 + 15 hand-expanded table rows of
                     NDT:329-392;
  * arithmetic       float64 throughout, where the reference (and the oracle, and the product) evaluate the NDT terms in
                     float32: the comparison tolerance (1e-6) is what that difference leaves.

What it deliberately keeps, because it IS the algorithm: pcl::VoxelGrid / VoxelGridCovariance index arithmetic in f32
(VGC:67-103, 218-223), the single-pass covariance and the eigenvalue inflation (VGC:282-367), the DIRECT7 neighbourhood
(VGC:373-433), the Gaussian constants (NDT:86-93), the term rejection rule (NDT:505-506), the Newton / More-Thuente driver
(NDT:103-171, 771-931), fast_gicp's covariance regularisation (FG:273-293), its Mahalanobis fusion (FG:146-150) and the
Levenberg-Marquardt loop of LsqRegistration (LSQ:53-172).

tests/golden/make_reference_fixtures.py runs this on the bundled Velodyne pair and freezes the results;
tests/test_oracle_cpu.py compares the C++ oracle with them (and re-runs the cheap parts live).
"""
import numpy as np
from scipy.spatial import cKDTree
from scipy.spatial.transform import Rotation

F32 = np.float32


# --------------------------------------------------------------------------------------------------------------------------
# pcl::VoxelGrid (centroids in ascending voxel index), the preprocessing of the reference's gtest (gicp_test.cpp:55-65)
def voxel_grid(pts, leaf):
    p = np.asarray(pts, F32)
    inv = F32(1.0) / F32(leaf)
    xyz = p[:, :3]
    mn, mx = xyz.min(0), xyz.max(0)
    min_b = np.floor(mn * inv).astype(np.int64)
    max_b = np.floor(mx * inv).astype(np.int64)
    div_b = max_b - min_b + 1
    mul = np.array([1, div_b[0], div_b[0] * div_b[1]], np.int64)
    ijk = (np.floor(xyz * inv) - min_b.astype(F32)).astype(np.int64)
    idx = ijk @ mul
    order = np.argsort(idx, kind="stable")
    sidx = idx[order]
    starts = np.flatnonzero(np.r_[True, sidx[1:] != sidx[:-1]])
    out = np.empty((len(starts), 4), F32)
    ends = np.r_[starts[1:], len(sidx)]
    for k, (b, e) in enumerate(zip(starts, ends)):  # f32 running sums in ascending point index (CentroidPoint)
        acc = np.zeros(4, F32)
        for row in p[order[b:e]]:
            acc = acc + row
        out[k] = acc / F32(e - b)
    return out


# --------------------------------------------------------------------------------------------------------------------------
# NDT
def gauss_constants(outlier_ratio, resolution):  # NDT:86-93
    c1 = 10.0 * (1 - outlier_ratio)
    c2 = outlier_ratio / resolution ** 3
    d3 = -np.log(c2)
    d1 = -np.log(c1 + c2) - d3
    d2 = -2 * np.log((-np.log(c1 * np.exp(-0.5) + c2) - d3) / d1)
    return d1, d2, d3


class VoxelGridCovariance:
    """VGC:48-370 with LAPACK's symmetric eigensolver; min_points_per_voxel = 6, eigenvalue inflation 0.01."""

    def __init__(self, target, resolution):
        xyz = np.asarray(target, F32)[:, :3]
        self.res = F32(resolution)
        inv = F32(1.0) / self.res
        mn, mx = xyz.min(0), xyz.max(0)
        self.min_b = np.floor(mn * inv).astype(np.int64)
        self.max_b = np.floor(mx * inv).astype(np.int64)
        div_b = self.max_b - self.min_b + 1
        self.mul = np.array([1, div_b[0], div_b[0] * div_b[1]], np.int64)
        ijk = (np.floor(xyz * inv) - self.min_b.astype(F32)).astype(np.int64)
        idx = ijk @ self.mul
        order = np.argsort(idx, kind="stable")
        sidx = idx[order]
        starts = np.flatnonzero(np.r_[True, sidx[1:] != sidx[:-1]])
        ends = np.r_[starts[1:], len(sidx)]
        x = xyz[order].astype(np.float64)
        self.leaf_idx = sidx[starts]
        self.n = (ends - starts).astype(np.int64)
        self.mean = np.zeros((len(starts), 3))
        self.icov = np.zeros((len(starts), 3, 3))
        self.valid = np.zeros(len(starts), bool)
        for k, (b, e) in enumerate(zip(starts, ends)):
            n = e - b
            pts = x[b:e]
            s = np.zeros(3)
            ss = np.zeros((3, 3))
            for q in pts:  # serial accumulation in point order, as the reference's first pass
                s = s + q
                ss = ss + np.outer(q, q)
            mean = s / n
            self.mean[k] = mean
            if n < 6:
                continue
            cov = (ss - 2 * np.outer(s, mean)) / n + np.outer(mean, mean)  # VGC:329
            cov = cov * ((n - 1.0) / n)                                    # VGC:330
            w, V = np.linalg.eigh(cov)
            if w[0] < 0 or w[1] < 0 or w[2] <= 0:
                continue
            min_ev = 0.01 * w[2]
            if w[0] < min_ev:
                w[0] = min_ev
                if w[1] < min_ev:
                    w[1] = min_ev
                cov = V @ np.diag(w) @ np.linalg.inv(V)
            icov = np.linalg.inv(cov)
            if not np.isfinite(icov).all():
                continue
            self.icov[k] = icov
            self.valid[k] = True
        self.lookup = {int(i): k for k, i in enumerate(self.leaf_idx) if self.valid[k]}

    OFFSETS7 = np.array([[0, 0, 0], [1, 0, 0], [-1, 0, 0], [0, 1, 0], [0, -1, 0], [0, 0, 1], [0, 0, -1]], np.int64)  # VGC:423-430

    def neighbours7(self, xt32):
        """(point index, voxel slot) pairs of getNeighborhoodAtPoint7 for transformed f32 points."""
        cell = np.floor(xt32 / self.res).astype(np.int64)  # f32 division, VGC:379-381
        pi, vs = [], []
        for off in self.OFFSETS7:
            c = cell + off
            inside = np.all((c >= self.min_b) & (c <= self.max_b), axis=1)
            lin = (c - self.min_b) @ self.mul
            slots = np.array([self.lookup.get(int(v), -1) if ok else -1 for v, ok in zip(lin, inside)], np.int64)
            hit = slots >= 0
            pi.append(np.flatnonzero(hit))
            vs.append(slots[hit])
        return np.concatenate(pi), np.concatenate(vs)


def _rot_and_derivatives(angles, snap):
    """R = Rx Ry Rz with its first and second derivatives w.r.t. (roll, pitch, yaw) as products of derivative matrices.
    snap: the reference replaces (cos, sin) by (1, 0) for |angle| < 1e-4 in the DERIVATIVES (NDT:293-326)."""
    def axis(i, a, order):
        c, s = np.cos(a), np.sin(a)
        if snap and abs(a) < 10e-5:
            c, s = 1.0, 0.0
        if order == 1:
            c, s = -s, c
        elif order == 2:
            c, s = -c, -s
        M = np.eye(3) if order == 0 else np.zeros((3, 3))
        j, k = (i + 1) % 3, (i + 2) % 3
        M[j, j], M[j, k], M[k, j], M[k, k] = c, -s, s, c
        return M
    R = [[axis(i, angles[i], o) for o in range(3)] for i in range(3)]
    def prod(o):
        return R[0][o[0]] @ R[1][o[1]] @ R[2][o[2]]
    d1 = [prod([1 if i == a else 0 for i in range(3)]) for a in range(3)]
    d2 = [[prod([(1 if i == a else 0) + (1 if i == b else 0) for i in range(3)]) for b in range(3)] for a in range(3)]
    # A quirk of the reference that this independent derivation exposed: row d1 of its table (NDT:360, 382) reads
    # (-cy cz, cy sz, +sy) where the analytic d2/dpitch2 of the first row of R is (-cy cz, cy sz, -sy).  The registration the
    # reference performs uses +sy, so parity means +sy; everything else of the 8 + 15 rows equals the analytic derivatives.
    sy = 0.0 if (snap and abs(angles[1]) < 10e-5) else np.sin(angles[1])
    assert abs(d2[1][1][0, 2] + sy) < 1e-15
    d2[1][1][0, 2] = sy
    return d1, d2


def convert_transform(p):
    """NDT.h:214-231 in f32: Translation * AngleAxis(roll, X) * AngleAxis(pitch, Y) * AngleAxis(yaw, Z)."""
    def ax(i, a):
        a = F32(a)
        c, s = F32(np.cos(a, dtype=F32)), F32(np.sin(a, dtype=F32))
        M = np.eye(3, dtype=F32)
        j, k = (i + 1) % 3, (i + 2) % 3
        M[j, j], M[j, k], M[k, j], M[k, k] = c, -s, s, c
        return M
    R = (ax(0, p[3]) @ ax(1, p[4])).astype(F32) @ ax(2, p[5])
    T = np.eye(4, dtype=F32)
    T[:3, :3] = R
    T[:3, 3] = np.asarray(p[:3], F32)
    return T


def transform_points_f32(T, xyz):  # pcl::transformPointCloud's order: c0 x + (c1 y + (c2 z + c3))
    x, y, z = xyz[:, 0], xyz[:, 1], xyz[:, 2]
    out = np.empty_like(xyz)
    for r in range(3):
        out[:, r] = T[r, 0] * x + (T[r, 1] * y + (T[r, 2] * z + T[r, 3]))
    return out


class NDT:
    def __init__(self, resolution=1.0, step_size=0.1, trans_eps=0.1, max_iter=35, outlier_ratio=0.55):
        self.res, self.step_size, self.trans_eps, self.max_iter, self.outlier_ratio = resolution, step_size, trans_eps, max_iter, outlier_ratio

    def set_target(self, target):
        self.grid = VoxelGridCovariance(target, self.res)

    def set_source(self, source):
        self.src = np.asarray(source, F32)[:, :3]

    def derivatives(self, p, T, want_hessian=True):
        """computeDerivatives (NDT:179-285) in f64: score, gradient (6), Hessian (6x6), accepted terms."""
        d1, d2, _ = self.gauss
        xt32 = transform_points_f32(T, self.src)
        pi, vs = self.grid.neighbours7(xt32)
        x0 = self.src[pi].astype(np.float64)
        xx = xt32[pi].astype(np.float64) - self.grid.mean[vs]
        C = self.grid.icov[vs]
        dR, ddR = _rot_and_derivatives(p[3:6], snap=True)
        J = np.zeros((len(pi), 3, 6))
        J[:, 0, 0] = J[:, 1, 1] = J[:, 2, 2] = 1.0
        for a in range(3):
            J[:, :, 3 + a] = x0 @ dR[a].T
        Cx = np.einsum("nij,nj->ni", C, xx)
        q = np.einsum("ni,ni->n", xx, Cx)
        e = np.exp(-d2 * q / 2)
        e2 = d2 * e
        ok = ~((e2 > 1) | (e2 < 0) | np.isnan(e2))  # NDT:505-506
        score = float(np.sum(-d1 * e[ok]))
        w = (d1 * e2)[ok]
        xCJ = np.einsum("ni,nij->nj", Cx[ok], J[ok])
        g = np.einsum("n,nj->j", w, xCJ)
        H = np.zeros((6, 6))
        if want_hessian:
            CJ = np.einsum("nij,njk->nik", C[ok], J[ok])
            JCJ = np.einsum("nia,nib->nab", J[ok], CJ)
            xCH = np.zeros((int(ok.sum()), 6, 6))
            for a in range(3):
                for b in range(3):
                    xCH[:, 3 + a, 3 + b] = np.einsum("ni,ni->n", Cx[ok], x0[ok] @ ddR[a][b].T)
            H = np.einsum("n,nab->ab", w, -d2 * xCJ[:, :, None] * xCJ[:, None, :] + xCH + JCJ)
        return score, g, H, int(ok.sum())

    def align(self, guess=None):
        """computeTransformation (NDT:80-171) + computeStepLengthMT (NDT:771-931)."""
        self.gauss = gauss_constants(self.outlier_ratio, self.res)
        T = np.eye(4, dtype=F32) if guess is None else np.asarray(guess, F32)
        R = Rotation.from_matrix(T[:3, :3].astype(np.float64))
        p = np.r_[T[:3, 3].astype(np.float64), R.as_euler("XYZ")]  # eulerAngles(0, 1, 2): intrinsic x-y'-z''
        st = dict(evals=0, trials=0, hess=0)
        score, g, H, _ = self.derivatives(p, T)
        st["evals"] += 1
        it, converged = 0, False
        n_in = float(len(self.src))
        while not converged:
            U, s, Vt = np.linalg.svd(H)
            thr = max(s[0] * 6 * np.finfo(float).eps, np.finfo(float).tiny)
            rank = int((s >= thr).sum())
            delta = Vt[:rank].T @ ((U[:, :rank].T @ (-g)) / s[:rank])
            nrm = np.linalg.norm(delta)
            if nrm == 0 or np.isnan(nrm):
                converged = not np.isnan(nrm)
                break
            d = delta / nrm
            # ---- More-Thuente (NDT:771-931)
            step_max, step_min = self.step_size, self.trans_eps / 2
            phi0, dphi0 = -score, -(g @ d)
            step = None
            if dphi0 >= 0:
                if dphi0 == 0:
                    step = 0.0
                else:
                    dphi0, d = -dphi0, -d
            if step is None:
                mu, nu = 1e-4, 0.9
                a_l = a_u = 0.0
                f_l = f_u = 0.0
                g_l = g_u = dphi0 - mu * dphi0
                interval_converged, open_interval = (step_max - step_min) < 0, True
                a_t = max(min(nrm, step_max), step_min)
                x_t = p + d * a_t
                T = convert_transform(x_t)
                score, g, H, _ = self.derivatives(x_t, T)
                st["evals"] += 1
                phi_t, dphi_t = -score, -(g @ d)
                psi_t, dpsi_t = phi_t - phi0 - mu * dphi0 * a_t, dphi_t - mu * dphi0
                k = 0
                while not interval_converged and k < 10 and not (psi_t <= 0 and dphi_t <= -nu * dphi0):
                    st["trials"] += 1
                    ft, gt = (psi_t, dpsi_t) if open_interval else (phi_t, dphi_t)
                    a_t = _trial_value(a_l, f_l, g_l, a_u, f_u, g_u, a_t, ft, gt)
                    a_t = max(min(a_t, step_max), step_min)
                    x_t = p + d * a_t
                    T = convert_transform(x_t)
                    score, g, _, _ = self.derivatives(x_t, T, want_hessian=False)
                    st["evals"] += 1
                    phi_t, dphi_t = -score, -(g @ d)
                    psi_t, dpsi_t = phi_t - phi0 - mu * dphi0 * a_t, dphi_t - mu * dphi0
                    if open_interval and psi_t <= 0 and dpsi_t >= 0:
                        open_interval = False
                        f_l, g_l = f_l + phi0 - mu * dphi0 * a_l, g_l + mu * dphi0
                        f_u, g_u = f_u + phi0 - mu * dphi0 * a_u, g_u + mu * dphi0
                    ft, gt = (psi_t, dpsi_t) if open_interval else (phi_t, dphi_t)
                    # updateIntervalMT (NDT:647-685)
                    if ft > f_l:
                        a_u, f_u, g_u = a_t, ft, gt
                    elif gt * (a_l - a_t) > 0:
                        a_l, f_l, g_l = a_t, ft, gt
                    elif gt * (a_l - a_t) < 0:
                        a_u, f_u, g_u = a_l, f_l, g_l
                        a_l, f_l, g_l = a_t, ft, gt
                    else:
                        interval_converged = True
                    k += 1
                if k:
                    _, _, H, _ = self.derivatives(x_t, T)  # computeHessian (NDT:539-644), f64 in the reference as well
                    st["hess"] += 1
                step = a_t
            p = p + d * step
            if it > self.max_iter or (it and abs(step) < self.trans_eps):
                converged = True
            it += 1
        self.final_T, self.iterations, self.converged, self.stats = T, it, converged, st
        self.trans_probability = score / n_in
        return T


def _cubic(a0, f0, g0, a1, f1, g1):
    z = 3 * (f1 - f0) / (a1 - a0) - g1 - g0
    w = np.sqrt(z * z - g1 * g0)
    return a0 + (a1 - a0) * (w - g0 - z) / (g1 - g0 + 2 * w)


def _trial_value(a_l, f_l, g_l, a_u, f_u, g_u, a_t, f_t, g_t):  # trialValueSelectionMT (NDT:688-768)
    if f_t > f_l:
        a_c = _cubic(a_l, f_l, g_l, a_t, f_t, g_t)
        a_q = a_l - 0.5 * (a_l - a_t) * g_l / (g_l - (f_l - f_t) / (a_l - a_t))
        return a_c if abs(a_c - a_l) < abs(a_q - a_l) else 0.5 * (a_q + a_c)
    if g_t * g_l < 0:
        a_c = _cubic(a_l, f_l, g_l, a_t, f_t, g_t)
        a_s = a_l - (a_l - a_t) / (g_l - g_t) * g_l
        return a_c if abs(a_c - a_t) >= abs(a_s - a_t) else a_s
    if abs(g_t) <= abs(g_l):
        a_c = _cubic(a_l, f_l, g_l, a_t, f_t, g_t)
        a_s = a_l - (a_l - a_t) / (g_l - g_t) * g_l
        a_next = a_c if abs(a_c - a_t) < abs(a_s - a_t) else a_s
        lim = a_t + 0.66 * (a_u - a_t)
        return min(lim, a_next) if a_t > a_l else max(lim, a_next)
    return _cubic(a_u, f_u, g_u, a_t, f_t, g_t)


# --------------------------------------------------------------------------------------------------------------------------
# fast_gicp::FastGICP
def gicp_covariances(pts, k=20):
    """calculate_covariances (FG:241-298), PLANE regularisation: U diag(1, 1, 1e-3) V^T of the k-NN covariance."""
    xyz = np.asarray(pts, F32)[:, :3].astype(np.float64)
    _, nb = cKDTree(xyz).query(xyz, k=k)
    P = xyz[nb]
    Pc = P - P.mean(axis=1, keepdims=True)
    cov = np.einsum("nki,nkj->nij", Pc, Pc) / k
    U, _, Vt = np.linalg.svd(cov)
    return np.einsum("nij,j,njk->nik", U, np.array([1.0, 1.0, 1e-3]), Vt)


class FastGICP:
    def __init__(self, k=20, max_corr=np.finfo(F32).max, max_iter=64, rot_eps=2e-3, trans_eps=5e-4):
        self.k, self.max_corr, self.max_iter, self.rot_eps, self.trans_eps = k, float(max_corr), max_iter, rot_eps, trans_eps

    def set_target(self, pts):
        self.tgt = np.asarray(pts, F32)[:, :3]
        self.tgt_tree = cKDTree(self.tgt.astype(np.float64))
        self.tgt_cov = gicp_covariances(pts, self.k)

    def set_source(self, pts):
        self.src = np.asarray(pts, F32)[:, :3]
        self.src_cov = gicp_covariances(pts, self.k)

    def _linearize(self, T):
        Tf = T.astype(F32)
        q = (self.src @ Tf[:3, :3].T + Tf[:3, 3]).astype(np.float64)  # update_correspondences: f32 transform (FG:128-131)
        d, j = self.tgt_tree.query(q, k=1)
        ok = d * d < self.max_corr ** 2
        self.corr = np.where(ok, j, -1)
        R = T[:3, :3]
        RCR = self.tgt_cov[j[ok]] + R @ self.src_cov[ok] @ R.T
        self.M = np.linalg.inv(RCR)
        return self._terms(T, True)

    def _terms(self, T, want_hb):
        ok = self.corr >= 0
        a = self.src[ok].astype(np.float64)
        b = self.tgt[self.corr[ok]].astype(np.float64)
        ta = a @ T[:3, :3].T + T[:3, 3]
        e = b - ta
        Me = np.einsum("nij,nj->ni", self.M, e)
        cost = float(np.einsum("ni,ni->", e, Me))
        if not want_hb:
            return cost, None, None
        J = np.zeros((len(a), 3, 6))
        J[:, 0, 1], J[:, 0, 2] = -ta[:, 2], ta[:, 1]     # skew(T a): rotation block first (FG:185-187)
        J[:, 1, 0], J[:, 1, 2] = ta[:, 2], -ta[:, 0]
        J[:, 2, 0], J[:, 2, 1] = -ta[:, 1], ta[:, 0]
        J[:, 0, 3] = J[:, 1, 4] = J[:, 2, 5] = -1.0
        MJ = np.einsum("nij,njk->nik", self.M, J)
        H = np.einsum("nia,nib->ab", J, MJ)
        bb = np.einsum("nia,ni->a", J, Me)
        return cost, H, bb

    def _converged(self, delta):
        r = np.abs(delta[:3, :3] - np.eye(3)).max() / self.rot_eps
        t = np.abs(delta[:3, 3]).max() / self.trans_eps
        return max(r, t) < 1

    def align(self, guess=None):
        x0 = np.eye(4) if guess is None else np.asarray(guess, np.float64)
        lam, it_out, converged = -1.0, 0, False
        n_lin = n_err = 0
        for i in range(self.max_iter):
            if converged:
                break
            it_out = i
            y0, H, b = self._linearize(x0)
            n_lin += 1
            if lam < 0:
                lam = 1e-9 * np.abs(np.diag(H)).max()
            nu, ok, delta = 2.0, False, None
            for _ in range(10):
                d = np.linalg.solve(H + lam * np.eye(6), -b)
                delta = np.eye(4)
                delta[:3, :3] = Rotation.from_rotvec(d[:3]).as_matrix()
                delta[:3, 3] = d[3:]
                xi = delta @ x0
                yi, _, _ = self._terms(xi, False)
                n_err += 1
                rho = (y0 - yi) / (d @ (lam * d - b))
                if rho < 0:
                    if self._converged(delta):
                        ok = True
                        break
                    lam, nu = nu * lam, 2 * nu
                    continue
                x0 = xi
                lam = lam * max(1.0 / 3.0, 1 - (2 * rho - 1) ** 3)
                ok = True
                break
            if not ok:
                break
            converged = self._converged(delta)
        self.final_T, self.iterations, self.converged = x0.astype(F32), it_out, converged
        self.stats = dict(linearize=n_lin, compute_error=n_err)
        return self.final_T


def fitness(src, tgt, T):
    """pcl::Registration::getFitnessScore: mean squared 1-NN distance of T * src in tgt."""
    s = transform_points_f32(np.asarray(T, F32), np.asarray(src, F32)[:, :3]).astype(np.float64)
    d, _ = cKDTree(np.asarray(tgt, F32)[:, :3].astype(np.float64)).query(s, k=1)
    return float(np.mean(d * d))
