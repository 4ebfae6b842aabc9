"""Independent numpy restatement of the reference arithmetic, used ONLY to cross-check the C++ oracle
(a second, separately written implementation on top of LAPACK: np.linalg.eig for
Eigen::EigenSolver, np.linalg.inv for MatrixXd::inverse()). Citations: /root/reference/legkilo/src/...

This is test infrastructure like oracle/: never imported by the product."""
import numpy as np


def skew(v):  # common/math_utils.hpp:13-17
    return np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0]], float)


def exp3(v, thr=1e-5):  # math_utils.hpp:55-68 (thr 1e-5) / :20-32 (thr 1e-7)
    n = np.sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2])
    if n > thr:
        K = skew(np.asarray(v, float) / n)
        return np.eye(3) + np.sin(n) * K + (1 - np.cos(n)) * K @ K
    return np.eye(3)


def log_so3(R):  # math_utils.hpp:72-76
    tr = np.trace(R)
    th = 0.0 if tr > 3.0 - 1e-6 else np.arccos(0.5 * (tr - 1))
    K = np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]])
    return 0.5 * K if abs(th) < 0.001 else 0.5 * th / np.sin(th) * K


def calc_body_cov(pb, range_inc, degree_inc):  # core/slam/voxel_map.cc:22-40
    pb = np.array(pb, float)
    if pb[2] == 0:
        pb[2] = 0.0001
    rng = np.float32(np.sqrt(pb @ pb))
    range_var = np.float32(range_inc) * np.float32(range_inc)
    dv = np.sin(float(np.float32(degree_inc)) * 0.017453293) ** 2
    d = pb / np.linalg.norm(pb)
    dh = skew(d)
    b1 = np.array([1.0, 1.0, -(d[0] + d[1]) / d[2]])
    b1 /= np.linalg.norm(b1)
    b2 = np.cross(b1, d)
    b2 /= np.linalg.norm(b2)
    N = np.stack([b1, b2], 1)
    A = float(rng) * dh @ N
    return np.outer(d, d) * float(range_var) + A @ (dv * np.eye(2)) @ A.T, pb


def init_plane(pw, var, planer_threshold=0.01):  # voxel_map.cc:42-117
    pw = np.asarray(pw, float)
    n = len(pw)
    c = pw.sum(0) / n
    cov = (pw.T @ pw) / n - np.outer(c, c)
    w, V = np.linalg.eig(cov)
    w = w.real
    V = V.real
    imin, imax = int(np.argmin(w)), int(np.argmax(w))
    out = dict(center=c, is_plane=bool(w[imin] < np.float32(planer_threshold)), eig=w)
    if not out["is_plane"]:
        return out
    pv = np.zeros((6, 6))
    for i in range(n):
        F = np.zeros((3, 3))
        for m in range(3):
            if m != imin:
                F[m] = (pw[i] - c) / (n * (w[imin] - w[m])) @ (np.outer(V[:, m], V[:, imin]) + np.outer(V[:, imin], V[:, m]))
        J = np.zeros((6, 3))
        J[:3] = V @ F
        J[3:] = np.eye(3) / n
        pv += J @ var[i] @ J.T
    nrm = V[:, imin]
    out.update(normal=nrm, plane_var=pv, radius=np.float32(np.sqrt(w[imax])), d=np.float32(-(nrm @ c)))
    return out


def plane_residual(pw, var, plane, sigma_num=3.0):  # voxel_map.cc:370-411, plane branch
    n, c = plane["normal"], plane["center"]
    s = n @ pw + float(plane["d"])
    dis = np.float32(abs(s))
    dc = np.float32(((c - pw) ** 2).sum())
    with np.errstate(invalid="ignore"):
        rd = np.sqrt(np.float32(dc - dis * dis))
    if not (float(rd) <= 3.0 * float(plane["radius"])):
        return None
    J = np.concatenate([pw - c, -n])
    sigma_l = J @ plane["plane_var"] @ J + n @ var @ n
    if not (float(dis) < sigma_num * np.sqrt(sigma_l)):
        return None
    prob = 1.0 / np.sqrt(sigma_l) * np.exp(-0.5 * float(dis) * float(dis) / sigma_l)
    return dict(dis_to_plane=np.float32(s), prob=prob, J=J)


def point_var(R, Re, te, pb_cov, pi, P):  # core/slam/KILO.cc:136-140
    M = R @ Re
    G = R @ skew(pi)
    return M @ pb_cov @ M.T + G @ P[0:3, 0:3] @ G.T + P[3:6, 3:6]


def obs_row(R, Re, pi, body_cov, plane, pw, res, ratio):  # KILO.cc:192-209
    n = plane["normal"]
    h = np.concatenate([skew(pi) @ R.T @ n, n])
    z = -float(res["dis_to_plane"])
    var = R @ Re @ body_cov @ Re.T @ R.T
    Rk = ratio * (res["J"] @ plane["plane_var"] @ res["J"] + n @ var @ n)
    return h, z, Rk


def update_by_points_literal(P, h, z, r):  # core/slam/eskf.cc:91-113
    h = np.atleast_2d(h)
    N = len(z)
    if N == 1:
        PHT = P[:, :6] @ h.T
        s = 1.0 / (0.0001 + (h @ PHT[:6])[0, 0] + r[0])
        K = s * PHT
    else:
        PHT = P[:, :6] @ h.T
        S = h @ PHT[:6] + np.diag(r)
        K = PHT @ np.linalg.inv(S)
    delta = K @ np.asarray(z)
    return delta, P - K @ h @ P[:6, :]


def fx(R, imu_a, imu_w, dt):  # eskf.cc:72-81
    F = np.eye(30)
    F[0:3, 0:3] = exp3(-dt * np.asarray(imu_w), thr=1e-7)
    F[0:3, 21:24] = dt * np.eye(3)
    F[3:6, 6:9] = dt * np.eye(3)
    F[6:9, 0:3] = -dt * R @ skew(imu_a)
    F[6:9, 15:18] = dt * np.eye(3)
    F[6:9, 18:21] = dt * R
    return F


def update_by_imu(P, z, r):  # eskf.cc:125-135
    PHT = P[:, 9:15] + P[:, 18:24]
    HP = P[9:15, :] + P[18:24, :]
    HPHT = PHT[9:15, :] + PHT[18:24, :] + np.diag(r)
    K = PHT @ np.linalg.inv(HPHT)
    return K @ z, P - K @ HP


def update_by_kinimu(P, H, z, r):  # eskf.cc:137-145
    PHT = P @ H.T
    K = PHT @ np.linalg.inv(H @ PHT + np.diag(r))
    return K @ z, P - K @ H @ P
