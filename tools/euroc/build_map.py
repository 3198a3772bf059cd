"""Offline front-end for BASELINE config 1: turns the bundled data/euroc_V1 stereo keyframes into a real bundle-
adjustment problem (the stages of src/sfm.cpp:1191-1571 and :1575-1880, head-less, with OpenCV's
Python module standing in for the reference's keypoint / matching code and opengv's RANSAC).

Stages (reference counterpart):
  keypoints + descriptors   cv2.goodFeaturesToTrack(1500, 0.01, 8) + ORB descriptors      keypoints.h:133-278
  stereo matches            Hamming <= 70, next-best ratio 1.2, mutual; epipolar test       matching_utils.h:51-79, sfm.cpp:1229-1262
                            |x_L^T E x_R| <= 1e-3 with E from the calibrated T_0_1
  sequential matches        same descriptor test against the next frames; RANSAC essential   matching_utils.h:81-175, sfm.cpp:1264-1440
                            matrix on the unprojected rays (cv2.findEssentialMat)
  feature tracks            union-find over all inlier matches, conflicting tracks dropped,  tracks.h:53-221, sfm.cpp:1511-1525
                            min length 3
  initial map               first stereo pair, cameras I and T_i_c[1], landmarks by          map_utils.h:204-228
                            add_new_landmarks_between_cams (oracle restatement)
  next cameras              PnP + RANSAC on the landmarks' points (cv2.solvePnPRansac on     map_utils.h:242-302, sfm.cpp:1575-1780
                            normalised rays), the stereo partner from the calibration
  new landmarks             between the stereo pair of every new frame                      map_utils.h:121-195
  optimize                  the reference's own visnav::bundle_adjustment (oracle/_ref,      sfm.cpp:1883-1925
                            Huber 1, cameras of frame 0 fixed) every few frames
  outlier removal           reprojection error > 3 px (normal) / 40 px (huge), z < 0.05      sfm.cpp:1928-2132
The inverse distance of a landmark is taken along its HOST's ray (host = first observation), which is what the
reference's parameterisation needs (its own initialisation measures it in the triangulating camera, map_utils.h:190).

Output: tests/golden/euroc_v1_map.npz — the flat geometric BA problem of the final map BEFORE the last optimisation
(poses perturbed by the reference's own drift, nothing synthetic) + the result of the reference's bundle_adjustment on
it; and tests/golden/euroc_v1_photo.npz — a photometric sub-problem on the first keyframes with their real images.

    python tools/euroc/calibrate.py && python tools/euroc/build_map.py [--frames 100]
Needs /root/reference (offline input generator, like tests/golden/make_golden*.py).
"""
import argparse
import os
import sys
import time

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle_ffi as of  # noqa: E402
import pba_b200 as pb  # noqa: E402
from pba_b200 import _ffi  # noqa: E402
import ctypes as C  # noqa: E402

DATA = "/root/reference/data/euroc_V1"
GOLDEN = os.path.join(ROOT, "tests", "golden")


# ------------------------------------------------------------------ geometry helpers (numpy) --
def quat_to_rot(q):
    x, y, z, w = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


def rot_to_quat(R):
    q = np.empty(4)
    t = np.trace(R)
    if t > 0:
        s = np.sqrt(t + 1.0) * 2
        q[3] = 0.25 * s
        q[0] = (R[2, 1] - R[1, 2]) / s; q[1] = (R[0, 2] - R[2, 0]) / s; q[2] = (R[1, 0] - R[0, 1]) / s
    else:
        i = int(np.argmax(np.diag(R)))
        j, k = (i + 1) % 3, (i + 2) % 3
        s = np.sqrt(1.0 + R[i, i] - R[j, j] - R[k, k]) * 2
        q[i] = 0.25 * s
        q[3] = (R[k, j] - R[j, k]) / s
        q[j] = (R[j, i] + R[i, j]) / s
        q[k] = (R[k, i] + R[i, k]) / s
    return q / np.linalg.norm(q)


def pose7(R, t):
    return np.r_[rot_to_quat(R), t]


def pose_Rt(T):
    return quat_to_rot(T[:4]), T[4:].copy()


def compose(Ta, Tb):
    Ra, ta = pose_Rt(Ta); Rb, tb = pose_Rt(Tb)
    return pose7(Ra @ Rb, Ra @ tb + ta)


def inverse(T):
    R, t = pose_Rt(T)
    return pose7(R.T, -R.T @ t)


def bearings(model_id, intr, uv):
    """normalize(unproject(z)) through the oracle's camera models."""
    uv = np.ascontiguousarray(uv, np.float64).reshape(-1, 2)
    out = np.zeros((uv.shape[0], 3))
    intr = np.ascontiguousarray(intr, np.float64)
    of.oracle().pba_oracle_unproject(model_id, _ffi.ptr(intr, C.c_double), uv.shape[0], _ffi.ptr(uv, C.c_double),
                                     _ffi.ptr(out, C.c_double))
    return out / np.linalg.norm(out, axis=1, keepdims=True)


def project(model_id, intr, X):
    X = np.ascontiguousarray(X, np.float64).reshape(-1, 3)
    uv = np.zeros((X.shape[0], 2))
    intr = np.ascontiguousarray(intr, np.float64)
    of.oracle().pba_oracle_project(model_id, _ffi.ptr(intr, C.c_double), X.shape[0], _ffi.ptr(X, C.c_double),
                                   _ffi.ptr(uv, C.c_double), None)
    return uv


# ------------------------------------------------------------------ front-end ------------------
def detect(img, n=1500):
    pts = cv2.goodFeaturesToTrack(img, n, 0.01, 8)  # keypoints.h:133-150
    pts = pts.reshape(-1, 2)
    kps = [cv2.KeyPoint(float(x), float(y), 31) for x, y in pts]
    orb = cv2.ORB_create()
    kps, desc = orb.compute(img, kps)  # orientation by intensity centroid + rotated BRIEF (keypoints.h:152-278)
    return np.array([k.pt for k in kps], np.float64), desc


def match(d0, d1, max_dist=70, ratio=1.2):
    """matchDescriptors (matching_utils.h / keypoints.h): threshold, next-best ratio, mutual consistency."""
    if d0 is None or d1 is None or len(d0) < 2 or len(d1) < 2:
        return np.zeros((0, 2), np.int64)
    bf = cv2.BFMatcher(cv2.NORM_HAMMING)

    def one_way(a, b):
        best = {}
        for m in bf.knnMatch(a, b, k=2):
            if len(m) == 2 and m[0].distance <= max_dist and m[0].distance * ratio <= m[1].distance:
                best[m[0].queryIdx] = m[0].trainIdx
        return best
    f, b = one_way(d0, d1), one_way(d1, d0)
    return np.array([(i, j) for i, j in f.items() if b.get(j) == i], np.int64).reshape(-1, 2)


class UnionFind:
    def __init__(self):
        self.p = {}

    def find(self, a):
        self.p.setdefault(a, a)
        while self.p[a] != a:
            self.p[a] = self.p[self.p[a]]
            a = self.p[a]
        return a

    def union(self, a, b):
        ra, rb = self.find(a), self.find(b)
        if ra != rb:
            self.p[rb] = ra


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=100)
    ap.add_argument("--calib", default=os.path.join(ROOT, "tools", "euroc", "opt_calib.json"))
    ap.add_argument("--neighbours", type=int, default=1000,
                    help="match every image against this many following frames (default: all pairs, like the reference)")
    a = ap.parse_args()
    assert of.have_ref(), "build oracle/_ref first (make ref)"
    cal = pb.load_calibration(a.calib)
    model = _ffi.CAM_NAMES[cal.models[0]]
    intr = np.ascontiguousarray(cal.intrinsics, np.float64)
    T_0_1 = compose(inverse(cal.T_i_c[0]), cal.T_i_c[1])  # camera 1 in camera 0
    ts = [l.strip() for l in open(os.path.join(DATA, "timestamps.txt")) if l.strip()][:a.frames]
    t0 = time.time()
    imgs, kd = {}, {}
    for f, stamp in enumerate(ts):
        for c in (0, 1):
            im = cv2.imread(os.path.join(DATA, "%s_%d.jpg" % (stamp, c)), 0)
            imgs[(f, c)] = im
            kd[(f, c)] = detect(im)
    print("keypoints: %d images, %.1f s" % (len(kd), time.time() - t0), flush=True)
    rays = {k: bearings(model, intr[k[1]], v[0]) for k, v in kd.items()}

    # ---- matches ----
    R01, t01 = pose_Rt(T_0_1)
    tn = t01 / np.linalg.norm(t01)
    E = np.array([[0, -tn[2], tn[1]], [tn[2], 0, -tn[0]], [-tn[1], tn[0], 0]]) @ R01  # matching_utils.h:51-60
    uf = UnionFind()
    n_stereo = n_seq = 0
    for f in range(len(ts)):
        m = match(kd[(f, 0)][1], kd[(f, 1)][1])
        if len(m):
            err = np.abs(np.einsum("ij,jk,ik->i", rays[(f, 0)][m[:, 0]], E, rays[(f, 1)][m[:, 1]]))
            for i, j in m[err <= 1e-3]:
                uf.union((f, 0, int(i)), (f, 1, int(j)))
                n_stereo += 1
    for f in range(len(ts)):
        for g in range(f + 1, min(len(ts), f + 1 + a.neighbours)):
            for c in (0, 1):
                m = match(kd[(f, c)][1], kd[(g, c)][1])
                if len(m) < 20:
                    continue
                x0 = rays[(f, c)][m[:, 0]]; x1 = rays[(g, c)][m[:, 1]]
                p0 = (x0[:, :2] / x0[:, 2:3]); p1 = (x1[:, :2] / x1[:, 2:3])
                Em, mask = cv2.findEssentialMat(p0, p1, np.eye(3), method=cv2.RANSAC, prob=0.999, threshold=2.0 / 350.0)
                if mask is None or mask.sum() < 20:
                    continue
                for (i, j) in m[mask.ravel() > 0]:
                    uf.union((f, c, int(i)), (g, c, int(j)))
                    n_seq += 1
    print("matches: %d stereo inliers, %d sequential inliers, %.1f s" % (n_stereo, n_seq, time.time() - t0), flush=True)

    # ---- tracks (tracks.h: a track may hold one feature per image; min length 3) ----
    groups = {}
    for node in list(uf.p):
        groups.setdefault(uf.find(node), []).append(node)
    tracks = []
    for nodes in groups.values():
        cams = [(f, c) for f, c, _ in nodes]
        if len(nodes) >= 3 and len(set(cams)) == len(cams):
            tracks.append({(f, c): i for f, c, i in nodes})
    print("tracks: %d (of %d groups)" % (len(tracks), len(groups)), flush=True)

    # ---- incremental map ----
    cams = {}          # (f, c) -> pose7 T_w_c
    lms = {}           # track index -> dict(obs = {(f,c): feature}, p_w)
    in_image = {}
    for ti, tr in enumerate(tracks):
        for k in tr:
            in_image.setdefault(k, []).append(ti)

    def corner(k, feat):
        return kd[k][0][feat]

    def add_stereo_landmarks(f):
        k0, k1 = (f, 0), (f, 1)
        new = [ti for ti in in_image.get(k0, []) if ti not in lms and k1 in tracks[ti]]
        if not new:
            return 0
        uv0 = np.array([corner(k0, tracks[ti][k0]) for ti in new])
        uv1 = np.array([corner(k1, tracks[ti][k1]) for ti in new])
        p, rho = of.triangulate("oracle", model, intr[0], model, intr[1], cams[k0], cams[k1], uv0, uv1)
        R, t = pose_Rt(cams[k0])
        cnt = 0
        for ti, pc in zip(new, p):
            if not np.all(np.isfinite(pc)) or pc[2] < 0.1 or np.linalg.norm(pc) > 30.0:
                continue
            lms[ti] = {"obs": {k: ft for k, ft in tracks[ti].items() if k in cams}, "p_w": R @ pc + t}
            cnt += 1
        return cnt

    def flatten():
        """Flat problem in the layout of include/pba.h; host = smallest (frame, cam)."""
        keys = sorted(cams)
        idx = {k: i for i, k in enumerate(keys)}
        poses = np.array([cams[k] for k in keys])
        fixed = np.array([1 if k[0] == 0 else 0 for k in keys], np.uint8)
        calib_idx = np.array([k[1] for k in keys], np.int32)
        rho, host, host_uv, ptr, tgt, uv, ids = [], [], [], [0], [], [], []
        for ti in sorted(lms):
            ob = sorted(k for k in lms[ti]["obs"] if k in cams)
            if len(ob) < 2:
                continue
            h = ob[0]
            R, t = pose_Rt(cams[h])
            ph = R.T @ (lms[ti]["p_w"] - t)
            rho.append(1.0 / np.linalg.norm(ph)); host.append(idx[h]); host_uv.append(corner(h, lms[ti]["obs"][h]))
            for k in ob[1:]:
                tgt.append(idx[k]); uv.append(corner(k, lms[ti]["obs"][k]))
            ptr.append(len(tgt)); ids.append(ti)
        prob = pb.Problem(pb.MODE_GEOMETRIC, poses, fixed, calib_idx, np.array([model, model], np.int32), intr,
                          np.array(rho), np.array(host, np.int32), np.array(host_uv), np.array(ptr, np.int64),
                          np.array(tgt, np.int32), np.array(uv))
        return prob, keys, ids

    def write_back(prob, keys, ids):
        for k, T in zip(keys, prob.poses):
            cams[k] = T.copy()
        pw = of.landmark_positions("oracle", prob)
        for ti, p in zip(ids, pw):
            lms[ti]["p_w"] = p

    def remove_outliers(prob, keys, ids):
        pr = of.compute_projections("oracle", prob)
        flags = pr["outlier_flags"]
        removed = 0
        for l, ti in enumerate(ids):
            s0, s1 = int(prob.lm_obs_ptr[l]) + l, int(prob.lm_obs_ptr[l + 1]) + l + 1
            obs_keys = [keys[prob.lm_host[l]]] + [keys[t] for t in prob.obs_target[prob.lm_obs_ptr[l]:prob.lm_obs_ptr[l + 1]]]
            for s, k in zip(range(s0, s1), obs_keys):
                if flags[s]:
                    lms[ti]["obs"].pop(k, None)
                    removed += 1
            if len(lms[ti]["obs"]) < 2:
                del lms[ti]
        return removed

    def optimize(max_it=20):
        prob, keys, ids = flatten()
        s = of.solve("ref", prob, of.default_options(huber_parameter=1.0, max_num_iterations=max_it), use_reference_entry=True)
        write_back(prob, keys, ids)
        prob2, keys2, ids2 = flatten()
        return remove_outliers(prob2, keys2, ids2), prob.n_obs

    cams[(0, 0)] = np.array([0, 0, 0, 1, 0, 0, 0], np.float64)
    cams[(0, 1)] = T_0_1.copy()
    print("initial stereo pair: %d landmarks" % add_stereo_landmarks(0), flush=True)
    remaining = set(range(1, len(ts)))
    step = 0
    while remaining:
        # next camera = the one sharing most tracks with the map (src/sfm.cpp:1575-1700 picks its candidates the same way)
        f = max(remaining, key=lambda g: sum(1 for ti in in_image.get((g, 0), []) if ti in lms))
        remaining.discard(f)
        step += 1
        k0 = (f, 0)
        shared = [ti for ti in in_image.get(k0, []) if ti in lms]
        if len(shared) < 15:
            print("frame %d: only %d shared tracks; %d frames left out" % (f, len(shared), len(remaining) + 1))
            break
        P = np.array([lms[ti]["p_w"] for ti in shared])
        b = rays[k0][[tracks[ti][k0] for ti in shared]]
        xy = (b[:, :2] / b[:, 2:3]).astype(np.float64)
        ok, rvec, tvec, inl = cv2.solvePnPRansac(P.astype(np.float64), xy, np.eye(3), None, iterationsCount=300,
                                                 reprojectionError=3.0 / 350.0, confidence=0.999, flags=cv2.SOLVEPNP_EPNP)
        if not ok or inl is None or len(inl) < 12:
            print("frame %d: localisation failed (%s inliers), skipped" % (f, None if inl is None else len(inl)))
            continue
        inl = inl.ravel()
        ok, rvec, tvec = cv2.solvePnP(P[inl], xy[inl], np.eye(3), None, rvec, tvec, True, cv2.SOLVEPNP_ITERATIVE)
        Rcw, _ = cv2.Rodrigues(rvec)
        cams[k0] = pose7(Rcw.T, -Rcw.T @ tvec.ravel())
        cams[(f, 1)] = compose(cams[k0], T_0_1)
        for ti in np.array(shared)[inl]:
            lms[ti]["obs"][k0] = tracks[ti][k0]
        # the stereo partner's observations of existing landmarks: accepted when they reproject within 3 px
        k1 = (f, 1)
        R1, t1 = pose_Rt(cams[k1])
        cand = [ti for ti in in_image.get(k1, []) if ti in lms]
        if cand:
            Xc = (np.array([lms[ti]["p_w"] for ti in cand]) - t1) @ R1
            uvp = project(model, intr[1], Xc)
            for ti, u, xc in zip(cand, uvp, Xc):
                if xc[2] > 0.05 and np.linalg.norm(u - corner(k1, tracks[ti][k1])) < 3.0:
                    lms[ti]["obs"][k1] = tracks[ti][k1]
        n_new = add_stereo_landmarks(f)
        if step % 5 == 0 and len(remaining) > 6:  # the last batch of cameras is left to the final optimisation (the fixture)
            removed, n_obs = optimize()
            print("frame %3d: %d cameras, %d landmarks, %d residual blocks, %d inliers of %d shared, +%d new, -%d outliers"
                  % (f, len(cams), len(lms), n_obs, len(inl), len(shared), n_new, removed), flush=True)

    # ---- fixtures ----
    prob, keys, ids = flatten()
    before = prob.copy()
    sol = prob.copy()
    s = of.solve("ref", sol, of.default_options(huber_parameter=1.0), use_reference_entry=True)   # unmodified bundle_adjustment()
    trace = prob.copy()
    st = of.solve("ref", trace, of.default_options(huber_parameter=1.0))                           # same problem, with the summary
    np.savez_compressed(
        os.path.join(GOLDEN, "euroc_v1_map.npz"), mode=before.mode, poses=before.poses, pose_fixed=before.pose_fixed,
        pose_calib=before.pose_calib, calib_model=before.calib_model, intrinsics=before.intrinsics, inv_depth=before.inv_depth,
        lm_host=before.lm_host, lm_host_uv=before.lm_host_uv, lm_obs_ptr=before.lm_obs_ptr, obs_target=before.obs_target,
        obs_uv=before.obs_uv, frame_cam=np.array(keys, np.int32), track_id=np.array(ids, np.int64),
        entry_poses=sol.poses, entry_inv_depth=sol.inv_depth,
        sol_poses=trace.poses, sol_inv_depth=trace.inv_depth, sol_initial_cost=st.initial_cost, sol_final_cost=st.final_cost,
        sol_termination=st.termination_type, sol_iter_cost=np.array([i["cost"] for i in st.iterations]),
        sol_iter_success=np.array([i["step_is_successful"] for i in st.iterations]),
        huber=1.0, T_0_1=T_0_1)
    slot, ns, bw0, bw1, nb = pb.analyze_structure(before)
    print("final map: %d cameras, %d landmarks, %d residual blocks; RCS %d slots, half-bandwidth %d (natural) / %d, %d blocks; "
          "reference BA %d iterations, cost %.6e -> %.6e" % (before.n_poses, before.n_landmarks, before.n_obs, ns, bw0, bw1, nb,
                                                             st.num_iterations, st.initial_cost, st.final_cost), flush=True)

    # ---- photometric sub-problem on real images: the first keyframes, landmarks whose whole pattern stays inside
    #      every image that sees them; a small state perturbation so that the LM has something to do ----
    n_photo = 12
    keep_pose = [i for i, k in enumerate(keys) if k[0] < n_photo // 2]
    pmap = {i: j for j, i in enumerate(keep_pose)}
    rho, host, host_uv, ptr, tgt = [], [], [], [0], []
    for l in range(before.n_landmarks):
        members = [int(before.lm_host[l])] + [int(t) for t in before.obs_target[before.lm_obs_ptr[l]:before.lm_obs_ptr[l + 1]]]
        members = [m for m in members if m in pmap]
        if len(members) < 2 or members[0] != int(before.lm_host[l]):
            continue
        u, v = before.lm_host_uv[l]
        if not (12 <= u < 752 - 12 and 12 <= v < 480 - 12):
            continue
        rho.append(sol.inv_depth[l]); host.append(pmap[members[0]]); host_uv.append((u, v))
        tgt += [pmap[m] for m in members[1:]]
        ptr.append(len(tgt))
    images = np.stack([imgs[keys[i]] for i in keep_pose])
    rng = np.random.default_rng(3)
    poses = sol.poses[keep_pose].copy()
    photo = pb.Problem(pb.MODE_PHOTOMETRIC, poses, np.array([1 if keys[i][0] == 0 else 0 for i in keep_pose], np.uint8),
                       before.pose_calib[keep_pose], before.calib_model, before.intrinsics,
                       np.array(rho) * (1 + rng.normal(0, 0.01, len(rho))), np.array(host, np.int32), np.array(host_uv),
                       np.array(ptr, np.int64), np.array(tgt, np.int32), None, images, np.zeros((len(keep_pose), 2)))
    psol = photo.copy()
    sp = of.solve("ref", psol, of.default_options(huber_parameter=9.0, max_num_iterations=10))
    cost, r, J = of.evaluate("ref", photo, True, 9.0)
    sel = np.arange(0, photo.n_obs, max(1, photo.n_obs // 400))
    np.savez_compressed(
        os.path.join(GOLDEN, "euroc_v1_photo.npz"), mode=photo.mode, poses=photo.poses, pose_fixed=photo.pose_fixed,
        pose_calib=photo.pose_calib, calib_model=photo.calib_model, intrinsics=photo.intrinsics, inv_depth=photo.inv_depth,
        lm_host=photo.lm_host, lm_host_uv=photo.lm_host_uv, lm_obs_ptr=photo.lm_obs_ptr, obs_target=photo.obs_target,
        images=photo.images, affine=photo.affine, huber=9.0, ref_cost=cost, ref_sel=sel, ref_residuals=r[sel],
        ref_jacobians=J[sel], sol_poses=psol.poses, sol_inv_depth=psol.inv_depth, sol_affine=psol.affine,
        sol_initial_cost=sp.initial_cost, sol_final_cost=sp.final_cost, sol_termination=sp.termination_type,
        sol_iter_cost=np.array([i["cost"] for i in sp.iterations]),
        sol_iter_success=np.array([i["step_is_successful"] for i in sp.iterations]), max_num_iterations=10)
    print("photometric sub-problem: %d keyframes, %d landmarks, %d residual blocks; reference %d iterations, cost %.6e -> %.6e (%s)"
          % (photo.n_poses, photo.n_landmarks, photo.n_obs, sp.num_iterations, sp.initial_cost, sp.final_cost, sp.message))


if __name__ == "__main__":
    main()
