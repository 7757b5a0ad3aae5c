"""Synthetic inputs for tests and bench.py: the reference's example robots, scenes and problem sets.

Robots and scenes are the ones the reference's examples build (data tables restated from
matlab/+gpmp2/generateArm.m, generate2Ddataset.m, generate3Ddataset.m); signed distance fields
follow matlab/+gpmp2/signedDistanceField{2D,3D}.m: (bwdist(map) - bwdist(1-map)) * cell_size.
Host-side input generation only -- nothing here is on the hot path.
"""
import numpy as np

from .api import (Arm, ArmModel, BodySphere, PlanarSDF, Pose3, SignedDistanceField,
                  TrajOptimizerSetting, straight_line_traj)


# ------------------------------------------------------------------------------------------------
# robots (matlab/+gpmp2/generateArm.m:20-117)
# ------------------------------------------------------------------------------------------------
def simple_two_links_arm(base_pose=None):
    arm = Arm(2, [0.5, 0.5], [0, 0], [0, 0], base_pose)
    data = [(0, x) for x in (-0.5, -0.4, -0.3, -0.2, -0.1)] + [(1, x) for x in (-0.5, -0.4, -0.3, -0.2, -0.1, 0.0)]
    return ArmModel(arm, [BodySphere(l, 0.01, [x, 0, 0]) for l, x in data])


def simple_three_links_arm(base_pose=None):
    arm = Arm(3, [0.5, 0.5, 0.5], [0, 0, 0], [0, 0, 0], base_pose)
    data = ([(0, x) for x in (-0.5, -0.4, -0.3, -0.2, -0.1)] + [(1, x) for x in (-0.5, -0.4, -0.3, -0.2, -0.1)]
            + [(2, x) for x in (-0.5, -0.4, -0.3, -0.2, -0.1, 0.0)])
    return ArmModel(arm, [BodySphere(l, 0.01, [x, 0, 0]) for l, x in data])


WAM_ALPHA = np.array([-np.pi / 2, np.pi / 2, -np.pi / 2, np.pi / 2, -np.pi / 2, np.pi / 2, 0])
WAM_A = np.array([0, 0, 0.045, -0.045, 0, 0, 0])
WAM_D = np.array([0, 0, 0.55, 0, 0.3, 0, 0.06])
WAM_SPHERES = [
    (0, 0.0, 0.0, 0.0, 0.15), (1, 0.0, 0.0, 0.2, 0.06), (1, 0.0, 0.0, 0.3, 0.06), (1, 0.0, 0.0, 0.4, 0.06),
    (1, 0.0, 0.0, 0.5, 0.06), (2, 0.0, 0.0, 0.0, 0.06), (3, 0.0, 0.0, 0.1, 0.06), (3, 0.0, 0.0, 0.2, 0.06),
    (3, 0.0, 0.0, 0.3, 0.06), (5, 0.0, 0.0, 0.1, 0.06), (6, 0.1, -0.025, 0.08, 0.04), (6, 0.1, 0.025, 0.08, 0.04),
    (6, -0.1, 0, 0.08, 0.04), (6, 0.15, -0.025, 0.13, 0.04), (6, 0.15, 0.025, 0.13, 0.04), (6, -0.15, 0, 0.13, 0.04)]
# joint range used for random start/goal sampling (Barrett WAM nominal limits, rad)
WAM_Q_LO = np.array([-2.6, -2.0, -2.8, -0.9, -4.76, -1.6, -3.0])
WAM_Q_HI = np.array([2.6, 2.0, 2.8, 3.1, 1.24, 1.6, 3.0])
# the example's query (matlab/WAMPlannerExample.m:32-33)
WAM_START = np.array([-0.8, -1.70, 1.64, 1.29, 1.1, -0.106, 2.2])
WAM_END = np.array([-0.0, 0.94, 0, 1.6, 0, -0.919, 1.55])


def wam_arm(base_pose=None):
    arm = Arm(7, WAM_A, WAM_ALPHA, WAM_D, base_pose, np.zeros(7))
    return ArmModel(arm, [BodySphere(l, r, [x, y, z]) for l, x, y, z, r in WAM_SPHERES])


# ------------------------------------------------------------------------------------------------
# scenes
# ------------------------------------------------------------------------------------------------
def _signed_edt(occ, cell):
    from scipy.ndimage import distance_transform_edt
    occ = occ > 0.75
    if not occ.any():
        return np.full(occ.shape, 1000.0)
    return (distance_transform_edt(~occ) - distance_transform_edt(occ)) * cell


def _add_box(occ, center, size):
    """add_obstacle of generate{2D,3D}dataset.m (1-based centre index, odd-rounded size)."""
    sl = []
    for c, s in zip(center, size):
        h = (s - 1) // 2
        sl.append(slice(max(c - 1 - h, 0), min(c - 1 + h + 1, 10 ** 9)))
    occ[tuple(sl)] = 1


def planar_dataset(name="OneObstacleDataset"):
    """-> PlanarSDF.  matlab/+gpmp2/generate2Ddataset.m:18-42 (300x300, cell 0.01, origin (-1,-1))."""
    rows = cols = 300
    occ = np.zeros((rows, cols))
    if name == "OneObstacleDataset":
        _add_box(occ, (190, 160), (60, 80))
    elif name == "TwoObstaclesDataset":
        _add_box(occ, (200, 200), (80, 100))
        _add_box(occ, (160, 80), (30, 80))
    elif name == "Empty":
        pass
    else:
        raise ValueError(name)
    return PlanarSDF([-1.0, -1.0], 0.01, _signed_edt(occ, 0.01))


def mobile_map(name="MobileMap1"):
    """500x500 map, cell 0.01, origin (-5,-5): 1 x 5 m block at the centre plus 1 m walls
    (generate2Ddataset.m:60-76; get_center/get_dim convert metres to [row, col] cells)."""
    cell, n, o = 0.01, 500, -5.0
    occ = np.zeros((n, n))

    def add(cx, cy, w, h):
        r0, r1 = int(round((cy - h / 2 - o) / cell)), int(round((cy + h / 2 - o) / cell))
        c0, c1 = int(round((cx - w / 2 - o) / cell)), int(round((cx + w / 2 - o) / cell))
        occ[max(r0, 0):min(r1, n), max(c0, 0):min(c1, n)] = 1

    add(0, 0, 1, 5)
    add(0, 4.5, 10, 1)
    add(0, -4.5, 10, 1)
    add(4.5, 0, 1, 10)
    add(-4.5, 0, 1, 10)
    return PlanarSDF([o, o], cell, _signed_edt(occ, cell))


WAM_DESK_BOXES = [((170, 220, 130), (140, 60, 5)), ((105, 195, 90), (10, 10, 80)), ((235, 195, 90), (10, 10, 80)),
                  ((105, 245, 90), (10, 10, 80)), ((235, 245, 90), (10, 10, 80)), ((250, 190, 145), (60, 5, 190)),
                  ((250, 90, 145), (60, 5, 190)), ((200, 190, 145), (40, 5, 190)), ((250, 140, 240), (60, 100, 5)),
                  ((250, 140, 190), (60, 100, 5)), ((250, 140, 140), (60, 100, 5)), ((250, 140, 90), (60, 100, 5))]


def wam_desk_dataset(n=300):
    """-> SignedDistanceField.  WAMDeskDataset (generate3Ddataset.m:38-67): 300^3, cell 0.01, origin
    (-1.5,-1.5,-1.5).  `n` < 300 builds the same scene at a coarser grid (cell = 3/n) for tests.
    The example passes field(:,:,z)' to initFieldData (WAMPlannerExample.m:25-27), so the
    dataset's first map index is the SDF column (x) index."""
    scale = n / 300.0
    cell = 3.0 / n
    occ = np.zeros((n, n, n))  # indexed [x][y][z] like dataset.map(row, col, z)
    for c, s in WAM_DESK_BOXES:
        cc = tuple(max(1, int(round(v * scale))) for v in c)
        ss = tuple(max(1, int(round(v * scale))) for v in s)
        _add_box(occ, cc, ss)
    field = _signed_edt(occ, cell)          # [x][y][z]
    data = np.ascontiguousarray(field.transpose(2, 1, 0))  # (nz, rows=y, cols=x)
    return SignedDistanceField([-1.5, -1.5, -1.5], cell, data)


def random_boxes_dataset(n=300, n_boxes=12, seed=3, extent=3.0):
    """Seeded random-box scene (SURVEY.md section 8d, config 3 variant)."""
    rng = np.random.default_rng(seed)
    cell = extent / n
    occ = np.zeros((n, n, n))
    for _ in range(n_boxes):
        size = rng.integers(max(2, n // 60), max(3, n // 5), size=3)
        c = rng.integers(n // 6, n - n // 6, size=3)
        # keep the robot base free
        if np.all(np.abs(c - n // 2) < size // 2 + n // 10):
            continue
        _add_box(occ, tuple(int(v) for v in c), tuple(int(v) for v in size))
    field = _signed_edt(occ, cell)
    data = np.ascontiguousarray(field.transpose(2, 1, 0))
    return SignedDistanceField([-extent / 2] * 3, cell, data)


# ------------------------------------------------------------------------------------------------
# settings + problem sets
# ------------------------------------------------------------------------------------------------
def bench_setting(dof, total_time=2.0, total_step=10, cost_sigma=0.02, epsilon=0.2, inter=5, max_iter=10):
    """The fixed optimizer setting of SURVEY.md section 8d: LM, max_iter 10, rel_thresh 0."""
    s = TrajOptimizerSetting(dof)
    s.set_total_step(total_step)
    s.set_total_time(total_time)
    s.set_epsilon(epsilon)
    s.set_cost_sigma(cost_sigma)
    s.set_obs_check_inter(inter)
    s.set_conf_prior_model(0.0001)
    s.set_vel_prior_model(0.0001)
    s.set_Qc_model(np.eye(dof))
    s.setLM()
    s.set_rel_thresh(0.0)
    s.set_max_iter(max_iter)
    return s


def wam_problems(B, total_step=10, seed=3, mode="restart", sigma=0.3):
    """B WAM problems -> dict(start_conf, start_vel, end_conf, end_vel, init_traj).

    mode "restart": the example's (start, goal) with straight-line init + Gaussian perturbation of
    the interior support states (random restarts); mode "random": uniform random (start, goal) in
    the joint range with straight-line init."""
    rng = np.random.default_rng(seed)
    D = 7
    if mode == "restart":
        s = np.tile(WAM_START, (B, 1))
        e = np.tile(WAM_END, (B, 1))
    else:
        s = rng.uniform(WAM_Q_LO, WAM_Q_HI, size=(B, D))
        e = rng.uniform(WAM_Q_LO, WAM_Q_HI, size=(B, D))
    t = straight_line_traj(s, e, total_step).reshape(B, 2 * (total_step + 1), D)
    if mode == "restart":
        t[:, 1:total_step] += sigma * rng.standard_normal((B, total_step - 1, D))
    z = np.zeros((B, D))
    return {"start_conf": s, "start_vel": z, "end_conf": e, "end_vel": z.copy(),
            "init_traj": np.ascontiguousarray(t.reshape(B, -1))}


def planar_problems(B, dof, total_step=10, seed=1):
    rng = np.random.default_rng(seed)
    s = rng.uniform(-np.pi, np.pi, size=(B, dof))
    e = rng.uniform(-np.pi, np.pi, size=(B, dof))
    z = np.zeros((B, dof))
    return {"start_conf": s, "start_vel": z, "end_conf": e, "end_vel": z.copy(),
            "init_traj": straight_line_traj(s, e, total_step)}


# ------------------------------------------------------------------------------------------------
# Pose2MobileArm (config 4): matlab/+gpmp2/generateMobileArm.m:20-51, MobileArm2FactorGraphExample.m
# ------------------------------------------------------------------------------------------------
def mobile_two_links_arm(base_T_arm=None):
    from .api import Pose2MobileArm, Pose2MobileArmModel
    arm = Arm(2, [0.3, 0.3], [0, 0], [0, 0])
    marm = Pose2MobileArm(arm, base_T_arm)
    data = [(0, -0.1, 0.12), (0, 0.0, 0.12), (0, 0.1, 0.12), (1, -0.3, 0.05), (1, -0.2, 0.05), (1, -0.1, 0.05),
            (2, -0.3, 0.05), (2, -0.2, 0.05), (2, -0.1, 0.05), (2, 0.0, 0.05)]
    return Pose2MobileArmModel(marm, [BodySphere(l, r, [x, 0, 0]) for l, x, r in data])


def other_mobile_robot(kind):
    """Small instances of the other Pose2Vector robots of gpmp2/planner/BatchTrajOptimizer.cpp:92-128 that fit the
    kernels' system dof <= 7: 'two_arms' Pose2Mobile2Arms (3 + 2 + 2), 'vetlin' Pose2MobileVetLinArm (3 + 1 + 2),
    'vetlin_two_arms' Pose2MobileVetLin2Arms (3 + 1 + 2 + 1), 'vetlin_reversed' (reverse_linact).  Spheres on every link."""
    from . import api
    from .api import Pose3
    rz = lambda t: np.array([[np.cos(t), -np.sin(t), 0], [np.sin(t), np.cos(t), 0], [0, 0, 1.0]])
    arm2 = lambda: Arm(2, [0.3, 0.3], [0, 0], [0, 0])
    arm1 = lambda: Arm(1, [0.35], [0], [0])
    veh = [(0, -0.1, 0.12), (0, 0.0, 0.12), (0, 0.1, 0.12)]
    links = lambda first, n: [(first + k, x, 0.05) for k in range(n) for x in (-0.25, -0.12, 0.0)]
    if kind == "two_arms":
        marm = api.Pose2Mobile2Arms(arm2(), arm2(), Pose3(R=rz(0.4), t=[0.1, 0.1, 0.05]), Pose3(R=rz(-0.5), t=[0.1, -0.1, 0.1]))
        data, cls = veh + links(1, 2) + links(3, 2), api.Pose2Mobile2ArmsModel
    elif kind in ("vetlin", "vetlin_reversed"):
        marm = api.Pose2MobileVetLinArm(arm2(), Pose3(R=rz(0.2), t=[0.05, 0.0, 0.1]), Pose3(R=rz(0.3), t=[0.1, 0.0, 0.05]),
                                        kind == "vetlin_reversed")
        data, cls = veh + [(1, 0.0, 0.1)] + links(2, 2), api.Pose2MobileVetLinArmModel
    elif kind == "vetlin_two_arms":
        marm = api.Pose2MobileVetLin2Arms(arm2(), arm1(), Pose3(R=rz(0.2), t=[0.05, 0.0, 0.1]), Pose3(R=rz(0.5), t=[0.1, 0.1, 0.05]),
                                          Pose3(R=rz(-0.6), t=[0.1, -0.1, 0.0]))
        data, cls = veh + [(1, 0.0, 0.1)] + links(2, 2) + links(4, 1), api.Pose2MobileVetLin2ArmsModel
    else:
        raise ValueError(kind)
    return cls(marm, [BodySphere(l, r, [x, 0.02 * l, 0.01 * l]) for l, x, r in data])


def other_mobile_problems(model, B, seed=4, extent=3.5, total_step=10, lift=(0.0, 0.4)):
    """Problems for other_mobile_robot(): base poses as mobile_problems, the lift (if any) U(lift), joints U(-pi/2, pi/2).
    The heading changes by 1 .. 2.5 rad from start to end, so no interval of the initial trajectory has the tiny heading
    step at which Pose2::LogmapDerivative is ill-conditioned (tests/test_gpu_parity.py::
    test_mobile_linearize_logmap_derivative_conditioning covers that case on its own)."""
    rng = np.random.default_rng(seed)
    D = model.dof()
    has_lift = hasattr(model.fk_model(), "reverse_linact")
    sc, ec = np.zeros((B, D)), np.zeros((B, D))
    for k in range(B):
        for conf in (sc, ec):
            conf[k, :2] = rng.uniform(-extent, extent, 2)
            conf[k, 2] = rng.uniform(-np.pi, np.pi)
            conf[k, 3:] = rng.uniform(-np.pi / 2, np.pi / 2, D - 3)
            if has_lift:
                conf[k, 3] = rng.uniform(*lift)
        ec[k, 2] = sc[k, 2] + rng.choice([-1.0, 1.0]) * rng.uniform(1.0, 2.5)
    tr = init_pose2vector_traj_straight_line_batch(sc, ec, total_step)
    z = np.zeros((B, D))
    return {"start_conf": sc, "start_vel": z, "end_conf": ec, "end_vel": z.copy(), "init_traj": tr}


def _pose2_expmap(v):
    w = v[2]
    if abs(w) < 1e-10:
        return np.array([v[0], v[1], w])
    c, s = np.cos(w), np.sin(w)
    ox, oy = -v[1], v[0]
    return np.array([(ox - (c * ox - s * oy)) / w, (oy - (s * ox + c * oy)) / w, w])


def _pose2_logmap(p):
    w = p[2]
    if abs(w) < 1e-10:
        return np.array([p[0], p[1], w])
    c, s = np.cos(w), np.sin(w)
    det = (c - 1) ** 2 + s * s
    ux, uy = c * p[0] + s * p[1], -s * p[0] + c * p[1]
    return np.array([(w / det) * -(uy - p[1]), (w / det) * (ux - p[0]), w])


def _pose2_compose(a, b):
    c, s = np.cos(a[2]), np.sin(a[2])
    th = a[2] + b[2]
    return np.array([a[0] + c * b[0] - s * b[1], a[1] + s * b[0] + c * b[1], np.arctan2(np.sin(th), np.cos(th))])


def _pose2_between(a, b):
    c, s = np.cos(a[2]), np.sin(a[2])
    dx, dy = b[0] - a[0], b[1] - a[1]
    th = b[2] - a[2]
    return np.array([c * dx + s * dy, -s * dx + c * dy, np.arctan2(np.sin(th), np.cos(th))])


def init_pose2vector_traj_straight_line(init_pose, init_conf, end_pose, end_conf, total_step):
    """initPose2VectorTrajStraightLine (gpmp2/planner/TrajUtils.cpp:53-73): interpolate<Pose2>(a, b, t) =
    a * Expmap(t * Logmap(a^-1 b)) for the base, linear for the arm, avg_vel = (end - init) / total_step.
    -> wire-layout trajectory (2*N*D,)"""
    init_pose, end_pose = np.asarray(init_pose, float), np.asarray(end_pose, float)
    init_conf, end_conf = np.asarray(init_conf, float), np.asarray(end_conf, float)
    D, N = 3 + init_conf.size, total_step + 1
    t = np.empty((2 * N, D))
    lg = _pose2_logmap(_pose2_between(init_pose, end_pose))
    for i in range(N):
        ratio = float(i) / float(total_step)
        t[i, :3] = _pose2_compose(init_pose, _pose2_expmap(ratio * lg))
        t[i, 3:] = (1.0 - ratio) * init_conf + ratio * end_conf
    t[N:] = np.concatenate([end_pose - init_pose, end_conf - init_conf]) / float(total_step)
    return t.reshape(-1)


def init_pose2vector_traj_straight_line_batch(start, end, total_step):
    """initPose2VectorTrajStraightLine for B (start, end) Pose2Vector pairs at once (rows (x, y, theta, q...)) -> (B, 2*N*D)."""
    start, end = np.asarray(start, float), np.asarray(end, float)
    B, D = start.shape
    N = total_step + 1
    wrap = lambda a: np.arctan2(np.sin(a), np.cos(a))   # noqa: E731
    c, s_ = np.cos(start[:, 2]), np.sin(start[:, 2])
    dx, dy = end[:, 0] - start[:, 0], end[:, 1] - start[:, 1]
    bx, by, bw = c * dx + s_ * dy, -s_ * dx + c * dy, wrap(end[:, 2] - start[:, 2])          # between(start, end)
    small = np.abs(bw) < 1e-10
    w = np.where(small, 1.0, bw)
    cw, sw = np.cos(w), np.sin(w)
    det = (cw - 1) ** 2 + sw * sw
    ux, uy = cw * bx + sw * by, -sw * bx + cw * by
    lx = np.where(small, bx, (w / det) * -(uy - by))                                          # Logmap
    ly = np.where(small, by, (w / det) * (ux - bx))
    t = np.empty((B, 2 * N, D))
    for i in range(N):
        ratio = float(i) / float(total_step)
        vx, vy, vw = ratio * lx, ratio * ly, ratio * bw
        sm = np.abs(vw) < 1e-10
        ww = np.where(sm, 1.0, vw)
        ce, se = np.cos(ww), np.sin(ww)
        ox, oy = -vy, vx
        ex = np.where(sm, vx, (ox - (ce * ox - se * oy)) / ww)                                 # Expmap
        ey = np.where(sm, vy, (oy - (se * ox + ce * oy)) / ww)
        t[:, i, 0] = start[:, 0] + c * ex - s_ * ey                                            # compose
        t[:, i, 1] = start[:, 1] + s_ * ex + c * ey
        t[:, i, 2] = wrap(start[:, 2] + vw)
        t[:, i, 3:] = (1.0 - ratio) * start[:, 3:] + ratio * end[:, 3:]
    t[:, N:, :] = ((end - start) / float(total_step))[:, None, :]
    return np.ascontiguousarray(t.reshape(B, -1))


def mobile_problems(B, sdf_free_fn=None, total_step=10, seed=4, extent=4.0):
    """B mobile-manipulator problems: base poses U([-extent, extent]^2 x (-pi, pi]),
    arm U(-pi/2, pi/2)^2, zero end velocities, straight-line initialisation."""
    rng = np.random.default_rng(seed)
    D = 5

    def sample():
        while True:
            p = np.concatenate([rng.uniform(-extent, extent, 2), rng.uniform(-np.pi, np.pi, 1)])
            if sdf_free_fn is None or sdf_free_fn(p[:2]):
                return p
    sc, ec = np.zeros((B, D)), np.zeros((B, D))
    for k in range(B):       # (the draw order defines the problem set of a seed)
        ps, pe = sample(), sample()
        qs, qe = rng.uniform(-np.pi / 2, np.pi / 2, 2), rng.uniform(-np.pi / 2, np.pi / 2, 2)
        sc[k], ec[k] = np.concatenate([ps, qs]), np.concatenate([pe, qe])
    tr = init_pose2vector_traj_straight_line_batch(sc, ec, total_step)
    z = np.zeros((B, D))
    return {"start_conf": sc, "start_vel": z, "end_conf": ec, "end_vel": z.copy(), "init_traj": tr}


# ------------------------------------------------------------------------------------------------
# The BASELINE.json configs, one registry for tests, bench.py and scripts
# ------------------------------------------------------------------------------------------------
BASELINE_CONFIGS = ("wam", "planar2", "planar3gp", "mobile", "sweep")


def baseline_config(name, sdf_cells=300, inter=None):
    """-> dict(label, model, sdf, setting, problems(B, seed, mode=...), batch, seed, D, S, P, N, K, ndim, lie).

    wam       config 3: WAM 7-DOF, WAMDeskDataset 300^3, K = 5, random restarts of the example's query, B = 65536
    planar2   config 1: SimpleTwoLinksArm, OneObstacleDataset 300 x 300, K = 4, B = 4096 (B = 1 is the reference's CPU case)
    planar3gp config 2: SimpleThreeLinksArm, TwoObstaclesDataset, K = 5, B = 4096
    mobile    config 4: Pose2MobileArm (2-link arm) in MobileMap1 500 x 500, K = 5, B = 16384
    sweep     config 5: the WAM inputs at B in {1k .. 1M}
    P = number of (sphere, joint) pairs of the robot: sum over spheres of the joints its link depends on."""
    if name in ("wam", "sweep"):
        model, D, K = wam_arm(), 7, 5 if inter is None else inter
        cfg = dict(label="WAM 7-DOF ArmModel (16 spheres), WAMDeskDataset %d^3 SDF" % sdf_cells, model=model,
                   sdf=wam_desk_dataset(sdf_cells), setting=bench_setting(7, inter=K),
                   problems=lambda B, seed, mode="restart": wam_problems(B, seed=seed, mode=mode),
                   batch=65536, seed=3, ndim=3, lie=False)
        links = [s[0] for s in WAM_SPHERES]
        cfg["P"] = sum(l + 1 for l in links)
    elif name == "planar2":
        model, D, K = simple_two_links_arm(), 2, 4 if inter is None else inter
        cfg = dict(label="planar 2-link SimpleTwoLinksArm (11 spheres), OneObstacleDataset 300x300 PlanarSDF", model=model,
                   sdf=planar_dataset("OneObstacleDataset"),
                   setting=bench_setting(2, total_time=10.0, cost_sigma=0.1, epsilon=0.1, inter=K),
                   problems=lambda B, seed, mode=None: planar_problems(B, 2, seed=seed), batch=4096, seed=1, ndim=2, lie=False)
        cfg["P"] = 5 * 1 + 6 * 2
    elif name == "planar3gp":
        model, D, K = simple_three_links_arm(), 3, 5 if inter is None else inter
        cfg = dict(label="planar 3-link SimpleThreeLinksArm (16 spheres) with ObstaclePlanarSDFFactorGP, TwoObstaclesDataset 300x300",
                   model=model, sdf=planar_dataset("TwoObstaclesDataset"),
                   setting=bench_setting(3, total_time=10.0, cost_sigma=0.1, epsilon=0.2, inter=K),
                   problems=lambda B, seed, mode=None: planar_problems(B, 3, seed=seed), batch=4096, seed=2, ndim=2, lie=False)
        cfg["P"] = 5 * 1 + 5 * 2 + 6 * 3
    elif name == "mobile":
        model, D, K = mobile_two_links_arm(), 5, 5 if inter is None else inter
        cfg = dict(label="Pose2MobileArm (vehicle + 2-link arm, 10 spheres), MobileMap1 500x500 PlanarSDF", model=model,
                   sdf=mobile_map(), setting=bench_setting(5, total_time=5.0, cost_sigma=0.1, epsilon=0.1, inter=K),
                   problems=lambda B, seed, mode=None: mobile_problems(B, seed=seed, extent=3.5), batch=16384, seed=4,
                   ndim=2, lie=True)
        cfg["P"] = int(sum(3 + int(l) for l in model._link))   # link 0 = vehicle (3 base dof), link k adds k arm joints
    else:
        raise ValueError("unknown config %r (one of %s)" % (name, ", ".join(BASELINE_CONFIGS)))
    cfg.update(name=name, D=D, K=K, N=cfg["setting"].total_step + 1, S=cfg["model"].nr_body_spheres())
    return cfg


def algorithmic_work(cfg):
    """SURVEY.md section 8(d) per-unit figures for a config: MFLOP per linearization / solve / error evaluation and
    algorithmic L2 bytes per SDF lookup (8 doubles in 3-D, 4 in 2-D) x lookups per linearization / error evaluation.
    The formulas are the survey's, recounted for (D, S, P, N, K, ndim); for WAM they give 0.32 / 0.08 / 0.08 MFLOP."""
    D, S, P, N, K, nd = cfg["D"], cfg["S"], cfg["P"], cfg["N"], cfg["K"], cfg["ndim"]
    b, C = 2 * D, N + (N - 1) * K
    lin_cfg = 63 * D + 18 * S + 12 * P + (100 if nd == 3 else 30) * S + (5 if nd == 3 else 3) * P + S * D * (D + 1) + 2 * S * D
    lin = C * lin_cfg + (N - 1) * K * (10 * D * (D + 1) + 8 * D) + 20 * D * (N - 1)
    solve = N * (b ** 3 / 3.0 + 2 * b ** 3) + 4 * N * b * b
    err = C * (63 * D + 18 * S + (40 if nd == 3 else 12) * S) + 10 * D * N
    return {"mflop_linearize": lin / 1e6, "mflop_solve": solve / 1e6, "mflop_error_eval": err / 1e6,
            "lookups_per_pass": S * C, "l2_bytes_per_lookup": 64.0 if nd == 3 else 32.0, "configurations": C}
