"""Pin the CPU oracle against the reference's own known-answer vectors (tests/golden/, lifted from
the reference unit tests) and the reference's analytic-vs-numeric Jacobian checks.  CPU only."""
import numpy as np
import pytest

import gpmp2_b200 as G


def _rot_z(y):
    c, s = np.cos(y), np.sin(y)
    return np.array([[c, -s, 0], [s, c, 0], [0, 0, 1.0]])


def _num_jac(f, x, h=1e-6):
    x = np.asarray(x, dtype=float)
    f0 = np.asarray(f(x))
    J = np.zeros((f0.size, x.size))
    for k in range(x.size):
        d = np.zeros_like(x)
        d[k] = h
        J[:, k] = (np.asarray(f(x + d)) - np.asarray(f(x - d))).ravel() / (2 * h)
    return J


def _model(arm_cfg, spheres, base_R=None):
    base = G.Pose3(R=base_R, t=arm_cfg.get("base_t", [0, 0, 0]))
    n = len(arm_cfg["a"])
    arm = G.Arm(n, arm_cfg["a"], arm_cfg["alpha"], arm_cfg["d"], base)
    return G.ArmModel(arm, [G.BodySphere(l, r, c) for l, r, c in spheres])


def test_arm_2link_poses(golden, oracle):
    g = golden["arm_2link_rotated_base"]
    m = _model({"a": g["a"], "alpha": g["alpha"], "d": g["d"], "base_t": g["base_t"]}, [[0, 0.1, [0, 0, 0]]],
               base_R=_rot_z(g["base_ypr"][0]))
    for case in g["cases"]:
        poses, J = oracle.forward_kinematics(m, case["q"])
        for i in range(2):
            assert np.allclose(poses[i][:3, 3], case["link_t"][i], atol=case["tol"])
            assert np.allclose(poses[i][:3, :3], _rot_z(case["link_yaw"][i]), atol=case["tol"])


def test_arm_wam_positions(golden, oracle):
    g = golden["arm_wam"]
    arm = G.Arm(7, g["a"], np.array(g["alpha_over_pi"]) * np.pi, g["d"])
    m = G.ArmModel(arm, [G.BodySphere(6, 0.05, [0, 0, 0])])
    poses, _ = oracle.forward_kinematics(m, g["q"])
    assert np.allclose(poses[:, :3, 3], np.array(g["link_t"]), atol=g["tol"])


def test_arm_pose_jacobian_numeric(oracle):
    # testArm.cpp:369-427 -- analytic pose Jacobians vs numerical derivative (on the translation
    # part in the world frame: d t / dq = R * v-rows)
    m = G.synth.wam_arm() if hasattr(G, "synth") else None
    from gpmp2_b200 import synth
    m = synth.wam_arm()
    q = np.array([1.58, 1.1, 0, 1.7, 0, -1.24, 1.57])
    poses, J = oracle.forward_kinematics(m, q)
    for link in range(7):
        R = poses[link][:3, :3]
        Jt = _num_jac(lambda x: oracle.forward_kinematics(m, x, False)[0][link][:3, 3], q)
        assert np.allclose(R @ J[link][3:6], Jt, atol=1e-7)
        # rotation part: body-frame omega: skew = R^T dR/dq
        for j in range(7):
            d = np.zeros(7)
            d[j] = 1e-6
            Rp = oracle.forward_kinematics(m, q + d, False)[0][link][:3, :3]
            Rm = oracle.forward_kinematics(m, q - d, False)[0][link][:3, :3]
            S = R.T @ (Rp - Rm) / 2e-6
            assert np.allclose([S[2, 1], S[0, 2], S[1, 0]], J[link][0:3, j], atol=1e-6)


def test_arm_model_sphere_centers(golden, oracle):
    g = golden["arm_model_2link"]
    m = _model(g, g["spheres"])
    for case in g["cases"]:
        c, J = oracle.sphere_centers(m, case["q"])
        assert np.allclose(c, case["centers"], atol=g["center_tol"])
        for s in range(len(g["spheres"])):
            Jn = _num_jac(lambda x: oracle.sphere_centers(m, x, False)[0][s], case["q"])
            assert np.allclose(J[s], Jn, atol=1e-8)


def test_sdf3_values_and_gradient(golden, oracle):
    g = golden["sdf3"]
    sdf = G.SignedDistanceField(g["origin"], g["cell_size"], np.array(g["data"]))
    for v in g["values"]:
        ok, d, _ = oracle.sdf_query(sdf, v["p"])
        assert ok and abs(d - v["d"]) < g["value_tol"]
    for p in g["gradient_points"]:
        ok, d, grad = oracle.sdf_query(sdf, p)
        gn = _num_jac(lambda x: [oracle.sdf_query(sdf, x)[1]], p).ravel()
        assert np.allclose(grad, gn, atol=g["gradient_numeric_tol"])
    # SDFQueryOutOfRange
    assert not oracle.sdf_query(sdf, [1.0, 0, 0])[0]
    assert not oracle.sdf_query(sdf, [0, 0, -0.11])[0]


def test_sdf2_values_and_gradient(golden, oracle):
    g = golden["sdf2"]
    sdf = G.PlanarSDF(g["origin"], g["cell_size"], np.array(g["data"]))
    for v in g["values"]:
        ok, d, _ = oracle.sdf_query(sdf, v["p"])
        assert ok and abs(d - v["d"]) < g["value_tol"]
    for p in g["gradient_points"]:
        ok, d, grad = oracle.sdf_query(sdf, p)
        gn = _num_jac(lambda x: [oracle.sdf_query(sdf, x)[1]], p).ravel()
        assert np.allclose(grad, gn, atol=g["gradient_numeric_tol"])
    assert not oracle.sdf_query(sdf, [0.3, 0])[0]


def _sdf_to_err(sdf_exp, eps):
    e = eps - np.asarray(sdf_exp)
    return np.where(e > 0, e, 0.0)


@pytest.mark.parametrize("key,planar", [("obstacle_sdf_factor_arm", False), ("obstacle_planar_sdf_factor_arm", True)])
def test_obstacle_factors(golden, oracle, key, planar):
    g = golden[key]
    field = np.array(g["field"])
    sdf = G.PlanarSDF(g["origin"], g["cell_size"], field) if planar else \
        G.SignedDistanceField(g["origin"], g["cell_size"], field)
    m = _model(g["arm"], g["spheres"])
    eps, r = g["obs_eps"], g["sphere_r"]
    for case in g["unary_cases"]:
        e, H = oracle.obstacle_factor(m, sdf, case["q"], eps)
        assert np.allclose(e, _sdf_to_err(case["sdf_exp"], eps + r), atol=g["err_tol"])
        Hn = _num_jac(lambda x: oracle.obstacle_factor(m, sdf, x, eps, False)[0], case["q"])
        assert np.allclose(H, Hn, atol=g["jacobian_numeric_tol"])
    gp = g["gp"]
    Qc = gp["Qc_sigma"] ** 2 * np.eye(2)
    for case in g["gp_cases"]:
        args = [np.array(case[k], dtype=float) for k in ("q1", "qdot1", "q2", "qdot2")]
        e, H = oracle.obstacle_gp_factor(m, sdf, Qc, gp["delta_t"], gp["tau"], *args, eps)
        assert np.allclose(e, _sdf_to_err(case["sdf_exp"], eps + r), atol=g["err_tol"])
        for k in range(4):
            def f(x, k=k):
                a = list(args)
                a[k] = x
                return oracle.obstacle_gp_factor(m, sdf, Qc, gp["delta_t"], gp["tau"], *a, eps, False)[0]
            assert np.allclose(H[k], _num_jac(f, args[k]), atol=g["jacobian_numeric_tol"])


def test_gp_interpolator_linear(golden, oracle):
    g = golden["gp_interpolator_linear"]
    Qc = g["Qc_scale"] * np.eye(g["dof"])
    for case in g["cases"]:
        args = [np.array(case[k], dtype=float) for k in ("p1", "v1", "p2", "v2")]
        p, H = oracle.gp_interpolate(g["dof"], False, Qc, g["delta_t"], g["tau"], *args)
        assert np.allclose(p, case["expect"], atol=g["tol"])
        for k in range(4):
            def f(x, k=k):
                a = list(args)
                a[k] = x
                return oracle.gp_interpolate(g["dof"], False, Qc, g["delta_t"], g["tau"], *a, False)[0]
            assert np.allclose(H[k], _num_jac(f, args[k]), atol=g["tol"])
    # closed-form 2x2 weights for dt=0.1, tau=0.03 -- and every D x D block of
    # Lambda/Psi is a scalar multiple of I for any SPD Qc (what the CUDA path relies on)
    L, P = oracle.gp_lambda_psi(3, Qc, 0.1, 0.03)
    assert np.allclose([L[0, 0], L[0, 3], P[0, 0], P[0, 3]], [0.784, 0.0147, 0.216, -0.0063], atol=1e-12)
    rng = np.random.default_rng(0)
    A = rng.standard_normal((3, 3))
    L2, P2 = oracle.gp_lambda_psi(3, A @ A.T + 0.5 * np.eye(3), 0.1, 0.03)
    for M in (L2, P2):
        for bi in range(2):
            for bj in range(2):
                blk = M[3 * bi:3 * bi + 3, 3 * bj:3 * bj + 3]
                assert np.allclose(blk, blk[0, 0] * np.eye(3), atol=1e-9)
    assert np.allclose(L2, L, atol=1e-9) and np.allclose(P2, P, atol=1e-9)


def test_gp_prior_linear(golden, oracle):
    g = golden["gp_prior_linear"]
    for case in g["zero_error_cases"]:
        e, H = oracle.gp_prior(g["dof"], False, g["delta_t"], case["p1"], case["v1"], case["p2"], case["v2"])
        assert np.allclose(e, 0, atol=1e-12)
    rng = np.random.default_rng(1)
    args = [rng.standard_normal(3) for _ in range(4)]
    e, H = oracle.gp_prior(3, False, 0.1, *args)
    for k in range(4):
        def f(x, k=k):
            a = list(args)
            a[k] = x
            return oracle.gp_prior(3, False, 0.1, *a, False)[0]
        assert np.allclose(H[k], _num_jac(f, args[k]), atol=1e-7)


def test_two_state_gauss_newton(golden, oracle):
    """Planner-level analogue of testGaussianProcessPriorLinear.cpp:144-202: a 2-state chain with a
    constant-velocity-consistent start/goal converges to zero error under Gauss-Newton."""
    o = golden["gp_prior_linear"]["optimization"]
    from gpmp2_b200 import synth
    arm = G.Arm(3, [0.5, 0.5, 0.5], [0, 0, 0], [0, 0, 0])
    model = G.ArmModel(arm, [G.BodySphere(2, 0.01, [0, 0, 0])])
    sdf = synth.planar_dataset("Empty")  # constant +1000 field: no obstacle cost
    st = G.TrajOptimizerSetting(3)
    st.set_total_step(1)
    st.set_total_time(golden["gp_prior_linear"]["delta_t"])
    st.set_Qc_model(golden["gp_prior_linear"]["Qc_scale"] * np.eye(3))
    st.set_conf_prior_model(o["prior_sigma"])
    st.set_vel_prior_model(o["prior_sigma"])
    st.setGaussNewton()
    st.set_max_iter(100)
    st.set_rel_thresh(1e-5)
    init = np.concatenate([o["p1init"], o["p2init"], o["v1init"], o["v2init"]])
    r = oracle.batch_optimize(model, sdf, o["p1"], o["v1"], o["p2"], o["v2"], init, st, dense=True)
    exp = np.concatenate([o["p1"], o["p2"], o["v1"], o["v2"]])
    assert r["error"][0] < o["tol"]
    assert np.allclose(r["traj"][0], exp, atol=o["tol"])


# ---------------------------------------------------------------------------------------------
# Pose2Vector (mobile manipulator) pieces of the oracle
# ---------------------------------------------------------------------------------------------
def _p2v_retract(x, d):
    """Pose2Vector chart retract: Pose2 part x * Pose2(d0, d1, d2), vector part add (ProductDynamicLieGroup.h:84-90)."""
    x = np.array(x, dtype=float)
    d = np.asarray(d, dtype=float)
    c, s = np.cos(x[2]), np.sin(x[2])
    out = x + d
    out[0] = x[0] + c * d[0] - s * d[1]
    out[1] = x[1] + s * d[0] + c * d[1]
    return out


def _p2v_local(a, b):
    """Pose2Vector local coordinates of b at a (Pose2 chart: between.x, between.y, between.theta)."""
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    c, s = np.cos(a[2]), np.sin(a[2])
    dx, dy = b[0] - a[0], b[1] - a[1]
    th = np.arctan2(np.sin(b[2] - a[2]), np.cos(b[2] - a[2]))
    return np.concatenate([[c * dx + s * dy, -s * dx + c * dy, th], b[3:] - a[3:]])


def _num_jac_lie(f, x, h=1e-6, out_lie=False):
    x = np.asarray(x, dtype=float)
    f0 = np.asarray(f(x))
    J = np.zeros((f0.size, x.size))
    for k in range(x.size):
        d = np.zeros(x.size)
        d[k] = h
        fp, fm = np.asarray(f(_p2v_retract(x, d))), np.asarray(f(_p2v_retract(x, -d)))
        J[:, k] = (_p2v_local(fm, fp) if out_lie else (fp - fm)) / (2 * h)
    return J


def _mobile_model(g, spheres):
    arm = G.Arm(2, g["a"], g["alpha"], g["d"])
    marm = G.Pose2MobileArm(arm, G.Pose3(R=_rot_z(g["base_T_arm_ypr"][0]), t=g["base_T_arm_t"]))
    return G.Pose2MobileArmModel(marm, [G.BodySphere(l, r, c) for l, r, c in spheres])


def test_pose2_mobile_arm_fk(golden, oracle):
    g = golden["pose2_mobile_arm"]
    m = _mobile_model(g, [[0, 0.2, [0.1, 0, 0.3]], [1, 0.1, [-0.5, 0, 0]], [2, 0.1, [0, 0.1, 0]], [2, 0.1, [-0.3, 0, 0.2]]])
    for case in g["cases"]:
        conf = np.concatenate([case["pose2"], case["q"]])
        poses, _ = oracle.forward_kinematics(m, conf)
        for i in range(3):
            assert np.allclose(poses[i][:3, 3], case["link_t"][i], atol=g["tol"])
            assert np.allclose(poses[i][:3, :3], _rot_z(case["link_yaw"][i]), atol=g["tol"])
    # sphere-centre Jacobians w.r.t. the Pose2Vector chart, numerically (the reference checks pose Jacobians
    # with numericalDerivativeDynamic at the same random state)
    js = g["jacobian_state"]
    conf = np.concatenate([js["pose2"], js["q"]])
    c, J = oracle.sphere_centers(m, conf)
    for s in range(4):
        Jn = _num_jac_lie(lambda x: oracle.sphere_centers(m, x, False)[0][s], conf)
        assert np.allclose(J[s], Jn, atol=g["jacobian_numeric_tol"])


def _ypr(y, p, r):
    """Rot3::Ypr(y, p, r) = Rz(y) Ry(p) Rx(r)  [GTSAM Rot3]."""
    cy, sy, cp, sp, cr, sr = np.cos(y), np.sin(y), np.cos(p), np.sin(p), np.cos(r), np.sin(r)
    Rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1.0]])
    Ry = np.array([[cp, 0, sp], [0, 1.0, 0], [-sp, 0, cp]])
    Rx = np.array([[1.0, 0, 0], [0, cr, -sr], [0, sr, cr]])
    return Rz @ Ry @ Rx


def _yaw_pose(d):
    return G.Pose3(R=_rot_z(d["yaw"]), t=d["t"])


def _other_mobile_model(name, g, jacobian_variant=False):
    """The robots of testPose2Mobile2Arms.cpp / testPose2MobileVetLinArm.cpp / testPose2MobileVetLin2Arms.cpp with one
    sphere on every link (+ an off-axis one on the last) so that the sphere-centre Jacobians cover every column."""
    arm = lambda: G.Arm(2, g["a"], g["alpha"], g["d"])
    bt = None
    if jacobian_variant and "jacobian_base_T_torso" in g:
        bt = G.Pose3(R=_ypr(*g["jacobian_base_T_torso"]["ypr"]), t=g["jacobian_base_T_torso"]["t"])
    if name == "pose2_mobile_2arms":
        marm = G.Pose2Mobile2Arms(arm(), arm(), _yaw_pose(g["base_T_arm1"]), _yaw_pose(g["base_T_arm2"]))
        cls = G.Pose2Mobile2ArmsModel
    elif name == "pose2_mobile_vetlin_arm":
        marm = G.Pose2MobileVetLinArm(arm(), bt, _yaw_pose(g["torso_T_arm"]), False)
        cls = G.Pose2MobileVetLinArmModel
    else:
        marm = G.Pose2MobileVetLin2Arms(arm(), arm(), bt, _yaw_pose(g["torso_T_arm1"]), _yaw_pose(g["torso_T_arm2"]), False)
        cls = G.Pose2MobileVetLin2ArmsModel
    L = marm.nr_links()
    spheres = [G.BodySphere(l, 0.1, [0.1 * (l + 1), -0.05 * l, 0.2]) for l in range(L)] + [G.BodySphere(L - 1, 0.1, [-0.3, 0.2, 0.1])]
    return marm, cls(marm, spheres)


@pytest.mark.parametrize("name", ["pose2_mobile_2arms", "pose2_mobile_vetlin_arm", "pose2_mobile_vetlin_2arms"])
def test_other_mobile_robots_fk(golden, oracle, name):
    """Link poses of Pose2Mobile2Arms / Pose2MobileVetLinArm / Pose2MobileVetLin2Arms against the reference's expected
    values, and the sphere-centre Jacobians against numerical derivatives in the Pose2Vector chart at the reference's
    random state (the reference checks its pose Jacobians the same way, with numericalDerivativeDynamic)."""
    g = golden[name]
    marm, m = _other_mobile_model(name, g)
    for case in g["cases"]:
        conf = np.concatenate([case["pose2"], case["q"]])
        assert conf.size == marm.dof()
        poses, _ = oracle.forward_kinematics(m, conf)
        assert len(poses) == marm.nr_links() == len(case["link_t"])
        for i in range(marm.nr_links()):
            assert np.allclose(poses[i][:3, 3], case["link_t"][i], atol=g["tol"]), (name, i)
            assert np.allclose(poses[i][:3, :3], _rot_z(case["link_yaw"][i]), atol=g["tol"]), (name, i)
    marm, m = _other_mobile_model(name, g, jacobian_variant=True)
    js = g["jacobian_state"]
    conf = np.concatenate([js["pose2"], js["q"]])
    c, J = oracle.sphere_centers(m, conf)
    for s in range(m.nr_body_spheres()):
        Jn = _num_jac_lie(lambda x: oracle.sphere_centers(m, x, False)[0][s], conf)
        assert np.allclose(J[s], Jn, atol=g["jacobian_numeric_tol"]), (name, s)
    # a reversed linear actuator lowers the torso instead (liftBasePose3, mobileBaseUtils.cpp:58-62)
    if "vetlin" in name:
        arm = G.Arm(2, g["a"], g["alpha"], g["d"])
        if name == "pose2_mobile_vetlin_arm":
            rm = G.Pose2MobileVetLinArmModel(G.Pose2MobileVetLinArm(arm, None, _yaw_pose(g["torso_T_arm"]), True), [G.BodySphere(1, 0.1, [0, 0, 0])])
        else:
            rm = G.Pose2MobileVetLin2ArmsModel(G.Pose2MobileVetLin2Arms(arm, arm, None, _yaw_pose(g["torso_T_arm1"]), _yaw_pose(g["torso_T_arm2"]), True),
                                               [G.BodySphere(1, 0.1, [0, 0, 0])])
        case = g["cases"][1]
        conf = np.concatenate([case["pose2"], case["q"]])
        poses, _ = oracle.forward_kinematics(rm, conf)
        assert np.isclose(poses[1][2, 3], -case["q"][0]) and np.allclose(poses[1][:2, 3], case["link_t"][1][:2])
        cc, Jr = oracle.sphere_centers(rm, conf)
        Jn = _num_jac_lie(lambda x: oracle.sphere_centers(rm, x, False)[0][0], conf)
        assert np.allclose(Jr[0], Jn, atol=1e-6)


def test_gp_interpolator_pose2vector(golden, oracle):
    g = golden["gp_interpolator_pose2vector"]
    D = g["dof"]
    Qc = g["Qc_scale"] * np.eye(D)
    for case in g["cases"]:
        args = [np.array(case[k], dtype=float) for k in ("p1", "v1", "p2", "v2")]
        p, H = oracle.gp_interpolate(D, True, Qc, g["delta_t"], g["tau"], *args)
        assert np.allclose(p, case["expect"], atol=g["tol"])
    rng = np.random.default_rng(5)
    args = [np.array([0.3, -0.2, 0.4, 0.1, 0.2, -0.3]), rng.standard_normal(D), np.array([0.5, 0.1, 0.9, -0.2, 0.4, 0.1]),
            rng.standard_normal(D)]
    p, H = oracle.gp_interpolate(D, True, Qc, g["delta_t"], g["tau"], *args)
    for k in range(4):
        def f(x, k=k):
            a = list(args)
            a[k] = x
            return oracle.gp_interpolate(D, True, Qc, g["delta_t"], g["tau"], *a, False)[0]
        if k in (0, 2):
            Jn = _num_jac_lie(f, args[k], out_lie=True)
        else:
            Jn = np.zeros((D, D))
            for c in range(D):
                d = np.zeros(D)
                d[c] = 1e-6
                Jn[:, c] = _p2v_local(f(args[k] - d), f(args[k] + d)) / 2e-6
        assert np.allclose(H[k], Jn, atol=g["jacobian_numeric_tol"]), k


def test_gp_prior_pose2vector(golden, oracle):
    g = golden["gp_prior_pose2vector"]
    D = g["dof"]
    for case in g["zero_error_cases"]:
        e, _ = oracle.gp_prior(D, True, g["delta_t"], case["p1"], case["v1"], case["p2"], case["v2"])
        assert np.allclose(e, 0, atol=1e-12)
    rng = np.random.default_rng(6)
    args = [np.array([0.3, -0.2, 0.4, 0.1, 0.2, -0.3]), rng.standard_normal(D), np.array([0.5, 0.1, 0.9, -0.2, 0.4, 0.1]),
            rng.standard_normal(D)]
    e, H = oracle.gp_prior(D, True, g["delta_t"], *args)
    for k in range(4):
        def f(x, k=k):
            a = list(args)
            a[k] = x
            return oracle.gp_prior(D, True, g["delta_t"], *a, False)[0]
        Jn = _num_jac_lie(f, args[k]) if k in (0, 2) else _num_jac(f, args[k])
        assert np.allclose(H[k], Jn, atol=1e-6), k


def test_two_state_gauss_newton_pose2vector(golden, oracle):
    """Planner-level analogue of testGaussianProcessPriorPose2Vector.cpp:147-196 (Gauss-Newton reaches zero error)."""
    g = golden["gp_prior_pose2vector"]
    o = g["optimization"]
    from gpmp2_b200 import synth
    arm = G.Arm(3, [0.3, 0.3, 0.3], [0, 0, 0], [0, 0, 0])
    model = G.Pose2MobileArmModel(G.Pose2MobileArm(arm), [G.BodySphere(0, 0.1, [0, 0, 0]), G.BodySphere(3, 0.05, [0, 0, 0])])
    sdf = synth.planar_dataset("Empty")
    st = G.TrajOptimizerSetting(6)
    st.set_total_step(1)
    st.set_total_time(g["delta_t"])
    st.set_Qc_model(g["Qc_scale"] * np.eye(6))
    st.set_conf_prior_model(o["prior_sigma"])
    st.set_vel_prior_model(o["prior_sigma"])
    st.setGaussNewton()
    st.set_max_iter(100)
    st.set_rel_thresh(1e-5)
    init = np.concatenate([o["pose1"], o["pose2"], o["v1"], o["v2init"]])
    r = oracle.batch_optimize(model, sdf, o["pose1"], o["v1"], o["pose2"], o["v1"], init, st, dense=True)
    assert r["error"][0] < o["tol"]
    assert np.allclose(r["traj"][0], np.concatenate([o["pose1"], o["pose2"], o["v1"], o["v1"]]), atol=o["tol"])


def test_traj_utils_interpolate_linear_golden(oracle):
    """testTrajUtils.cpp:28-53: constant velocity 10 over dt = 0.1, Qc = 0.01 I, 4 interpolated states."""
    traj = np.array([0, 0, 1, 0, 10, 0, 10, 0], dtype=float)   # [x0 x1 | v0 v1]
    out = oracle.interpolate_traj(False, 2, 1, 0.1, 0.01 * np.eye(2), 4, traj)[0].reshape(2, 6, 2)
    assert np.allclose(out[0, :, 0], [0, 0.2, 0.4, 0.6, 0.8, 1.0], atol=1e-6)
    assert np.allclose(out[0, :, 1], 0, atol=1e-6)
    assert np.allclose(out[1, :, 0], 10, atol=1e-6) and np.allclose(out[1, :, 1], 0, atol=1e-6)
    # the start_index / end_index overload over the same single interval is the same list of states
    assert np.array_equal(oracle.interpolate_traj(False, 2, 1, 0.1, 0.01 * np.eye(2), 4, traj, 0, 1)[0], out.ravel())


def test_traj_utils_init_straight_line(oracle):
    """testTrajUtils.cpp:56-66 (first state = the start Pose2Vector) and TrajUtils.cpp:23-48 for vectors."""
    s = np.array([1, 3, np.pi - 0.5, 2, 4]); e = np.array([3, 7, -np.pi + 0.5, 4, 8])
    t = oracle.init_straight_line(True, 5, 5, s, e)[0].reshape(2, 6, 5)
    assert np.allclose(t[0, 0], s, atol=1e-6)
    assert np.allclose(t[0, 5, :2], e[:2], atol=1e-9) and np.allclose(t[0, 5, 3:], e[3:], atol=1e-12)
    # the commented-out expectation of the reference test (:63-64) is the Euclidean one; the Lie interpolation it
    # actually calls takes the short way round through theta = pi (1.0 rad in 5 steps) instead
    assert np.allclose(np.unwrap(t[0, :, 2]), np.pi - 0.5 + 0.2 * np.arange(6), atol=1e-9)
    assert np.allclose(t[1], (e - s) / 5.0)
    v = oracle.init_straight_line(False, 3, 4, [0, 1, 2], [4, 5, -2])[0].reshape(2, 5, 3)
    assert np.array_equal(v[0, 0], [0, 1, 2]) and np.array_equal(v[0, 4], [4, 5, -2])
    assert np.allclose(v[0, 2], [2, 3, 0]) and np.allclose(v[1], [[1, 1, -1]] * 5)
    assert np.allclose(v, G.straight_line_traj(np.array([[0, 1, 2.0]]), np.array([[4, 5, -2.0]]), 4).reshape(2, 5, 3))


from conftest import limit_problem as _limit_problem  # noqa: E402


@pytest.mark.parametrize("lie", [False, True])
def test_joint_limit_factor_golden(oracle, lie):
    """testJointLimitFactorVector.cpp:25-64 / testJointLimitFactorPose2Vector.cpp:25-66: limits (-5,-10)..(5,10),
    threshold 2, unit sigma: error (0,0) at the origin, (7,2) at (-10,-10) and at (10,10)."""
    for conf, e in (((0.0, 0.0), (0.0, 0.0)), ((-10.0, -10.0), (7.0, 2.0)), ((10.0, 10.0), (7.0, 2.0))):
        model, sdf, st, x, traj = _limit_problem(conf, lie)
        z = np.zeros_like(x)
        err = oracle.graph_error(model, sdf, x, z, x, z, traj, st)[0]
        assert abs(err - 2 * 0.5 * (e[0] ** 2 + e[1] ** 2)) < 1e-9      # two states, 0.5 |e|^2 each


def test_pose2vector_group_ops_golden(oracle):
    """testPose2Vector.cpp:61-113 (pose part; the vector part is plain addition): compose, between, inverse."""
    h = np.pi / 2
    def same(p, q):
        return np.allclose(p[:2], q[:2], atol=1e-9) and abs(np.angle(np.exp(1j * (p[2] - q[2])))) < 1e-9
    assert same(oracle.pose2_op("compose", [1, 1, h], [1, 1, h]), [0, 2, np.pi])
    assert same(oracle.pose2_op("between", [1, 1, h], [0, 2, np.pi]), [1, 1, h])
    assert same(oracle.pose2_op("inverse", [1, 1, h]), [-1, 1, -h])


def test_pose2_expmap_logmap_roundtrip(oracle):
    """testPose2Vector.cpp:34-46: logmap(expmap(d)) = d for d = (0.1, 0.2, 0.3 | 4, 5, 6) (pose part)."""
    import ctypes as C
    v = np.array([0.1, 0.2, 0.3]); p = np.zeros(3); back = np.zeros(3)
    lib = oracle.lib()
    assert lib.orc_pose2_expmap(v.ctypes.data_as(C.c_void_p), p.ctypes.data_as(C.c_void_p)) == 0
    assert lib.orc_pose2_logmap(p.ctypes.data_as(C.c_void_p), back.ctypes.data_as(C.c_void_p)) == 0
    assert np.allclose(back, v, atol=1e-12)
    # Pose2::Expmap in closed form: theta = w, t = (v_ortho - R v_ortho) / w with v_ortho = (-v_y, v_x)
    w = 0.3; c, s = np.cos(w), np.sin(w); ox, oy = -0.2, 0.1
    assert np.allclose(p, [(ox - (c * ox - s * oy)) / w, (oy - (s * ox + c * oy)) / w, w], atol=1e-12)


def test_mobile_base_utils_golden(oracle):
    """testMobileBaseUtils.cpp:25-66 (computeBaseTransPose3): the arm base of a Pose2MobileArm is
    Pose3(Yaw(theta), (x, y, 0)) * base_T_arm.  Read off as link 1 of a one-joint arm with trivial DH parameters at q = 0."""
    for bta_yaw, bta_t, p2, yaw, t in ((0.0, [0, 0, 0], [0, 0, 0], 0.0, [0, 0, 0]),
                                       (0.0, [0, 0, 0], [1.3, 4.5, -0.3], -0.3, [1.3, 4.5, 0]),
                                       (-0.3, [1, 1, 2], [0, 0, 0], -0.3, [1, 1, 2]),
                                       (-0.3, [1, 1, 2], [2, -2, np.pi / 2], np.pi / 2 - 0.3, [1, -1, 2])):
        marm = G.Pose2MobileArm(G.Arm(1, [0.0], [0.0], [0.0]), G.Pose3(R=_rot_z(bta_yaw), t=bta_t))
        m = G.Pose2MobileArmModel(marm, [G.BodySphere(0, 0.1, [0, 0, 0])])
        poses, _ = oracle.forward_kinematics(m, np.array(p2 + [0.0]))
        assert np.allclose(poses[0][:3, 3], [p2[0], p2[1], 0], atol=1e-9) and np.allclose(poses[0][:3, :3], _rot_z(p2[2]), atol=1e-9)
        assert np.allclose(poses[1][:3, 3], t, atol=1e-9) and np.allclose(poses[1][:3, :3], _rot_z(yaw), atol=1e-9)


def test_joint_limit_optimization_golden(oracle):
    """testJointLimitFactorVector.cpp:71-157: Gauss-Newton lands on the limit (threshold included): (0,0) stays,
    (-10,-10) -> (-3,-8), (10,10) -> (3,8), tol 1e-6."""
    from conftest import limit_optimization_problem
    for conf, want in (((0.0, 0.0), (0.0, 0.0)), ((-10.0, -10.0), (-3.0, -8.0)), ((10.0, 10.0), (3.0, 8.0))):
        model, sdf, st, x, traj = limit_optimization_problem(conf)
        z = np.zeros(2)
        r = oracle.batch_optimize(model, sdf, x, z, x, z, traj, st)
        t = r["traj"][0].reshape(2, 2, 2)
        assert np.allclose(t[0], [want, want], atol=1e-6) and np.allclose(t[1], 0.0, atol=1e-6)


# ---------------------------------------------------------------------------------------------
# workspace goal factor: GoalFactorArm / GaussianPriorWorkspacePositionArm
# ---------------------------------------------------------------------------------------------
def test_goal_factor_golden(golden, oracle):
    """testGoalFactorArm.cpp:26-73 / testGaussianPriorWorkspacePosition.cpp:26-76: errors and numerical Jacobians."""
    g = golden["goal_factor_arm"]
    model = _model(g, [(0, 0.1, [0, 0, 0])])
    for c in g["cases"]:
        for link in (-1, 1):      # GoalFactorArm (last joint frame) and GaussianPriorWorkspacePositionArm(joint = 1)
            e, H = oracle.goal_factor(model, c["q"], c["goal"], link)
            assert np.allclose(e, c["expect"], atol=g["tol"])
            Hn = _num_jac(lambda q: oracle.goal_factor(model, q, c["goal"], link, want_H=False)[0], c["q"])
            assert np.allclose(H, Hn, atol=g["tol"])
    # an inner joint frame of a 3-D arm: analytic vs numerical Jacobian
    wam = golden["arm_wam"]
    arm = G.Arm(7, wam["a"], [x * np.pi for x in wam["alpha_over_pi"]], wam["d"])
    m7 = G.ArmModel(arm, [G.BodySphere(0, 0.1, [0, 0, 0])])
    for link in (2, 4, 6):
        e, H = oracle.goal_factor(m7, wam["q"], [0.1, 0.2, 0.3], link)
        assert np.allclose(e, np.asarray(wam["link_t"][link]) - [0.1, 0.2, 0.3], atol=1e-3)   # testArm.cpp:283-309 (4 digits)
        assert np.allclose(e, oracle.forward_kinematics(m7, np.asarray(wam["q"]))[0][link][:3, 3] - [0.1, 0.2, 0.3], atol=1e-12)
        Hn = _num_jac(lambda q: oracle.goal_factor(m7, q, [0.1, 0.2, 0.3], link, want_H=False)[0], wam["q"])
        assert np.allclose(H, Hn, atol=1e-6) and np.all(H[:, link + 1:] == 0.0)


from conftest import goal_ik_problem as _goal_ik_problem  # noqa: E402


def test_goal_factor_lm_inverse_kinematics(golden, oracle):
    """testGoalFactorArm.cpp:77-107 / testGaussianPriorWorkspacePosition.cpp:80-108: LM solves the inverse kinematics of
    the 2-link arm from q = 0 (the only LevenbergMarquardtOptimizer known answer the reference holds for this path).
    Planner-level analogue: see conftest.goal_ik_problem."""
    o = golden["goal_factor_arm"]["optimization"]
    model, sdf, st, start, end, init = _goal_ik_problem(golden["goal_factor_arm"])
    r = oracle.batch_optimize(model, sdf, start, np.zeros(2), end, np.zeros(2), init, st)
    q = r["traj"][0].reshape(2, 2, 2)[0, 1]
    tip, _ = oracle.goal_factor(model, q, o["goal"], -1, want_H=False)
    assert r["error"][0] < o["tol"]                      # EXPECT_DOUBLES_EQUAL(0, graph.error(results), 1e-3)
    assert np.linalg.norm(tip) < o["tol"]
    # the goal sits at full reach, where the cost is quartic in q_2: gpmp2::optimize stops on its fixed absolute
    # tolerance 1e-5 (the reference test sets 1e-12 to get q to 1e-3), so q is only checked to the resulting accuracy
    assert np.allclose(q, o["q"], atol=5e-2)


# ---------------------------------------------------------------------------------------------
# self-collision factor: SelfCollisionArm
# ---------------------------------------------------------------------------------------------
def test_self_collision_golden(golden, oracle):
    """testSelfCollision.cpp:26-56: error vector and numerical Jacobian of the 3-link arm's two sphere pairs."""
    g = golden["self_collision_arm"]
    model = _model(g, g["spheres"])
    e, H = oracle.self_collision_factor(model, g["q"], g["data"])
    assert np.allclose(e, g["expect"], atol=g["tol"])
    Hn = _num_jac(lambda q: oracle.self_collision_factor(model, q, g["data"], want_H=False)[0], g["q"])
    assert np.allclose(H, Hn, atol=g["tol"])
    # hinge off: pairs farther apart than r_A + r_B + epsilon contribute nothing
    far = [[0, 3, 0.1, 0.1]]
    e, H = oracle.self_collision_factor(model, g["q"], far)
    assert e[0] == 0.0 and np.all(H == 0.0)


def test_self_collision_in_graph(golden, oracle):
    """The factor sits on every support state with Diagonal::Sigmas(data.col(3)): graph error of two identical states."""
    g = golden["self_collision_arm"]
    model = _model(g, g["spheres"])
    sdf = G.PlanarSDF([-20.0, -20.0], 1.0, np.full((40, 40), 1000.0))
    st = G.TrajOptimizerSetting(3)
    st.set_total_step(1); st.set_total_time(1.0); st.set_obs_check_inter(0); st.setLM()
    x = np.asarray(g["q"]); z = np.zeros(3)
    traj = np.concatenate([x, x, z, z])
    e0 = oracle.graph_error(model, sdf, x, z, x, z, traj, st)[0]
    st.set_self_collision(g["data"])
    e1 = oracle.graph_error(model, sdf, x, z, x, z, traj, st)[0]
    want = 2 * 0.5 * sum((e / row[3]) ** 2 for e, row in zip(g["expect"], g["data"]))
    assert abs((e1 - e0) - want) < 1e-6 * want


# ---------------------------------------------------------------------------------------------
# vehicle dynamics factor: VehicleDynamicsFactorPose2Vector
# ---------------------------------------------------------------------------------------------
def test_vehicle_dynamics_golden(golden, oracle):
    """testVehicleDynamics.cpp:23-95: simple2DVehicleDynamicsPose2(p, v) = v(1) whatever the pose.  Read off the graph
    error of a two-state Pose2MobileArm problem with and without the factor (both states carry it, Isotropic sigma)."""
    from gpmp2_b200 import synth
    model = synth.mobile_two_links_arm()
    sdf = G.PlanarSDF([-50.0, -50.0], 1.0, np.full((100, 100), 1000.0))
    sigma = 0.5
    for c in golden["vehicle_dynamics_pose2"]["cases"]:
        st = G.TrajOptimizerSetting(5)
        st.set_total_step(1); st.set_total_time(1.0); st.set_obs_check_inter(0); st.setLM()
        x = np.array(c["p"] + [0.1, -0.2]); v = np.array(c["v"] + [0.3, 0.4])
        traj = np.concatenate([x, x, v, v])
        e0 = oracle.graph_error(model, sdf, x, v, x, v, traj, st)[0]
        st.set_vehicle_dynamics(sigma)
        e1 = oracle.graph_error(model, sdf, x, v, x, v, traj, st)[0]
        assert abs((e1 - e0) - 2 * 0.5 * (c["c"] / sigma) ** 2) < 1e-9


# ---------------------------------------------------------------------------------------------
# replanning re-solve: ISAM2TrajOptimizer::fixConfigAndVel as a PriorFactor pair on one support state
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("lie", [False, True])
def test_fixed_state_prior_in_the_oracle(oracle, lie):
    """ISAM2TrajOptimizer-inl.h:160-168: PriorFactor<Pose>(x_k, conf_fix, conf_prior_model) + PriorFactor<Velocity>(v_k,
    vel_fix, vel_prior_model).  PriorFactor's error is Local(x, prior) whitened by the isotropic sigma: read off the graph
    error with and without the pair, its Hessian off the dense normal equations (+ 1/sigma^2 on the diagonal of state k),
    per problem (one target row each)."""
    from gpmp2_b200 import synth
    model = synth.mobile_two_links_arm() if lie else synth.simple_two_links_arm()
    D = 5 if lie else 2
    sdf = G.PlanarSDF([-50.0, -50.0], 1.0, np.full((100, 100), 1000.0))
    st = G.TrajOptimizerSetting(D)
    st.set_total_step(3); st.set_total_time(1.5); st.set_obs_check_inter(1); st.setLM()
    st.set_conf_prior_model(0.2); st.set_vel_prior_model(0.5)
    rng = np.random.default_rng(7)
    B, N, k = 3, 4, 2
    traj = 0.3 * rng.standard_normal((B, 2 * N * D))
    sc, ec = traj[:, :D].copy(), traj[:, (N - 1) * D:N * D].copy()
    z = np.zeros((B, D))
    e0 = oracle.graph_error(model, sdf, sc, z, ec, z, traj, st)
    lin0 = oracle.linearize(model, sdf, sc, z, ec, z, traj, st, want_dense=True)
    fc, fv = 0.3 * rng.standard_normal((B, D)), 0.3 * rng.standard_normal((B, D))
    st.fix_config_and_vel(k, fc, fv)
    e1 = oracle.graph_error(model, sdf, sc, z, ec, z, traj, st)
    lin1 = oracle.linearize(model, sdf, sc, z, ec, z, traj, st, want_dense=True)
    for p in range(B):
        x, v = traj[p, k * D:(k + 1) * D], traj[p, (N + k) * D:(N + k + 1) * D]
        ex = _p2v_local(x, fc[p]) if lie else fc[p] - x            # |Local(x, prior)| (the sign does not matter for the error)
        want = 0.5 * (np.dot(ex, ex) / 0.2 ** 2 + np.dot(v - fv[p], v - fv[p]) / 0.5 ** 2)
        assert abs((e1[p] - e0[p]) - want) < 1e-9 * max(1.0, want)
    dH = lin1["dense_H"][0] - lin0["dense_H"][0]
    want_H = np.zeros_like(dH)
    b = 2 * D
    for d in range(D):
        want_H[k * b + d, k * b + d] = 1.0 / 0.2 ** 2
        want_H[k * b + D + d, k * b + D + d] = 1.0 / 0.5 ** 2
    assert np.abs(dH - want_H).max() < 1e-9
    st.clear_fixed_state()
    assert np.allclose(oracle.graph_error(model, sdf, sc, z, ec, z, traj, st), e0, rtol=0, atol=0)


# ---------------------------------------------------------------------------------------------
# workspace orientation prior: GaussianPriorWorkspaceOrientationArm
# ---------------------------------------------------------------------------------------------
def test_workspace_orientation_golden(golden, oracle):
    """testGaussianPriorWorkspaceOrientation.cpp:26-45: error vector and numerical Jacobian."""
    g = golden["workspace_orientation_arm"]
    model = _model(g, [(0, 0.1, [0, 0, 0])])
    des = _rot_z(g["des_yaw"])
    e, H = oracle.orientation_factor(model, g["q"], des, g["link"])
    assert np.allclose(e, g["expect"], atol=g["tol"])
    Hn = _num_jac(lambda q: oracle.orientation_factor(model, q, des, g["link"], want_H=False)[0], g["q"])
    assert np.allclose(H, Hn, atol=g["tol"])
    # WAM, random desired rotations: analytic vs numerical Jacobian; zero error at the frame's own orientation
    wam = golden["arm_wam"]
    m7 = G.ArmModel(G.Arm(7, wam["a"], [x * np.pi for x in wam["alpha_over_pi"]], wam["d"]), [G.BodySphere(0, 0.1, [0, 0, 0])])
    rng = np.random.default_rng(1)
    for link in (3, 6):
        Q, _ = np.linalg.qr(rng.standard_normal((3, 3)))
        Q *= np.sign(np.linalg.det(Q))
        e, H = oracle.orientation_factor(m7, wam["q"], Q, link)
        Hn = _num_jac(lambda q: oracle.orientation_factor(m7, q, Q, link, want_H=False)[0], wam["q"])
        assert np.allclose(H, Hn, atol=1e-6)
        R = oracle.forward_kinematics(m7, np.asarray(wam["q"]))[0][link][:3, :3]
        assert np.allclose(oracle.orientation_factor(m7, wam["q"], R, link, want_H=False)[0], 0.0, atol=1e-7)


# ---------------------------------------------------------------------------------------------
# 6-D workspace pose prior: GaussianPriorWorkspacePoseArm
# ---------------------------------------------------------------------------------------------
def test_workspace_pose_golden(golden, oracle):
    """testGaussianPriorWorkspacePose.cpp:26-46: error vector [omega; u] and numerical Jacobian (pins the restated
    Pose3::Logmap / LogmapDerivative)."""
    g = golden["workspace_pose_arm"]
    model = _model(g, [(0, 0.1, [0, 0, 0])])
    e, H = oracle.pose_factor(model, g["q"], np.eye(3), [0, 0, 0], g["link"])
    assert np.allclose(e, g["expect"], atol=g["tol"])
    Hn = _num_jac(lambda q: oracle.pose_factor(model, q, np.eye(3), [0, 0, 0], g["link"], want_H=False)[0], g["q"])
    assert np.allclose(H, Hn, atol=g["tol"])
    wam = golden["arm_wam"]
    m7 = G.ArmModel(G.Arm(7, wam["a"], [x * np.pi for x in wam["alpha_over_pi"]], wam["d"]), [G.BodySphere(0, 0.1, [0, 0, 0])])
    rng = np.random.default_rng(2)
    for link in (3, 6):
        Q, _ = np.linalg.qr(rng.standard_normal((3, 3)))
        Q *= np.sign(np.linalg.det(Q))
        t = rng.standard_normal(3) * 0.3
        e, H = oracle.pose_factor(m7, wam["q"], Q, t, link)
        Hn = _num_jac(lambda q: oracle.pose_factor(m7, q, Q, t, link, want_H=False)[0], wam["q"])
        assert np.allclose(H, Hn, atol=1e-6)
        # the rotational part is the orientation prior's error
        assert np.allclose(e[:3], oracle.orientation_factor(m7, wam["q"], Q, link, want_H=False)[0], atol=1e-12)


from conftest import pose_ik_problem as _pose_ik_problem  # noqa: E402


def test_workspace_pose_lm_inverse_kinematics(golden, oracle):
    """testGaussianPriorWorkspacePose.cpp:50-78: LM drives the 2-link arm from (pi/2, pi/2) to the pose Pose3(I, (2,0,0)),
    i.e. q = (0, 0), graph error < 1e-3.  Planner-level analogue: conftest.pose_ik_problem."""
    o = golden["workspace_pose_arm"]["optimization"]
    model, sdf, st, start, end, init = _pose_ik_problem(o)
    r = oracle.batch_optimize(model, sdf, start, np.zeros(2), end, np.zeros(2), init, st)
    q = r["traj"][0].reshape(2, 2, 2)[0, 1]
    assert r["error"][0] < o["tol"] and np.allclose(q, o["q"], atol=o["tol"])


# ------------------------------------------------------------------------------------------------
# a second, independently written optimizer over the oracle's pinned factor level
# ------------------------------------------------------------------------------------------------
def _numpy_optimize(oracle, model, sdf, sc, ec, init, st, opt, lie=False):
    """gpmp2::optimize (gpmp2/planner/BatchTrajOptimizer.cpp:212-308) over GTSAM's LM / GN as SURVEY.md App. B describes
    them, written a second time: dense normal equations from the oracle's linearize (itself pinned to the reference's
    factor vectors), numpy's LU solve instead of the oracle's banded Cholesky, the accept / reject logic in Python.
    One vector-space problem.  Returns (trajectory, error, accepted iterations)."""
    D, N = st.dof, st.total_step + 1
    z = np.zeros(D)

    def to_theta(traj):      # [x/v][state][D] -> state-grouped unknowns [state][x; v]
        return traj.reshape(2, N, D).transpose(1, 0, 2).reshape(-1).copy()

    def to_traj(theta):
        return theta.reshape(N, 2, D).transpose(1, 0, 2).reshape(1, -1).copy()

    def err(theta):
        return float(oracle.graph_error(model, sdf, sc, z, ec, z, to_traj(theta), st)[0])

    def retract(theta, d):
        """Values::retract: vectors add; Pose2Vector (ProductDynamicLieGroup.h:84-90) composes the pose with
        Pose2(dx, dy, dtheta) -- GTSAM's default Pose2 chart -- and adds the rest."""
        if not lie:
            return theta + d
        t, dd = theta.reshape(N, 2, D).copy(), d.reshape(N, 2, D)
        c, s_ = np.cos(t[:, 0, 2]), np.sin(t[:, 0, 2])
        t[:, 0, 0] += c * dd[:, 0, 0] - s_ * dd[:, 0, 1]
        t[:, 0, 1] += s_ * dd[:, 0, 0] + c * dd[:, 0, 1]
        th = t[:, 0, 2] + dd[:, 0, 2]
        t[:, 0, 2] = np.arctan2(np.sin(th), np.cos(th))
        t[:, 0, 3:] += dd[:, 0, 3:]
        t[:, 1] += dd[:, 1]
        return t.reshape(-1)

    theta = to_theta(np.asarray(init, dtype=np.float64))
    lam, delta, iters = 100.0, 0.2, 0                       # lambdaInitial: BatchTrajOptimizer.cpp:226
    error = err(theta)
    while True:
        cur, last_theta = error, theta
        lin = oracle.linearize(model, sdf, sc, z, ec, z, to_traj(theta), st, want_dense=True)
        H, g = lin["dense_H"][0], lin["g"][0].reshape(-1)
        if opt == "gn":
            theta = retract(theta, np.linalg.solve(H, -g))
            error = err(theta)
            iters += 1
        elif opt == "dogleg":
            # DoglegOptimizerImpl::Iterate, ONE_STEP_PER_ITERATION (App. B); Delta_0 = 0.2: BatchTrajOptimizer.cpp:219-222
            dx_n = np.linalg.solve(H, -g)
            dx_u = -(g @ g) / (g @ H @ g) * g
            f = error
            while True:
                if delta ** 2 < dx_u @ dx_u:
                    d = np.sqrt(delta ** 2 / (dx_u @ dx_u)) * dx_u
                elif delta ** 2 < dx_n @ dx_n:
                    a = dx_u @ dx_u - 2.0 * (dx_u @ dx_n) + dx_n @ dx_n
                    bq, c = 2.0 * (dx_u @ dx_n - dx_u @ dx_u), dx_u @ dx_u - delta ** 2
                    sq = np.sqrt(bq * bq - 4.0 * a * c)
                    t1, t2 = (-bq + sq) / (2.0 * a), (-bq - sq) / (2.0 * a)
                    tau = t1 if 0.0 <= t1 <= 1.0 else t2
                    d = (1.0 - tau) * dx_u + tau * dx_n
                else:
                    d = dx_n
                new_f = err(retract(theta, d))
                new_m = f + g @ d + 0.5 * d @ H @ d
                rho = 0.5 if abs(f - new_f) < 1e-15 or abs(f - new_m) < 1e-15 else (f - new_f) / (f - new_m)
                if rho >= 0.75:
                    delta = max(delta, 3.0 * np.sqrt(d @ d))
                    break
                if rho >= 0.25:
                    break
                if rho >= 0.0:
                    if delta > 1e-5:
                        delta *= 0.5
                    break
                if delta > 1e-5:
                    delta *= 0.5
                else:
                    d, new_f = np.zeros_like(d), f
                    break
            theta, error, iters = retract(theta, d), new_f, iters + 1
        else:
            while True:                                      # tryLambda
                d = np.linalg.solve(H + lam * np.eye(H.shape[0]), -g)
                lin_change = -(g @ d + 0.5 * d @ H @ d)      # linear.error(0) - linear.error(d), undamped system
                ok = False
                if lin_change >= 0:
                    new_error = err(retract(theta, d))
                    if lin_change > np.finfo(float).eps * abs(error):
                        ok = (error - new_error) / lin_change > 1e-3
                    stop = abs(error - new_error) < st.rel_thresh * error
                else:
                    stop = False
                if ok:
                    theta, error, lam, iters = retract(theta, d), new_error, lam / 10.0, iters + 1
                    break
                if stop:
                    break
                lam *= 10.0
                if lam >= 1e5:
                    break
        # checkConvergence(rel, abs = 1e-5, err = 0) and the iteration cap
        dec = cur - error
        if iters >= st.max_iter or error <= 0.0 or (st.rel_thresh and dec / cur <= st.rel_thresh) or dec <= 1e-5:
            break
    if error > cur:                                          # final_iter_no_increase: BatchTrajOptimizer.cpp:297-307
        theta, error = last_theta, cur
    return to_traj(theta)[0], error, iters


@pytest.mark.parametrize("opt", ["lm", "gn", "dogleg"])
def test_optimizer_loop_against_an_independent_dense_restatement(oracle, opt):
    """The oracle's optimizer level (block-banded Cholesky + LM / GN control flow in C++) against a second restatement
    in numpy that shares only the factor level with it.  This does not pin the level to GTSAM (nothing here can), it
    removes implementation slips of the oracle as an explanation: trajectories agree to solver round-off, iteration
    counts exactly."""
    from gpmp2_b200 import synth
    model, sdf = synth.wam_arm(), synth.wam_desk_dataset(60)
    st = synth.bench_setting(7, max_iter=10)
    if opt == "gn":
        st.setGaussNewton()
    elif opt == "dogleg":
        st.setDogleg()
    pr = synth.wam_problems(6, mode="restart", seed=22, sigma=0.6)     # perturbed restarts: LM rejects steps here
    z = np.zeros((6, 7))
    ref = oracle.batch_optimize(model, sdf, pr["start_conf"], z, pr["end_conf"], z, pr["init_traj"], st)
    assert ref["iters"].max() >= 2            # the loop is exercised, including rejected lambdas for LM
    if opt == "lm":
        assert (ref["counts"][:, 1] > ref["iters"]).any()          # more solves than accepted steps
        assert (ref["status"] & 8).any()                           # and LM giving up at lambda >= 1e5 (ST_LAMBDA_MAXED)
    elif opt == "gn":
        assert (ref["status"] & 64).any()                          # ST_ERR_INCREASED: the previous iterate is returned
    else:
        assert (ref["counts"][:, 2] > ref["iters"] + 1).any()      # trust-region shrinks: more error evaluations than steps
    for k in range(6):
        traj, error, iters = _numpy_optimize(oracle, model, sdf, pr["start_conf"][k], pr["end_conf"][k], pr["init_traj"][k], st, opt)
        assert iters == ref["iters"][k]
        assert np.abs(traj - ref["traj"][k]).max() < 1e-8, k
        assert abs(error - ref["error"][k]) <= 1e-9 * max(1.0, abs(error))


@pytest.mark.parametrize("opt", ["lm", "gn", "dogleg"])
def test_optimizer_loop_independent_restatement_pose2vector(oracle, opt):
    """The same second restatement for Pose2Vector states (config 4's Pose2MobileArm in MobileMap1): the retraction is the
    Pose2 chart composed per support state."""
    from gpmp2_b200 import synth
    cfg = synth.baseline_config("mobile")
    model, sdf, st = cfg["model"], cfg["sdf"], cfg["setting"]
    if opt == "gn":
        st.setGaussNewton()
    elif opt == "dogleg":
        st.setDogleg()
    pr = cfg["problems"](6, 9)
    z = np.zeros((6, 5))
    ref = oracle.batch_optimize(model, sdf, pr["start_conf"], z, pr["end_conf"], z, pr["init_traj"], st)
    assert ref["iters"].max() >= 2
    for k in range(6):
        traj, error, iters = _numpy_optimize(oracle, model, sdf, pr["start_conf"][k], pr["end_conf"][k], pr["init_traj"][k], st, opt, lie=True)
        assert iters == ref["iters"][k], k
        d = traj - ref["traj"][k]
        d.reshape(2, -1, 5)[0, :, 2] = np.arctan2(np.sin(d.reshape(2, -1, 5)[0, :, 2]), np.cos(d.reshape(2, -1, 5)[0, :, 2]))
        assert np.abs(d).max() < 1e-8, k
        assert abs(error - ref["error"][k]) <= 1e-9 * max(1.0, abs(error))
