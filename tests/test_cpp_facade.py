"""The header-only C++ facade (include/gpmp2b/gpmp2.hpp): compiles and links against the C-ABI library on CPU;
on the GPU box its output matches the oracle."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "test_facade")


def _build():
    lib = os.path.join(ROOT, "gpmp2_b200", "csrc")
    src = os.path.join(ROOT, "tests", "cpp", "test_facade.cpp")
    deps = [src, os.path.join(ROOT, "include", "gpmp2b", "gpmp2.hpp"), os.path.join(ROOT, "include", "gpmp2b.h")]
    if not os.path.exists(EXE) or os.path.getmtime(EXE) < max(os.path.getmtime(d) for d in deps):
        subprocess.check_call(["g++", "-std=c++14", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), src, "-o", EXE,
                               "-L", lib, "-lgpmp2b", "-Wl,-rpath," + lib])
    return EXE


def test_facade_compiles_and_fails_loudly_without_gpu():
    exe = _build()
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    p = subprocess.run([exe], capture_output=True, text=True)
    assert p.returncode == 3 and "no CPU fallback" in p.stderr


@pytest.mark.gpu
def test_facade_matches_oracle(oracle, golden):
    import gpmp2_b200 as G
    exe = _build()
    p = subprocess.run([exe], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    out = dict((ln.split()[0], ln.split()[1:]) for ln in p.stdout.strip().splitlines())
    assert "ok" in out and "velocity limit <= 0" in " ".join(out["threw"]) and out.get("trajutils") == ["ok"] and out.get("othermobile") == ["ok"]
    g = golden["obstacle_planar_sdf_factor_arm"]
    sdf = G.PlanarSDF(g["origin"], g["cell_size"], np.array(g["field"]))
    model = G.ArmModel(G.Arm(2, g["arm"]["a"], g["arm"]["alpha"], g["arm"]["d"], G.Pose3(t=g["arm"]["base_t"])),
                       [G.BodySphere(l, r, c) for l, r, c in g["spheres"]])
    st = G.TrajOptimizerSetting(2)
    st.set_total_step(4); st.set_total_time(2.0); st.set_epsilon(1.0); st.set_cost_sigma(0.5)
    st.set_obs_check_inter(2); st.setLM(); st.set_max_iter(8); st.set_rel_thresh(1e-6)
    s, e, z = np.array([0.0, 0.0]), np.array([np.pi / 2, 0.0]), np.zeros(2)
    ref = oracle.batch_optimize(model, sdf, s, z, e, z, G.straight_line_traj(s[None], e[None], 4), st)
    t = ref["traj"][0].reshape(2, 5, 2)
    for i in range(5):
        assert np.allclose([float(v) for v in out["x%d" % i]], t[0, i], atol=1e-6)
        assert np.allclose([float(v) for v in out["v%d" % i]], t[1, i], atol=1e-6)
    assert abs(float(out["coll_cost"][0]) - ref["coll_cost"][0]) < 1e-6
    # the same call with the optional factors set through the C++ setters
    st.set_workspace_pose_goal([[0, -1, 0], [1, 0, 0], [0, 0, 1]], [0.6, 3.4, 0.0], 0.2)
    st.set_self_collision([[0, 3, 0.2, 0.5], [1, 2, 0.1, 0.3]])
    st.set_workspace_orientation(np.eye(3), 2.0, 0, 1, 3)
    rex = oracle.batch_optimize(model, sdf, s, z, e, z, G.straight_line_traj(s[None], e[None], 4), st)
    tex = rex["traj"][0].reshape(2, 5, 2)
    assert np.abs(tex - t).max() > 1e-3          # the factors change the answer
    for i in range(5):
        assert np.allclose([float(v) for v in out["ex%d" % i]], tex[0, i], atol=1e-6)
