"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol the header
declares, fails loudly without a GPU, and the host-side mirror packs arguments like the reference."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import gpmp2_b200 as G
from gpmp2_b200 import _abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "gpmp2b.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gpmp2b_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = _abi.load_library()
    syms = _header_symbols()
    assert len(syms) >= 14
    for s in syms:
        assert hasattr(lib, s), "missing export " + s
        assert s in _abi.PROTOTYPES, "ctypes mirror lacks " + s
    assert b"gpmp2b" in lib.gpmp2b_version()


def test_struct_sizes_match_header():
    # compile-time sizes of the header's structs, through a tiny C program
    import subprocess
    import tempfile
    code = '#include <stdio.h>\n#include "gpmp2b.h"\nint main(){printf("%zu %zu %zu\\n", sizeof(gpmp2b_robot_desc), sizeof(gpmp2b_sdf_desc), sizeof(gpmp2b_setting));return 0;}\n'
    with tempfile.TemporaryDirectory() as td:
        open(os.path.join(td, "t.c"), "w").write(code)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", os.path.join(td, "t"), os.path.join(td, "t.c")])
        out = subprocess.check_output([os.path.join(td, "t")]).split()
    assert [int(x) for x in out] == [C.sizeof(_abi.RobotDesc), C.sizeof(_abi.SdfDesc), C.sizeof(_abi.Setting)]


def test_no_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = _abi.load_library()
    h = C.c_void_p()
    assert lib.gpmp2b_create(0, C.byref(h)) == _abi.ERR_NO_DEVICE
    with pytest.raises(RuntimeError):
        G.Context(0)


def test_missing_library_fails_loudly(tmp_path):
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _abi.load_library(str(tmp_path / "nope.so"))


def test_values_roundtrip_and_straight_line():
    # initArmTrajStraightLine, gpmp2/planner/TrajUtils.cpp:25-50 (avg_vel = (end-init)/total_step)
    s, e = np.array([0.0, 1.0, -1.0]), np.array([1.0, 3.0, 0.0])
    vals = G.initArmTrajStraightLine(s, e, 4)
    assert vals.size() == 10
    assert np.allclose(vals.atVector(G.symbol('x', 2)), 0.5 * (s + e))
    assert np.allclose(vals.atVector(G.symbol('v', 3)), (e - s) / 4.0)
    t = G.api.values_to_traj(vals, 4, 3)
    assert np.allclose(t, G.straight_line_traj(s[None], e[None], 4)[0])
    back = G.api.traj_to_values(t, 4, 3)
    for k in vals:
        assert np.allclose(vals[k], back[k])
    with pytest.raises(KeyError):
        vals.atVector(G.symbol('x', 99))


def test_setting_defaults_match_reference():
    # gpmp2/planner/TrajOptimizerSetting.cpp:44-68
    s = G.TrajOptimizerSetting(7)
    assert (s.total_step, s.total_time, s.obs_check_inter, s.max_iter) == (10, 1.0, 5, 50)
    assert (s.epsilon, s.cost_sigma, s.rel_thresh) == (0.2, 0.1, 1e-2)
    assert s.opt_type == G.TrajOptimizerSetting.Dogleg and not s.flag_pos_limit and not s.flag_vel_limit
    assert s.conf_prior_sigma == 1e-4 and s.vel_prior_sigma == 1e-4
    p, keep = s.pack()
    assert p.dof == 7 and p.total_step == 10 and p.opt_type == _abi.OPT_DOGLEG
    s.set_vel_limits([1, 2, 3])
    with pytest.raises(RuntimeError, match="dim does not fit"):
        s.pack()


def test_sdf_wire_layout():
    # data[z][col][row] == the reference's per-slice column-major Matrix
    data = np.arange(2 * 3 * 4, dtype=float).reshape(2, 3, 4)  # (nz, rows, cols)
    sdf = G.SignedDistanceField([0, 0, 0], 0.1, data)
    assert (sdf.desc.rows, sdf.desc.cols, sdf.desc.nz) == (3, 4, 2)
    flat = np.ctypeslib.as_array(sdf.desc.data, shape=(24,))
    z, r, c = 1, 2, 3
    assert flat[(z * 4 + c) * 3 + r] == data[z, r, c]
    sdf2 = G.SignedDistanceField([0, 0, 0], 0.1, 3, 4, 2)
    sdf2.initFieldData(1, data[1])
    assert np.ctypeslib.as_array(sdf2.desc.data, shape=(24,))[(1 * 4 + c) * 3 + r] == data[1, r, c]
    with pytest.raises(RuntimeError, match="out of index"):
        sdf2.initFieldData(2, data[1])


def test_pose2vector_values_helpers():
    """insertPose2VectorInValues / atPose2VectorValues (gpmp2/utils/matlabUtils.cpp:14-22), as the toolboxes call them."""
    import gpmp2_b200 as G
    v = G.Values()
    p = G.Pose2Vector(G.Pose2(1.0, -2.0, 0.5), [0.25, 0.75])
    G.insertPose2VectorInValues(G.symbol("x", 3), p, v)
    q = G.atPose2VectorValues(G.symbol("x", 3), v)
    assert (q.pose().x(), q.pose().y(), q.pose().theta()) == (1.0, -2.0, 0.5) and list(q.configuration()) == [0.25, 0.75]
    v.insert(G.symbol("x", 4), p.flat())                       # a flat wire vector under the key reads back as a Pose2Vector
    assert np.array_equal(G.atPose2VectorValues(G.symbol("x", 4), v).flat(), p.flat())
    with pytest.raises(KeyError, match="ValuesKeyAlreadyExists"):
        G.insertPose2VectorInValues(G.symbol("x", 3), p, v)
    with pytest.raises(KeyError, match="ValuesKeyDoesNotExist"):
        G.atPose2VectorValues(G.symbol("x", 9), v)
    with pytest.raises(TypeError):
        G.insertPose2VectorInValues(G.symbol("x", 5), [1.0, 2.0, 3.0], v)


def test_sdf_accessors_follow_the_reference_indexing():
    """signed_distance(r, c, z) = data_[z](r, c); origin(), x_count = columns, y_count = rows (SignedDistanceField.h:170-179)."""
    import gpmp2_b200 as G
    d = np.arange(24.0).reshape(2, 3, 4)          # (nz, rows, cols)
    s = G.SignedDistanceField([1.0, 2.0, 3.0], 0.1, d)
    assert (s.x_count(), s.y_count(), s.z_count(), s.cell_size()) == (4, 3, 2, 0.1) and list(s.origin()) == [1.0, 2.0, 3.0]
    assert all(s.signed_distance(r, c, z) == d[z, r, c] for z in range(2) for r in range(3) for c in range(4))
    s.initFieldData(1, -d[1])
    assert s.signed_distance(2, 3, 1) == -d[1, 2, 3]
    p = G.PlanarSDF([1.0, 2.0], 0.1, d[0])                     # PlanarSDF.h:119-126
    assert (p.x_count(), p.y_count()) == (4, 3) and list(p.origin()) == [1.0, 2.0]
    assert all(p.signed_distance(r, c) == d[0, r, c] for r in range(3) for c in range(4))
