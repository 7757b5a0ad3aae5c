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


def _build_archive_tool():
    lib = os.path.join(ROOT, "gpmp2_b200", "csrc")
    src = os.path.join(ROOT, "tests", "cpp", "test_sdf_archive.cpp")
    exe = os.path.join(ROOT, "tests", "cpp", "test_sdf_archive")
    deps = [src, os.path.join(ROOT, "include", "gpmp2b", "gpmp2.hpp"), os.path.join(ROOT, "include", "gpmp2b.h")]
    if not os.path.exists(exe) or os.path.getmtime(exe) < max(os.path.getmtime(d) for d in deps):
        subprocess.check_call(["g++", "-std=c++14", "-O1", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), src, "-o", exe,
                               "-L", lib, "-lgpmp2b", "-Wl,-rpath," + lib])
    return exe


@pytest.mark.parametrize("ext", ["txt", "bin"])
def test_facade_sdf_archives_agree_with_python(tmp_path, ext):
    """saveSDF / loadSDF of the C++ facade and of the Python facade (gpmp2/obstacle/SignedDistanceField.cpp:14-50) write
    the same bytes and read each other's files; host-side only."""
    import gpmp2_b200 as G
    exe = _build_archive_tool()
    rng = np.random.default_rng(5)
    src = G.SignedDistanceField([0.1, -2.5, 1.0 / 3.0], 0.025, rng.standard_normal((4, 5, 3)) * 10.0 ** rng.integers(-8, 8, (4, 5, 3)))
    a, b = str(tmp_path / ("py." + ext)), str(tmp_path / ("cpp." + ext))
    src.saveSDF(a)
    p = subprocess.run([exe, "rewrite", a, b], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    head = p.stdout.split()
    assert [int(v) for v in head[:3]] == [5, 3, 4] and [float(v) for v in head[3:]] == [0.025] + src._origin
    assert open(a, "rb").read() == open(b, "rb").read()             # byte for byte
    # a field built through the reference's constructor + initFieldData in C++, read in Python
    c = str(tmp_path / ("made." + ext))
    assert subprocess.run([exe, "make", c]).returncode == 0
    dst = G.SignedDistanceField([0, 0, 0], 1.0, 1, 1, 1)
    dst.loadSDF(c)
    z, r, col = np.meshgrid(np.arange(4), np.arange(3), np.arange(2), indexing="ij")
    ref = G.SignedDistanceField([-0.2, 0.4, 1.0 / 3.0], 0.01, (100.0 * z + 10.0 * r + col) / 7.0)
    assert np.array_equal(dst._wire, ref._wire) and dst._origin == ref._origin and dst.cell_size() == 0.01
    # the other text cross-conversion: binary read in C++, text written, read in Python
    if ext == "bin":
        t = str(tmp_path / "conv.txt")
        assert subprocess.run([exe, "rewrite", a, t], capture_output=True).returncode == 0
        dst.loadSDF(t)
        assert np.array_equal(dst._wire, src._wire)


def test_facade_sdf_archive_errors(tmp_path):
    exe = _build_archive_tool()
    p = subprocess.run([exe, "rewrite", str(tmp_path / "missing.txt"), str(tmp_path / "o.txt")], capture_output=True, text=True)
    assert p.returncode == 1 and "does not exist" in p.stderr
    open(str(tmp_path / "junk.bin"), "wb").write(b"\x00" * 64)
    p = subprocess.run([exe, "rewrite", str(tmp_path / "junk.bin"), str(tmp_path / "o.txt")], capture_output=True, text=True)
    assert p.returncode == 1 and "signature" in p.stderr
    p = subprocess.run([exe, "make", str(tmp_path / "o.xml")], capture_output=True, text=True)
    assert p.returncode == 1 and "*this" in p.stderr


@pytest.mark.parametrize("ext", ["txt", "bin"])
def test_facade_reads_and_rewrites_the_boost_runtime_fixture(tmp_path, ext):
    """tests/golden/sdf_boost178_2x3x2.*: framing written by a real Boost 1.78 runtime (oracle/boost_probe).  The C++
    facade loads it and writes it back identically except for the library-version stamp (17, so that older Boosts read it)."""
    exe = _build_archive_tool()
    gold = os.path.join(ROOT, "tests", "golden", "sdf_boost178_2x3x2." + ext)
    out = str(tmp_path / ("o." + ext))
    p = subprocess.run([exe, "rewrite", gold, out], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    assert p.stdout.split() == ["2", "3", "2", "0.5", "-1.5", "0.25", "2"]
    a, b = open(gold, "rb").read(), open(out, "rb").read()
    if ext == "txt":
        assert a.replace(b"serialization::archive 19 ", b"serialization::archive 17 ", 1) == b
    else:
        assert a[:30] + bytes([17, 0]) + a[32:] == b


def test_facade_pose2vector_values_helpers():
    exe = _build_archive_tool()
    p = subprocess.run([exe, "values"], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    assert p.stdout.splitlines() == ["1 -2 0.5 2 0.25 0.75", "ValuesKeyDoesNotExist", "ValuesKeyAlreadyExists"]


def test_facade_sdf_archive_rejects_implausible_sizes(tmp_path):
    """A header that announces 2^31 - 1 rows, columns and layers in a 200-byte file is refused before anything is allocated
    or copied (both facades)."""
    import struct
    import gpmp2_b200 as G
    from gpmp2_b200 import boost_archive as BA
    exe = _build_archive_tool()
    cls, big = struct.pack("<BI", 0, 0), 2 ** 31 - 1
    raw = (BA._BIN_HEAD + struct.pack("<H", 17) + BA._BIN_SIZES + cls * 3 + struct.pack("<QQ3d", 3, 1, 0, 0, 0)
           + struct.pack("<QQQd", big, big, big, 0.1) + cls + struct.pack("<QI", big, 0) + b"\0" * 64)
    fn = str(tmp_path / "evil.bin")
    open(fn, "wb").write(raw)
    p = subprocess.run([exe, "rewrite", fn, str(tmp_path / "o.bin")], capture_output=True, text=True)
    assert p.returncode == 1 and "any known layout" in p.stderr
    with pytest.raises(RuntimeError, match="any known layout"):
        G.SignedDistanceField([0, 0, 0], 1.0, 1, 1, 1).loadSDF(fn)
    txt = "22 serialization::archive 17 0 0 0 0 0 0 3 1 0 0 0 %d %d %d 0.1 0 0 %d 0 0 0 %d %d 1 2 3" % (big, big, big, big, big, big)
    fn = str(tmp_path / "evil.txt")
    open(fn, "w").write(txt)
    p = subprocess.run([exe, "rewrite", fn, str(tmp_path / "o.txt")], capture_output=True, text=True)
    assert p.returncode == 1 and "any known layout" in p.stderr
    with pytest.raises(RuntimeError, match="any known layout"):
        G.SignedDistanceField([0, 0, 0], 1.0, 1, 1, 1).loadSDF(fn)
