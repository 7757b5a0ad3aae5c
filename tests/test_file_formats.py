"""SDF file formats either side of the path: the `.vol` text volume of gpmp2::readSDFvolfile
(gpmp2/utils/fileUtils.cpp:17-62).  CPU only (host-side parsing; the field is uploaded like any other SDF)."""
import numpy as np

import gpmp2_b200 as G


def test_read_vol_index_order(tmp_path):
    """The reference loops x (columns) outermost, then y (rows), z innermost, and stores vmat[z](y, x)."""
    pre = str(tmp_path / "tiny")
    cols, rows, nz = 2, 3, 4
    with open(pre + ".vol.head", "w") as f:
        f.write("%d %d %d\n-1.5 0.25 2\n0.05\n" % (cols, rows, nz))
    with open(pre + ".vol.data", "w") as f:
        for x in range(cols):
            for y in range(rows):
                f.write(" ".join(str(100 * x + 10 * y + z) for z in range(nz)) + "\n")
    sdf = G.readSDFvolfile(pre)
    assert (sdf.x_count(), sdf.y_count(), sdf.z_count()) == (cols, rows, nz)
    assert sdf.cell_size() == 0.05 and np.allclose(sdf._origin, [-1.5, 0.25, 2.0])
    for z in range(nz):
        for y in range(rows):
            for x in range(cols):
                assert sdf._wire[z, x, y] == 100 * x + 10 * y + z      # wire layout [z][col][row]
    assert G.readSDFvolfile(str(tmp_path / "missing")) is None         # the reference returns false


def test_vol_round_trip(tmp_path):
    rng = np.random.default_rng(0)
    data = rng.standard_normal((5, 7, 6))
    pre = str(tmp_path / "field")
    G.writeSDFvolfile(pre, [0.1, -0.2, 0.3], 0.02, data)
    sdf = G.readSDFvolfile(pre)
    ref = G.SignedDistanceField([0.1, -0.2, 0.3], 0.02, data)
    assert np.array_equal(sdf._wire, ref._wire) and sdf._origin == ref._origin and sdf.cell_size() == ref.cell_size()


# ------------------------------------------------------------------------------------------------
# Boost archives of SignedDistanceField::saveSDF / loadSDF (gpmp2/obstacle/SignedDistanceField.cpp:14-50).
# The literals below are written by hand from the text-archive grammar (the three gtsam::Point3 layouts); the framing itself
# is pinned to a real Boost runtime further down (tests/golden/sdf_boost178_*); see gpmp2_b200/boost_archive.py.
# ------------------------------------------------------------------------------------------------
import struct

import pytest

from gpmp2_b200 import boost_archive as BA


def _field(seed=0, shape=(3, 4, 2)):
    rng = np.random.default_rng(seed)
    return G.SignedDistanceField([0.1, -0.2, 0.3], 0.02, rng.standard_normal(shape))


@pytest.mark.parametrize("ext", ["txt", "bin", "sdf", "noext"])
def test_boost_archive_round_trip_is_bit_exact(tmp_path, ext):
    """17 significant digits in the text archive (Boost's own precision) bring every double back exactly."""
    src = _field()
    fn = str(tmp_path / ("field." + ext if ext != "noext" else "field"))
    src.saveSDF(fn)
    dst = G.SignedDistanceField([0, 0, 0], 1.0, 1, 1, 1)
    dst.loadSDF(fn)
    assert (dst.y_count(), dst.x_count(), dst.z_count()) == (4, 2, 3)
    assert np.array_equal(dst._wire, src._wire) and dst._origin == src._origin and dst.cell_size() == src.cell_size()
    assert dst.desc.rows == 4 and dst.desc.cols == 2 and dst.desc.nz == 3      # the descriptor the library uploads follows


# one 2 x 3 layer pair, column-major per layer (Eigen): layer0 = [[1, 3, 5], [2, 4, 6]], layer1 = 10 x that
_TAIL = "2 3 2 5.00000000000000000e-01 0 0 2 0 0 0 2 3 1 2 3 4 5 6 2 3 10 20 30 40 50 60"
_LITERALS = {
    "A (GTSAM 4.0: class Point3 : Vector3)": "22 serialization::archive 15 0 0 0 0 0 0 3 1 -1.5 0.25 2 " + _TAIL,
    "B (Point3 = Vector3)": "22 serialization::archive 17 0 0 0 0 3 1 -1.5 0.25 2 " + _TAIL,
    "C (GTSAM 3: x_, y_, z_)": "22 serialization::archive 12 0 0 0 0 -1.5 0.25 2 " + _TAIL,
}


@pytest.mark.parametrize("name", sorted(_LITERALS))
def test_boost_text_archive_literal(tmp_path, name):
    fn = str(tmp_path / "lit.txt")
    with open(fn, "w") as f:
        f.write(_LITERALS[name].replace(" 0 0 2 0 0 0 ", " 0 0 2 0\n0 0 ") + "\n")     # Boost breaks lines at class info
    sdf = G.SignedDistanceField([0, 0, 0], 1.0, 1, 1, 1)
    sdf.loadSDF(fn)
    assert sdf._origin == [-1.5, 0.25, 2.0] and sdf.cell_size() == 0.5
    assert (sdf.y_count(), sdf.x_count(), sdf.z_count()) == (2, 3, 2)
    layer0 = np.array([[1.0, 3.0, 5.0], [2.0, 4.0, 6.0]])
    ref = G.SignedDistanceField([-1.5, 0.25, 2.0], 0.5, np.stack([layer0, 10 * layer0]))
    assert np.array_equal(sdf._wire, ref._wire)


def test_boost_text_archive_as_written(tmp_path):
    """Token stream of the writer: header, class infos, members in SignedDistanceField.h:201-208 order."""
    layer0 = np.array([[1.0, 3.0, 5.0], [2.0, 4.0, 6.0]])
    G.SignedDistanceField([-1.5, 0.25, 2.0], 0.5, np.stack([layer0, 10 * layer0])).saveSDF(str(tmp_path / "w.txt"))
    tok = open(str(tmp_path / "w.txt")).read().split()
    exp = _LITERALS["A (GTSAM 4.0: class Point3 : Vector3)"].split()
    assert tok[:2] == exp[:2] and int(tok[2]) == BA.LIBRARY_VERSION
    assert [float(a) for a in tok[3:]] == [float(a) for a in exp[3:]]
    assert tok[11] == "-1.50000000000000000e+00"          # scientific, 17 digits: basic_text_oprimitive<double>


def test_boost_binary_archive_layout(tmp_path):
    """Header = length-prefixed signature, uint16 library version, sizeof(int, long, float, double), int 1; coefficient
    blocks raw (array optimisation of binary archives)."""
    src = _field(1, (2, 3, 5))
    fn = str(tmp_path / "f.bin")
    src.saveSDF(fn)
    raw = open(fn, "rb").read()
    assert raw[:8] == struct.pack("<Q", 22) and raw[8:30] == b"serialization::archive"
    assert struct.unpack_from("<H", raw, 30)[0] == BA.LIBRARY_VERSION and raw[32:36] == bytes([4, 8, 4, 8])
    assert len(raw) == 40 + 5 + (5 + 5 + 16 + 24) + 24 + 8 + (5 + 8 + 4) + 5 + 2 * (16 + 8 * 15)
    assert raw[-8 * 15:] == src._wire[1].tobytes()
    # layouts B and C of the same field are accepted too (drop one class info / the Vector3 shape)
    body = raw[40:]
    for cut in (body[:5] + body[10:], body[:5] + body[10:15] + body[31:]):
        open(fn, "wb").write(raw[:40] + cut)
        dst = G.SignedDistanceField([0, 0, 0], 1.0, 1, 1, 1)
        dst.loadSDF(fn)
        assert np.array_equal(dst._wire, src._wire) and dst._origin == src._origin


def test_boost_archive_errors(tmp_path):
    sdf = _field()
    with pytest.raises(RuntimeError, match="does not exist"):
        sdf.loadSDF(str(tmp_path / "missing.txt"))
    with pytest.raises(RuntimeError, match=r"\*this"):
        sdf.saveSDF(str(tmp_path / "f.xml"))
    open(str(tmp_path / "junk.txt"), "w").write("1 2 3\n")
    with pytest.raises(RuntimeError, match="signature"):
        sdf.loadSDF(str(tmp_path / "junk.txt"))
    fn = str(tmp_path / "cut.txt")
    sdf.saveSDF(fn)
    txt = open(fn).read().split()
    open(fn, "w").write(" ".join(txt[:-1]))                 # one coefficient short
    with pytest.raises(RuntimeError, match="any known layout"):
        sdf.loadSDF(fn)
    fnb = str(tmp_path / "cut.bin")
    sdf.saveSDF(fnb)
    raw = open(fnb, "rb").read()
    open(fnb, "wb").write(raw[:-3])
    with pytest.raises(RuntimeError, match="any known layout"):
        sdf.loadSDF(fnb)
    open(fnb, "wb").write(raw[:32] + bytes([4, 4, 4, 8]) + raw[36:])      # written where long is 4 bytes
    with pytest.raises(RuntimeError, match="not portable"):
        sdf.loadSDF(fnb)
    # a failed load leaves the field as it was
    assert sdf.z_count() == 3


# ------------------------------------------------------------------------------------------------
# The archive framing pinned to a real Boost runtime: tests/golden/sdf_boost178_2x3x2.{txt,bin} were written by
# oracle/boost_probe/boost_sdf_probe.cpp linked against Boost 1.78's libboost_serialization (header, class-info emission,
# delimiters, item_version and the binary header / class-info byte widths are the runtime's own; the member lists of
# SignedDistanceField / gtsam::Point3 / Matrix are restated -- see that file's header).
# ------------------------------------------------------------------------------------------------
import os
import subprocess

_GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
_PROBE = os.path.join(os.path.dirname(_GOLD), os.pardir, "oracle", "_ref", "boost_sdf_probe")


def _probe_field(rows=2, cols=3, nz=2):
    z, r, c = np.meshgrid(np.arange(nz), np.arange(rows), np.arange(cols), indexing="ij")
    return G.SignedDistanceField([-1.5, 0.25, 2.0], 0.5, (100.0 * z + 10.0 * r + c) / 7.0)


@pytest.mark.parametrize("ext", ["txt", "bin"])
def test_boost_runtime_fixture_reads_and_is_reproduced_byte_for_byte(tmp_path, ext, monkeypatch):
    gold = open(os.path.join(_GOLD, "sdf_boost178_2x3x2." + ext), "rb").read()
    ref = _probe_field()
    dst = G.SignedDistanceField([0, 0, 0], 1.0, 1, 1, 1)
    dst.loadSDF(os.path.join(_GOLD, "sdf_boost178_2x3x2." + ext))
    assert np.array_equal(dst._wire, ref._wire) and dst._origin == ref._origin and dst.cell_size() == 0.5
    # our writer, stamped with the runtime's library version (19 = Boost 1.76+), writes the same bytes
    monkeypatch.setattr(BA, "LIBRARY_VERSION", 19)
    fn = str(tmp_path / ("w." + ext))
    ref.saveSDF(fn)
    mine = open(fn, "rb").read()
    assert mine == gold            # including the newline the text archive's destructor appends


def _run_probe(mode, rows, cols, nz):
    if not os.path.exists(_PROBE):
        pytest.skip("no Boost serialization runtime in this image (oracle/_ref/boost_sdf_probe not built)")
    p = subprocess.run([_PROBE, mode, str(rows), str(cols), str(nz)], capture_output=True)
    if p.returncode != 0:
        pytest.skip("boost_sdf_probe cannot run here (its Boost runtime is missing)")
    return p.stdout


@pytest.mark.parametrize("mode,ext", [("text", "txt"), ("bin", "bin")])
def test_boost_runtime_live(tmp_path, mode, ext, monkeypatch):
    """Where the Boost runtime exists (this image): the committed fixtures are what it writes, and a larger field written
    by it equals our writer's output and loads back exactly."""
    assert _run_probe(mode, 2, 3, 2) == open(os.path.join(_GOLD, "sdf_boost178_2x3x2." + ext), "rb").read()
    live = _run_probe(mode, 7, 5, 4)
    ref = _probe_field(7, 5, 4)
    monkeypatch.setattr(BA, "LIBRARY_VERSION", 19)
    fn = str(tmp_path / ("w." + ext))
    ref.saveSDF(fn)
    mine = open(fn, "rb").read()
    assert mine == live
    open(fn, "wb").write(live)
    dst = G.SignedDistanceField([0, 0, 0], 1.0, 1, 1, 1)
    dst.loadSDF(fn)
    assert np.array_equal(dst._wire, ref._wire)


def test_boost_runtime_rejects_the_reference_xml_tag():
    """SignedDistanceField.cpp:19-20 saves BOOST_SERIALIZATION_NVP(*this): the real runtime's xml archive throws on that
    tag name, so the reference has no XML form of a field -- which is why `.xml` raises in the facades."""
    if not os.path.exists(_PROBE):
        pytest.skip("no Boost serialization runtime in this image")
    run = lambda name: subprocess.run([_PROBE, "xmlname", name], capture_output=True, text=True)
    if run("origin_").returncode != 0:
        pytest.skip("boost_sdf_probe cannot run here")
    assert run("origin_").stdout.strip() == "accepted"
    assert run("*this").stdout.strip() == "threw: Invalid XML tag name"


@pytest.mark.parametrize("mode,ext", [("text", "txt"), ("bin", "bin")])
def test_boost_runtime_reads_what_save_sdf_writes(tmp_path, mode, ext):
    """The other direction: the real Boost 1.78 reader (basic_iarchive::load_object through the probe) accepts a file written
    by saveSDF with its library-version stamp 17 and finds the same field, bit for bit."""
    if not os.path.exists(_PROBE):
        pytest.skip("no Boost serialization runtime in this image")
    rng = np.random.default_rng(3)
    src = G.SignedDistanceField([0.1, -2.5, 1.0 / 3.0], 0.025, rng.standard_normal((3, 4, 6)) * 10.0 ** rng.integers(-6, 6, (3, 4, 6)))
    fn = str(tmp_path / ("f." + ext))
    src.saveSDF(fn)
    p = subprocess.run([_PROBE, "read", mode, fn], capture_output=True, text=True)
    if p.returncode not in (0, 1):
        pytest.skip("boost_sdf_probe cannot run here")
    assert p.returncode == 0, p.stdout
    lines = p.stdout.strip().splitlines()
    head = [float(v) for v in lines[0].split()]
    assert head == [4, 6, 3, 0.025] + src._origin                     # rows, cols, nz, cell, origin
    got = np.array([[float(v) for v in ln.split()] for ln in lines[1:]])
    assert np.array_equal(got, src._wire.reshape(3, -1))              # column-major layers = the wire layout
    # and it rejects a damaged file (cut inside the last layer) instead of returning garbage
    raw = open(fn, "rb").read()
    open(fn, "wb").write(raw[:-12] if ext == "bin" else b" ".join(raw.split()[:-2]))
    p = subprocess.run([_PROBE, "read", mode, fn], capture_output=True, text=True)
    assert p.returncode == 1 and p.stdout.startswith("threw: ")
