"""Boost.Serialization archives of gpmp2::SignedDistanceField -- SignedDistanceField::saveSDF / loadSDF
(gpmp2/obstacle/SignedDistanceField.cpp:14-50; members in the order of SignedDistanceField.h:201-208: origin_,
field_rows_, field_cols_, field_z_, cell_size_, data_).  Host-side data format on the input side of the path
(SURVEY.md 8f-4); the field is uploaded like any other SDF afterwards.

PARITY: PARTLY PINNED.  No GTSAM / Boost headers exist in this image and the reference ships no saved field, so no
archive written by the reference itself could be committed.  The archive FRAMING, however, is pinned to a real Boost
runtime: Nsight Compute bundles libboost_serialization.so.1.78.0, and oracle/boost_probe/boost_sdf_probe.cpp drives it
(basic_oarchive::save_object, init, end_preamble, newtoken, the binary header and class-info bytes are the runtime's own)
to write tests/golden/sdf_boost178_2x3x2.{txt,bin}; this module reads them and reproduces them byte for byte
(tests/test_file_formats.py).  What stays restated from published sources: the member list of each class --
SignedDistanceField.h:201-208, GTSAM's Eigen serialization (gtsam/base/Matrix.h: `rows`, `cols`, then the coefficients
as an array, column-major) -- and Boost's header-inline wrapper widths (collection_size_type = size_t, item_version_type
= unsigned int).  Grammar: text_oarchive / binary_oarchive, library versions >= 7 (Boost >= 1.44), little-endian LP64.
What a class costs in an archive the first time it appears: its tracking level and its version (text: two integers;
binary: one byte + a uint32).  `gtsam::Point3`
changed between GTSAM releases, so the reader accepts the three layouts it has had and takes the one that accounts for
every token / byte of the file:
  A  class Point3 : public Vector3   (GTSAM 4.0, the branch gpmp2's README pins)   [cls][cls] 3 1 x y z
  B  typedef Vector3 Point3          (GTSAM >= 4.1, or 4.0 with GTSAM_TYPEDEF_POINTS_TO_VECTORS)   [cls] 3 1 x y z
  C  class Point3 { double x_, y_, z_; }   (GTSAM 3.x)                              [cls] x y z
The writer emits layout A.  The `.xml` branch of the reference wraps the object as BOOST_SERIALIZATION_NVP(*this),
whose tag name `*this` Boost's xml_oarchive rejects (xml_archive_exception "Invalid XML tag name" -- confirmed with the
real 1.78 runtime, test_boost_runtime_rejects_the_reference_xml_tag), so there is no
XML file of the reference's to be compatible with: `.xml` raises here too.
"""
import struct

import numpy as np

SIGNATURE = "serialization::archive"
LIBRARY_VERSION = 17          # what the writer stamps (Boost 1.71-1.73, so that older Boosts accept the file; the 1.78
                              # runtime stamps 19 -- the only difference from its output); the reader takes any version >= 7
_POINT_LAYOUTS = ("A", "B", "C")


class ArchiveError(RuntimeError):
    pass


def _ext(filename):
    return filename[filename.rfind(".") + 1:]     # SignedDistanceField.cpp:17 (find_last_of("."))


# ------------------------------------------------------------------------------------------------
# text_oarchive / text_iarchive
# ------------------------------------------------------------------------------------------------
def _fmt(v):
    # basic_text_oprimitive::save(double): scientific with max_digits10 = 17 digits of precision
    return "%.17e" % v


def _write_text(f, origin, rows, cols, nz, cell, wire):
    out = ["%d %s %d" % (len(SIGNATURE), SIGNATURE, LIBRARY_VERSION), "0 0",            # header, SignedDistanceField
           "0 0 0 0 3 1 " + " ".join(_fmt(v) for v in origin),                           # Point3 : Vector3
           "%d %d %d %s" % (rows, cols, nz, _fmt(cell)),
           "0 0 %d 0" % nz]                                                              # std::vector<Matrix>, count, item_version
    f.write(" ".join(out))
    for z in range(nz):
        f.write((" 0 0 %d %d " if z == 0 else " %d %d ") % (rows, cols))               # Matrix class info once
        f.write(" ".join(_fmt(v) for v in wire[z].ravel()))                              # column-major = [col][row]
    f.write("\n")


class _Tok:
    def __init__(self, tok):
        self.t, self.i = tok, 0

    def int(self):
        v = int(self.t[self.i])
        self.i += 1
        return v

    def flt(self, n):
        v = np.array(self.t[self.i:self.i + n], dtype=np.float64)
        if v.size != n:
            raise ValueError("short")
        self.i += n
        return v

    def cls(self):
        # tracking level, class version: both 0 for every class on this path
        if self.int() != 0 or self.int() != 0:
            raise ValueError("class info")


def _parse_text(tok, version, layout):
    p = _Tok(tok)
    p.cls()
    p.cls()
    if layout == "A":
        p.cls()
    if layout in ("A", "B"):
        if (p.int(), p.int()) != (3, 1):
            raise ValueError("Vector3 shape")
    origin = p.flt(3)
    rows, cols, nz = p.int(), p.int(), p.int()
    cell = float(p.flt(1)[0])
    p.cls()
    if p.int() != nz:
        raise ValueError("layer count")
    if version > 3:
        p.int()                                  # item_version
    if min(rows, cols, nz) < 0 or (len(tok) - p.i) != nz * (2 + rows * cols) + (2 if nz else 0):
        raise ValueError("length")
    wire = np.empty((nz, cols, rows))
    for z in range(nz):
        if z == 0:
            p.cls()
        if (p.int(), p.int()) != (rows, cols):
            raise ValueError("layer shape")
        wire[z] = p.flt(rows * cols).reshape(cols, rows)
    return origin, rows, cols, nz, cell, wire


def _read_text(raw):
    tok = raw.split()
    try:
        if int(tok[0]) != len(SIGNATURE) or tok[1].decode() != SIGNATURE:
            raise ValueError
        version = int(tok[2])
    except (ValueError, IndexError, UnicodeDecodeError):
        raise ArchiveError("[loadSDF] not a Boost text archive (signature missing)") from None
    for layout in _POINT_LAYOUTS:
        try:
            return _parse_text(tok[3:], version, layout)
        except (ValueError, IndexError):
            continue
    raise ArchiveError("[loadSDF] text archive (library version %d) does not hold a SignedDistanceField in any known layout" % version)


# ------------------------------------------------------------------------------------------------
# binary_oarchive / binary_iarchive (native little-endian LP64, as the reference writes on x86-64 / aarch64 Linux)
# ------------------------------------------------------------------------------------------------
_BIN_HEAD = struct.pack("<Q", len(SIGNATURE)) + SIGNATURE.encode()
_BIN_SIZES = bytes([4, 8, 4, 8]) + struct.pack("<i", 1)       # sizeof int, long, float, double; endian probe
_CLS = struct.pack("<BI", 0, 0)                                 # tracking_type (bool), version_type (uint32)


def _write_binary(f, origin, rows, cols, nz, cell, wire):
    f.write(_BIN_HEAD + struct.pack("<H", LIBRARY_VERSION) + _BIN_SIZES)
    f.write(_CLS + _CLS + _CLS + struct.pack("<QQ3d", 3, 1, *origin))
    f.write(struct.pack("<QQQd", rows, cols, nz, cell))
    f.write(_CLS + struct.pack("<QI", nz, 0))
    for z in range(nz):
        f.write((_CLS if z == 0 else b"") + struct.pack("<QQ", rows, cols))
        f.write(np.ascontiguousarray(wire[z], dtype="<f8").tobytes())


def _parse_binary(buf, layout):
    off = [0]

    def take(fmt):
        v = struct.unpack_from(fmt, buf, off[0])
        off[0] += struct.calcsize(fmt)
        return v

    def cls():
        if take("<BI") != (0, 0):
            raise ValueError("class info")

    cls()
    cls()
    if layout == "A":
        cls()
    if layout in ("A", "B") and take("<QQ") != (3, 1):
        raise ValueError("Vector3 shape")
    origin = np.array(take("<3d"))
    rows, cols, nz, cell = take("<QQQd")
    cls()
    count, _item_version = take("<QI")
    if count != nz or len(buf) - off[0] != nz * (16 + 8 * rows * cols) + (5 if nz else 0):
        raise ValueError("length")
    wire = np.empty((nz, cols, rows))
    for z in range(nz):
        if z == 0:
            cls()
        if take("<QQ") != (rows, cols):
            raise ValueError("layer shape")
        wire[z] = np.frombuffer(buf, dtype="<f8", count=rows * cols, offset=off[0]).reshape(cols, rows)
        off[0] += 8 * rows * cols
    return origin, rows, cols, nz, cell, wire


def _read_binary(raw):
    n = len(_BIN_HEAD)
    if raw[:n] != _BIN_HEAD or len(raw) < n + 10:
        raise ArchiveError("[loadSDF] not a Boost binary archive (signature missing)")
    version, = struct.unpack_from("<H", raw, n)
    if version < 7:
        raise ArchiveError("[loadSDF] binary archive of Boost library version %d (< 7, before Boost 1.44) is not supported" % version)
    if raw[n + 2:n + 10] != _BIN_SIZES:
        raise ArchiveError("[loadSDF] binary archive written on a platform with other type sizes or byte order "
                           "(Boost binary archives are not portable)")
    body = raw[n + 10:]
    for layout in _POINT_LAYOUTS:
        try:
            return _parse_binary(body, layout)
        except (ValueError, struct.error):
            continue
    raise ArchiveError("[loadSDF] binary archive (library version %d) does not hold a SignedDistanceField in any known layout" % version)


# ------------------------------------------------------------------------------------------------
def _xml_error(what):
    return ArchiveError("[%s] .xml: the reference serializes BOOST_SERIALIZATION_NVP(*this); Boost's xml archive rejects the "
                        "tag name '*this' (xml_archive_exception), so the reference has no XML form of the field either" % what)


def save_sdf(filename, origin, rows, cols, nz, cell, wire):
    """wire: [z][col][row] (each layer column-major, as Eigen stores the reference's Matrix)."""
    ext = _ext(filename)
    if ext == "xml":
        raise _xml_error("saveSDF")
    with open(filename, "wb" if ext == "bin" else "w") as f:
        (_write_binary if ext == "bin" else _write_text)(f, [float(v) for v in origin], int(rows), int(cols), int(nz), float(cell), wire)


def load_sdf(filename):
    """-> origin (3,), rows, cols, nz, cell_size, wire [z][col][row]"""
    ext = _ext(filename)
    if ext == "xml":
        raise _xml_error("loadSDF")
    try:
        with open(filename, "rb") as f:
            raw = f.read()
    except OSError:
        # SignedDistanceField.cpp:36-37 prints this and then fails inside the archive constructor
        raise ArchiveError("File '%s' does not exist!" % filename) from None
    return (_read_binary if ext == "bin" else _read_text)(raw)
