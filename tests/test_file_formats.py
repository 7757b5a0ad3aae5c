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
