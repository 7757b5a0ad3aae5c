"""Regenerates tests/golden/sdf_boost178_2x3x2.{txt,bin}: SignedDistanceField archives whose framing was written by the
real Boost 1.78 serialization runtime found in this image (see boost_sdf_probe.cpp for what is and is not the runtime's).
Run from the repo root after `make -C oracle`:  python oracle/boost_probe/make_fixtures.py"""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
PROBE = os.path.join(ROOT, "oracle", "_ref", "boost_sdf_probe")
SHAPE = ("2", "3", "2")     # rows, cols, nz


def generate(mode):
    return subprocess.run([PROBE, mode, *SHAPE], check=True, capture_output=True).stdout


if __name__ == "__main__":
    for mode, ext in (("text", "txt"), ("bin", "bin")):
        with open(os.path.join(ROOT, "tests", "golden", "sdf_boost178_2x3x2." + ext), "wb") as f:
            f.write(generate(mode))
