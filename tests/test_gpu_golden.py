"""GPU parity against the committed golden pins of the REFERENCE (tests/golden/golden.json), through the C ABI."""
import json
import os

import pytest

from tests.golden import cases

pytestmark = pytest.mark.gpu
GOLDEN = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "golden.json")))
GOLDEN_FULL = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "golden_full.json")))


@pytest.mark.parametrize("batch,path", [(1, 2), (8, 2), (8, 0), (8, 1)])
@pytest.mark.parametrize("name", cases.ALL_CASES)
def test_gpu_matches_reference_golden(chad_lib, name, batch, path):
    from chad_tsdf_b200 import TSDFMap
    m, d = cases.run_case(lambda r, t: TSDFMap(r, t, max_batch_scans=batch, pair_path=path), name)
    g = GOLDEN[name]
    for k, v in g["verbatim"]["before_finalize"].items():  # tier A: the reference exactly as written
        assert d["before_finalize"][k] == v, f"tier A {k}"
    assert d["before_finalize"] == g["stable"]["before_finalize"]  # tier B: bit-exact floats too
    assert d["final"]["roots"] == g["stable"]["final"]["roots"]
    for lv, (a, b) in enumerate(zip(d["final"]["levels"], g["stable"]["final"]["levels"])):
        assert a == b, f"DAG level {lv}"
    m.close()


@pytest.mark.parametrize("name", list(cases.FULL_CASES))
def test_gpu_matches_reference_golden_at_full_size(chad_lib, name):
    """BASELINE.json's configs at full size (configs[1] = the bench workload whole, with the library's defaults: the exact
    configuration bench.py times) against pins taken from the reference build: voxels of the last submap bit for bit, every
    word of all 21 DAG levels, counters and the roots of every submap."""
    from chad_tsdf_b200 import TSDFMap
    m, d = cases.run_case(lambda r, t: TSDFMap(r, t), name)
    g = GOLDEN_FULL[name]
    for k, v in g["verbatim"]["before_finalize"].items():
        assert d["before_finalize"][k] == v, f"tier A {k}"
    assert d["before_finalize"] == g["stable"]["before_finalize"]
    assert d["final"]["roots"] == g["stable"]["final"]["roots"]
    for lv, (a, b) in enumerate(zip(d["final"]["levels"], g["stable"]["final"]["levels"])):
        assert a == b, f"DAG level {lv}"
    st = m.stats()
    assert st["points"] == g["points"] and st["scans"] == g["scans"]
    m.close()
