"""Are the synthetic scans bit-identical on this host and on the host the pins were taken on? Per-scan sha256 of configs[3]'s first 1000
scans against profiles/inputs_cfg3_here.json (written where the pins were generated: python profiles/probe_inputs.py write)."""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from chad_tsdf_b200 import synth  # noqa: E402

w = synth.WORKLOADS["cfg3_urban_5km"].truncated(1000)
scans = bench.generate_scans(w)
digests = [hashlib.sha256(np.ascontiguousarray(p).tobytes()).hexdigest()[:16] for p, _ in scans]
path = os.path.join(ROOT, "profiles", "inputs_cfg3_here.json")
if len(sys.argv) > 1 and sys.argv[1] == "write":
    json.dump({"numpy": np.__version__, "digests": digests, "points": [int(len(p)) for p, _ in scans]}, open(path, "w"))
    print("written", len(digests))
else:
    ref = json.load(open(path))
    bad = [i for i, (a, b) in enumerate(zip(digests, ref["digests"])) if a != b]
    print("numpy", np.__version__, "vs", ref["numpy"], "; scans that differ:", len(bad), bad[:20])
    for i in bad[:3]:
        p = scans[i][0]
        print(" scan", i, "points here", len(p), "there", ref["points"][i])
    try:
        import numpy.core._multiarray_umath as mu
        print("cpu features:", getattr(mu, "__cpu_features__", None) and [k for k, v in mu.__cpu_features__.items() if v][-12:])
    except Exception as ex:  # noqa: BLE001
        print("no cpu feature table:", ex)
