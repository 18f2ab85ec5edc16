"""Generates tests/golden/stream_management_golden.json from the reference's StreamManagement
(/root/reference/src/sionna/phy/mimo/stream_management.py, pure NumPy; its only sionna import, `Object`, is stubbed).
Run in the build container:  python tests/golden/make_stream_management_golden.py"""
import json
import os
import sys
import types
import numpy as np

stub = types.ModuleType("sionna.phy.block")
stub.Object = object
for name in ("sionna", "sionna.phy"):
    sys.modules.setdefault(name, types.ModuleType(name))
sys.modules["sionna.phy.block"] = stub
src = open("/root/reference/src/sionna/phy/mimo/stream_management.py").read()
mod = types.ModuleType("ref_sm")
exec(compile(src, "stream_management.py", "exec"), mod.__dict__)

CASES = [([[1]], 1), ([[1]], 4), ([[1, 0], [0, 1]], 2), ([[1, 1, 0, 0], [0, 0, 1, 1]], 1),
         ([[1, 0, 1, 0], [0, 1, 0, 1]], 3), ([[1, 1, 1, 1]], 2), ([[1], [1]], 4), ([[0, 1], [1, 0]], 1),
         ([[1, 0, 0], [0, 0, 1], [0, 1, 0]], 2), ([[1], [1], [1], [1]], 8)]
FIELDS = ["num_rx", "num_tx", "num_streams_per_tx", "num_streams_per_rx", "num_interfering_streams_per_rx",
          "num_tx_per_rx", "num_rx_per_tx", "precoding_ind", "stream_association", "detection_desired_ind",
          "detection_undesired_ind", "tx_stream_ids", "rx_stream_ids", "stream_ind"]
out = []
for a, s in CASES:
    sm = mod.StreamManagement(np.array(a), s)
    out.append({"rx_tx_association": a, "num_streams_per_tx": s,
                **{f: np.asarray(getattr(sm, f)).tolist() for f in FIELDS}})
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "stream_management_golden.json")
json.dump(out, open(path, "w"))
print(len(out), "cases ->", path)
