"""Builds tests/golden/pusch_golden.npz from the reference's PUSCH test vectors (run where /root/reference exists).

Sources (all produced by an independent 5G toolbox and shipped with the reference's unit tests):
  * test/unit/nr/pusch_test_configs/test_{0..82}.{json,npy}: PUSCH configuration, payload bits b and the complete
    transmit resource grid (used by test_pusch_transmitter.py:tests_against_reference).
  * test/unit/nr/reference_dmrs_{1,2}.npy: DMRS sequences (test_pusch_config.py:test_against_reference_{1,2}).
  * test/unit/nr/pusch_dmrs_precoded_{L}_layer_{P}_ports.npy: precoded DMRS grids for every TPMI of the six codebooks
    (test_pusch_config.py:test_precoding_against_reference); stored as complex64 arrays [num_tpmi, ports, 12, 14].

The 83 grids are 80 MB of complex128. Stored here per case: the configuration, the payload (bit-packed), NPROJ seeded
random linear functionals of the grid (a 64-number fingerprint that any wrong resource element changes), and for the
cases in FULL the grid itself as complex64.
The .npy files pickle tf tensors; a stub module turns them back into ndarrays (TensorFlow is not installed).
"""
import json
import os
import sys
import types
import numpy as np

REF = "/root/reference/test/unit/nr"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "pusch_golden.npz")
NPROJ = 32
FULL = [0, 4, 6, 9, 10, 19, 20, 25, 26, 35, 36, 41]


def projection_vectors(n, seed):
    rng = np.random.default_rng(seed)
    return rng.standard_normal((NPROJ, n)) + 1j * rng.standard_normal((NPROJ, n))


def main():
    for name in ["tensorflow", "tensorflow.python", "tensorflow.python.framework", "tensorflow.python.framework.ops"]:
        sys.modules[name] = types.ModuleType(name)
    sys.modules["tensorflow.python.framework.ops"].convert_to_tensor = lambda x, *a, **k: np.asarray(x)
    out = {"nproj": np.int32(NPROJ), "full_cases": np.array(FULL, np.int32)}
    cfgs = []
    for i in range(83):
        b, g = np.load(f"{REF}/pusch_test_configs/test_{i}.npy", allow_pickle=True)
        with open(f"{REF}/pusch_test_configs/test_{i}.json") as f:
            cfgs.append(json.load(f))
        b = np.asarray(b).reshape(-1).astype(np.uint8)
        g = np.asarray(g, complex)
        if g.ndim == 2:
            g = g[..., None]                                  # [subcarriers, symbols, antenna ports]
        out[f"b_{i}"] = np.packbits(b)
        out[f"nb_{i}"] = np.int32(b.size)
        out[f"shape_{i}"] = np.array(g.shape, np.int32)
        out[f"proj_{i}"] = projection_vectors(g.size, 1000 + i) @ g.reshape(-1)
        if i in FULL:
            out[f"grid_{i}"] = g.astype(np.complex64)
    out["configs_json"] = np.frombuffer(json.dumps(cfgs).encode(), np.uint8)
    for k in (1, 2):
        out[f"reference_dmrs_{k}"] = np.load(f"{REF}/reference_dmrs_{k}.npy").astype(np.complex64)
    for layers, ports in ((1, 2), (1, 4), (2, 2), (2, 4), (3, 4), (4, 4)):
        a = np.load(f"{REF}/pusch_dmrs_precoded_{layers}_layer_{ports}_ports.npy", allow_pickle=True)
        out[f"dmrs_precoded_{layers}_{ports}"] = np.stack([np.asarray(v) for v in a]).astype(np.complex64)
    np.savez_compressed(OUT, **out)
    print(OUT, os.path.getsize(OUT) / 1e6, "MB")


if __name__ == "__main__":
    main()
