"""CPU oracle -- TEST INFRASTRUCTURE ONLY.

Restates the reference's algorithms (NVlabs/sionna v1.2.1, /root/reference/src/sionna/phy) for the hot path in
plain C / NumPy so the CUDA kernels can be checked without TensorFlow. Only ``tests/``,
``__graft_entry__.smoke()``, ``bench.py``'s cpu_baseline / ``--impl reference`` leg and the checker script
``tools/ber_sweep.py`` (which decodes the same inputs beside the GPU) may import this package; nothing under
``sionna_b200/`` does. PARITY STATUS: pinned against the reference's own golden vectors and
known-answer tests (see tests/test_oracle_*.py); bit-level parity with TensorFlow's kernels is UNPINNED because
TensorFlow is not installable in the build container.
"""
