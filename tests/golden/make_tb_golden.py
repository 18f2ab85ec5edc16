"""tests/golden/tb_golden.npz = the reference's 8 transport-block encoder test cases (/root/reference/test/unit/nr/tb_refs/
tb_testcase_*.npz: information bits u_ref, scrambled coded bits c_ref, unscrambled c_ref_no_scr and the parameters; used by
test/unit/nr/test_tb_encoder.py:17-63) with the bit arrays packed. Needs /root/reference."""
import numpy as np
out = {}
for i in range(8):
    d = np.load(f"/root/reference/test/unit/nr/tb_refs/tb_testcase_{i}.npz")
    out[f"u_{i}"] = np.packbits(d["u_ref"].astype(np.uint8), axis=1)
    out[f"c_{i}"] = np.packbits(d["c_ref"].astype(np.uint8), axis=1)
    out[f"cn_{i}"] = np.packbits(d["c_ref_no_scr"].astype(np.uint8), axis=1)
    out[f"p_{i}"] = np.array([d["u_ref"].shape[1], d["c_ref"].shape[1], int(d["n_id"]), int(d["n_rnti"]), int(d["num_bits_per_symbol"]),
                              int(d["num_layers"])], np.int64)
    out[f"r_{i}"] = np.array(float(d["coderate"]))
    print(i, out[f"p_{i}"], float(d["coderate"]), d["u_ref"].shape)
np.savez_compressed(__file__.replace("make_tb_golden.py", "tb_golden.npz"), **out)
