"""tests/golden/crc_golden.npz = the reference's CRC known-answer vectors (/root/reference/test/codes/crc/crc_u_<POL>.npy,
crc_x_ref_np_<POL>.npy; used by test/unit/fec/test_crc.py:177-199) repacked into one small file. Needs /root/reference."""
import numpy as np
gold = {}
for deg in ("CRC24A", "CRC24B", "CRC24C", "CRC16", "CRC11", "CRC6"):
    gold[f"u_{deg}"] = np.load(f"/root/reference/test/codes/crc/crc_u_{deg}.npy").astype(np.uint8)
    gold[f"x_{deg}"] = np.load(f"/root/reference/test/codes/crc/crc_x_ref_np_{deg}.npy").astype(np.uint8)
np.savez_compressed(__file__.replace("make_crc_golden.py", "crc_golden.npz"), **gold)
