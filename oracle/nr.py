"""Oracle: CRC and 5G scrambling sequences (NumPy). TEST INFRASTRUCTURE. Restates
/root/reference/src/sionna/phy/fec/crc.py:99-156, 175-215 (generator-matrix CRC == polynomial long division) and
/root/reference/src/sionna/phy/nr/utils.py:16-76 (38.211 5.2.1 Gold sequence), fec/scrambling.py:442-468.
Pinned by the reference's CRC known-answer vectors (tests/golden/crc_golden.npz)."""
import numpy as np

CRC_POLYS = {"CRC24A": (24, [24, 23, 18, 17, 14, 11, 10, 7, 6, 5, 4, 3, 1, 0]), "CRC24B": (24, [24, 23, 6, 5, 1, 0]),
             "CRC24C": (24, [24, 23, 21, 20, 17, 15, 13, 12, 8, 4, 2, 1, 0]), "CRC16": (16, [16, 12, 5, 0]),
             "CRC11": (11, [11, 10, 9, 5, 0]), "CRC6": (6, [6, 5, 0])}


def crc_parity(bits, degree):
    """Parity bits of one bit vector by polynomial long division over GF(2) (38.212 5.1)."""
    L, coeffs = CRC_POLYS[degree]
    pol = np.zeros(L + 1, np.uint8)
    pol[[L - c for c in coeffs]] = 1
    reg = np.concatenate([np.asarray(bits, np.uint8), np.zeros(L, np.uint8)])
    for i in range(len(bits)):
        if reg[i]:
            reg[i:i + L + 1] ^= pol
    return reg[-L:]


def crc_encode(bits, degree):
    b = np.asarray(bits).astype(np.uint8)
    flat = b.reshape(-1, b.shape[-1])
    out = np.stack([np.concatenate([r, crc_parity(r, degree)]) for r in flat])
    return out.reshape(b.shape[:-1] + (out.shape[-1],)).astype(np.float32)


def generate_prng_seq(length, c_init):
    """nr/utils.py:51-76, literal."""
    n_seq, n_c = 31, 1600
    c = np.zeros(length)
    x1 = np.zeros(length + n_c + n_seq)
    x2 = np.zeros(length + n_c + n_seq)
    bin_ = format(int(c_init), f"0{n_seq}b")
    ci = np.flip([int(x) for x in bin_[-n_seq:]])
    x1[0] = 1
    x2[0:n_seq] = ci
    for idx in range(length + n_c):
        x1[idx + 31] = np.mod(x1[idx + 3] + x1[idx], 2)
        x2[idx + 31] = np.mod(x2[idx + 3] + x2[idx + 2] + x2[idx + 1] + x2[idx], 2)
    for idx in range(length):
        c[idx] = np.mod(x1[idx + n_c] + x2[idx + n_c], 2)
    return c
