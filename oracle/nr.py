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


# ---- transport block chain (nr/tb_encoder.py:381-435, nr/utils.py:473-811) -------------------------------------------
def tb_params(target_tb_size, num_coded_bits, target_coderate, m, num_layers):
    """Literal float32 restatement of calculate_tb_size for one TB (utils.py:560-811)."""
    f = np.float32
    tbs_t, r = f(target_tb_size), f(target_coderate)
    if tbs_t <= 3824:
        n = np.maximum(f(3.0), f(np.floor(np.log(tbs_t) / f(np.log(2.0))) - 6))
        n_info_q = np.maximum(f(24.0), f(2 ** n * np.floor(tbs_t / 2 ** n)))
    else:
        n = np.floor(np.log(tbs_t - f(24)) / np.log(f(2.0))) - 5.0
        n_info_q = np.maximum(f(3840.0), f(2 ** n * np.round((tbs_t - 24) / 2 ** n)))
    tab = [24, 32, 40, 48, 56, 64, 72, 80, 88, 96, 104, 112, 120, 128, 136, 144, 152, 160, 168, 176, 184, 192, 208, 224, 240,
           256, 272, 288, 304, 320, 336, 352, 368, 384, 408, 432, 456, 480, 504, 528, 552, 576, 608, 640, 672, 704, 736, 768,
           808, 848, 888, 928, 984, 1032, 1064, 1128, 1160, 1192, 1224, 1256, 1288, 1320, 1352, 1416, 1480, 1544, 1608, 1672,
           1736, 1800, 1864, 1928, 2024, 2088, 2152, 2216, 2280, 2408, 2472, 2536, 2600, 2664, 2728, 2792, 2856, 2976, 3104,
           3240, 3368, 3496, 3624, 3752, 3824]
    if n_info_q <= 3824:
        num_cb, tb_size = 1, min(t for t in tab if t >= n_info_q)
    else:
        num_cb = int(np.ceil((n_info_q + 24) / 3816)) if r <= 0.25 else (int(np.ceil((n_info_q + 24) / 8424)) if n_info_q > 8424 else 1)
        tb_size = int(8 * num_cb * np.ceil((n_info_q + 24) / (8 * num_cb)) - 24)
    tb_crc = 24 if tb_size > 3824 else 16
    cb_crc = 24 if num_cb > 1 else 0
    cb_size = int((tb_size + tb_crc) / num_cb) + cb_crc
    q = num_layers * m
    n_last = int(num_coded_bits / q) % num_cb
    l_last = q * int(np.ceil(num_coded_bits / (q * num_cb)))
    l_first = q * int(np.floor(num_coded_bits / (q * num_cb)))
    return tb_size, cb_size, num_cb, tb_crc, cb_crc, [l_first] * (num_cb - n_last) + [l_last] * n_last


def tb_encode(u, num_coded_bits, target_coderate, m, num_layers, n_rnti, n_id, scramble=True):
    """[B, tb_size] -> [B, num_coded_bits] for one stream (tb_encoder.py:381-435)."""
    from .ldpc import LDPC5GEncoderRef
    u = np.asarray(u).astype(np.uint8)
    tb_size, cb_size, num_cb, tb_crc, cb_crc, cw = tb_params(u.shape[-1], num_coded_bits, target_coderate, m, num_layers)
    assert tb_size == u.shape[-1]
    x = crc_encode(u, "CRC16" if tb_crc == 16 else "CRC24A").astype(np.uint8)
    x = x.reshape(u.shape[0], num_cb, cb_size - cb_crc)
    if cb_crc:
        x = crc_encode(x, "CRC24B").astype(np.uint8)
    n_max, n_min = max(cw), min(cw)
    enc = LDPC5GEncoderRef(cb_size, n_max)
    c = enc(x.reshape(-1, cb_size)).reshape(u.shape[0], num_cb * n_max)
    def out_int(n):                                                  # encoding.py:238-244
        perm = np.zeros(n, int)
        for j in range(n // m):
            for i in range(m):
                perm[i + j * m] = i * (n // m) + j
        return perm
    perm, punc, pos = [], [], 0
    for l in cw:
        if l == n_min:
            perm.append(out_int(n_min) + pos); punc.append(np.arange(pos + n_min, pos + n_max)); pos += n_max
        else:
            perm.append(out_int(n_max) + pos); pos += l
    perm = np.concatenate(perm + punc).astype(int)
    c = c[:, perm][:, :sum(cw)]
    if scramble:
        c = np.abs(c - generate_prng_seq(sum(cw), n_rnti * 2 ** 15 + n_id))
    return c.astype(np.float32)


# ---- PUSCH -----------------------------------------------------------------------------------------
def layer_map(x, num_layers):
    """[..., n] -> [..., num_layers, n / num_layers], symbol i to layer i mod num_layers (layer_mapping.py:176-181,199).
    5..8 layers (dual codeword mode, :182-198): `x` is a pair of codewords; the first takes floor(num_layers / 2) layers."""
    if num_layers > 4:
        l0 = num_layers // 2
        y0, y1 = layer_map(np.asarray(x[0]), l0), layer_map(np.asarray(x[1]), num_layers - l0)
        return np.concatenate([y0, y1], axis=-2)
    n = x.shape[-1]
    return np.swapaxes(x.reshape(x.shape[:-1] + (n // num_layers, num_layers)), -1, -2)


def layer_demap(llr, num_bits_per_symbol):
    """[..., num_layers, n] -> [..., num_layers * n] keeping the LLRs of a symbol together (layer_mapping.py:268-280)."""
    q = num_bits_per_symbol
    x = llr.reshape(llr.shape[:-1] + (llr.shape[-1] // q, q))
    return np.swapaxes(x, -2, -3).reshape(llr.shape[:-2] + (-1,))


def pusch_transmit(cfg, b, tb_size, num_coded_bits):
    """Frequency-domain PUSCH slot of ONE transmitter (pusch_transmitter.py:199-243): b [B, tb_size] ->
    [B, num_antenna_ports, num_symbols, num_subcarriers].

    `cfg` carries host-side configuration results only: m (bits/symbol), target_coderate, num_layers, n_rnti, n_id,
    dmrs_mask [subcarriers, symbols] and dmrs_grid [layers, subcarriers, symbols] restricted to the allocation, and the
    precoding matrix w [ports, layers] or None."""
    from . import mapping as M
    c = tb_encode(b, num_coded_bits, cfg["target_coderate"], cfg["m"], cfg["num_layers"], cfg["n_rnti"], cfg["n_id"])
    x = M.mapper(c, M.qam(cfg["m"]))[0]                                # [B, symbols] (second output: symbol indices)
    xl = layer_map(x, cfg["num_layers"])                               # [B, L, n]
    mask = np.asarray(cfg["dmrs_mask"]).T                              # [S, F]
    dm = np.transpose(np.asarray(cfg["dmrs_grid"]), (0, 2, 1))         # [L, S, F]
    bsz, nl = xl.shape[0], cfg["num_layers"]
    grid = np.zeros((bsz, nl) + mask.shape, complex)
    for l in range(nl):                                                # resource_grid.py:394-412: row-major fill
        grid[:, l][:, ~mask] = xl[:, l]
        grid[:, l][:, mask] = dm[l][mask]
    if cfg.get("w") is not None:                                       # pusch_precoder.py:75-95
        grid = np.einsum("pl,blsf->bpsf", np.asarray(cfg["w"]), grid)
    return grid


def pusch_ls_combine(h, err, num_dmrs_syms, dmrs_length, num_cdm_groups_without_data):
    """CDM de-spreading of the LS estimates h [..., P] at the DMRS REs (pusch_channel_estimation.py:138-169)."""
    shp = h.shape
    pps = shp[-1] // num_dmrs_syms
    h = h.reshape(shp[:-1] + (num_dmrs_syms, pps))
    err = np.array(err, dtype=float)
    if dmrs_length == 2:
        h = (h[..., 0::2, :] + h[..., 1::2, :]) / 2
        h = np.repeat(h, 2, axis=-2)
        err = err / 2
    n = 2 * num_cdm_groups_without_data
    h = h.reshape(shp[:-1] + (shp[-1] // n, n))
    cond = np.abs(h) > 0
    avg = np.sum(h, axis=-1, keepdims=True) / 2
    h = np.where(cond, np.repeat(avg, n, axis=-1), 0)
    return h.reshape(shp), err / 2
