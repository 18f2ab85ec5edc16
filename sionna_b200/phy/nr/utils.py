"""NR utilities (mirror of /root/reference/src/sionna/phy/nr/utils.py:473-811): transport-block size determination of
38.214 5.1.3.2 / 6.1.4.2 and code-block segmentation of 38.212 5.2.2, scalar host arithmetic in float32 as the reference."""
import numpy as np

_TAB51321 = [24, 32, 40, 48, 56, 64, 72, 80, 88, 96, 104, 112, 120, 128, 136, 144, 152, 160, 168, 176, 184, 192, 208, 224,
             240, 256, 272, 288, 304, 320, 336, 352, 368, 384, 408, 432, 456, 480, 504, 528, 552, 576, 608, 640, 672, 704,
             736, 768, 808, 848, 888, 928, 984, 1032, 1064, 1128, 1160, 1192, 1224, 1256, 1288, 1320, 1352, 1416, 1480,
             1544, 1608, 1672, 1736, 1800, 1864, 1928, 2024, 2088, 2152, 2216, 2280, 2408, 2472, 2536, 2600, 2664, 2728,
             2792, 2856, 2976, 3104, 3240, 3368, 3496, 3624, 3752, 3824]


def calculate_tb_size(modulation_order, target_coderate, target_tb_size=None, num_coded_bits=None, num_prbs=None,
                      num_ofdm_symbols=None, num_dmrs_per_prb=None, num_layers=1, num_ov=0, tb_scaling=1.0,
                      return_cw_length=True, verbose=False, precision=None):
    """Returns ``(tb_size, cb_size, num_cb, tb_crc_length, cb_crc_length[, cw_lengths])`` (utils.py:473-811)."""
    f = np.float32
    m = int(modulation_order)
    num_layers = int(num_layers)
    if num_coded_bits is None:
        assert None not in (num_prbs, num_ofdm_symbols, num_dmrs_per_prb), \
            "If num_coded_bits is None then num_prbs, num_ofdm_symbols, num_dmrs_per_prb must be specified."
        n_re_per_prb = 12 * int(num_ofdm_symbols) - int(num_dmrs_per_prb) - int(num_ov)
        n_re = min(156, n_re_per_prb) * int(num_prbs)
        num_coded_bits = int(float(tb_scaling) * n_re * m * num_layers)
    num_coded_bits = int(num_coded_bits)
    assert num_coded_bits % m == 0, "num_coded_bits must be a multiple of modulation_order."
    assert num_coded_bits % num_layers == 0, "num_coded_bits must be a multiple of num_layers."
    r = f(target_coderate)
    if target_tb_size is not None:
        tbs_t = f(target_tb_size)
        assert tbs_t < num_coded_bits, "target_tb_size must be less than num_coded_bits."
    else:
        tbs_t = r * f(num_coded_bits)
    if tbs_t <= 3824:
        n = max(f(3.0), f(np.floor(np.log(tbs_t) / f(np.log(2.0))) - 6))
        n_info_q = max(f(24.0), f(2 ** n * np.floor(tbs_t / 2 ** n)))
    else:
        n = np.floor(np.log(tbs_t - f(24)) / np.log(f(2.0))) - 5.0
        n_info_q = max(f(3840.0), f(2 ** n * np.round((tbs_t - 24) / 2 ** n)))
    if n_info_q <= 3824:
        num_cb = 1
    elif r <= 1 / 4:
        num_cb = int(np.ceil((n_info_q + 24) / 3816))
    elif n_info_q > 8424:
        num_cb = int(np.ceil((n_info_q + 24) / 8424))
    else:
        num_cb = 1
    if n_info_q <= 3824:
        tb_size = next((t for t in _TAB51321 if t >= n_info_q), _TAB51321[-1])
    else:
        tb_size = int(8 * num_cb * np.ceil((n_info_q + 24) / (8 * num_cb)) - 24)
    tb_crc_length = 24 if tb_size > 3824 else 16
    cb_crc_length = 24 if num_cb > 1 else 0
    cb_size = int((tb_size + tb_crc_length) / num_cb) + cb_crc_length
    if verbose:
        print(f"Modulation order: {m}\nTarget coderate: {float(r):.3f}\nEffective coderate: {tb_size / num_coded_bits:.3f}")
        print(f"Number of layers: {num_layers}\nInfo bits per TB: {tb_size}\nTB CRC length: {tb_crc_length}")
        print(f"Total number of coded TB bits: {num_coded_bits}\nInfo bits per CB: {cb_size}\nNumber of CBs: {num_cb}")
    if not return_cw_length:
        return tb_size, cb_size, num_cb, tb_crc_length, cb_crc_length
    q = num_layers * m
    num_last = int(num_coded_bits / q) % num_cb
    len_last = q * int(np.ceil(num_coded_bits / (q * num_cb)))
    num_first = num_cb - num_last
    len_first = q * int(np.floor(num_coded_bits / (q * num_cb)))
    cw_length = np.array([len_first] * num_first + [len_last] * num_last, np.int64)
    return tb_size, cb_size, num_cb, tb_crc_length, cb_crc_length, cw_length


_NR_TABLES = None


def nr_tables():
    """3GPP table data (MCS tables, PUSCH codebooks) of codes/nr_tables.npz (tools/make_nr_tables.py)."""
    global _NR_TABLES
    if _NR_TABLES is None:
        import os
        with np.load(os.path.join(os.path.dirname(__file__), "codes", "nr_tables.npz")) as z:
            _NR_TABLES = {k: z[k] for k in z.files}
    return _NR_TABLES


def decode_mcs_index(mcs_index, table_index=1, is_pusch=True, transform_precoding=False, pi2bpsk=False,
                     check_index_validity=True, verbose=False):
    """(modulation_order, target_coderate) of an MCS index (utils.py:80-304). Table set: TS 38.214 5.1.3.1-1..4 for
    PDSCH and for PUSCH without transform precoding, 6.1.4.1-1/2 for PUSCH with transform precoding (where the first
    entries scale with q = 1 (pi/2-BPSK) or 2). Accepts scalars or arrays (broadcast together); rates are /1024."""
    idx = np.asarray(mcs_index)
    shape = np.broadcast(idx, np.asarray(table_index), np.asarray(is_pusch), np.asarray(transform_precoding),
                         np.asarray(pi2bpsk)).shape
    mcs = np.broadcast_to(idx, shape).astype(np.int64)
    tab = np.broadcast_to(np.asarray(table_index), shape).astype(np.int64)
    pus = np.broadcast_to(np.asarray(is_pusch, bool), shape)
    tpr = np.broadcast_to(np.asarray(transform_precoding, bool), shape)
    q = np.where(np.broadcast_to(np.asarray(pi2bpsk, bool), shape), 1, 2)
    assert np.all(mcs >= 0), "mcs_index must be non-negative"
    assert np.all(mcs <= 28), "mcs_index must be <= 28"
    assert np.all(np.isin(tab, (1, 2, 3, 4))), "table_index must contain values in [1,2,3,4]"
    t = nr_tables()
    family = np.where(~pus | ~tpr, 1, 0)                      # 0: PUSCH with transform precoding
    mod = t["mcs_mod_orders"][family, tab - 1, mcs].astype(np.int64)
    rate = t["mcs_target_rates"][family, tab - 1, mcs].astype(np.float64)
    if check_index_validity:
        assert np.all(mod >= 0), "Invalid MCS index"
    scaled = (family == 0) & (((tab == 1) & (mcs < 2)) | ((tab == 2) & (mcs < 6)))
    mod = np.where(scaled, mod * q, mod)
    rate = np.where(scaled, rate / q, rate) / 1024
    if verbose:
        print(f"Modulation order: {mod}")
        print(f"Target code rate: {rate}")
    if shape == ():
        return int(mod), np.float32(rate)
    return mod.astype(np.int32), rate.astype(np.float32)
