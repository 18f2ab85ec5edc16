"""5G NR LDPC encoder (mirror of /root/reference/src/sionna/phy/fec/ldpc/encoding.py:14-668).

Host side (NumPy/SciPy, construction time only): base-graph selection (encoding.py:248-282), lifting-size
selection (:354-409), base-matrix tables (38.212 Tab. 5.3.2-2/-3, shipped as ``codes/bg_tables.npz``),
lifting to the sparse parity-check matrix (:322-352) and the Richardson-Urbanke split
``H = [[A, B, 0], [C1, C2, I]]`` with the closed-form ``B^-1`` (:411-522).
Device side: ``sb_ldpc5g_encode`` (``csrc/ldpc_enc.cu``) evaluates ``p_a = B^-1 (A s)``,
``p_b = C1 s + C2 p_a`` over GF(2) and applies filler removal, 2Z puncturing, truncation to ``n`` and the
optional 38.212 5.4.2.2 interleaver in one kernel (reference: ``call`` :599-668, ``_encode_fast`` :572-591).
"""
import numbers
import os
import numpy as np
import scipy.sparse as sp
import torch

from ...block import Block
from ... import config as _cfg_mod  # noqa: F401
from ...._lib import lib, check, ptr, current_stream

_CODES = os.path.join(os.path.dirname(os.path.abspath(__file__)), "codes")
_BG_CACHE = {}

# 38.212 Tab. 5.3.2-1: lifting-size sets, index = i_LS
_LIFTING_SETS = [[2, 4, 8, 16, 32, 64, 128, 256],
                 [3, 6, 12, 24, 48, 96, 192, 384],
                 [5, 10, 20, 40, 80, 160, 320],
                 [7, 14, 28, 56, 112, 224],
                 [9, 18, 36, 72, 144, 288],
                 [11, 22, 44, 88, 176, 352],
                 [13, 26, 52, 104, 208],
                 [15, 30, 60, 120, 240]]


def _bg_tables():
    if not _BG_CACHE:
        with np.load(os.path.join(_CODES, "bg_tables.npz")) as d:
            for k in d.files:
                _BG_CACHE[k] = d[k]
    return _BG_CACHE


def sel_basegraph(k, r, bg=None):
    """Base-graph choice of encoding.py:248-282 (same thresholds, same error conditions)."""
    if bg is None:
        if k <= 292:
            bg = "bg2"
        elif k <= 3824 and r <= 0.67:
            bg = "bg2"
        elif r <= 0.25:
            bg = "bg2"
        else:
            bg = "bg1"
    elif bg not in ("bg1", "bg2"):
        raise ValueError("Basegraph must be bg1, bg2 or None.")
    if bg == "bg1" and k > 8448:
        raise ValueError("K is not supported by BG1 (too large).")
    if bg == "bg2" and k > 3840:
        raise ValueError(f"K is not supported by BG2 (too large) k ={k}.")
    if bg == "bg1" and r < 1 / 3:
        raise ValueError("Only coderate>1/3 supported for BG1. Remark: Repetition coding is currently not supported.")
    if bg == "bg2" and r < 1 / 5:
        raise ValueError("Only coderate>1/5 supported for BG2. Remark: Repetition coding is currently not supported.")
    return bg


def sel_lifting(k, bg):
    """Lifting size Z, set index i_LS and k_b (encoding.py:354-409): smallest k_b*Z >= k, first set wins ties."""
    if bg == "bg1":
        k_b = 22
    elif k > 640:
        k_b = 10
    elif k > 560:
        k_b = 9
    elif k > 192:
        k_b = 8
    else:
        k_b = 6
    best, z, i_ls = 100000, 0, 0
    for i, s in enumerate(_LIFTING_SETS):
        for s1 in s:
            x = k_b * s1
            if x >= k and x < best:
                best, z, i_ls = x, s1, i
    return z, i_ls, (22 if bg == "bg1" else 10)


def load_basegraph(i_ls, bg):
    """Dense base matrix of shifts (-1 = empty) for set index ``i_ls`` (encoding.py:284-320)."""
    if i_ls > 7:
        raise ValueError("i_ls too large.")
    if i_ls < 0:
        raise ValueError("i_ls cannot be negative.")
    if bg not in ("bg1", "bg2"):
        raise ValueError("Basegraph not supported.")
    t = _bg_tables()
    shape = (46, 68) if bg == "bg1" else (42, 52)
    bm = np.full(shape, -1, dtype=np.int64)
    bm[t[f"{bg}_row"].astype(int), t[f"{bg}_col"].astype(int)] = t[f"{bg}_shift"][:, i_ls]
    return bm


def lift_basegraph(bm, z):
    """Replace every entry s >= 0 by the ZxZ identity cyclically shifted by ``s mod Z`` (encoding.py:322-352)."""
    r, c = np.nonzero(bm >= 0)
    s = bm[r, c]
    im = np.arange(z)
    rows = (r[:, None] * z + im[None, :]).ravel()
    cols = (c[:, None] * z + np.mod(im[None, :] + s[:, None], z)).ravel()
    return sp.csr_matrix((np.ones(rows.size), (rows, cols)), shape=(z * bm.shape[0], z * bm.shape[1]))


def _csr_lists(mat):
    """CSR (indptr int32, indices int32 ascending per row) of a binary sparse matrix with even entries removed."""
    m = sp.csr_matrix(mat)
    m.data = np.mod(np.round(m.data), 2)
    m.eliminate_zeros()
    m.sort_indices()
    return m.indptr.astype(np.int32), m.indices.astype(np.int32)


class LDPC5GEncoder(Block):
    """LDPC5GEncoder(k, n, num_bits_per_symbol=None, bg=None, precision=None)

    5G NR LDPC encoder with rate matching (38.212 5.3.2 / 5.4.2), same constructor, properties and
    ``[..., k] -> [..., n]`` float 0/1 interface as the reference (encoding.py:61-67, 143-190, 599-668).
    """

    def __init__(self, k, n, num_bits_per_symbol=None, bg=None, precision=None, **kwargs):
        super().__init__(precision=precision, **kwargs)
        if not isinstance(k, numbers.Number):
            raise TypeError("k must be a number.")
        if not isinstance(n, numbers.Number):
            raise TypeError("n must be a number.")
        k, n = int(k), int(n)
        if k > 8448:
            raise ValueError("Unsupported code length (k too large).")
        if k < 12:
            raise ValueError("Unsupported code length (k too small).")
        if n > (316 * 384):
            raise ValueError("Unsupported code length (n too large).")
        if n < 0:
            raise ValueError("Unsupported code length (n negative).")
        self._k, self._n = k, n
        self._coderate = k / n
        self._check_input = True
        if self._coderate > (948 / 1024):
            print(f"Warning: effective coderate r>948/1024 for n={n}, k={k}.")
        if self._coderate > 0.95:
            raise ValueError(f"Unsupported coderate (r>0.95) for n={n}, k={k}.")
        if self._coderate < (1 / 5):
            raise ValueError("Unsupported coderate (r<1/5).")

        self._bg = sel_basegraph(k, self._coderate, bg)
        self._z, self._i_ls, self._k_b = sel_lifting(k, self._bg)
        self._bm = load_basegraph(self._i_ls, self._bg)
        self._n_ldpc = self._bm.shape[1] * self._z
        self._k_ldpc = self._k_b * self._z
        self._pcm = lift_basegraph(self._bm, self._z)

        self._num_bits_per_symbol = num_bits_per_symbol
        if num_bits_per_symbol is not None:
            self._out_int, self._out_int_inv = self.generate_out_int(self._n, self._num_bits_per_symbol)
        else:
            self._out_int, self._out_int_inv = None, None
        if 2 * self._z + self._n > self._n_ldpc - (self._k_ldpc - self._k):
            raise ValueError("n exceeds the number of transmittable codeword bits for this k.")
        self._enc_handle = None

    # ---- properties (encoding.py:143-190) -------------------------------------------------------
    @property
    def k(self):
        return self._k

    @property
    def n(self):
        return self._n

    @property
    def coderate(self):
        return self._coderate

    @property
    def k_ldpc(self):
        return self._k_ldpc

    @property
    def n_ldpc(self):
        return self._n_ldpc

    @property
    def pcm(self):
        return self._pcm

    @property
    def z(self):
        return self._z

    @property
    def num_bits_per_symbol(self):
        return self._num_bits_per_symbol

    @property
    def out_int(self):
        return self._out_int

    @property
    def out_int_inv(self):
        return self._out_int_inv

    # ---- utilities ----------------------------------------------------------------------------------
    def generate_out_int(self, n, num_bits_per_symbol):
        """Rate-matching output interleaver of 38.212 5.4.2.2 (encoding.py:196-246): the coded bits are written
        row-wise into a ``num_bits_per_symbol x n/num_bits_per_symbol`` array and read column-wise."""
        if n % 1 != 0:
            raise ValueError("n must be int.")
        if num_bits_per_symbol % 1 != 0:
            raise ValueError("num_bits_per_symbol must be int.")
        n = int(n)
        if n <= 0:
            raise ValueError("n must be a positive integer.")
        if num_bits_per_symbol <= 0:
            raise ValueError("num_bits_per_symbol must be a positive integer.")
        m = int(num_bits_per_symbol)
        if n % m != 0:
            raise ValueError("n must be a multiple of num_bits_per_symbol.")
        j, i = np.meshgrid(np.arange(n // m), np.arange(m), indexing="ij")
        perm_seq = (i * (n // m) + j).reshape(-1).astype(int)
        perm_seq_inv = np.argsort(perm_seq)
        return perm_seq, perm_seq_inv

    def _ru_submatrices(self):
        """A, B^-1, C1, C2 of the Richardson-Urbanke split with gap g = 4 (encoding.py:411-522)."""
        g, z, k_b, bm = 4, self._z, self._k_b, self._bm
        mb = bm.shape[0]
        hm_a = lift_basegraph(bm[0:g, 0:k_b], z)
        hm_c1 = lift_basegraph(bm[g:mb, 0:k_b], z)
        hm_c2 = lift_basegraph(bm[g:mb, k_b:k_b + g], z)
        bm_b = bm[0:g, k_b:k_b + g]
        # B = [[P_A I 0 0],[P_B I I 0],[0 0 I I],[P_A 0 0 I]] (bg1; bg2 swaps rows 1/2 roles): its inverse is
        # built from P_B^-1 and P_A P_B^-1 (encoding.py:436-522).
        pm_a = int(bm_b[0, 0])
        pm_b_inv = int(-bm_b[1, 0]) if self._bg == "bg1" else int(-bm_b[2, 0])
        im = np.eye(z)
        b_inv = np.roll(im, pm_b_inv, axis=1)
        ab_inv = np.roll(im, pm_a, axis=1) @ b_inv
        ia = im + ab_inv
        if self._bg == "bg1":
            blocks = [[b_inv, b_inv, b_inv, b_inv], [ia, ab_inv, ab_inv, ab_inv],
                      [ab_inv, ab_inv, ia, ia], [ab_inv, ab_inv, ab_inv, ia]]
        else:
            blocks = [[b_inv, b_inv, b_inv, b_inv], [ia, ab_inv, ab_inv, ab_inv],
                      [ia, ia, ab_inv, ab_inv], [ab_inv, ab_inv, ab_inv, ia]]
        hm_b_inv = sp.csr_matrix(np.block(blocks))
        return hm_a, hm_b_inv, hm_c1, hm_c2

    def _tx_vn(self):
        """VN index (in the n_ldpc codeword) transmitted at output position j (encoding.py:645-661)."""
        i = np.arange(self._n) if self._out_int is None else np.asarray(self._out_int)
        q = 2 * self._z + i
        return np.where(q < self._k, q, q + (self._k_ldpc - self._k)).astype(np.int32)

    def build(self, input_shape):
        if input_shape[-1] != self._k:
            raise ValueError("Last dimension must be of length k.")

    def _ensure_handle(self):
        if self._enc_handle is None:
            from . import _enc_native
            self._enc_handle = _enc_native.EncoderHandle(self)
        return self._enc_handle

    def call(self, bits):
        """Encode ``[..., k]`` information bits (float 0/1) into ``[..., n]`` codeword bits."""
        shape = list(bits.shape)
        u = bits.reshape(-1, shape[-1]).contiguous()
        if self._check_input:
            if not bool(torch.all((u == 0) | (u == 1))):
                raise ValueError("Input must be binary.")
            self._check_input = False
        c = self._ensure_handle().encode(u)
        return c.reshape(shape[:-1] + [self._n])
