"""Oracle: LDPC BP decoding and 5G rate matching (NumPy glue around oracle/ldpc_bp_ref.c). TEST INFRASTRUCTURE.

Follows /root/reference/src/sionna/phy/fec/ldpc/decoding.py:
  edge lists            :277-345     -> ref_edges()
  LDPCBPDecoder.call    :544-637     -> bp_decode()          (C: sbo_bp_decode)
  node updates          :681-1166    -> cn_update()/vn_update() (C: sbo_cn_update / sbo_vn_update)
  LDPC5GDecoder         :1302-1403   -> LDPC5GDecoderRef.__init__ (pruning, layered schedule)
  LDPC5GDecoder.call    :1427-1536   -> LDPC5GDecoderRef.__call__ (literal concat / slice sequence)
and encoding.py (:61-137, 248-409, 572-668) -> LDPC5GEncoderRef (gather/reduce_sum RU encoding restated with
scipy.sparse products mod 2; pinned by the 28 generator-matrix goldens of /root/reference/test/codes/ldpc).
"""
import ctypes as C
import os
import numpy as np
import scipy.sparse as sp

from . import build as _build

CN_RULES = {"boxplus-phi": 0, "boxplus": 1, "minsum": 2, "min": 2, "offset-minsum": 3, "identity": 4}
VN_RULES = {"sum": 0, "identity": 1}
_lib = None
_lib_pure = None


def _decl(h):
    h.sbo_bp_decode.restype = C.c_int
    h.sbo_bp_decode.argtypes = [C.c_int] * 3 + [C.c_void_p] * 4 + [C.c_int] * 3 + [C.c_void_p, C.c_int, C.c_int,
                                C.c_int, C.c_int, C.c_float, C.c_float, C.c_int, C.c_void_p, C.c_void_p,
                                C.c_void_p, C.c_int, C.c_int]
    h.sbo_cn_update.restype = None
    h.sbo_cn_update.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_float, C.c_int]
    h.sbo_vn_update.restype = C.c_float
    h.sbo_vn_update.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_int, C.c_float]
    h.sbo_phi.restype = C.c_float
    h.sbo_phi.argtypes = [C.c_float, C.c_int]
    h.sbo_is_pure_libm.restype = C.c_int
    return h


def lib():
    global _lib
    if _lib is None:
        _lib = _decl(C.CDLL(_build.build()))
    return _lib


def lib_pure():
    """ldpc_bp_ref.c compiled WITHOUT the product's sb_math.h (libm only; math_mode must be 0)."""
    global _lib_pure
    if _lib_pure is None:
        _build.build()
        _lib_pure = _decl(C.CDLL(_build.OUT_LIBM))
        assert _lib_pure.sbo_is_pure_libm() == 1
    return _lib_pure


def _p(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


def phi(x, math_mode=0):
    return float(lib().sbo_phi(float(np.float32(x)), math_mode))


def cn_update(rule, x, llr_clipping=None, offset=0.5, math_mode=0):
    """c2v messages of ONE check node with incoming v2c ``x`` (1-D float32)."""
    x = np.ascontiguousarray(x, np.float32)
    out = np.empty_like(x)
    lib().sbo_cn_update(CN_RULES[rule], _p(x), _p(out), len(x), int(llr_clipping is not None),
                        float(llr_clipping or 0.0), float(offset), math_mode)
    return out


def vn_update(rule, c2v, llr_ch, llr_clipping=None):
    c2v = np.ascontiguousarray(c2v, np.float32)
    out = np.empty_like(c2v)
    xt = lib().sbo_vn_update(VN_RULES[rule], _p(c2v), _p(out), len(c2v), float(np.float32(llr_ch)),
                             int(llr_clipping is not None), float(llr_clipping or 0.0))
    return out, np.float32(xt)


def ref_edges(pcm):
    """cn_idx, vn_idx of every edge in the reference's VN order (decoding.py:277-292, same NumPy calls)."""
    if isinstance(pcm, np.ndarray):
        pcm = sp.csr_matrix(pcm)
    cn_idx, vn_idx, _ = sp.find(pcm)
    idx = np.argsort(vn_idx)
    return cn_idx[idx], vn_idx[idx]


def bp_decode(pcm, llr_ch, num_iter=20, cn_update="boxplus-phi", vn_update="sum", llr_max=20.0, hard_out=True,
              msg_v2c=None, return_state=False, cn_schedule=None, offset=0.5, math_mode=0, order="reference",
              num_threads=None, edges=None, pure=False):
    """LDPCBPDecoder.call on ``llr_ch [B, N]`` (logits). ``msg_v2c``/returned state are ``[E, B]`` in the
    reference's edge order. ``order``: "reference" sums node inputs in the reference's own list orders
    (argsort results, decoding.py:286, 329); "kernel" in ascending neighbour index, the CUDA kernels' order."""
    cn_idx, vn_idx = ref_edges(pcm) if edges is None else edges
    C_, N_ = pcm.shape
    E_ = len(vn_idx)
    llr_ch = np.ascontiguousarray(llr_ch, np.float32).reshape(-1, N_)
    B_ = llr_ch.shape[0]
    if order == "kernel":
        ren = np.lexsort((cn_idx, vn_idx))           # oracle edge e' = reference edge ren[e']
        cn_o, vn_o = cn_idx[ren], vn_idx[ren]
        v2c_perm = np.lexsort((vn_o, cn_o))
    else:
        ren = np.arange(E_)
        cn_o, vn_o = cn_idx, vn_idx
        v2c_perm = np.argsort(cn_o)                   # decoding.py:329
    vn_ptr = np.zeros(N_ + 1, np.int32)
    np.cumsum(np.bincount(vn_o, minlength=N_), out=vn_ptr[1:])
    cn_ptr = np.zeros(C_ + 1, np.int32)
    np.cumsum(np.bincount(cn_o, minlength=C_), out=cn_ptr[1:])
    cn_edge = np.ascontiguousarray(v2c_perm, np.int32)
    if cn_schedule is None:
        sched = np.arange(C_, dtype=np.int32)[None, :]
        flooding = 1
    else:
        sched = np.ascontiguousarray(cn_schedule, np.int32)
        flooding = 0
    st_in = None
    if msg_v2c is not None:
        st_in = np.ascontiguousarray(np.asarray(msg_v2c, np.float32)[ren, :])
    st_out = np.empty((E_, B_), np.float32) if return_state else None
    x = np.empty((B_, N_), np.float32)
    if num_threads is None:
        num_threads = os.cpu_count() or 1
    rc = (lib_pure() if pure else lib()).sbo_bp_decode(C_, N_, E_, _p(vn_ptr), _p(cn_ptr), _p(cn_edge), _p(sched), sched.shape[0],
                             sched.shape[1], flooding, _p(llr_ch), B_, int(num_iter), CN_RULES[cn_update],
                             VN_RULES[vn_update], float(offset), float(llr_max), int(hard_out), _p(st_in),
                             _p(st_out), _p(x), int(math_mode), int(num_threads))
    assert rc == 0
    if return_state:
        full = np.empty_like(st_out)
        full[ren, :] = st_out
        return x, full
    return x


# ---------------------------------------------------------------------------------------------------------
# 5G NR LDPC code construction + encoder restatement (encoding.py)
# ---------------------------------------------------------------------------------------------------------
_S_VAL = [[2, 4, 8, 16, 32, 64, 128, 256], [3, 6, 12, 24, 48, 96, 192, 384], [5, 10, 20, 40, 80, 160, 320],
          [7, 14, 28, 56, 112, 224], [9, 18, 36, 72, 144, 288], [11, 22, 44, 88, 176, 352], [13, 26, 52, 104, 208],
          [15, 30, 60, 120, 240]]
_TABLES = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "sionna_b200", "phy", "fec", "ldpc",
                       "codes", "bg_tables.npz")


class LDPC5GEncoderRef:
    """Restatement of LDPC5GEncoder (encoding.py:61-137 construction, :599-668 call)."""

    def __init__(self, k, n, num_bits_per_symbol=None, bg=None):
        k, n = int(k), int(n)
        self.k, self.n = k, n
        r = k / n
        if bg is None:                                               # :252-260
            if k <= 292:
                bg = "bg2"
            elif k <= 3824 and r <= 0.67:
                bg = "bg2"
            elif r <= 0.25:
                bg = "bg2"
            else:
                bg = "bg1"
        self.bg = bg
        if bg == "bg1":                                              # :375-385
            k_b = 22
        elif k > 640:
            k_b = 10
        elif k > 560:
            k_b = 9
        elif k > 192:
            k_b = 8
        else:
            k_b = 6
        min_val, z, i_ls = 100000, 0, 0                              # :388-401
        for i, s in enumerate(_S_VAL):
            for s1 in s:
                x = k_b * s1
                if x >= k and x < min_val:
                    min_val, z, i_ls = x, s1, i
        k_b = 22 if bg == "bg1" else 10                              # :404-407
        self.z, self.i_ls, self.k_b = z, i_ls, k_b
        with np.load(_TABLES) as t:
            rows, cols, sh = t[f"{bg}_row"].astype(int), t[f"{bg}_col"].astype(int), t[f"{bg}_shift"][:, i_ls].astype(int)
        shape = (46, 68) if bg == "bg1" else (42, 52)
        bm = -np.ones(shape, dtype=int)
        bm[rows, cols] = sh
        self.bm = bm
        self.n_ldpc = shape[1] * z
        self.k_ldpc = k_b * z
        self.pcm = self._lift(bm, z)
        self.num_bits_per_symbol = num_bits_per_symbol
        self.out_int = self.out_int_inv = None
        if num_bits_per_symbol is not None:                          # :238-244
            m = int(num_bits_per_symbol)
            perm = np.zeros(n, dtype=int)
            for j in range(n // m):
                for i in range(m):
                    perm[i + j * m] = i * (n // m) + j
            self.out_int, self.out_int_inv = perm, np.argsort(perm)
        # RU sub-matrices (:411-434), B^-1 obtained by solving over GF(2) instead of the closed form so that the
        # restatement is an independent check of the product's closed-form construction.
        g = 4
        mb = bm.shape[0]
        self.A = self._lift(bm[0:g, 0:k_b], z)
        self.C1 = self._lift(bm[g:mb, 0:k_b], z)
        self.C2 = self._lift(bm[g:mb, k_b:k_b + g], z)
        self.Bm = self._lift(bm[0:g, k_b:k_b + g], z)

    @staticmethod
    def _lift(bm, z):                                                # :322-352
        rr, cc, dd = [], [], []
        im = np.arange(z)
        for r in range(bm.shape[0]):
            for c in range(bm.shape[1]):
                if bm[r, c] != -1:
                    rr.append(r * z + im)
                    cc.append(c * z + np.mod(im + bm[r, c], z))
        rr = np.concatenate(rr) if rr else np.zeros(0, int)
        cc = np.concatenate(cc) if cc else np.zeros(0, int)
        return sp.csr_matrix((np.ones(len(rr)), (rr, cc)), shape=(z * bm.shape[0], z * bm.shape[1]))

    @staticmethod
    def _gf2_solve(Bm, rhs):
        """Solve Bm x = rhs over GF(2) (dense Gaussian elimination on the 4Z x 4Z block)."""
        a = (np.asarray(Bm.todense()) % 2).astype(np.uint8)
        n = a.shape[0]
        x = (rhs.T % 2).astype(np.uint8)                              # [n, batch]
        aug = np.concatenate([a, x], axis=1)
        row = 0
        for col in range(n):
            piv = np.nonzero(aug[row:, col])[0]
            assert len(piv), "B is singular"
            p = piv[0] + row
            if p != row:
                aug[[row, p]] = aug[[p, row]]
            others = np.nonzero(aug[:, col])[0]
            others = others[others != row]
            aug[others] ^= aug[row]
            row += 1
        return aug[:, n:].T                                           # [batch, n]

    def encode_full(self, u):
        """[B, k] bits -> [B, n_ldpc] full codeword before rate matching (:636-643, :572-591)."""
        u = np.asarray(u).astype(np.int64).reshape(-1, self.k)
        s = np.concatenate([u, np.zeros((u.shape[0], self.k_ldpc - self.k), np.int64)], axis=1)
        As = (self.A @ s.T).T.astype(np.int64) % 2
        p_a = self._gf2_solve(self.Bm, As).astype(np.int64)
        p_b = ((self.C1 @ s.T).T + (self.C2 @ p_a.T).T).astype(np.int64) % 2
        return np.concatenate([s, p_a, p_b], axis=1)

    def __call__(self, u):
        shape = list(np.shape(u))
        c = self.encode_full(u)
        c_no_filler = np.concatenate([c[:, :self.k], c[:, self.k_ldpc:]], axis=1)    # :646-651
        c_short = c_no_filler[:, 2 * self.z: 2 * self.z + self.n]                    # :655
        if self.out_int is not None:
            c_short = c_short[:, self.out_int]                                       # :661
        return c_short.reshape(shape[:-1] + [self.n]).astype(np.float32)


class LDPC5GDecoderRef:
    """Restatement of LDPC5GDecoder (decoding.py:1302-1403 and :1427-1536) around bp_decode()."""

    def __init__(self, enc, cn_update="boxplus-phi", vn_update="sum", cn_schedule="flooding", hard_out=True,
                 return_infobits=True, num_iter=20, llr_max=20.0, prune_pcm=True, return_state=False):
        self.enc, self.cn_update, self.vn_update = enc, cn_update, vn_update
        self.hard_out, self.return_infobits, self.num_iter = hard_out, return_infobits, num_iter
        self.llr_max, self.return_state = float(llr_max), return_state
        pcm = enc.pcm
        if prune_pcm:                                                # :1344-1378
            dv = np.asarray(pcm.sum(axis=0))
            last_pos = enc.n_ldpc
            for idx in range(enc.n_ldpc - 1, 0, -1):
                if dv[0, idx] == 1:
                    last_pos = idx
                else:
                    break
            k_filler = enc.k_ldpc - enc.k
            nb_punc_bits = (enc.n_ldpc - k_filler) - enc.n - 2 * enc.z
            if cn_schedule == "layered":
                nb_punc_bits = int(np.floor(nb_punc_bits / enc.z) * enc.z)
            self.n_pruned = int(np.max((last_pos, enc.n_ldpc - nb_punc_bits)))
            self.nb_pruned_nodes = enc.n_ldpc - self.n_pruned
            if self.nb_pruned_nodes > 0:
                pcm = pcm[:-self.nb_pruned_nodes, :-self.nb_pruned_nodes]
        else:
            self.nb_pruned_nodes = 0
            self.n_pruned = enc.n_ldpc
        self.pcm = sp.csr_matrix(pcm)
        self.schedule = None
        if isinstance(cn_schedule, str) and cn_schedule == "layered":    # :1384-1390
            z = enc.z
            self.schedule = np.stack([np.arange(z) + i * z for i in range(int(self.pcm.shape[0] / z))], axis=0)
        elif not isinstance(cn_schedule, str):
            self.schedule = np.asarray(cn_schedule)
        self.edges = ref_edges(self.pcm)

    def __call__(self, llr_ch, num_iter=None, msg_v2c=None, math_mode=0, order="reference", num_threads=None,
                 pure=False):
        enc = self.enc
        llr_ch = np.asarray(llr_ch, np.float32)
        shape = list(llr_ch.shape)
        x = llr_ch.reshape(-1, enc.n)
        bs = x.shape[0]
        if enc.out_int_inv is not None:                              # :1438-1441
            x = x[:, enc.out_int_inv]
        llr_5g = np.concatenate([np.zeros((bs, 2 * enc.z), np.float32), x], axis=1)          # :1444
        k_filler = enc.k_ldpc - enc.k
        nb_punc_bits = (enc.n_ldpc - k_filler) - enc.n - 2 * enc.z
        llr_5g = np.concatenate([llr_5g, np.zeros((bs, nb_punc_bits - self.nb_pruned_nodes), np.float32)], axis=1)
        x1 = llr_5g[:, :enc.k]                                                               # :1462
        nb_par_bits = enc.n_ldpc - k_filler - enc.k - self.nb_pruned_nodes
        x2 = llr_5g[:, enc.k: enc.k + nb_par_bits]
        zf = -np.float32(self.llr_max) * np.ones((bs, k_filler), np.float32)                 # :1472
        llr_5g = np.concatenate([x1, zf, x2], axis=1)
        out = bp_decode(self.pcm, llr_5g, num_iter=self.num_iter if num_iter is None else num_iter,
                        cn_update=self.cn_update, vn_update=self.vn_update, llr_max=self.llr_max,
                        hard_out=self.hard_out, msg_v2c=msg_v2c, return_state=self.return_state,
                        cn_schedule=self.schedule, math_mode=math_mode, order=order, num_threads=num_threads,
                        edges=self.edges, pure=pure)
        x_hat, st = out if self.return_state else (out, None)
        if self.return_infobits:                                                             # :1486-1499
            res = x_hat[:, :enc.k].reshape(shape[:-1] + [enc.k])
        else:                                                                                # :1501-1536
            xx = x_hat.reshape(bs, self.n_pruned)
            x_no_filler = np.concatenate([xx[:, :enc.k], xx[:, enc.k_ldpc: self.n_pruned]], axis=1)
            x_short = x_no_filler[:, 2 * enc.z: 2 * enc.z + enc.n]
            if enc.out_int is not None:
                x_short = x_short[:, enc.out_int]
            res = x_short.reshape(shape)
        return (res, st) if self.return_state else res
