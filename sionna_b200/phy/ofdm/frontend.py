"""Fused OFDM receive front-end: received resource grid -> bit LLRs in one kernel launch
(``sb_ofdm_frontend``, csrc/frontend.cu). An extension of the reference's block set: it computes what

    LSChannelEstimator(rg, interpolation_type)   (or PUSCHLSChannelEstimator)          ofdm/channel_estimation.py:138-285
    LinearDetector("lmmse", "bit", method, rg, sm, "qam", m)                            ofdm/detection.py:740-847

compute one after the other, without writing the channel estimate, its error variance, the equalised symbols or the
effective noise variance to HBM. `PUSCHReceiver` switches to it on its own when its estimator and detector are the
defaults; link models can use `FusedLSLinearDetector` directly.

Supported: no interfering streams (every receiver sees only its own transmitters' streams), at most 4 streams per
receiver, square-QAM constellations up to 1024 points without prior, nearest-neighbour / linear / time-averaged linear
interpolation of any pilot pattern whose estimates combine at most `MAX_TERMS` received pilots per resource element.
`fusable(...)` tells whether a configuration qualifies; otherwise use the separate blocks.
"""
import numpy as np
import torch

from ..block import Block
from ..mapping import Constellation, separable_levels_np
from ..._lib import lib, check, ptr, current_stream
from .channel_estimation import NearestNeighborInterpolator, LinearInterpolator
from .equalization import _sm_tables, _strides_for

MAX_TERMS = 16


def _interp_operator(interp, num_streams, num_re, num_pilots):
    """Dense [streams, RE, pilots] float64 matrix of the (linear) interpolation ``interp`` applies to pilot estimates."""
    if isinstance(interp, NearestNeighborInterpolator):
        w = np.zeros((num_streams, num_re, num_pilots))
        g = interp._gather_ind
        for a in range(num_streams):
            w[a, np.arange(num_re), g[a]] = 1.0
        return w
    if not isinstance(interp, LinearInterpolator):
        return None
    x0, x1, y0p, y1p, t0, t1, npil = interp._tabs_np
    a_n, s_n, f_n = x0.shape
    eye = np.eye(num_pilots)
    zero = np.zeros((1, num_pilots))
    tab = np.concatenate([zero, eye], 0)                                  # index 0 = the zero pad (fy = pilot index + 1)
    f = np.arange(f_n, dtype=np.float64)[None, :, None]
    out = np.zeros((a_n, s_n * f_n, num_pilots))
    for a in range(a_n):
        y0, y1 = tab[y0p[a]], tab[y1p[a]]                                   # [S, F, P]
        dx = (x1[a] - x0[a]).astype(np.float64)[..., None]
        with np.errstate(divide="ignore", invalid="ignore"):
            slope = np.where(dx == 0, 0.0, (y1 - y0) / np.where(dx == 0, 1.0, dx))   # divide_no_nan
        v = (f - x0[a].astype(np.float64)[..., None]) * slope + y0          # frequency interpolation on every symbol
        if interp._time_avg:
            o = np.broadcast_to((v.sum(0) / float(npil[a]))[None], (s_n, f_n, num_pilots))
        else:
            s = np.arange(s_n, dtype=np.float64)
            v0, v1 = v[t0[a]], v[t1[a]]                                     # [S, F, P]
            dt = (t1[a] - t0[a]).astype(np.float64)[:, None, None]
            with np.errstate(divide="ignore", invalid="ignore"):
                sl = np.where(dt == 0, 0.0, (v1 - v0) / np.where(dt == 0, 1.0, dt))
            o = (s - t0[a])[:, None, None] * sl + v0
        out[a] = o.reshape(s_n * f_n, num_pilots)
    return out


def _cdm_operator(pilots, pps, dmrs_length, group):
    """[streams, P, P] matrix of ``sb_pusch_ls_combine`` (nr/pusch_channel_estimation.py:131-167) and the factor it
    applies to the error variances."""
    ts, p = pilots.shape
    c = np.zeros((ts, p, p))
    units = (p // pps) // dmrs_length
    for a in range(ts):
        nz = np.abs(pilots[a]) > 0
        for u in range(units):
            for g in range(pps // group):
                p0 = u * dmrs_length * pps + g * group
                idx = [p0 + k for k in range(group)]
                src = [(i, 0.5) for i in idx] if dmrs_length == 1 else \
                      [(i, 0.25) for i in idx] + [(i + pps, 0.25) for i in idx]
                for k in idx:
                    on = nz[k] or (dmrs_length == 2 and nz[k + pps])
                    if not on:
                        continue
                    for tgt in ([k] if dmrs_length == 1 else [k, k + pps]):
                        for i, wgt in src:
                            c[a, tgt, i] = wgt
    return c, (0.5 if dmrs_length == 1 else 0.25)


def frontend_tables(resource_grid, estimator):
    """Host tables of the fused kernel for ``estimator`` (an `LSChannelEstimator` or `PUSCHLSChannelEstimator`), or
    ``None`` when its estimate is not a short linear combination of received pilots."""
    rg, pp = resource_grid, resource_grid.pilot_pattern
    ts = pp.num_tx * pp.num_streams_per_tx
    s_n, f_n = rg.num_ofdm_symbols, rg.num_effective_subcarriers
    num_re, num_p = s_n * f_n, pp.num_pilot_symbols
    pilots = np.asarray(pp.pilots).reshape(ts, num_p).astype(np.complex128)
    interp = getattr(estimator, "_interpol", None)
    if getattr(estimator, "_interpolation_type", None) is None or interp is None:
        return None
    w = _interp_operator(interp, ts, num_re, num_p)
    if w is None:
        return None
    e_ls = np.where(np.abs(pilots) > 0, 1.0 / np.maximum(np.abs(pilots) ** 2, 1e-300), 0.0)      # err_var / no at the pilots
    scale = 1.0
    if hasattr(estimator, "_num_cdm_groups_without_data"):                                          # PUSCH: CDM de-spreading
        cmb, scale = _cdm_operator(pilots, estimator._num_pilots_per_dmrs_sym, estimator._dmrs_length,
                                   2 * estimator._num_cdm_groups_without_data)
        w_h = np.einsum("arp,apq->arq", w, cmb)
    else:
        w_h = w
    pinv = np.where(np.abs(pilots) > 0, 1.0 / np.where(np.abs(pilots) > 0, pilots, 1.0), 0.0)     # divide_no_nan
    t_dense = w_h * pinv[:, None, :]                                                                # coefficient of y[pilot i]
    e_fac = np.maximum(scale * np.einsum("arp,ap->ar", w, e_ls), 0.0)                               # err_var / no, floored (:171)
    nz = np.abs(t_dense) > 1e-12 * np.abs(t_dense).max()
    nt = int(nz.sum(-1).max())
    if nt == 0 or nt > MAX_TERMS:
        return None
    eff = np.asarray(rg.effective_subcarrier_ind, np.int64)
    re_full = (np.arange(s_n)[:, None] * rg.fft_size + eff[None, :]).reshape(-1)
    pilot_re = np.asarray(estimator._pilot_ind).reshape(ts, num_p)                                 # effective-grid RE of pilot i
    t_idx = np.full((ts, num_re, nt), -1, np.int32)
    t_w = np.zeros((ts, num_re, nt), np.complex64)
    order = np.argsort(~nz, axis=-1, kind="stable")[..., :nt]                                      # non-zero columns first
    take = np.take_along_axis(nz, order, -1)
    cols_re = re_full[pilot_re]                                                                    # [ts, P] full-grid position
    for a in range(ts):
        idx = np.where(take[a], cols_re[a][order[a]], -1)
        t_idx[a] = idx
        t_w[a] = np.where(take[a], np.take_along_axis(t_dense[a], order[a], -1), 0)
    return {"t_idx": t_idx, "t_w": t_w, "e_sum": e_fac.sum(0).astype(np.float32), "re_full": re_full.astype(np.int32),
            "num_terms": nt}


def fusable(resource_grid, stream_management, estimator, constellation):
    """True if (estimator, LMMSE detection, demapping with `constellation`) can run as one `sb_ofdm_frontend` launch."""
    sm = stream_management
    if sm.num_interfering_streams_per_rx != 0 or not 1 <= sm.num_streams_per_rx <= 4:
        return False
    if constellation is not None:
        if separable_levels_np(constellation.points.numpy(), constellation.num_bits_per_symbol) is None:
            return False
    return frontend_tables(resource_grid, estimator) is not None


class FusedLSLinearDetector(Block):
    """FusedLSLinearDetector(channel_estimator, resource_grid, stream_management, demapping_method, constellation_type=None, num_bits_per_symbol=None, constellation=None, hard_out=False, precision=None)

    ``call(y, no)``: ``y [batch, num_rx, num_rx_ant, num_ofdm_symbols, fft_size]``, ``no`` as for `LSChannelEstimator` ->
    LLRs ``[batch, num_tx, num_streams_per_tx, num_data_symbols * num_bits_per_symbol]``: the output of
    ``LinearDetector("lmmse", "bit", demapping_method, ...)(y, *channel_estimator(y, no), no)``.
    ``equalize(y, no)`` returns ``(x_hat, no_eff)`` of the LMMSE equaliser instead (no demapping)."""

    def __init__(self, channel_estimator, resource_grid, stream_management, demapping_method="app", constellation_type=None,
                 num_bits_per_symbol=None, constellation=None, hard_out=False, precision=None, **kwargs):
        super().__init__(precision=precision, **kwargs)
        assert demapping_method in ("app", "maxlog"), "Unknown demapping method"
        self._rg, self._sm = resource_grid, stream_management
        self._method = 0 if demapping_method == "app" else 1
        self._hard_out = bool(hard_out)
        self._constellation = Constellation.check_or_create(constellation_type=constellation_type,
                                                            num_bits_per_symbol=num_bits_per_symbol,
                                                            constellation=constellation, precision=precision)
        lev = separable_levels_np(self._constellation.points.numpy(), self._constellation.num_bits_per_symbol)
        if lev is None:
            raise ValueError("FusedLSLinearDetector needs a separable (square QAM) constellation of at most 1024 points")
        if stream_management.num_interfering_streams_per_rx != 0 or not 1 <= stream_management.num_streams_per_rx <= 4:
            raise ValueError("FusedLSLinearDetector: no interfering streams and at most 4 streams per receiver")
        tabs = frontend_tables(resource_grid, channel_estimator)
        if tabs is None:
            raise ValueError("FusedLSLinearDetector: this estimator / pilot pattern cannot be fused")
        des, _, out_ts, data_pos = _sm_tables(resource_grid, stream_management)
        # Device layout: only the REs that carry data for at least one stream are listed (pilot-only OFDM symbols are not
        # walked), and the term tables are term-major [streams, terms, listed REs] so that a warp's table loads coalesce.
        data_pos = np.asarray(data_pos)
        keep = np.nonzero((data_pos >= 0).any(0))[0]
        self._np = dict(des=des, out_ts=out_ts, data_pos=np.ascontiguousarray(data_pos[:, keep]),
                        re_full=np.ascontiguousarray(tabs["re_full"][keep]), e_sum=np.ascontiguousarray(tabs["e_sum"][keep]),
                        t_idx=np.ascontiguousarray(tabs["t_idx"][:, keep, :].transpose(0, 2, 1)),
                        t_w=np.ascontiguousarray(tabs["t_w"][:, keep, :].transpose(0, 2, 1)))
        self._num_listed, self._num_terms = int(len(keep)), int(tabs["num_terms"])
        self._lev = (np.ascontiguousarray(lev[0], np.float32), np.ascontiguousarray(lev[1], np.float32))   # host arrays
        self._dev = None

    def _tables(self, dev):
        if self._dev is None or self._dev["des"].device != dev:
            self._dev = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in self._np.items()
                         if isinstance(v, np.ndarray)}
        return self._dev

    def _run(self, y, no, want_llr):
        if self.precision != "single":
            raise NotImplementedError("sb_ofdm_frontend is a complex64 kernel.")
        rg, sm = self._rg, self._sm
        dev = self.device
        y = y.to(device=dev, dtype=torch.complex64).contiguous()
        b, rx, ant, s_, nf = y.shape
        assert s_ == rg.num_ofdm_symbols and nf == rg.fft_size, "y must be the full resource grid"
        t = self._tables(dev)
        no_t = torch.as_tensor(no).to(device=dev, dtype=torch.float32)
        no_t = no_t.reshape(list(no_t.shape) + [1] * (3 - no_t.dim()))
        no_t, no_st = _strides_for(no_t, [b, rx, ant])
        txs = sm.num_tx * sm.num_streams_per_tx
        nd = rg.pilot_pattern.num_data_symbols
        m = self._constellation.num_bits_per_symbol
        llr = xh = ne = None
        if want_llr:
            llr = torch.zeros((b, sm.num_tx, sm.num_streams_per_tx, nd * m), dtype=torch.float32, device=dev)
        else:
            xh = torch.zeros((b, sm.num_tx, sm.num_streams_per_tx, nd), dtype=torch.complex64, device=dev)
            ne = torch.zeros((b, sm.num_tx, sm.num_streams_per_tx, nd), dtype=torch.float32, device=dev)
        no_arr = np.asarray(no_st, np.int64)
        check(lib().sb_ofdm_frontend(ptr(y), ptr(no_t), ptr(no_arr), ptr(t["des"]), ptr(t["out_ts"]), ptr(t["data_pos"]),
                                     ptr(t["re_full"]), ptr(t["t_idx"]), ptr(t["t_w"]), ptr(t["e_sum"]), ptr(self._lev[0]),
                                     ptr(self._lev[1]), ptr(llr), ptr(xh), ptr(ne), b, rx, ant, txs,
                                     self._num_listed, s_ * nf, sm.num_streams_per_rx, self._num_terms, nd, m // 2, self._method, int(self._hard_out),
                                     current_stream()), "sb_ofdm_frontend")
        return llr if want_llr else (xh, ne)

    def call(self, y, no):
        return self._run(y, no, True)

    def equalize(self, y, no):
        return self._run(y, no, False)
