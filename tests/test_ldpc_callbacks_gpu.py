"""Decoder plug points (reference decoding.py:79-126, 484-486, 513-515 and ldpc/utils.py:12-260):
`v2c_callbacks` / `c2v_callbacks` and callable node updates run the unfused path (csrc/ldpc_bp_flat.cu). Its kernels
execute the same rule code (ldpc_rules.cuh) in the reference's list orders, so with callbacks that do not change the
messages the result must equal the fused generic kernel with sum_order="reference" - and the oracle - bit for bit."""
import numpy as np
import pytest
import torch

from oracle import ldpc as O

pytestmark = pytest.mark.gpu
RULES = ["boxplus-phi", "boxplus", "minsum", "offset-minsum"]


def _llr(c, ebno_db, rate, rng):
    no = 1.0 / (10 ** (ebno_db / 10) * rate)
    y = (2.0 * c - 1.0) + rng.normal(size=c.shape) * np.sqrt(no / 2)
    return (4 * y / no).astype(np.float32)


class _Spy:
    """Records what the decoder hands to a callback and returns the messages unchanged."""

    def __init__(self):
        self.calls = []

    def __call__(self, msg, it, *args):
        self.calls.append((it, msg.shape, len(args), None if not args else tuple(args[0].shape)))
        return msg


@pytest.mark.parametrize("rule", RULES)
def test_noop_callbacks_equal_fused_reference_order_and_oracle(cuda_device, rule):
    from sionna_b200.phy.fec.ldpc import LDPC5GEncoder, LDPC5GDecoder
    rng = np.random.default_rng(31)
    k, n, bs, it = 400, 1000, 37, 8
    enc_r = O.LDPC5GEncoderRef(k, n)
    llr = _llr(enc_r(rng.integers(0, 2, (bs, k))), 1.0, k / n, rng)
    enc = LDPC5GEncoder(k, n)
    x = torch.from_numpy(llr).to(cuda_device)
    for hard, info in ((False, False), (True, True)):
        c2v, v2c = _Spy(), _Spy()
        dec = LDPC5GDecoder(enc, cn_update=rule, hard_out=hard, return_infobits=info, num_iter=it, return_state=True,
                            c2v_callbacks=[c2v], v2c_callbacks=[v2c])
        fused = LDPC5GDecoder(enc, cn_update=rule, hard_out=hard, return_infobits=info, num_iter=it, return_state=True,
                              sum_order="reference")
        y, st = dec(x)
        yf, stf = fused(x)
        assert torch.equal(y, yf) and torch.equal(st, stf)
        ref = O.LDPC5GDecoderRef(enc_r, cn_update=rule, hard_out=hard, return_infobits=info, num_iter=it, return_state=True)
        pure = rule in ("minsum", "offset-minsum")
        yr, sr = ref(llr, math_mode=0 if pure else 1, order="reference", pure=pure)
        assert np.array_equal(y.cpu().numpy(), yr) and np.array_equal(st.cpu().numpy(), sr)
        # call protocol: c2v callbacks see (msg [num_cns, None, B], it), v2c callbacks (msg [num_vns, None, B], it + 1, x_hat)
        assert [c[0] for c in c2v.calls] == list(range(it)) and [c[0] for c in v2c.calls] == list(range(1, it + 1))
        assert c2v.calls[0][1] == (dec.num_cns, None, bs) and c2v.calls[0][2] == 0
        assert v2c.calls[0][1] == (dec.num_vns, None, bs) and v2c.calls[0][3] == (dec.num_vns, bs)
    # state hand-over through the unfused path: 2 x 4 iterations == 8 iterations
    dec4 = LDPC5GDecoder(enc, cn_update=rule, hard_out=False, num_iter=4, return_state=True, c2v_callbacks=[_Spy()])
    y1, s1 = dec4(x)
    y2, s2 = dec4(x, msg_v2c=s1)
    y8, s8 = LDPC5GDecoder(enc, cn_update=rule, hard_out=False, num_iter=8, return_state=True, sum_order="reference")(x)
    assert torch.equal(y2, y8) and torch.equal(s2, s8)


def test_generic_pcm_layered_schedule_and_identity_rules(cuda_device):
    from sionna_b200.phy.fec.ldpc import LDPCBPDecoder
    from sionna_b200.phy.fec.utils import load_parity_check_examples
    pcm = load_parity_check_examples(3)[0].astype(np.float64)
    rng = np.random.default_rng(5)
    llr = (rng.normal(size=(21, pcm.shape[1])) * 2 + 1).astype(np.float32)
    x = torch.from_numpy(llr).to(cuda_device)
    sched = np.stack([np.arange(10) + 10 * i for i in range(pcm.shape[0] // 10)])
    for rule in ("boxplus-phi", "minsum"):
        spy = _Spy()
        dec = LDPCBPDecoder(pcm, cn_update=rule, cn_schedule=sched, hard_out=False, num_iter=3, c2v_callbacks=[spy])
        fused = LDPCBPDecoder(pcm, cn_update=rule, cn_schedule=sched, hard_out=False, num_iter=3, sum_order="reference")
        assert torch.equal(dec(x), fused(x))
        assert len(spy.calls) == 3 * sched.shape[0] and spy.calls[0][1] == (10, None, 21)      # only the active CNs
    y = LDPCBPDecoder(pcm, cn_update="identity", vn_update="identity", hard_out=False, num_iter=2, v2c_callbacks=[_Spy()])(x)
    yf = LDPCBPDecoder(pcm, cn_update="identity", vn_update="identity", hard_out=False, num_iter=2)(x)
    assert torch.equal(y, yf)


def test_statistics_exit_and_weighted_bp_callbacks(cuda_device):
    """The reference's three callbacks (ldpc/utils.py): decoder statistics (all checks satisfied per iteration), EXIT
    mutual information (all-zero codeword), weighted BP (unit weights = plain BP; weights < 1 damp the messages)."""
    from sionna_b200.phy.fec.ldpc import (LDPC5GEncoder, LDPC5GDecoder, LDPCBPDecoder, DecoderStatisticsCallback,
                                          EXITCallback, WeightedBPCallback)
    from sionna_b200.phy.fec.utils import GaussianPriorSource, load_parity_check_examples
    from sionna_b200.phy import config
    config.seed = 8
    # The statistics callback counts codewords whose check nodes all see an even number of negative messages. In a 5G
    # graph a degree-1 parity VN keeps sending its channel LLR, so one wrong-sign channel observation keeps its check
    # "unsatisfied" for ever although the information bits decode: the callback is meaningful on graphs without
    # degree-1 VNs, e.g. the 802.11n code of the reference's example set (same remark applies to the reference).
    pcm = load_parity_check_examples(4)[0].astype(np.float64)
    n, bs, it = pcm.shape[1], 300, 12
    llr = GaussianPriorSource()([bs, n], no=0.5)                      # all-zero codeword, logits ~ N(-2/no, 4/no)
    stats, exit_c, exit_v = DecoderStatisticsCallback(it), EXITCallback(it), EXITCallback(it)
    dec = LDPCBPDecoder(pcm, num_iter=it, hard_out=True, c2v_callbacks=[stats, exit_c], v2c_callbacks=[exit_v])
    u_hat = dec(llr)
    assert stats.num_samples.tolist() == [bs] * it
    succ = stats.num_decoded_cws.numpy()
    assert succ[-1] >= 0.8 * bs and succ[0] < succ[-1] and np.all(np.diff(succ) >= -3)
    assert 0.0 < float(stats.avg_number_iterations) < it
    # a codeword whose checks are all satisfied at the end decodes to the all-zero word (or another codeword: rare)
    assert float((u_hat != 0).any(dim=-1).float().mean()) <= 1 - succ[-1] / bs + 0.02
    mi_c, mi_v = exit_c.mi.numpy()[:it], exit_v.mi.numpy()[1:it + 1]
    assert mi_c[-1] > 0.9 and mi_v[-1] > 0.95 and mi_c[-1] > mi_c[0] and mi_v[-1] > mi_v[0]
    k, n = 500, 1000
    enc = LDPC5GEncoder(k, n)
    llr = GaussianPriorSource()([64, n], no=0.55)
    dec = LDPC5GDecoder(enc, num_iter=5, hard_out=False)
    # weighted BP: unit weights reproduce plain BP bit for bit; damping changes the soft outputs
    plain = LDPC5GDecoder(enc, num_iter=5, hard_out=False, sum_order="reference")(llr)
    wcb = WeightedBPCallback(dec.num_edges)
    w1 = LDPC5GDecoder(enc, num_iter=5, hard_out=False, v2c_callbacks=[wcb], c2v_callbacks=[wcb])(llr)
    assert torch.equal(plain, w1)
    wcb.weights.mul_(0.8)
    w2 = LDPC5GDecoder(enc, num_iter=5, hard_out=False, v2c_callbacks=[wcb])(llr)
    assert not torch.equal(plain, w2) and torch.isfinite(w2).all()


def test_callable_node_updates(cuda_device):
    """User-supplied node updates on ragged messages (decoding.py:79-92): a Python restatement of the sum VN update and
    of the min-sum CN update, written with RaggedMessages operations, reproduce the built-in rules."""
    from sionna_b200.phy.fec.ldpc import LDPC5GEncoder, LDPC5GDecoder
    rng = np.random.default_rng(77)
    k, n, bs = 300, 600, 25
    enc_r = O.LDPC5GEncoderRef(k, n)
    llr = _llr(enc_r(rng.integers(0, 2, (bs, k))), 2.0, k / n, rng)
    x = torch.from_numpy(llr).to(cuda_device)
    enc = LDPC5GEncoder(k, n)

    def vn_sum(msg, llr_ch, llr_clipping):                               # vn_update_sum, decoding.py:714-732
        x_tot = msg.reduce_sum() + llr_ch
        out = msg.gather_rows(x_tot) - msg.flat_values
        return msg.with_flat_values(out.clamp(-llr_clipping, llr_clipping)), x_tot.clamp(-llr_clipping, llr_clipping)

    def cn_minsum(msg, llr_clipping):                                    # cn_update_minsum, decoding.py:911-953
        v = msg.flat_values
        sgn = torch.where(v < 0, -torch.ones_like(v), torch.ones_like(v))
        node_sgn = msg.with_flat_values(sgn).reduce_prod()
        a = v.abs()
        ra = msg.with_flat_values(a)
        m1 = ra.reduce_min()
        is_min = a == ra.gather_rows(m1)
        m2 = msg.with_flat_values(torch.where(is_min, torch.full_like(a, float("inf")), a)).reduce_min()
        n_min = msg.with_flat_values(is_min.float()).reduce_sum()
        m2 = torch.where(n_min > 1, m1, m2)                              # duplicated minimum -> everybody gets m1
        mag = torch.where(is_min, ra.gather_rows(m2), ra.gather_rows(m1))
        return msg.with_flat_values((sgn * ra.gather_rows(node_sgn) * mag).clamp(-llr_clipping, llr_clipping))

    ref = LDPC5GDecoder(enc, cn_update="minsum", hard_out=False, num_iter=6, sum_order="reference")(x)
    got_v = LDPC5GDecoder(enc, cn_update="minsum", vn_update=vn_sum, hard_out=False, num_iter=6)(x)
    assert torch.allclose(got_v, ref, rtol=1e-5, atol=1e-4)             # torch sums in a different order
    got_c = LDPC5GDecoder(enc, cn_update=cn_minsum, hard_out=False, num_iter=6)(x)
    assert torch.allclose(got_c, ref, rtol=1e-5, atol=1e-4)
    got_both = LDPC5GDecoder(enc, cn_update=cn_minsum, vn_update=vn_sum, hard_out=True, num_iter=6)(x)
    assert torch.equal(got_both, LDPC5GDecoder(enc, cn_update="minsum", hard_out=True, num_iter=6)(x))


def test_ragged_messages_reductions(cuda_device):
    from sionna_b200.phy.fec.ldpc import RaggedMessages
    rng = np.random.default_rng(1)
    lens = np.array([3, 1, 5, 2, 4])
    splits = np.concatenate([[0], np.cumsum(lens)])
    v = rng.normal(size=(lens.sum(), 6)).astype(np.float32)
    r = RaggedMessages(torch.from_numpy(v).to(cuda_device), torch.from_numpy(splits).to(cuda_device))
    seg = [v[splits[i]:splits[i + 1]] for i in range(len(lens))]
    assert r.shape == (5, None, 6)
    np.testing.assert_allclose(r.reduce_sum().cpu().numpy(), np.stack([s.sum(0) for s in seg]), rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(r.reduce_prod().cpu().numpy(), np.stack([s.prod(0) for s in seg]), rtol=1e-5)
    np.testing.assert_array_equal(r.reduce_min().cpu().numpy(), np.stack([s.min(0) for s in seg]))
    np.testing.assert_array_equal(r.reduce_max().cpu().numpy(), np.stack([s.max(0) for s in seg]))
    np.testing.assert_array_equal(r.gather_rows(r.reduce_max()).cpu().numpy(), np.concatenate([np.repeat(s.max(0)[None], len(s), 0) for s in seg]))
    assert r.value_rowids().tolist() == np.repeat(np.arange(5), lens).tolist()
