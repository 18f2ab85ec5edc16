"""N > 1 path of the Monte-Carlo driver on CPU: two gloo ranks, per-rank random streams, ONE all-reduce of the four
int64 counters per batch, identical stopping decisions on every rank, max_mc_iter divided by the replica count
(reference semantics: /root/reference/src/sionna/phy/utils/misc.py:614-655; its test: test/unit/utils/test_utils.py:80-128)."""
import os
import socket
import sys
import numpy as np
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from sionna_b200.phy.utils import sim_ber
    calls = []

    def mc_fun(batch_size, ebno_db):
        # BPSK over AWGN decided on the host; every rank owns a different stream
        g = torch.Generator().manual_seed(1000 + 17 * rank + len(calls))
        calls.append(float(ebno_db))
        no = 1.0 / (10 ** (float(ebno_db) / 10))
        b = torch.randint(0, 2, (batch_size, 50), generator=g).float()
        y = (2 * b - 1) + torch.randn(b.shape, generator=g) * np.sqrt(no / 2)
        return b, (y > 0).float()

    ber, bler = sim_ber(mc_fun, [0.0, 4.0], batch_size=200, max_mc_iter=8, num_target_block_errors=10 ** 9,
                        early_stop=False, distribute="all", verbose=False)
    q.put((rank, ber.tolist(), bler.tolist(), len(calls)))
    dist.destroy_process_group()


def test_sim_ber_two_replicas_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, ber0, bler0, n0), (_, ber1, bler1, n1) = res
    assert ber0 == ber1 and bler0 == bler1                  # identical reduced statistics on every rank
    assert n0 == n1 == 2 * 4                                # max_mc_iter 8 -> 4 per replica, two SNR points
    from scipy.special import erfc
    for e_db, ber in zip((0.0, 4.0), ber0):
        theory = 0.5 * erfc(np.sqrt(10 ** (e_db / 10)))
        assert abs(ber - theory) < 0.25 * theory + 1e-3     # 2 ranks x 4 iters x 200 x 50 bits


def test_sim_ber_single_process_stopping_rules():
    sys.path.insert(0, ROOT)
    from sionna_b200.phy.utils import sim_ber
    n_calls = []

    def mc_fun(batch_size, ebno_db):
        n_calls.append(1)
        b = torch.zeros(batch_size, 10)
        b_hat = b.clone()
        if float(ebno_db) < 1.0:
            b_hat[0, 0] = 1.0                                # exactly one bit / block error per batch
        return b, b_hat

    ber, bler = sim_ber(mc_fun, [0.0, 2.0, 4.0], batch_size=5, max_mc_iter=10, num_target_bit_errors=3, verbose=False)
    assert len(n_calls) == 3 + 10                            # 3 batches reach the target, then 10 error-free -> early stop
    assert ber[0] == 3 / (3 * 50) and bler[0] == 3 / 15 and ber[1] == 0 and ber[2] == 0
    # callback-driven stop
    def cb(mc_iter, snr_idx, *stats):
        return sim_ber.CALLBACK_NEXT_SNR if mc_iter == 1 else sim_ber.CALLBACK_CONTINUE
    n_calls.clear()
    sim_ber(mc_fun, [0.0], batch_size=5, max_mc_iter=10, callback=cb, verbose=False)
    assert len(n_calls) == 2
