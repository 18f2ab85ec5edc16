"""Link-level helpers and the Monte-Carlo driver (mirror of /root/reference/src/sionna/phy/utils/misc.py).

``sim_ber`` keeps the reference's stopping rules, status codes, progress table and return values
(misc.py:329-860). Its multi-device mode replaces ``tf.distribute.MirroredStrategy`` + ``strategy.gather`` of the
full bit tensors (misc.py:540-548, 614-655) by the B200 design of SURVEY.md section 8(e): one process per GPU
(``torchrun``), every rank runs ``mc_fun`` on its own random stream, and after every batch ONE all-reduce (NCCL over
NVLink; 32 bytes) sums the device-resident int64 counters {bit errors, block errors, bits, blocks}; all ranks
therefore take identical stopping decisions and ``max_mc_iter`` is divided by the number of replicas exactly as the
reference does (misc.py:651-655). The host never stalls the GPU for a stopping decision: counters reach pinned host
memory through asynchronous copies and the next batch is already enqueued when they are read (see the three batch
modes in ``sim_ber``).
"""
import time
import numpy as np
import torch

from ..config import config, dtypes
from ..._lib import lib, check, ptr, current_stream
from .metrics import ErrorCounter


def complex_normal(shape, var=1.0, precision=None):
    """Complex normal tensor with total variance ``var`` (``var/2`` per real dimension), misc.py:19-54.
    Drawn on the device by ``sb_awgn`` (Philox4x32-10 + Box-Muller) from zeros."""
    from ..block import fallback_to_single
    if fallback_to_single("complex_normal", precision):
        return complex_normal(shape, var, "single").to(torch.complex128)
    shape = [int(s) for s in shape]
    dev = config.device
    n = int(np.prod(shape)) if len(shape) else 1
    x = torch.zeros(shape, dtype=torch.complex64, device=dev)
    no = torch.as_tensor(var, dtype=torch.float32, device=dev).reshape(1)
    seed, off = config.next_philox()
    check(lib().sb_awgn(ptr(x), ptr(no), max(n, 1), ptr(x), n, seed, off, current_stream()), "sb_awgn")
    return x


def db_to_lin(x, precision=None):
    return 10.0 ** (torch.as_tensor(x, dtype=torch.float64) / 10.0)


def lin_to_db(x, precision=None):
    return 10.0 * torch.log10(torch.as_tensor(x, dtype=torch.float64))


def ebnodb2no(ebno_db, num_bits_per_symbol, coderate, resource_grid=None, precision=None):
    r"""Noise variance :math:`N_o` for a given :math:`E_b/N_o` in dB (misc.py:171-251):
    ``no = 1 / (10^(ebno_db/10) * coderate * num_bits_per_symbol / E_s)`` with ``E_s = 1`` or, with a resource grid,
    the per-stream energy corrected for cyclic prefix and pilot overhead (misc.py:233-247). Computed in the block
    precision with the same operation order as the reference."""
    if precision is None:
        precision = config.precision
    rd = dtypes[precision]["np"]["rdtype"]
    is_tensor = isinstance(ebno_db, torch.Tensor)
    e = ebno_db.detach().cpu().numpy().astype(rd) if is_tensor else np.asarray(ebno_db, dtype=rd)
    ten = rd(10)
    ebno = np.power(ten, e / ten).astype(rd)
    energy_per_symbol = 1.0
    if resource_grid is not None:
        energy_per_symbol /= resource_grid.num_streams_per_tx
        cp_overhead = resource_grid.cyclic_prefix_length / resource_grid.fft_size
        num_syms = resource_grid.num_ofdm_symbols * (1 + cp_overhead) * resource_grid.num_effective_subcarriers
        energy_per_symbol *= num_syms / resource_grid.num_data_symbols
    no = (rd(1) / (ebno * rd(coderate) * rd(num_bits_per_symbol) / rd(energy_per_symbol))).astype(rd)
    if is_tensor:
        return torch.from_numpy(np.asarray(no)).to(ebno_db.device)
    return torch.as_tensor(no)


def hard_decisions(llr):
    """1 where ``llr > 0`` else 0, same dtype (misc.py:254-271; note 0 maps to 0, unlike the decoder)."""
    llr = torch.as_tensor(llr)
    return (llr > 0).to(llr.dtype)


# status codes of sim_ber (misc.py:469-476)
STATUS_NA, STATUS_MAX_IT, STATUS_NO_ERR, STATUS_TARGET_BIT, STATUS_TARGET_BLOCK = 0, 1, 2, 3, 4
STATUS_TARGET_BER, STATUS_TARGET_BLER, STATUS_CB_STOP = 5, 6, 7


def _count_batch(counter, b, b_hat):
    """Add the error counts of one batch to ``counter`` ([4] int64 on the tensors' device)."""
    if b.is_cuda:
        b2 = b.to(torch.float32).reshape(-1, b.shape[-1]).contiguous()
        h2 = b_hat.to(torch.float32).reshape(-1, b.shape[-1]).contiguous()
        check(lib().sb_count_errors(ptr(b2), ptr(h2), b2.shape[0], b2.shape[1], ptr(counter), current_stream()),
              "sb_count_errors")
    else:
        # host tensors only reach this driver from user-supplied mc_fun (and the gloo driver tests); the blocks of
        # this package always produce CUDA tensors
        e = b != b_hat.to(b.dtype)
        counter += torch.stack([e.sum(), e.reshape(-1, b.shape[-1]).any(dim=-1).sum(),
                                torch.tensor(b.numel()), torch.tensor(b.numel() // b.shape[-1])]).to(torch.int64)


def sim_ber(mc_fun, ebno_dbs, batch_size, max_mc_iter, soft_estimates=False, num_target_bit_errors=None,
            num_target_block_errors=None, target_ber=None, target_bler=None, early_stop=True, graph_mode=None,
            distribute=None, verbose=True, forward_keyboard_interrupt=True, callback=None, precision=None):
    # pylint: disable=line-too-long
    r"""Monte-Carlo BER/BLER simulation (reference: misc.py:329-860).

    ``mc_fun(batch_size=..., ebno_db=...)`` must return ``(b, b_hat, ...)``. Stopping rules are evaluated after every
    batch in the reference's order: callback, target bit errors, target block errors, max iterations; after each
    SNR point: ``early_stop`` on zero block errors, ``target_ber``, ``target_bler``.

    ``graph_mode`` is accepted for signature compatibility ("graph"/"xla" have no meaning without a tracing
    compiler; kernels are already fused). ``distribute``: `None` (single device), ``"all"`` or a
    ``torch.distributed`` process group: the calling processes (one per GPU, e.g. launched by ``torchrun``) act as
    the replicas; each keeps its own random stream and only the four int64 counters are all-reduced per batch.
    Returns ``(ber, bler)`` tensors of shape ``[len(ebno_dbs)]``.
    """
    if precision is None:
        precision = config.precision
    rdtype = dtypes[precision]["torch"]["rdtype"]
    status_levels = {STATUS_NA: "not simulated", STATUS_MAX_IT: "reached max iterations",
                     STATUS_NO_ERR: "no errors - early stop", STATUS_TARGET_BIT: "reached target bit errors",
                     STATUS_TARGET_BLOCK: "reached target block errors",
                     STATUS_TARGET_BER: "reached target BER - early stop",
                     STATUS_TARGET_BLER: "reached target BLER - early stop",
                     STATUS_CB_STOP: "callback triggered stopping"}
    if not isinstance(early_stop, bool):
        raise TypeError("early_stop must be bool.")
    if not isinstance(soft_estimates, bool):
        raise TypeError("soft_estimates must be bool.")
    if not isinstance(verbose, bool):
        raise TypeError("verbose must be bool.")
    if target_ber is not None:
        if not early_stop:
            print("Warning: early stop is deactivated. target_ber is ignored.")
    else:
        target_ber = -1.
    if target_bler is not None:
        if not early_stop:
            print("Warning: early stop is deactivated. target_bler is ignored.")
    else:
        target_bler = -1.
    if graph_mode is None:
        graph_mode = "default"
    if not isinstance(graph_mode, str):
        raise TypeError("graph_mode must be str.")
    if graph_mode not in ("default", "graph", "xla"):
        raise TypeError("Unknown graph_mode selected.")

    # ---- replicas: one process per GPU, counters all-reduced ---------------------------------------------------
    import torch.distributed as dist
    group, num_replicas = None, 1
    if distribute is not None:
        if not (dist.is_available() and dist.is_initialized()):
            if distribute == "all" or isinstance(distribute, (tuple, list)):
                distribute = None   # single process: nothing to distribute over (reference: no GPUs -> None)
            else:
                raise ValueError("Unknown value for distribute.")
        else:
            if distribute == "all" or isinstance(distribute, (tuple, list)):
                group = dist.group.WORLD
            elif isinstance(distribute, dist.ProcessGroup):
                group = distribute
            else:
                raise ValueError("Unknown value for distribute.")
            num_replicas = dist.get_world_size(group)
    run_multigpu = group is not None and num_replicas > 1
    is_rank0 = (not run_multigpu) or dist.get_rank(group) == 0
    verbose = verbose and is_rank0
    if run_multigpu:
        max_mc_iter = int(np.ceil(max_mc_iter / num_replicas))
        config.rank_offset = dist.get_rank(group) + 1
        if verbose:
            print(f"Distributing simulation across {num_replicas} devices.")
            print(f"Reducing max_mc_iter to {max_mc_iter}")

    ebno_np = np.atleast_1d(np.asarray(ebno_dbs.detach().cpu() if isinstance(ebno_dbs, torch.Tensor) else ebno_dbs,
                                       dtype=np.float64))
    num_points = len(ebno_np)
    bit_errors = np.zeros(num_points, np.int64)
    block_errors = np.zeros(num_points, np.int64)
    nb_bits = np.zeros(num_points, np.int64)
    nb_blocks = np.zeros(num_points, np.int64)
    status = np.zeros(num_points)
    runtime = np.zeros(num_points)
    header_text = ["EbNo [dB]", "BER", "BLER", "bit errors", "num bits", "block errors", "num blocks", "runtime [s]",
                   "status"]

    def _print_progress(is_final, rt, idx_snr, idx_it, header=None):
        end_str = "\n" if is_final else "\r"
        if header is not None:
            row_text, end_str = header, "\n"
        else:
            with np.errstate(divide="ignore", invalid="ignore"):
                ber_np = np.nan_to_num(bit_errors[idx_snr] / nb_bits[idx_snr])
                bler_np = np.nan_to_num(block_errors[idx_snr] / nb_blocks[idx_snr])
            if status[idx_snr] == STATUS_NA:
                status_txt = f"iter: {idx_it:.0f}/{max_mc_iter:.0f}"
            else:
                status_txt = status_levels[int(status[idx_snr])]
            row_text = [str(np.round(ebno_np[idx_snr], 3)), f"{ber_np:.4e}", f"{bler_np:.4e}", bit_errors[idx_snr],
                        nb_bits[idx_snr], block_errors[idx_snr], nb_blocks[idx_snr], np.round(rt, 1), status_txt]
        print("{: >9} |{: >11} |{: >11} |{: >12} |{: >12} |{: >13} |{: >12} |{: >12} |{: >10}".format(*row_text),
              end=end_str)

    # ---- batch execution ---------------------------------------------------------------------------------------
    # The counters of a batch live on the device (sb_count_errors); what differs between the modes is WHEN the host
    # reads them:
    #   "sync"  callback given, or mc_fun returns host tensors: read after every batch before anything else happens
    #           (the callback contract; exactly the reference's sequence of mc_fun calls).
    #   "lag"   CUDA tensors + a target number of bit / block errors: batch ii+1 is enqueued speculatively BEFORE the
    #           counters of batch ii are read (asynchronous copy to pinned memory + event), so the GPU never waits for
    #           the host's stopping decision; when batch ii reaches the target the speculative batch is discarded, i.e.
    #           the counted batches - and therefore BER/BLER - are those of the reference's rule (misc.py:742-753).
    #   "free"  CUDA tensors and no per-batch stopping rule: nothing is read until the SNR point is finished (running
    #           totals are copied asynchronously only to feed the progress line).
    slots = {}

    def _slots(device):
        if device not in slots:
            st = {"dev": [torch.zeros(4, dtype=torch.int64, device=device) for _ in range(2)]}
            if device.type == "cuda":
                st["host"] = [torch.zeros(4, dtype=torch.int64).pin_memory() for _ in range(2)]
                st["ev"] = [torch.cuda.Event() for _ in range(2)]
                st["run"] = torch.zeros(4, dtype=torch.int64, device=device)
            slots[device] = st
        return slots[device]

    def _mc(i):
        ebno_i = torch.as_tensor(ebno_np[i], dtype=rdtype)
        outputs = mc_fun(batch_size=batch_size, ebno_db=ebno_i)
        b, b_hat = outputs[0], outputs[1]
        if soft_estimates:
            b_hat = hard_decisions(b_hat)
        return b, b_hat

    def _post(st, k, src=None):
        """Reduce slot k over the replicas and start its copy to pinned host memory."""
        if src is not None:
            st["dev"][k].copy_(src)
        if run_multigpu:
            dist.all_reduce(st["dev"][k], op=dist.ReduceOp.SUM, group=group)
        st["host"][k].copy_(st["dev"][k], non_blocking=True)
        st["ev"][k].record()

    def _add(i, c):
        bit_errors[i] += c[0]
        block_errors[i] += c[1]
        nb_bits[i] += c[2]
        nb_blocks[i] += c[3]

    def _header_once(i, iter_count):
        if verbose and i == 0 and iter_count == 0:
            _print_progress(True, 0, 0, 0, header_text)
            print("-" * 135)

    def _targets_reached(i):
        if num_target_bit_errors is not None and bit_errors[i] >= num_target_bit_errors:
            return STATUS_TARGET_BIT
        if num_target_block_errors is not None and block_errors[i] >= num_target_block_errors:
            return STATUS_TARGET_BLOCK
        return None

    per_batch_rule = num_target_bit_errors is not None or num_target_block_errors is not None
    i = 0
    cb_state = sim_ber.CALLBACK_CONTINUE
    try:
        for i in range(num_points):
            runtime[i] = time.perf_counter()
            iter_count = -1
            mode = "none"
            if max_mc_iter >= 1:
                b, b_hat = _mc(i)                                        # batch 0 of this SNR point
                st = _slots(b.device)
                mode = "sync" if (callback is not None or not b.is_cuda) else ("lag" if per_batch_rule else "free")
            if mode == "free":
                st["run"].zero_()
                last = None
                for ii in range(max_mc_iter):
                    iter_count = ii
                    if ii > 0:
                        b, b_hat = _mc(i)
                    _count_batch(st["run"], b, b_hat)                    # running totals of this SNR point
                    k = ii & 1
                    if verbose and last is not None and st["ev"][last].query():
                        c = st["host"][last].tolist()                     # newest totals that have already arrived
                        bit_errors[i], block_errors[i], nb_bits[i], nb_blocks[i] = c
                    _post(st, k, st["run"])
                    last = k
                    _header_once(i, ii)
                    if verbose:
                        _print_progress(False, time.perf_counter() - runtime[i], i, ii)
                st["ev"][last].synchronize()
                c = st["host"][last].tolist()
                bit_errors[i], block_errors[i], nb_bits[i], nb_blocks[i] = c
                runtime[i] = time.perf_counter() - runtime[i]
                status[i] = STATUS_MAX_IT
            elif mode == "lag":
                k = 0
                st["dev"][k].zero_()
                _count_batch(st["dev"][k], b, b_hat)
                _post(st, k)
                ii = 0
                while True:
                    iter_count = ii
                    speculative = ii + 1 < max_mc_iter
                    if speculative:                                       # enqueue batch ii+1 before reading batch ii
                        b, b_hat = _mc(i)
                        st["dev"][1 - k].zero_()
                        _count_batch(st["dev"][1 - k], b, b_hat)
                        _post(st, 1 - k)
                    st["ev"][k].synchronize()
                    _add(i, st["host"][k].tolist())
                    _header_once(i, ii)
                    if verbose:
                        _print_progress(False, time.perf_counter() - runtime[i], i, ii)
                    reached = _targets_reached(i)
                    if reached is not None:                               # a speculative batch, if any, is not counted
                        status[i] = reached
                        runtime[i] = time.perf_counter() - runtime[i]
                        break
                    if not speculative:
                        runtime[i] = time.perf_counter() - runtime[i]
                        status[i] = STATUS_MAX_IT
                        break
                    ii += 1
                    k = 1 - k
            elif mode == "sync":
                for ii in range(max_mc_iter):
                    iter_count = ii
                    if ii > 0:
                        b, b_hat = _mc(i)
                    counter = st["dev"][0]
                    counter.zero_()
                    _count_batch(counter, b, b_hat)
                    if run_multigpu:
                        dist.all_reduce(counter, op=dist.ReduceOp.SUM, group=group)
                    _add(i, counter.cpu().tolist())

                    cb_state = sim_ber.CALLBACK_CONTINUE
                    if callback is not None:
                        cb_state = callback(ii, i, ebno_np, bit_errors, block_errors, nb_bits, nb_blocks)
                        if cb_state in (sim_ber.CALLBACK_STOP, sim_ber.CALLBACK_NEXT_SNR):
                            runtime[i] = time.perf_counter() - runtime[i]
                            status[i] = STATUS_CB_STOP
                            break
                    _header_once(i, ii)
                    if verbose:
                        _print_progress(False, time.perf_counter() - runtime[i], i, ii)
                    reached = _targets_reached(i)
                    if reached is not None:
                        status[i] = reached
                        runtime[i] = time.perf_counter() - runtime[i]
                        break
                    if ii == max_mc_iter - 1:
                        runtime[i] = time.perf_counter() - runtime[i]
                        status[i] = STATUS_MAX_IT
            if verbose:
                _print_progress(True, runtime[i], i, iter_count)
            if early_stop:
                if block_errors[i] == 0:
                    status[i] = STATUS_NO_ERR
                    if verbose:
                        print(f"\nSimulation stopped as no error occurred @ EbNo = {ebno_np[i]:.1f} dB.\n")
                    break
                ber_true = bit_errors[i] / nb_bits[i]
                bler_true = block_errors[i] / nb_blocks[i]
                if ber_true < target_ber:
                    status[i] = STATUS_TARGET_BER
                    if verbose:
                        print(f"\nSimulation stopped as target BER is reached@ EbNo = {ebno_np[i]:.1f} dB.\n")
                    break
                if bler_true < target_bler:
                    status[i] = STATUS_TARGET_BLER
                    if verbose:
                        print(f"\nSimulation stopped as target BLER is reached @ EbNo = {ebno_np[i]:.1f} dB.\n")
                    break
            if cb_state is sim_ber.CALLBACK_STOP:
                status[i] = STATUS_CB_STOP
                if verbose:
                    print(f"\nSimulation stopped by callback function @ EbNo = {ebno_np[i]:.1f} dB.\n")
                break
    except KeyboardInterrupt as e:
        if forward_keyboard_interrupt:
            raise e
        print(f"\nSimulation stopped by the user @ EbNo = {ebno_np[i]} dB.")
        for idx in range(i + 1, num_points):
            bit_errors[idx] += -1
            block_errors[idx] += -1
            nb_bits[idx] += 1
            nb_blocks[idx] += 1

    with np.errstate(divide="ignore", invalid="ignore"):
        ber = np.nan_to_num(bit_errors.astype(np.float64) / nb_bits.astype(np.float64), nan=0.0)
        bler = np.nan_to_num(block_errors.astype(np.float64) / nb_blocks.astype(np.float64), nan=0.0)
    return torch.as_tensor(ber).to(rdtype), torch.as_tensor(bler).to(rdtype)


sim_ber.CALLBACK_CONTINUE = None
sim_ber.CALLBACK_STOP = 2
sim_ber.CALLBACK_NEXT_SNR = 1
