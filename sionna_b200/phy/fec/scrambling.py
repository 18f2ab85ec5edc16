"""5G NR transport-block scrambling (mirror of /root/reference/src/sionna/phy/fec/scrambling.py:236-468 and
nr/utils.py:16-76), SURVEY.md section 8(f2)."""
import numpy as np
import torch

from ..block import Block
from ..._lib import lib, check, ptr, current_stream


def generate_prng_seq(length, c_init):
    """Length-31 Gold sequence of 38.211 5.2.1 with N_c = 1600 (nr/utils.py:16-76): x1(n+31) = x1(n+3) + x1(n),
    x2(n+31) = x2(n+3) + x2(n+2) + x2(n+1) + x2(n), c(n) = x1(n+1600) + x2(n+1600) mod 2."""
    assert length % 1 == 0 and int(length) > 0, "length must be a positive integer."
    assert c_init % 1 == 0 and 0 <= int(c_init) < 2 ** 32, "c_init must be in [0, 2^32-1]."
    length, c_init = int(length), int(c_init)
    n_c, total = 1600, int(length) + 1600 + 31
    x1 = np.zeros(total, np.uint8)
    x2 = np.zeros(total, np.uint8)
    x1[0] = 1
    x2[:31] = [(c_init >> i) & 1 for i in range(31)]
    for i in range(length + n_c):
        x1[i + 31] = x1[i + 3] ^ x1[i]
        x2[i + 31] = x2[i + 3] ^ x2[i + 2] ^ x2[i + 1] ^ x2[i]
    return (x1[n_c:n_c + length] ^ x2[n_c:n_c + length]).astype(np.float64)


class TB5GScrambler(Block):
    """TB5GScrambler(n_rnti=1, n_id=1, binary=True, channel_type="PUSCH", codeword_index=0): pseudo-random bit scrambling
    of 38.211 6.3.1.1 / 7.3.1.1; ``call(x, binary=None)`` flips bits (``binary``) or signs of soft values. Lists of
    ``n_rnti`` / ``n_id`` scramble axis -2 stream by stream (scrambling.py:236-468). Scrambling twice restores the input."""

    def __init__(self, n_rnti=1, n_id=1, binary=True, channel_type="PUSCH", codeword_index=0, precision=None, **kwargs):
        super().__init__(precision=precision, **kwargs)
        if not isinstance(binary, bool):
            raise TypeError("binary must be bool.")
        self._binary = binary
        if channel_type not in ("PDSCH", "PUSCH"):
            raise TypeError("Unsupported channel_type.")
        if codeword_index not in (0, 1):
            raise ValueError("codeword_index must be 0 or 1.")
        if isinstance(n_rnti, (list, tuple)):
            if not isinstance(n_id, (list, tuple)):
                raise TypeError("n_id must be a list of same length as n_rnti.")
            if len(n_rnti) != len(n_id):
                raise ValueError("n_rnti and n_id must be of same length.")
            self._multi_stream = True
            n_rnti, n_id = list(n_rnti), list(n_id)
        else:
            n_rnti, n_id, self._multi_stream = [n_rnti], [n_id], False
        for nr, ni in zip(n_rnti, n_id):
            if nr % 1 != 0 or int(nr) not in range(2 ** 16):
                raise ValueError("n_rnti must be in [0, 65535].")
            if ni % 1 != 0 or int(ni) not in range(2 ** 10):
                raise ValueError("n_id must be in [0, 1023].")
        if channel_type == "PUSCH":
            self._c_init = [int(nr) * 2 ** 15 + int(ni) for nr, ni in zip(n_rnti, n_id)]
        else:
            self._c_init = [int(nr) * 2 ** 15 + codeword_index * 2 ** 14 + int(ni) for nr, ni in zip(n_rnti, n_id)]
        self._n = None
        self._seq = None

    @property
    def keep_state(self):
        return True

    def build(self, input_shape, **kwargs):
        if self._multi_stream:
            assert input_shape[-2] == len(self._c_init), "Dimension of axis=-2 must be equal to len(n_rnti)."
        self._n = input_shape[-1]
        self._seq = None

    def call(self, x, /, *, binary=None):
        if binary is None:
            binary = self._binary
        elif not isinstance(binary, bool):
            raise TypeError("binary must be bool.")
        if x.shape[-1] != self._n:
            self.build(x.shape)
        dev = self.device
        if self._seq is None or self._seq.device != dev:
            seq = np.stack([generate_prng_seq(self._n, c) for c in self._c_init]).astype(np.float32)
            self._seq = torch.from_numpy(seq).to(dev)
        xin = x.to(device=dev, dtype=torch.float32).contiguous()
        out = torch.empty_like(xin)
        check(lib().sb_scramble(ptr(xin), ptr(self._seq), int(binary), ptr(out), xin.numel() // self._n, self._n,
                                len(self._c_init), current_stream()), "sb_scramble")
        return out.to(x.dtype) if x.dtype.is_floating_point else out


class Scrambler(Block):
    """Scrambler(seed=None, keep_batch_constant=False, binary=True, sequence=None, keep_state=True, precision=None)

    Pseudo-random flipping of bits (``binary``) or of the signs of soft values (scrambling.py:20-262). ``call(x,
    seed=None, binary=None)``: the sequence is drawn on the device from the block's seed (kept constant over calls when
    ``keep_state``), from the ``seed`` given to the call, or from a fresh random seed; ``sequence`` overrides all of them;
    ``keep_batch_constant`` uses one sequence for every batch sample. Scrambling twice with the same seed restores x."""

    def __init__(self, seed=None, keep_batch_constant=False, binary=True, sequence=None, keep_state=True, precision=None,
                 **kwargs):
        super().__init__(precision=precision, **kwargs)
        if not isinstance(keep_batch_constant, bool):
            raise TypeError("keep_batch_constant must be bool.")
        if seed is not None:
            if sequence is not None:
                print("Note: explicit scrambling sequence provided. Seed will be ignored.")
            if not isinstance(seed, int):
                raise TypeError("seed must be int.")
        else:
            from ..config import config
            seed = int(config.np_rng.uniform(0, 2 ** 31 - 1))
        if not isinstance(binary, bool):
            raise TypeError("binary must be bool.")
        if not isinstance(keep_state, bool):
            raise TypeError("keep_state must be bool.")
        self._keep_batch_constant, self._binary, self._keep_state, self._seed = keep_batch_constant, binary, keep_state, seed
        self._sequence = None
        if sequence is not None:
            seq = np.asarray(sequence.cpu() if isinstance(sequence, torch.Tensor) else sequence, np.float32)
            if not np.all((seq == 0) | (seq == 1)):
                raise AssertionError("Scrambling sequence must be binary.")
            self._sequence = seq

    seed = property(lambda self: self._seed)
    keep_state = property(lambda self: self._keep_state)
    sequence = property(lambda self: self._sequence)

    def _draw(self, shape, seed, dev):
        """0/1 sequence of `shape` from the Philox stream keyed by (1337, seed) (scrambling.py:156-179)."""
        seq = torch.empty(shape, dtype=torch.float32, device=dev)
        key = ((1337 << 32) ^ (int(seed) & 0xFFFFFFFF)) & 0x7FFFFFFFFFFFFFFF
        check(lib().sb_binary_source(ptr(seq), seq.numel(), key, 0, current_stream()), "sb_binary_source")
        return seq

    def call(self, x, seed=None, binary=None):
        if binary is None:
            binary = self._binary
        elif not isinstance(binary, bool):
            raise TypeError("binary must be bool.")
        dev = self.device
        xin = x.to(device=dev, dtype=torch.float32).contiguous()
        if seed is None:
            if self._keep_state:
                seed = self._seed
            else:
                from ..config import config
                seed = int(config.np_rng.integers(0, 2 ** 31 - 1))
        if self._sequence is not None:
            seq = torch.from_numpy(np.ascontiguousarray(np.broadcast_to(self._sequence, tuple(xin.shape)))).to(dev)
        elif self._keep_batch_constant and xin.dim() > 1:
            seq = self._draw([1] + list(xin.shape[1:]), seed, dev)
        else:
            seq = self._draw(list(xin.shape), seed, dev)
        # kernel rows = leading dimension, so the int32 row length of the C-ABI never sees the whole-tensor size; a
        # sequence of the full input shape supplies one sequence row per input row, a batch-constant one a single row
        rows = xin.shape[0] if xin.dim() > 1 else 1
        n = xin.numel() // max(rows, 1)
        seq_rows = rows if seq.numel() == xin.numel() else 1
        if n >= 2 ** 31:
            raise ValueError("Scrambler: rows of 2^31 or more elements are not supported")
        out = torch.empty_like(xin)
        check(lib().sb_scramble(ptr(xin), ptr(seq), int(binary), ptr(out), rows, n, seq_rows, current_stream()),
              "sb_scramble")
        return out.to(x.dtype) if x.dtype.is_floating_point else out


class Descrambler(Block):
    """Descrambler(scrambler, binary=True, precision=None): inverse of an associated `Scrambler` / `TB5GScrambler`
    (scrambling.py:470-579); ``call(x, seed=None)`` re-applies its sequence, typically on LLRs (``binary=False``)."""

    def __init__(self, scrambler, binary=True, precision=None, **kwargs):
        super().__init__(precision=precision, **kwargs)
        if not isinstance(scrambler, (Scrambler, TB5GScrambler)):
            raise TypeError("scrambler must be an instance of Scrambler.")
        if not isinstance(binary, bool):
            raise TypeError("binary must be bool.")
        self._scrambler, self._binary = scrambler, binary
        if not scrambler.keep_state:
            print("Warning: scrambler uses random sequences that cannot be accessed by descrambler. Please use "
                  "keep_state=True or provide explicit random seed as input to call function.")

    scrambler = property(lambda self: self._scrambler)

    def call(self, x, /, *, seed=None):
        if isinstance(self._scrambler, Scrambler):
            s = seed if seed is not None else self._scrambler.seed
            return self._scrambler(x, seed=s, binary=self._binary)
        return self._scrambler(x, binary=self._binary)
