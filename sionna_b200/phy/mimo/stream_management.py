"""StreamManagement (mirror of /root/reference/src/sionna/phy/mimo/stream_management.py:9-246): which transmitter sends
how many streams to which receiver, and the gather indices derived from it. Host-side NumPy only."""
import numpy as np


class StreamManagement:
    """StreamManagement(rx_tx_association, num_streams_per_tx)

    ``rx_tx_association[i, j] = 1`` means receiver i gets one or more streams from transmitter j. All index arrays
    (``detection_desired_ind``, ``detection_undesired_ind``, ``stream_ind`` ...) have the reference's definitions."""

    def __init__(self, rx_tx_association, num_streams_per_tx):
        self._num_streams_per_tx = int(num_streams_per_tx)
        self.rx_tx_association = rx_tx_association

    @property
    def rx_tx_association(self):
        return self._rx_tx_association

    @property
    def num_rx(self):
        return self._num_rx

    @property
    def num_tx(self):
        return self._num_tx

    @property
    def num_streams_per_tx(self):
        return self._num_streams_per_tx

    @property
    def num_streams_per_rx(self):
        return int(self.num_tx * self.num_streams_per_tx / self.num_rx)

    @property
    def num_interfering_streams_per_rx(self):
        return int(self.num_tx * self.num_streams_per_tx - self.num_streams_per_rx)

    @property
    def num_tx_per_rx(self):
        return self._num_tx_per_rx

    @property
    def num_rx_per_tx(self):
        return self._num_rx_per_tx

    @property
    def precoding_ind(self):
        return self._precoding_ind

    @property
    def stream_association(self):
        return self._stream_association

    @property
    def detection_desired_ind(self):
        return self._detection_desired_ind

    @property
    def detection_undesired_ind(self):
        return self._detection_undesired_ind

    @property
    def tx_stream_ids(self):
        return self._tx_stream_ids

    @property
    def rx_stream_ids(self):
        return self._rx_stream_ids

    @property
    def stream_ind(self):
        return self._stream_ind

    @rx_tx_association.setter
    def rx_tx_association(self, rx_tx_association):
        """Derives every index array from the association matrix with array operations (the reference builds them
        with nested loops, stream_management.py:165-246; tests/test_host_logic.py compares all of them with outputs
        of the reference class stored in tests/golden/stream_management_golden.json)."""
        assoc = np.array(rx_tx_association, np.int32)
        if assoc.ndim != 2 or not np.isin(assoc, (0, 1)).all():
            raise AssertionError("All elements of `stream_association` must be 0 or 1")
        per_rx, per_tx = assoc.sum(axis=1), assoc.sum(axis=0)
        if per_rx.min() != per_rx.max():
            raise AssertionError("Each receiver needs to be associated with the same number of transmitters.")
        if per_tx.min() != per_tx.max():
            raise AssertionError("Each transmitter needs to be associated with the same number of receivers.")
        self._num_rx, self._num_tx = assoc.shape
        self._num_tx_per_rx, self._num_rx_per_tx = per_rx[0], per_tx[0]
        self._rx_tx_association = assoc
        s_tx, s_rx = self.num_streams_per_tx, self.num_streams_per_rx
        # receivers of every transmitter, ascending: column-major scan of the non-zeros
        self._precoding_ind = np.nonzero(assoc.T)[1].reshape(self.num_tx, self._num_rx_per_tx).astype(np.int32)
        # the q-th receiver of a transmitter is served by that transmitter's streams [q*s_rx, (q+1)*s_rx)
        q = np.cumsum(assoc, axis=0) - 1
        k = np.arange(s_tx)
        lo = (q * s_rx)[:, :, None]
        served = (assoc[:, :, None] == 1) & (k >= lo) & (k < lo + s_rx)
        if (served.sum(axis=(1, 2)) != s_rx).any():
            raise ValueError("could not distribute the transmitters' streams: a transmitter needs num_rx_per_tx * "
                             "num_streams_per_rx streams")
        self._stream_association = served.astype(np.int32)
        flat = served.reshape(-1)
        self._detection_desired_ind = np.flatnonzero(flat)
        self._detection_undesired_ind = np.flatnonzero(~flat)
        self._tx_stream_ids = np.arange(self.num_tx * s_tx).reshape(self.num_tx, s_tx)
        # global stream numbers (tx * s_tx + k) received by each receiver, ascending
        self._rx_stream_ids = np.nonzero(served.reshape(self.num_rx, -1))[1].reshape(self.num_rx, s_rx).astype(np.int32)
        self._stream_ind = np.argsort(self._rx_stream_ids.reshape(-1))
