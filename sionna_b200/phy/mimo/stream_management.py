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
        a = np.array(rx_tx_association, np.int32)
        assert all(x in [0, 1] for x in np.nditer(a)), "All elements of `stream_association` must be 0 or 1"
        self._num_rx, self._num_tx = np.shape(a)
        num_tx_per_rx = np.sum(a, 1)
        assert np.min(num_tx_per_rx) == np.max(num_tx_per_rx), \
            "Each receiver needs to be associated with the same number of transmitters."
        self._num_tx_per_rx = num_tx_per_rx[0]
        num_rx_per_tx = np.sum(a, 0)
        assert np.min(num_rx_per_tx) == np.max(num_rx_per_tx), \
            "Each transmitter needs to be associated with the same number of receivers."
        self._num_rx_per_tx = num_rx_per_tx[0]
        self._rx_tx_association = a
        self._precoding_ind = np.zeros([self.num_tx, self.num_rx_per_tx], np.int32)
        for i in range(self.num_tx):
            self._precoding_ind[i, :] = np.where(a[:, i])[0]
        sa = np.zeros([self.num_rx, self.num_tx, self.num_streams_per_tx], np.int32)
        n_streams = np.min([self.num_streams_per_rx, self.num_streams_per_tx])
        for j in range(self.num_tx):
            c = 0
            for i in range(self.num_rx):
                if a[i, j]:
                    sa[i, j, c:c + self.num_streams_per_rx] = np.ones([n_streams])
                    c += self.num_streams_per_rx
        self._stream_association = sa
        self._detection_desired_ind = np.where(np.reshape(sa, [-1]) == 1)[0]
        self._detection_undesired_ind = np.where(np.reshape(sa, [-1]) == 0)[0]
        self._tx_stream_ids = np.reshape(np.arange(0, self.num_tx * self.num_streams_per_tx),
                                         [self.num_tx, self.num_streams_per_tx])
        self._rx_stream_ids = np.zeros([self.num_rx, self.num_streams_per_rx], np.int32)
        for i in range(self.num_rx):
            c = []
            for j in range(self.num_tx):
                if a[i, j]:
                    tmp = np.where(sa[i, j])[0] + j * self.num_streams_per_tx
                    c += list(tmp)
            self._rx_stream_ids[i, :] = c
        self._stream_ind = np.argsort(np.reshape(self._rx_stream_ids, [-1]))
