"""Raw IQ file reader (reference: gnsstools/io.py:3-12)."""

import numpy as np


def get_samples_complex(fp, n):
    """Read n interleaved int8 I/Q pairs from `fp` as complex64; None on a short read."""
    raw = fp.read(2 * n)
    if len(raw) != 2 * n:
        return None
    iq = np.frombuffer(raw, dtype=np.int8).astype(np.float32)
    return iq.view(np.complex64)
