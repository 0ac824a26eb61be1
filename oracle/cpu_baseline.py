"""CPU baseline leg: the float32 NumPy restatement of the reference's CPU path (same
[N, n_tau, 129] materialisation TensorFlow performs, signals.py:169-171) run over voxel chunks on
all host threads.  TEST / BENCH INFRASTRUCTURE ONLY -- "restated reference CPU path
(TensorFlow unavailable offline)"; never part of the product path."""
import os
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from . import qbold_oracle as o


def _work(args):
    ph, x, g, backward = args
    if backward:
        return o.forward_backward(ph, x, g, True, True, np.float32, chunk=1024)
    return o.forward(ph, x, True, True, np.float32, chunk=1024)


def run(ph, oef_dbv, g_signal=None, threads=None, backward=True, piece=2048):
    """Forward (+VJP) of `oef_dbv` split into pieces over a thread pool (NumPy ufuncs release the GIL).
    Returns (seconds, threads_used)."""
    threads = threads or os.cpu_count() or 1
    n = oef_dbv.shape[0]
    jobs = [(ph, oef_dbv[i:i + piece], None if g_signal is None else g_signal[i:i + piece], backward)
            for i in range(0, n, piece)]
    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=threads) as ex:
        list(ex.map(_work, jobs))
    return time.perf_counter() - t0, threads
