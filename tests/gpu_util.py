"""Helpers shared by the GPU parity tests."""
import os

import numpy as np

LOG = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'gpurun_out', 'parity_report.txt')


def rel(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def report(name, **errs):
    """Append measured discrepancies to gpurun_out/parity_report.txt (travels back from the GPU box)."""
    try:
        os.makedirs(os.path.dirname(LOG), exist_ok=True)
        with open(LOG, 'a') as fh:
            fh.write(name + ' ' + ' '.join('%s=%.3e' % kv for kv in errs.items()) + '\n')
    except OSError:
        pass


def dev(a, dtype=None):
    import torch
    t = torch.from_numpy(np.ascontiguousarray(a)).cuda()
    return t if dtype is None else t.to(dtype)
