"""``MLPG`` with the reference's signature (``morgana/viz/synthesis.py:79``), batched on the device.

The reference loops over ``batch_size x feat_dim`` banded solves in fp64 on the CPU through ``bandmat`` and wraps them
in a device->host / host->device round trip inside ``predict()`` (models/RNN_SPSS.py:108-118).  Here every system is
one GPU WARP (fp64 L D L^T of <= 32 chunks of the sequence + a Schur complement over their separators, ``mg_mlpg.cu``), and
tensors never leave the device.  ``windows`` takes the reference's ``(l, u, win_coeff)`` tuples for windows that reach at most
one frame to either side (every 3-tap delta window); wider ones would leave the pentadiagonal solver: NotImplementedError.

Parity: ``bandmat`` (MattShannon/bandmat, unpinned in the reference's ``setup.py:12``) is neither vendored nor installable
here, so this op is pinned to the reference's own ``MLPG`` code running on a stand-in for the five bandmat entry points it
uses (``oracle/bandmat_standin.py``; solver = ``scipy.linalg.solveh_banded``, LAPACK's banded Cholesky -- the factorisation
``bandmat.linalg.solveh`` performs), not to bandmat's own bits.
"""
import numpy as np
import torch

from morgana_b200 import ops


def MLPG(means, variances, windows=None, padding_size=0, seq_len=None):
    r"""Maximum-likelihood parameter generation; arguments as in the reference.

    ``means`` (batch_size, seq_len, n_windows * feat_dim) or (seq_len, n_windows * feat_dim); ``variances`` same shape or
    (n_windows * feat_dim,); ``windows`` as in the reference (default: [1], [-0.5, 0, 0.5], [1, -2, 1], synthesis.py:122-127).
    NumPy inputs are accepted (as the reference's callers pass them) and moved to the current CUDA device; the result has
    the type of ``means`` (float32 tensor / float64 array as the reference returns).
    """
    table = None if windows is None else ops.mlpg_window_table(windows)
    as_numpy = isinstance(means, np.ndarray)
    device = means.device if isinstance(means, torch.Tensor) else torch.device('cuda', torch.cuda.current_device())

    def to_dev(x):
        if x is None or isinstance(x, torch.Tensor):
            return x
        return torch.as_tensor(np.asarray(x)).to(device)
    means_t = to_dev(means).to(torch.float32)
    var_t = to_dev(variances).to(torch.float32)
    single = means_t.dim() == 2
    if single:
        means_t = means_t[None]
        if var_t.dim() == 2:
            var_t = var_t[None]
    seq_t = to_dev(seq_len)
    out = ops.mlpg(means_t, var_t, padding_size=padding_size, seq_len=seq_t, windows=table)
    if single:
        out = out[0]
    return out.cpu().numpy().astype(np.float64) if as_numpy else out
