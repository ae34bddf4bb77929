"""Seeded synthetic inputs for the BASELINE.json configurations.

Generated on the CPU with ``torch.Generator().manual_seed(seed)`` so the CPU baseline, the oracle and the CUDA path
see identical bits.  Shapes only -- there is no dataset or checkpoint behind any of this.
"""
import torch

# 187-dim WORLD layout after torch.split in models/RNN_SPSS.py:86-88 with output_dims {'lf0':3,'vuv':1,'mcep':180,'bap':3}
LF0, VUV, MCEP, BAP = slice(0, 3), slice(3, 4), slice(4, 184), slice(184, 187)
ACOUSTIC_DIM = 187
LAB_DIM = 600


def linguistic_batch(batch_size=256, min_phones=40, max_phones=80, max_dur=30, feat_dim=LAB_DIM, seed=1234):
    """Config 2: phone-rate labels + durations.  dur ~ U{1..max_dur}, zero past each utterance's phone count."""
    g = torch.Generator().manual_seed(seed)
    n_phones = torch.randint(min_phones, max_phones + 1, (batch_size,), generator=g)
    P = int(n_phones.max())
    valid = torch.arange(P)[None, :] < n_phones[:, None]
    dur = torch.randint(1, max_dur + 1, (batch_size, P), generator=g) * valid
    lab = torch.rand(batch_size, P, feat_dim, generator=g) * valid[:, :, None]   # collate_fn zero-pads (data.py:189)
    mmin = torch.zeros(feat_dim)
    mmax = torch.rand(feat_dim, generator=g) + 0.5
    mmax[::97] = mmin[::97]                                  # constant dims: scale forced to 1 (data.py:581)
    mean = torch.randn(feat_dim, generator=g)
    std = torch.randn(feat_dim, generator=g).abs() + 0.1
    n_frames = dur.sum(dim=1)
    return {'lab': lab, 'dur': dur[:, :, None].contiguous(), 'n_phones': n_phones, 'n_frames': n_frames,
            'mmin': mmin, 'mmax': mmax, 'mean': mean, 'std_dev': std}


def acoustic_batch(n_frames, max_len=None, seed=1234):
    """Config 3 tensors for given utterance lengths: 187-dim targets and predictions (B, T, 187) + the V/UV target."""
    g = torch.Generator().manual_seed(seed + 1000003)
    B = n_frames.shape[0]
    T = int(n_frames.max()) if max_len is None else int(max_len)
    target = torch.randn(B, T, ACOUSTIC_DIM, generator=g)
    pred = target + 0.1 * torch.randn(B, T, ACOUSTIC_DIM, generator=g)
    target[:, :, 0] = 5. + 0.3 * target[:, :, 0]                              # log-F0 in log-Hz
    pred[:, :, 0] = target[:, :, 0] + 0.05 * torch.randn(B, T, generator=g)
    voiced = torch.rand(B, T, 1, generator=g) < 0.6
    target[:, :, 3:4] = voiced.float()
    pred[:, :, 3:4] = torch.sigmoid(torch.randn(B, T, 1, generator=g))       # a probability, as after torch.sigmoid
    return {'target': target, 'pred': pred, 'voiced': voiced, 'n_frames': n_frames.clone()}


def acoustic_lengths(batch_size=1024, min_frames=300, max_frames=1200, seed=1234):
    g = torch.Generator().manual_seed(seed)
    return torch.randint(min_frames, max_frames + 1, (batch_size,), generator=g)
