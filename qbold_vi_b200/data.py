"""Dataset plumbing either side of the hot path: the batch shapes and crop / shuffle semantics of
the reference's ``prepare_dataset`` / ``prepare_synthetic_dataset`` (train.py:17-104) on torch tensors.

Input layout is the one ``data_preprocessing.py`` writes (data_preprocessing.py:265-266): 5-D arrays
``[subject, X, Y, Z, n_tau + 1]`` whose last channel is the mask.
"""
from __future__ import annotations

import numpy as np
import torch


def prepare_synthetic_dataset(x, y, batch_size=512, generator=None):
    """train.py:82-104: reshape the generated voxels to [-1,10,10,5,C] blocks, keep the last 10 % for validation,
    return (iterator factory over shuffled training batches, (valid_x, valid_y))."""
    x = x.reshape(-1, 10, 10, 5, x.shape[-1])
    y = y[:x.shape[0] * 500].reshape(-1, 10, 10, 5, 3)      # x may be shorter than y (signals.py:283-287)
    n_valid = x.shape[0] // 10
    train_x, train_y = x[:-n_valid], y[:-n_valid]
    valid = (x[-n_valid:], y[-n_valid:])

    def batches():
        perm = torch.randperm(train_x.shape[0], device=train_x.device, generator=generator)
        for i in range(0, train_x.shape[0], batch_size):
            idx = perm[i:i + batch_size]
            yield train_x[idx], train_y[idx]

    return batches, valid


class FineTuneDataset:
    """prepare_dataset (train.py:17-72): random in-plane crops of every subject volume together with the prior
    predicted by the pre-trained model, masked data, endless shuffled batches of 38 crops (3 when not training).

    ``real_data`` [S,X,Y,Z,n_tau+1] (numpy or torch, last channel = mask); ``model(data)`` returns
    (q_voxelwise, q_spatial, sigma) like qbold_vi_b200.encoder.Encoder.  Iterating yields
    ((data [B,c,c,Z,n_tau], mask [B,c,c,Z,1]), {'predictions': [.., 6], 'predicted_images': [.., n_tau+1]})."""

    def __init__(self, real_data, model, crop_size=20, training=True, blank_crop=True, device=None, seed=0):
        real = torch.as_tensor(np.asarray(real_data, dtype=np.float32) if not torch.is_tensor(real_data) else real_data)
        if blank_crop:
            real = real[:, 17:-17, 10:-10, :, :]                               # train.py:20
        self.device = torch.device(device) if device is not None else real.device
        self.real = real.float().to(self.device).contiguous()
        self.crop = [min(crop_size, self.real.shape[1]), min(crop_size, self.real.shape[2])]
        masked = self.real[..., :-1] * self.real[..., -1:]
        with torch.no_grad():                                                   # train.py:26-31
            can_fuse = getattr(model, 'supports_voxelwise_fused', lambda: False)()
            q = model.voxelwise_fused(masked) if can_fuse and masked.is_cuda else model(masked)[0]
        self.prior = q[..., :5].contiguous()
        self.batch = 38 if training else 3                                      # train.py:68,70
        self.gen = torch.Generator(device='cpu').manual_seed(seed)

    def _crop(self, s):
        cx, cy = self.crop
        x0 = int(torch.randint(0, self.real.shape[1] - cx + 1, (1,), generator=self.gen))
        y0 = int(torch.randint(0, self.real.shape[2] - cy + 1, (1,), generator=self.gen))
        d = self.real[s, x0:x0 + cx, y0:y0 + cy]                                # one crop for data and prior (train.py:43-45)
        p = self.prior[s, x0:x0 + cx, y0:y0 + cy]
        mask = d[..., -1:]
        data = d[..., :-1] * mask                                               # train.py:56
        return data, mask, torch.cat([p, mask], -1), torch.cat([data, mask], -1)

    def _subjects(self):
        """Endless stream of subject indices: shuffled passes over ALL subjects (tf.data shuffle + repeat, train.py:64-70),
        not independent draws -- every subject is visited once per pass."""
        while True:
            for s in torch.randperm(self.real.shape[0], generator=self.gen).tolist():
                yield s

    def __iter__(self):
        stream = self._subjects()
        while True:                                                             # .repeat(-1), shuffled
            subj = [next(stream) for _ in range(self.batch)]
            parts = [self._crop(s) for s in subj]
            data, mask, pred, img = (torch.stack(t) for t in zip(*parts))
            yield (data, mask), {'predictions': pred, 'predicted_images': img}
