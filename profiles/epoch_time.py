#!/usr/bin/env python
"""Epoch time of the CLI hot loop beside the reference's DataLoader loop.

C1 = 2PL simulation, 10,000 persons x 100 items, ability-dim 1 (8,000-person train split), at the
CLI default batch size 16 and at 4096.

* ours: `python -m vibo_b200.vibo --cuda ...` for 3 epochs; epoch 0 includes CUDA-graph capture, the
  reported figure is the mean of the later epochs (|train_times.npy|, the file the CLI writes).
* reference: the UNMODIFIED reference modules from baseline/_ref (VIBO_2PL forward + elbo +
  backward + Adam, body of src/torch_core/vibo.py:232-278) driven by a torch DataLoader over a
  per-person `__getitem__` dataset shaped like src/datasets.py:928-940, on the host CPU.

Prints one JSON line per configuration.
"""
import json
import os
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def ours(batch_size, epochs=3):
    from vibo_b200 import vibo
    out = tempfile.mkdtemp()
    vibo.main(['--irt-model', '2pl', '--dataset', '2pl_simulation', '--num-person', '10000', '--num-item', '100',
               '--ability-dim', '1', '--epochs', str(epochs), '--batch-size', str(batch_size), '--cuda',
               '--no-test', '--no-marginal', '--no-predictive', '--no-infer-dict', '--out-dir', out])
    run = os.path.join(out, os.listdir(out)[0])
    t = np.abs(np.load(os.path.join(run, 'train_times.npy')))
    losses = np.load(os.path.join(run, 'train_losses.npy'))
    return float(t[1:].mean()), float(t[0]), losses.tolist()


class PersonDataset(torch.utils.data.Dataset):
    """per-person items like the reference's datasets (index, response, item_id, mask)"""

    def __init__(self, response, mask):
        self.response, self.mask = response, mask
        self.item_id = np.arange(response.shape[1])

    def __len__(self):
        return self.response.shape[0]

    def __getitem__(self, index):
        return index, self.response[index], self.item_id, self.mask[index]


def reference(batch_size, max_steps=None):
    ref = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.exists(os.path.join(ref, "src", "torch_core", "models.py")):
        return None
    sys.path.insert(0, ref)
    from src.torch_core.models import VIBO_2PL
    from vibo_b200 import vibo
    torch.set_num_threads(os.cpu_count() or 1)
    response, _, _ = vibo.simulate('2pl', 10000, 100, 1, 'cpu', 42)
    response = response[:8000].unsqueeze(2).numpy().astype(np.float32)
    mask = np.ones_like(response)
    loader = torch.utils.data.DataLoader(PersonDataset(response, mask), batch_size=batch_size, shuffle=True)
    torch.manual_seed(42)
    model = VIBO_2PL(1, 100, hidden_dim=64, ability_merge='product')
    opt = torch.optim.Adam(model.parameters(), lr=5e-3)
    model.train()
    n_steps = len(loader)
    t0 = time.time()
    done = 0
    for _, resp, _, msk in loader:
        msk = msk.long()
        opt.zero_grad()
        out = model(resp, msk)
        loss = model.elbo(*out, annealing_factor=1.0, use_kl_divergence=True)
        loss.backward()
        opt.step()
        loss.item()
        done += 1
        if max_steps and done >= max_steps:
            break
    dt = time.time() - t0
    return dt * n_steps / done, done, n_steps


if __name__ == "__main__":
    for bs in (16, 4096):
        mean_s, first_s, losses = ours(bs)
        rec = {"config": "C1 2PL 10000x100 (8000 train persons)", "batch_size": bs,
               "ours_epoch_s": mean_s, "ours_first_epoch_s_incl_capture": first_s,
               "ours_cells_per_s": 8000 * 100 / mean_s, "train_losses": losses}
        r = reference(bs, max_steps=200 if bs == 16 else None)
        if r is not None:
            rec.update(reference_epoch_s=r[0], reference_steps_timed=r[1], reference_steps_per_epoch=r[2],
                       reference_cores=os.cpu_count(), speedup=r[0] / mean_s)
        print(json.dumps(rec))
