"""CLI re-host (vibo_b200/vibo.py): flag surface, dataset masking semantics
(CPU) and an end-to-end tiny training run (GPU)."""
import os

import numpy as np
import pytest
import torch

from helpers import GOLDEN


def test_flags_match_reference_surface():
    from vibo_b200 import vibo
    a = vibo.build_parser().parse_args([])
    # defaults of the reference's vibo.py:25-100
    assert (a.irt_model, a.dataset, a.ability_dim, a.ability_merge) == ('1pl', '1pl_simulation', 1, 'product')
    assert (a.lr, a.batch_size, a.epochs, a.max_iters, a.seed) == (5e-3, 16, 100, -1, 42)
    assert (a.num_person, a.num_item, a.num_posterior_samples, a.hidden_dim) == (1000, 100, 400, 64)
    assert a.beta_kl == 1.0 and not a.anneal_kl and not a.conditional_posterior and a.n_norm_flows == 0
    for flag in ('--drop-missing', '--no-infer-dict', '--no-marginal', '--no-test', '--no-predictive', '--cuda'):
        assert getattr(vibo.build_parser().parse_args([flag]), flag[2:].replace('-', '_')) is True


def test_artificial_mask_matches_reference():
    """RandomState(42) choice over the observed pool, as src/datasets.py:46-78."""
    from vibo_b200 import vibo
    z = np.load(os.path.join(GOLDEN, "artificial_mask.npz"))
    ds = vibo.ResidentDataset(torch.from_numpy(z["response_in"]), torch.from_numpy(z["mask_in"] != 0))
    out = vibo.artificially_mask_dataset(ds, 0.2)
    assert np.array_equal(out.missing_indices, z["missing_indices"])
    assert np.array_equal(out.missing_labels, z["missing_labels"][:, 0])
    assert np.array_equal(out.response.numpy(), z["response_out"])
    assert np.array_equal(out.mask.numpy(), z["mask_out"] != 0)
    # the original dataset is untouched
    assert np.array_equal(ds.response.numpy(), z["response_in"])


def test_simulation_split_is_80_20():
    from vibo_b200 import vibo
    a = vibo.build_parser().parse_args(['--dataset', '2pl_simulation', '--num-person', '50', '--num-item', '7'])
    tr, te = vibo.load_resident(a, True, 'cpu'), vibo.load_resident(a, False, 'cpu')
    assert (tr.num_person, te.num_person, tr.num_item) == (40, 10, 7)
    assert tr.response.shape == (40, 7, 1) and tr.mask.dtype == torch.bool
    assert set(np.unique(tr.response.numpy())) <= {0.0, 1.0}


@pytest.mark.gpu
def test_cli_end_to_end(tmp_path):
    from vibo_b200 import vibo
    vibo.main(['--irt-model', '2pl', '--dataset', '2pl_simulation', '--num-person', '400', '--num-item', '20',
               '--epochs', '6', '--batch-size', '32', '--num-posterior-samples', '3', '--cuda',
               '--artificial-missing-perc', '0.1', '--out-dir', str(tmp_path)])
    run = os.path.join(str(tmp_path), os.listdir(str(tmp_path))[0])
    for f in ('checkpoint.pth.tar', 'model_best.pth.tar', 'train_losses.npy', 'train_times.npy', 'test_losses.npy'):
        assert os.path.exists(os.path.join(run, f)), f
    ck = torch.load(os.path.join(run, 'checkpoint.pth.tar'), weights_only=False)
    assert {'model_state_dict', 'epoch', 'args', 'infer_dict', 'missing_imputation_accuracy', 'train_logp',
            'test_logp'} <= set(ck)
    assert ck['infer_dict']['ability_mu'].shape == (320, 1)
    losses = np.load(os.path.join(run, 'train_losses.npy'))
    assert np.isfinite(losses).all() and losses[-1] < losses[0]        # it trains
    assert (np.load(os.path.join(run, 'train_times.npy')) < 0).all()   # negative, like the reference
    assert 0.4 < ck['missing_imputation_accuracy'] <= 1.0
