"""`vibo.py` command line re-hosted on the B200 engine.

Same flags, defaults, run-directory naming and output files as the reference's
src/torch_core/vibo.py:25-143, :478-488 (checkpoint.pth.tar / model_best.pth.tar
with 'model_state_dict', 'epoch', 'args'; train_losses.npy, train_times.npy,
test_losses.npy), with the hot loop replaced:

* the dataset lives on the GPU as two tensors; an epoch is a device-side
  permutation + row gathers instead of a DataLoader calling a per-person
  ``__getitem__``;
* a training step is ``model.fused_elbo`` (one pass over the batch rows);
* a training step is one CUDA-graph replay of ``ShardedElboTrainer`` (``--engine trainer``);
* ``log_marginal`` and the posterior-predictive mean run the S-sample loop inside one kernel;
* posterior-predictive imputation stores the mean of the decoded samples
  (``checkpoint['posterior_predict_mean']``) instead of stacking (S, P, I, 1) on the host --
  DEVIATION from the reference's ``checkpoint['posterior_predict_samples']`` (vibo.py:497-498);
  pass ``--save-predictive-samples`` to write that key as well.

    python -m vibo_b200.vibo --irt-model 2pl --dataset 2pl_simulation \\
        --num-person 10000 --num-item 100 --ability-dim 1 --cuda --epochs 5

Simulation datasets are generated in-process (plain-torch restatement of
src/simulate.py / src/pyro_core/models.py:25-225; pyro is not needed).  Real
datasets come from the reference's own src/datasets.py via --reference-root.
"""
from __future__ import annotations

import argparse
import copy
import math
import os
import shutil
import sys
import time
import types

import numpy as np
import torch

MISSING_DATA = -1  # reference src/config.py:14
IS_REAL_WORLD = {  # reference src/config.py:16-25 plus the keys load_dataset accepts (datasets.py:26-33)
    '1pl_simulation': False, '2pl_simulation': False, '3pl_simulation': False,
    '1pl_nonlinear': False, '2pl_nonlinear': False, '3pl_nonlinear': False,
    'critlangacq': True, 'duolingo': True, 'wordbank': True, 'pisa2015_science': True,
}


class AverageMeter:
    """reference src/utils.py:7-22"""

    def __init__(self):
        self.val = self.avg = self.sum = self.count = 0

    def update(self, val, n=1):
        self.val = val
        self.sum += val * n
        self.count += n
        self.avg = self.sum / self.count


def save_checkpoint(state, is_best, folder='./', filename='checkpoint.pth.tar'):
    """reference src/utils.py:25-31"""
    os.makedirs(folder, exist_ok=True)
    torch.save(state, os.path.join(folder, filename))
    if is_best:
        shutil.copyfile(os.path.join(folder, filename), os.path.join(folder, 'model_best.pth.tar'))


def build_parser():
    p = argparse.ArgumentParser()
    p.add_argument('--irt-model', type=str, default='1pl', choices=['1pl', '2pl', '3pl'])
    p.add_argument('--dataset', type=str, default='1pl_simulation',
                   choices=['1pl_simulation', '2pl_simulation', '3pl_simulation', 'critlangacq', 'duolingo',
                            'wordbank', 'pisa2015_science'])
    p.add_argument('--ability-dim', type=int, default=1)
    p.add_argument('--ability-merge', type=str, default='product', choices=['mean', 'product', 'transformer'])
    p.add_argument('--conditional-posterior', action='store_true', default=False)
    p.add_argument('--generative-model', type=str, default='irt', choices=['irt', 'link', 'deep', 'residual'])
    p.add_argument('--response-dist', type=str, default='bernoulli', choices=['gaussian', 'bernoulli'])
    p.add_argument('--drop-missing', action='store_true', default=False)
    p.add_argument('--artificial-missing-perc', type=float, default=0.)
    p.add_argument('--n-norm-flows', type=int, default=0)
    p.add_argument('--no-infer-dict', action='store_true', default=False)
    p.add_argument('--no-marginal', action='store_true', default=False)
    p.add_argument('--no-test', action='store_true', default=False)
    p.add_argument('--no-predictive', action='store_true', default=False)
    p.add_argument('--num-person', type=int, default=1000)
    p.add_argument('--num-item', type=int, default=100)
    p.add_argument('--num-posterior-samples', type=int, default=400)
    p.add_argument('--hidden-dim', type=int, default=64)
    p.add_argument('--max-num-person')
    p.add_argument('--max-num-item')
    p.add_argument('--out-dir', type=str, default=os.path.join(os.getcwd(), 'out'))
    p.add_argument('--lr', type=float, default=5e-3)
    p.add_argument('--batch-size', type=int, default=16, metavar='N')
    p.add_argument('--epochs', type=int, default=100, metavar='N')
    p.add_argument('--max-iters', type=int, default=-1, metavar='N')
    p.add_argument('--num-workers', type=int, default=0)
    p.add_argument('--anneal-kl', action='store_true', default=False)
    p.add_argument('--beta-kl', type=float, default=1.0)
    p.add_argument('--seed', type=int, default=42, metavar='S')
    p.add_argument('--gpu-device', type=int, default=0)
    p.add_argument('--cuda', action='store_true', default=False)
    # additions
    p.add_argument('--world-size', type=int, default=int(os.environ.get('WORLD_SIZE', '1')),
                   help='person-sharded data parallelism: launch with torchrun, one process per GPU; every '
                        'minibatch is split over the ranks, one all-reduce of [loss | grads] per step')
    p.add_argument('--engine', type=str, default='trainer', choices=['trainer', 'autograd'],
                   help="'trainer': ShardedElboTrainer (CUDA-graph step, fused Adam kernel, peer-memory "
                        "all-reduce); 'autograd': model.fused_elbo + loss.backward() + torch.optim.Adam")
    p.add_argument('--save-predictive-samples', action='store_true', default=False,
                   help="also store checkpoint['posterior_predict_samples'] = {'response': (S, P, I, 1)} as the "
                        "reference does (vibo.py:497-498); default: only the mean over the S draws "
                        "('posterior_predict_mean'), computed on the GPU")
    p.add_argument('--reference-root', type=str, default=os.environ.get('VIBO_REF'),
                   help='checkout of the reference repo; needed only for the real-world datasets')
    return p


def simulate(irt_model, num_person, num_item, ability_dim, device, seed):
    """Plain-torch restatement of the reference simulator (src/simulate.py:35-58
    -> src/pyro_core/models.py:25-225): ability ~ N(0,1) (P,D), item_feat ~ N(0,1)
    (I,F), response ~ Bernoulli(link)."""
    g = torch.Generator(device=device).manual_seed(seed)
    D = ability_dim
    F = {'1pl': 1, '2pl': D + 1, '3pl': D + 2}[irt_model]
    ability = torch.randn(num_person, D, generator=g, device=device)
    item = torch.randn(num_item, F, generator=g, device=device)
    if irt_model == '1pl':
        z = ability.sum(1, keepdim=True) + item[:, 0][None, :]
    else:
        z = ability @ (-item[:, :D].T) + item[:, D][None, :]
    p = torch.sigmoid(z)
    if irt_model == '3pl':
        gs = torch.sigmoid(item[:, D + 1])[None, :]
        p = gs + (1 - gs) * p
    response = torch.bernoulli(p, generator=g)
    return response, ability, item


class ResidentDataset:
    """(P, I, 1) float32 responses + (P, I, 1) bool mask on the training device;
    80/20 person split with the train block first (reference datasets.py:901-911)."""

    def __init__(self, response, mask):
        self.response = response
        self.mask = mask
        self.num_person, self.num_item = response.shape[0], response.shape[1]
        self.missing_indices = None
        self.missing_labels = None

    def __len__(self):
        return self.num_person

    def index_batches(self, batch_size, shuffle, generator=None):
        """Row-index tensors of the epoch's minibatches (device side, one permutation per epoch)."""
        n = self.num_person
        order = torch.randperm(n, device=self.response.device, generator=generator) if shuffle \
            else torch.arange(n, device=self.response.device)
        for a in range(0, n, batch_size):
            yield order[a:a + batch_size]

    def batches(self, batch_size, shuffle, generator=None):
        n = self.num_person
        order = torch.randperm(n, device=self.response.device, generator=generator) if shuffle else None
        for a in range(0, n, batch_size):
            if order is None:
                yield self.response[a:a + batch_size], self.mask[a:a + batch_size]
            else:
                idx = order[a:a + batch_size]
                yield self.response.index_select(0, idx), self.mask.index_select(0, idx)

    def num_batches(self, batch_size):
        return (self.num_person + batch_size - 1) // batch_size


def artificially_mask_dataset(dataset, perc):
    """Hide `perc` of the observed cells (reference datasets.py:46-78: RandomState(42),
    sorted choice without replacement over the observed (row, col) pool)."""
    assert 0 <= perc <= 1
    out = copy.copy(dataset)
    mask = dataset.mask[:, :, 0].cpu().numpy().copy()
    response = dataset.response[:, :, 0].cpu().numpy().copy()
    row, col = np.where(mask != 0)
    num = int(perc * row.shape[0])
    rs = np.random.RandomState(42)
    pick = np.sort(rs.choice(np.arange(row.shape[0]), size=num, replace=False))
    r, c = row[pick], col[pick]
    out.missing_labels = response[r, c].copy()
    out.missing_indices = np.stack([r, c], 1)
    mask[r, c] = 0
    response[r, c] = MISSING_DATA
    dev = dataset.response.device
    out.response = torch.from_numpy(response).unsqueeze(2).to(dev)
    out.mask = torch.from_numpy(mask != 0).unsqueeze(2).to(dev)
    return out


def load_resident(args, train, device):
    if not IS_REAL_WORLD[args.dataset]:
        irt = args.dataset.split('_')[0]
        response, _, _ = simulate(irt, args.num_person, args.num_item, args.ability_dim, device, args.seed)
        n_train = int(0.8 * response.shape[0])
        response = response[:n_train] if train else response[n_train:]
        mask = response != MISSING_DATA
    else:
        if not args.reference_root:
            raise SystemExit(f"--dataset {args.dataset} needs --reference-root (the reference's src/datasets.py "
                             "parses the raw files)")
        sys.path.insert(0, args.reference_root)
        sys.modules.setdefault('nltk', types.SimpleNamespace(word_tokenize=None))  # dead import, datasets.py:8
        from src.datasets import load_dataset
        ds = load_dataset(args.dataset, train=train, num_person=args.num_person, num_item=args.num_item,
                          ability_dim=args.ability_dim, max_num_person=args.max_num_person,
                          max_num_item=args.max_num_item)
        response = torch.from_numpy(np.asarray(ds.response, dtype=np.float32)).reshape(len(ds), -1).to(device)
        mask = torch.from_numpy(np.asarray(ds.mask) != 0).reshape(len(ds), -1).to(device)
    return ResidentDataset(response.float().unsqueeze(2).contiguous(), mask.unsqueeze(2).contiguous())


def main(argv=None):
    from . import VIBO_1PL, VIBO_2PL, VIBO_3PL
    args = build_parser().parse_args(argv)
    if args.n_norm_flows > 0:
        args.no_infer_dict = True
        args.no_predictive = True
    if args.artificial_missing_perc > 0:
        args.no_predictive = False
    torch.manual_seed(args.seed)
    np.random.seed(args.seed)
    if IS_REAL_WORLD[args.dataset]:
        args.num_person = args.num_item = None
        args.max_num_person = int(args.max_num_person) if args.max_num_person is not None else None
        args.max_num_item = int(args.max_num_item) if args.max_num_item is not None else None
    else:
        args.max_num_person = args.max_num_item = None

    out_file = 'VIBO_{}_{}_{}_{}_{}person_{}item_{}maxperson_{}maxitem_{}maskperc_{}ability_{}_{}_seed{}'.format(
        args.irt_model, args.dataset, args.response_dist, args.generative_model, args.num_person, args.num_item,
        args.max_num_person, args.max_num_item, args.artificial_missing_perc, args.ability_dim,
        args.ability_merge, 'conditional_q' if args.conditional_posterior else 'unconditional_q', args.seed)
    args.out_dir = os.path.join(args.out_dir, out_file)
    os.makedirs(args.out_dir, exist_ok=True)

    if not args.cuda:
        raise SystemExit("the B200 engine needs --cuda (there is no CPU fallback)")
    world, rank = max(1, args.world_size), 0
    if world > 1:
        import torch.distributed as dist
        rank = int(os.environ.get('RANK', '0'))
        args.gpu_device = int(os.environ.get('LOCAL_RANK', str(rank)))
        torch.cuda.set_device(args.gpu_device)
        if not dist.is_initialized():
            dist.init_process_group('nccl', device_id=torch.device('cuda', args.gpu_device))
    torch.cuda.set_device(args.gpu_device)
    device = torch.device('cuda', args.gpu_device)

    train_dataset = load_resident(args, True, device)
    test_dataset = load_resident(args, False, device)
    if args.artificial_missing_perc > 0:
        train_dataset = artificially_mask_dataset(train_dataset, args.artificial_missing_perc)
    num_item = train_dataset.num_item
    n_batches = train_dataset.num_batches(args.batch_size)
    if args.max_iters != -1:
        args.epochs = int(math.ceil(args.max_iters / float(n_batches)))
        print(f'Found MAX_ITERS={args.max_iters}, setting EPOCHS={args.epochs}')

    model_class = {'1pl': VIBO_1PL, '2pl': VIBO_2PL, '3pl': VIBO_3PL}[args.irt_model]
    model = model_class(args.ability_dim, num_item, hidden_dim=args.hidden_dim, ability_merge=args.ability_merge,
                        conditional_posterior=args.conditional_posterior, generative_model=args.generative_model,
                        response_dist=args.response_dist, replace_missing_with_prior=not args.drop_missing,
                        n_norm_flows=args.n_norm_flows).to(device)
    optimizer = torch.optim.Adam(model.parameters(), lr=args.lr)
    use_kl = args.n_norm_flows == 0

    def annealing(epoch, which):
        # reference vibo.py:223-230
        if args.anneal_kl:
            return float(which + epoch * n_batches + 1) / float(args.epochs // 2 * n_batches)
        return args.beta_kl

    # ---- the hot loop (reference vibo.py:232-320) on the trainer: every minibatch is gathered into a
    # static device buffer (two index_select launches) and the step itself -- zero_grad, fused ELBO
    # forward/backward, all-reduce, Adam -- is ONE CUDA-graph replay; at world_size > 1 each rank
    # takes its contiguous share of the batch.
    from .distributed import ShardedElboTrainer, shard_bounds
    trainer = None
    if args.engine == 'trainer':
        trainer = ShardedElboTrainer(model, lr=args.lr, world_size=world, rank=rank, beta=args.beta_kl,
                                     use_kl_divergence=use_kl, cuda_graph=not args.anneal_kl, seed=args.seed)
    elif world > 1:
        raise SystemExit("--world-size > 1 needs --engine trainer")
    static = {}

    def local_rows(dataset, idx):
        """This rank's share of the minibatch `idx`, gathered into static buffers (stable addresses,
        so the captured graph of that batch shape is replayed)."""
        lo, hi = shard_bounds(idx.numel(), rank, world)
        n = hi - lo
        key = (id(dataset), n)
        if key not in static:
            static[key] = (torch.empty(n, dataset.num_item, 1, device=device),
                           torch.empty(n, dataset.num_item, 1, dtype=torch.bool, device=device))
        r, m = static[key]
        torch.index_select(dataset.response, 0, idx[lo:hi], out=r)
        torch.index_select(dataset.mask, 0, idx[lo:hi], out=m)
        return r, m, lo

    def run_epoch(dataset, train_mode, epoch):
        total = torch.zeros((), device=device, dtype=torch.float64)
        seen = 0
        # same permutation on every rank: torch's global generator is seeded identically (args.seed)
        for b, idx in enumerate(dataset.index_batches(args.batch_size, shuffle=train_mode)):
            response, mask, lo = local_rows(dataset, idx)
            if trainer is not None:
                trainer.person_offset = lo
                if train_mode:
                    trainer.beta = annealing(epoch, b)
                    loss = trainer.train_step(response, mask)
                else:
                    trainer.beta = 1.0
                    loss = trainer.eval_step(response, mask)
            elif train_mode:
                optimizer.zero_grad(set_to_none=True)
                loss = model.fused_elbo(response, mask, annealing_factor=annealing(epoch, b),
                                        use_kl_divergence=use_kl)
                loss.backward()
                optimizer.step()
                loss = loss.detach()
            else:
                with torch.no_grad():
                    loss = model.fused_elbo(response, mask, use_kl_divergence=use_kl)
            total += loss.double() * idx.numel()   # no per-step host sync
            seen += idx.numel()
        return float(total.item()) / seen

    def train(epoch):
        model.train()
        avg = run_epoch(train_dataset, True, epoch)
        if rank == 0:
            print('====> Train Epoch: {} Loss: {:.4f}'.format(epoch, avg))
        return avg

    def test(epoch):
        model.eval()
        avg = run_epoch(test_dataset, False, epoch)
        if rank == 0:
            print('====> Test Epoch: {} Loss: {:.4f}'.format(epoch, avg))
        return avg

    def get_log_marginal_density(dataset):
        # reference vibo.py:322-347
        model.eval()
        meter = AverageMeter()
        with torch.no_grad():
            for response, mask in dataset.batches(args.batch_size, shuffle=False):
                lm = model.log_marginal(response, mask, num_samples=args.num_posterior_samples)
                meter.update(float(lm.mean().item()), response.shape[0])
        print('====> Marginal: {:.4f}'.format(meter.avg))
        return meter.avg

    def get_infer_dict(dataset):
        # reference vibo.py:420-454 (abilities of every batch, item parameters of the last one)
        model.eval()
        mus, lvs = [], []
        with torch.no_grad():
            for response, mask in dataset.batches(args.batch_size, shuffle=False):
                _, a_mu, a_lv, _, i_mu, i_lv = model.encode(response, mask)
                mus.append(a_mu.cpu())
                lvs.append(a_lv.cpu())
        return {'ability_mu': torch.cat(mus), 'ability_logvar': torch.cat(lvs),
                'item_feat_mu': i_mu.cpu(), 'item_feat_logvar': i_lv.cpu()}

    def posterior_predictive_mean(dataset, num_samples):
        """mean over S posterior draws of decode(ability_s, item_feat_s) (reference vibo.py:349-390
        stacks the S samples on the host and averages later, :515): one kernel per batch with the
        sample loop inside (vibo_predictive_mean)."""
        model.eval()
        outs = []
        for response, mask in dataset.batches(max(args.batch_size, 4096), shuffle=False):
            outs.append(model.posterior_predictive_mean(response, mask, num_samples).cpu())
        return torch.cat(outs)

    def posterior_predictive_samples(dataset, num_samples):
        """reference-format output of sample_posterior_predictive (vibo.py:349-390):
        {'response': (S, P, I, 1)} on the host -- only with --save-predictive-samples (S x the
        dataset in host memory)."""
        model.eval()
        sets = []
        with torch.no_grad():
            for response, mask in dataset.batches(args.batch_size, shuffle=False):
                _, a_mu, a_lv, _, i_mu, i_lv = model.encode(response, mask)
                samples = []
                for _ in range(num_samples):
                    ability = a_mu + torch.exp(0.5 * a_lv) * torch.randn_like(a_mu)
                    item = i_mu + torch.exp(0.5 * i_lv) * torch.randn_like(i_mu)
                    samples.append(model.decode(ability, item).cpu())
                sets.append(torch.stack(samples))
        return {'response': torch.cat(sets, dim=1)}

    def posterior_mean_response(dataset):
        # reference vibo.py:392-418
        model.eval()
        outs = []
        with torch.no_grad():
            for response, mask in dataset.batches(args.batch_size, shuffle=False):
                _, a_mu, _, _, i_mu, _ = model.encode(response, mask)
                outs.append(model.decode(a_mu, i_mu).cpu())
        return torch.cat(outs)

    def imputation_accuracy(pred):
        idx, labels = train_dataset.missing_indices, train_dataset.missing_labels
        guess = torch.round(pred[:, :, 0])[idx[:, 0], idx[:, 1]].numpy()
        return float((guess == labels).mean())

    best_loss, is_best = np.inf, False
    train_losses, test_losses, train_times = np.zeros(args.epochs), np.zeros(args.epochs), np.zeros(args.epochs)
    for epoch in range(args.epochs):
        t0 = time.time()
        train_losses[epoch] = train(epoch)
        torch.cuda.synchronize()
        train_times[epoch] = t0 - time.time()  # negative, as the reference stores it (vibo.py:467)
        if not args.no_test:
            test_losses[epoch] = test(epoch)
            is_best = test_losses[epoch] < best_loss
            best_loss = min(test_losses[epoch], best_loss)
        else:
            is_best = train_losses[epoch] < best_loss
            best_loss = min(train_losses[epoch], best_loss)
        if rank != 0:
            continue
        save_checkpoint({'model_state_dict': model.state_dict(), 'epoch': epoch, 'args': args}, is_best,
                        folder=args.out_dir)
        np.save(os.path.join(args.out_dir, 'train_losses.npy'), train_losses)
        np.save(os.path.join(args.out_dir, 'train_times.npy'), train_times)
        if not args.no_test:
            np.save(os.path.join(args.out_dir, 'test_losses.npy'), test_losses)

    if trainer is not None:
        trainer.close()
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        if rank != 0:
            return
    for name in ['checkpoint.pth.tar', 'model_best.pth.tar']:
        path = os.path.join(args.out_dir, name)
        if not os.path.exists(path):
            continue
        checkpoint = torch.load(path, weights_only=False)
        model.load_state_dict(checkpoint['model_state_dict'])
        if not args.no_infer_dict:
            checkpoint['infer_dict'] = get_infer_dict(train_dataset)
        if not args.no_predictive:
            if args.save_predictive_samples:
                # the key the reference's analysis scripts read (vibo.py:497-498)
                checkpoint['posterior_predict_samples'] = posterior_predictive_samples(
                    train_dataset, args.num_posterior_samples)
                pred = checkpoint['posterior_predict_samples']['response'].mean(0)
            else:
                pred = posterior_predictive_mean(train_dataset, args.num_posterior_samples)
            checkpoint['posterior_predict_mean'] = pred
            if args.artificial_missing_perc > 0:
                acc = imputation_accuracy(pred)
                checkpoint['missing_imputation_accuracy'] = acc
                print(f'Missing Imputation Accuracy from samples: {acc}')
                acc = imputation_accuracy(posterior_mean_response(train_dataset))
                checkpoint['missing_imputation_accuracy_mean'] = acc
                print(f'Missing Imputation Accuracy from mean: {acc}')
        if not args.no_marginal:
            checkpoint['train_logp'] = get_log_marginal_density(train_dataset)
            if not args.no_test:
                checkpoint['test_logp'] = get_log_marginal_density(test_dataset)
        torch.save(checkpoint, path)
        print(f'Train time: {np.abs(train_times[:100]).sum()}')


if __name__ == '__main__':
    main()
