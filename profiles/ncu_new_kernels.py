"""One launch of each round-2 kernel family at a representative size, for `ncu --set full`
(profiles/collect_r02.sh)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vibo_b200  # noqa: E402
from vibo_b200 import kernels as K  # noqa: E402
from vibo_b200.distributed import ShardedElboTrainer  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)
# per-cell MLP on tcgen05: 20,000 x 500 cells
u, v = torch.randn(500, 64, device=dev), torch.randn(20000, 64, device=dev)
W2, c2, w4 = torch.randn(64, 64, device=dev) * 0.2, torch.randn(64, device=dev) * 0.1, torch.randn(64, device=dev)
for _ in range(2):
    K.percell_mlp(u, v, None, None, W2, c2, w4)
# sample loops: C1 train split, S = 400
model = vibo_b200.VIBO_2PL(1, 100, ability_merge="product").to(dev)
resp = (torch.rand(8000, 100, 1, device=dev) < 0.5).float()
mask = torch.ones(8000, 100, 1, dtype=torch.bool, device=dev)
model.log_marginal(resp, mask, 400, seed=1)
model.posterior_predictive_mean(resp, mask, 400, seed=1)
# the five-launch step at C2's shape (eager, so every kernel is a separate launch)
m2 = vibo_b200.VIBO_2PL(1, 500, ability_merge="product").to(dev)
r2 = (torch.rand(100000, 500, 1, device=dev) < 0.5).float()
k2 = torch.ones(100000, 500, 1, dtype=torch.bool, device=dev)
tr = ShardedElboTrainer(m2, cuda_graph=False)
for _ in range(2):
    tr.train_step(r2, k2)
# packed host entry: unpack kernel
pk = K.pack_rows(r2[:, :, 0].contiguous(), k2[:, :, 0].contiguous().view(torch.uint8)).cpu().pin_memory()
with torch.no_grad():
    m2.fused_elbo(pk, None)
torch.cuda.synchronize()
