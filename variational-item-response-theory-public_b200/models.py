"""Drop-in ``VIBO_1PL / VIBO_2PL / VIBO_3PL`` modules on the B200 kernels.

Mirrors the public surface of the reference's src/torch_core/models.py:246-548
(constructor arguments, ``forward`` / ``encode`` / ``decode`` / ``elbo`` /
``log_marginal`` signatures and tuple layouts, ``state_dict`` keys, seeded
initial weights) so that ``vibo.py``-style drivers and reference checkpoints
work unchanged, and adds ``fused_elbo`` -- the single-pass entry the training
loop should call instead of ``forward`` + ``elbo``.

How the encoder is evaluated.  The reference pushes every response cell
through ``Linear-ELU-Linear-ELU-Linear`` (models.py:575-582, :652-661).  For
Bernoulli responses a cell's input takes only 2 (unconditional) or 2*I
(conditional: ``[r, item_feat_j]``) distinct values, so the MLP is evaluated
on those rows only -- the *expert table* -- with ordinary PyTorch autograd, and
the kernels select an expert per cell (SURVEY.md finding 2).  Everything that
touches the (P, I) matrix runs in libvibo_b200.so.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F
from torch import nn
from torch.nn import init

from . import functional as VF
from .decoders import DeepIRT, LinkedIRT, ResidualIRT
from .flows import NormalizingFlows, PlanarFlow

LOG_SQRT_2PI = 0.5 * math.log(2.0 * math.pi)


def _encoder_mlp(input_dim, hidden_dim, output_dim):
    return nn.Sequential(
        nn.Linear(input_dim, hidden_dim),
        nn.ELU(inplace=True),
        nn.Linear(hidden_dim, hidden_dim),
        nn.ELU(inplace=True),
        nn.Linear(hidden_dim, output_dim),
    )


def _mean_mlps(input_dim, hidden_dim, output_dim):
    """reference models.py:584-594 (``_create_models_mean``): mlp1 then mlp2."""
    mlp1 = nn.Sequential(nn.Linear(input_dim, hidden_dim), nn.ELU(inplace=True),
                         nn.Linear(hidden_dim, hidden_dim))
    mlp2 = nn.Sequential(nn.Linear(hidden_dim, hidden_dim), nn.ELU(inplace=True),
                         nn.Linear(hidden_dim, output_dim))
    return mlp1, mlp2


class AbilityInferenceNetwork(nn.Module):
    """q(ability | responses) (reference models.py:551-661).

    ``product`` (the vibo.py default): per-cell Gaussian experts merged by a product of
    experts; holds ``self.mlp`` (``mlp.{0,2,4}.{weight,bias}``) and runs on the encode kernels.
    ``mean`` (the constructor default of the reference): per-cell hidden vectors averaged over a
    person's observed items, then a second MLP (``mlp1.{0,2}``, ``mlp2.{0,2}``).  For 0/1
    responses the per-cell MLP collapses to a table exactly as for ``product``, so the average
    is ``(counts or indicator matrix) x table`` -- composed from PyTorch reductions / GEMMs
    around the link kernel (no per-cell activations are ever materialised)."""

    conditional = False

    def __init__(self, ability_dim, response_dim, hidden_dim=64, ability_merge='mean',
                 replace_missing_with_prior=True):
        super().__init__()
        assert ability_merge in ('mean', 'product')
        self.ability_dim = ability_dim
        self.response_dim = response_dim
        self.hidden_dim = hidden_dim
        self.ability_merge = ability_merge
        self.replace_missing_with_prior = replace_missing_with_prior
        self._create_models(response_dim)
        # the two possible cell inputs; non-persistent so state_dict keys stay the reference's
        self.register_buffer("_cell_inputs", torch.tensor([[0.0], [1.0]]), persistent=False)

    def _create_models(self, input_dim):
        if self.ability_merge == 'product':
            self.mlp = _encoder_mlp(input_dim, self.hidden_dim, self.ability_dim * 2)
        else:
            self.mlp1, self.mlp2 = _mean_mlps(input_dim, self.hidden_dim, self.ability_dim * 2)

    def _cell_rows(self, item_feat):
        """The distinct per-cell MLP inputs: (2, 1) unconditional, (2 I, 1 + F) conditional."""
        if not self.conditional:
            return self._cell_inputs.to(self._first_weight().dtype), 1
        I = item_feat.shape[0]
        r = self._cell_inputs.to(item_feat.dtype).view(2, 1, 1).expand(2, I, 1)
        rows = torch.cat([r, item_feat.unsqueeze(0).expand(2, I, -1)], dim=2)
        return rows.reshape(2 * I, -1), I

    def _first_weight(self):
        return self.mlp[0].weight if self.ability_merge == 'product' else self.mlp1[0].weight

    def _forward_mean(self, resp, msk, item_feat):
        """reference models.py:631-650 with the table collapse: hidden table (2, It, H), per-person
        masked mean over items, mlp2.  resp (P, I) f32, msk (P, I) u8."""
        rows, It = self._cell_rows(item_feat)
        hid = F.elu(self.mlp1(rows)).reshape(2, It, self.hidden_dim)
        if It == 1 and resp.is_cuda:
            counts = VF.K.person_counts(resp, msk)   # one streaming pass: (ones, observed) per person
            n1, n_obs = counts[:, 0:1], counts[:, 1:2]
            hid_mean = ((n_obs - n1) * hid[0] + n1 * hid[1]) / n_obs
            mu, logvar = torch.chunk(self.mlp2(hid_mean), 2, dim=1)
            return mu, logvar
        obs = msk != 0
        one = (resp > 0.5) & obs
        n_obs = obs.sum(1, keepdim=True).to(hid.dtype)
        if It == 1:
            n1 = one.sum(1, keepdim=True).to(hid.dtype)
            hid_sum = (n_obs - n1) * hid[0] + n1 * hid[1]
        else:
            hid_sum = torch.zeros(resp.shape[0], self.hidden_dim, dtype=hid.dtype, device=hid.device)
            blk = 32768   # bounds the (rows, I) indicator temporaries
            parts = []
            for a in range(0, resp.shape[0], blk):
                w1 = one[a:a + blk].to(hid.dtype)
                w0 = obs[a:a + blk].to(hid.dtype) - w1
                parts.append(w1 @ hid[1] + w0 @ hid[0])
            hid_sum = torch.cat(parts) if parts else hid_sum
        hid_mean = hid_sum / n_obs
        mu, logvar = torch.chunk(self.mlp2(hid_mean), 2, dim=1)
        return mu, logvar

    @property
    def missing_policy(self):
        return VF.MISSING_PRIOR if self.replace_missing_with_prior else VF.MISSING_DROP

    # continuous responses (--response-dist gaussian): a cell's MLP input is no longer one of 2
    # values, so the expert table does not exist and the encoder runs PER CELL (reference
    # models.py:596-629 / :631-650 literally), in person chunks so activations stay bounded
    per_cell = False
    per_cell_chunk = 4096

    def _forward_per_cell(self, resp, msk, item_feat):
        P, I = resp.shape
        D = self.ability_dim
        mus, lvs = [], []
        for a in range(0, P, self.per_cell_chunk):
            r = resp[a:a + self.per_cell_chunk]
            o = msk[a:a + self.per_cell_chunk] != 0
            n = r.shape[0]
            flat = r.reshape(n * I, 1)
            if self.conditional:
                flat = torch.cat([flat, item_feat.unsqueeze(0).expand(n, I, -1).reshape(n * I, -1)], dim=1)
            if self.ability_merge == 'mean':
                hid = F.elu(self.mlp1(flat)).reshape(n, I, self.hidden_dim)
                w = o.unsqueeze(2).to(hid.dtype)
                hid_mean = (hid * w).sum(1) / w.sum(1)
                mu, lv = torch.chunk(self.mlp2(hid_mean), 2, dim=1)
            else:
                out = self.mlp(flat).reshape(n, I, 2 * D)
                mu_c, lv_c = out[:, :, :D], out[:, :, D:]
                # product of experts (utils.py:105-113) with prior experts N(0, 1) for the missing
                # cells, or none of them under --drop-missing (models.py:606-627)
                T = 1.0 / (torch.exp(lv_c) + 1e-8)
                obs = o.unsqueeze(2)
                prior_T = 1.0 / (1.0 + 1e-8) if self.replace_missing_with_prior else 0.0
                T = torch.where(obs, T, torch.full_like(T, prior_T))
                mu_c = torch.where(obs, mu_c, torch.zeros_like(mu_c))
                S = T.sum(1)
                mu = (mu_c * T).sum(1) / S
                lv = torch.log(1.0 / S)
            mus.append(mu)
            lvs.append(lv)
        return torch.cat(mus), torch.cat(lvs)

    def expert_table(self, item_feat=None):
        """(2, 1, 2D): the MLP on the two possible cell inputs r = 0, 1."""
        return self.mlp(self._cell_inputs.to(self.mlp[0].weight.dtype)).unsqueeze(1)

    def forward(self, response, mask, item_feat=None):
        resp, msk = VF.prepare_rows(response, mask)
        if self.per_cell:
            return self._forward_per_cell(resp, msk, item_feat)
        if self.ability_merge == 'mean':
            return self._forward_mean(resp, msk, item_feat)
        table = self.expert_table(item_feat)
        return VF.EncodePosterior.apply(resp, msk, table, self.conditional, self.missing_policy)


class ConditionalAbilityInferenceNetwork(AbilityInferenceNetwork):
    """q(ability | responses, items): cell input ``[r_ij, item_feat_j]``
    (reference models.py:664-710)."""

    conditional = True

    def __init__(self, ability_dim, response_dim, item_feat_dim, hidden_dim=64,
                 ability_merge='mean', replace_missing_with_prior=True):
        # The reference builds the unconditional MLP in the parent constructor
        # and then re-creates it with the wider input (models.py:675-693); the
        # same two constructions keep seeded initial weights identical.
        super().__init__(ability_dim, response_dim, hidden_dim=hidden_dim,
                         ability_merge=ability_merge,
                         replace_missing_with_prior=replace_missing_with_prior)
        self.item_feat_dim = item_feat_dim
        self._create_models(response_dim + item_feat_dim)

    def expert_table(self, item_feat=None):
        """(2, I, 2D): the MLP on [r, item_feat_j] for r = 0, 1 and every item."""
        I = item_feat.shape[0]
        r = self._cell_inputs.to(item_feat.dtype).view(2, 1, 1).expand(2, I, 1)
        rows = torch.cat([r, item_feat.unsqueeze(0).expand(2, I, -1)], dim=2)
        return self.mlp(rows.reshape(2 * I, -1)).reshape(2, I, -1)


class ItemInferenceNetwork(nn.Module):
    """Per-item Gaussian posterior tables (reference models.py:713-726)."""

    def __init__(self, num_item, item_feat_dim):
        super().__init__()
        self.mu_lookup = nn.Embedding(num_item, item_feat_dim)
        self.logvar_lookup = nn.Embedding(num_item, item_feat_dim)

    def forward(self, item_index=None):
        if item_index is None:
            return self.mu_lookup.weight, self.logvar_lookup.weight
        idx = item_index.reshape(-1).long()
        return self.mu_lookup(idx), self.logvar_lookup(idx)


def kl_divergence_standard_normal_prior(mu, logvar):
    """reference src/utils.py:85-88."""
    return torch.sum(-0.5 * (1 + logvar - mu.pow(2) - logvar.exp()), dim=1)


def normal_log_pdf(x, mu, logvar):
    """reference src/utils.py:59-61 (closed form of Normal.log_prob)."""
    return -(x - mu) ** 2 / (2 * torch.exp(logvar)) - 0.5 * logvar - LOG_SQRT_2PI


def standard_normal_log_pdf(x):
    """reference src/utils.py:64-67."""
    return -0.5 * x ** 2 - LOG_SQRT_2PI


class VIBO_1PL(nn.Module):
    irt_num = 1
    # run the parameter-side chain of the unconditional model (item reparameterisation, expert
    # table, item prior term and their backward) in two CUDA kernels instead of PyTorch ops
    fuse_param_chain = True

    def __init__(self, latent_dim, num_item, hidden_dim=64, ability_merge='mean',
                 conditional_posterior=False, generative_model='irt', response_dist='bernoulli',
                 replace_missing_with_prior=True, n_norm_flows=0):
        super().__init__()
        assert ability_merge in ['mean', 'product']
        assert generative_model in ['irt', 'link', 'deep', 'residual']
        assert response_dist in ['bernoulli', 'gaussian']

        self.latent_dim = latent_dim
        self.ability_dim = latent_dim
        self.response_dim = 1
        self.hidden_dim = hidden_dim
        self.num_item = num_item
        self.ability_merge = ability_merge
        self.conditional_posterior = conditional_posterior
        self.generative_model = generative_model
        self.response_dist = response_dist
        self.replace_missing_with_prior = replace_missing_with_prior
        self.n_norm_flows = n_norm_flows
        self._set_item_feat_dim()
        self._check_kernel_limits()

        # construction order == reference (models.py:281-329): RNG-compatible
        if conditional_posterior:
            self.ability_encoder = ConditionalAbilityInferenceNetwork(
                self.ability_dim, self.response_dim, self.item_feat_dim, self.hidden_dim,
                ability_merge=ability_merge, replace_missing_with_prior=replace_missing_with_prior)
        else:
            self.ability_encoder = AbilityInferenceNetwork(
                self.ability_dim, self.response_dim, self.hidden_dim,
                ability_merge=ability_merge, replace_missing_with_prior=replace_missing_with_prior)
        self.item_encoder = ItemInferenceNetwork(num_item, self.item_feat_dim)
        if n_norm_flows > 0:
            self.ability_norm_flows = NormalizingFlows(self.ability_dim, n_flows=n_norm_flows)
            self.item_norm_flows = NormalizingFlows(self.item_feat_dim, n_flows=n_norm_flows)
        # nonlinear generative models (reference models.py:311-327, :769-919)
        if generative_model == 'link':
            self.decoder = LinkedIRT(irt_model=f'{self.irt_num}pl', hidden_dim=self.hidden_dim)
        elif generative_model == 'deep':
            self.decoder = DeepIRT(self.ability_dim, irt_model=f'{self.irt_num}pl', hidden_dim=self.hidden_dim)
        elif generative_model == 'residual':
            self.decoder = ResidualIRT(self.ability_dim, irt_model=f'{self.irt_num}pl', hidden_dim=self.hidden_dim)
        self.ability_encoder.per_cell = response_dist == 'gaussian'
        self.apply(self.weights_init)

    # ------------------------------------------------------------------ setup
    def _check_kernel_limits(self):
        """Fail at construction, not at the first step, when the kernels cannot take this shape."""
        from . import _lib
        if self.ability_dim > _lib.MAX_ABILITY_DIM:
            raise ValueError(f"ability_dim {self.ability_dim} > {_lib.MAX_ABILITY_DIM} supported by the B200 kernels")
        try:
            lib = _lib.load()
        except Exception:
            return  # no library here (pure host-logic tests); the kernels check again at call time
        import ctypes
        d = _lib.Desc(0, self.num_item, self.ability_dim, self.irt_num, int(self.conditional_posterior), 0, 0, 0)
        limit = int(lib.vibo_max_items(ctypes.byref(d)))
        if self.num_item > limit:
            raise ValueError(f"num_item {self.num_item} exceeds the B200 kernels' limit of {limit} items for "
                             f"ability_dim {self.ability_dim} (item slabs are register-resident)")

    def _set_item_feat_dim(self):
        self.item_feat_dim = {1: 1, 2: self.latent_dim + 1, 3: self.latent_dim + 2}[self.irt_num]

    @staticmethod
    def weights_init(m):
        # reference models.py:512-518: xavier-normal (relu gain), zero bias
        if isinstance(m, (nn.Linear, nn.Conv2d)):
            init.xavier_normal_(m.weight.data, gain=init.calculate_gain('relu'))
            init.constant_(m.bias.data, 0)

    @staticmethod
    def reparameterize_gaussian(mean, logvar):
        # reference models.py:506-510
        std = torch.exp(0.5 * logvar)
        eps = torch.randn_like(std)
        return eps.mul(std).add_(mean)

    # ------------------------------------------------------- reference surface
    def encode(self, response, mask):
        """reference models.py:356-371 -> 6-tuple."""
        item_feat_mu, item_feat_logvar = self.item_encoder()
        item_feat = self.reparameterize_gaussian(item_feat_mu, item_feat_logvar)
        ability_mu, ability_logvar = self.ability_encoder(
            response, mask, item_feat if self.conditional_posterior else None)
        ability = self.reparameterize_gaussian(ability_mu, ability_logvar)
        return ability, ability_mu, ability_logvar, item_feat, item_feat_mu, item_feat_logvar

    def decode(self, ability, item_feat):
        """reference models.py:373-378 -> response_mu (P, I, 1)."""
        if self.generative_model == 'irt':
            return VF.Decode.apply(ability, item_feat, self.irt_num)
        return self.decoder(ability, item_feat)

    decode_chunk = 4096   # persons per decoder call in the loss paths (bounds the (P, I, H) activations)

    def _loglik(self, resp, msk, ability, item_feat):
        """sum_ij o_ij log p(x_ij | ability_i, item_feat_j) for the configured generative model and
        response distribution (reference models.py:397-404): the fused link kernel for the IRT /
        Bernoulli model, otherwise decoder (person chunks) + the materialised-probability kernels."""
        if self.generative_model == 'irt' and self.response_dist == 'bernoulli':
            return VF.LinkLogLik.apply(resp, msk, ability, item_feat, self.irt_num)
        total = 0.0
        for a in range(0, resp.shape[0], self.decode_chunk):
            b = a + self.decode_chunk
            mu = self.decode(ability[a:b], item_feat)
            total = total + self._loglik_given_mu(resp[a:b], msk[a:b], mu)
        return total

    def _loglik_given_mu(self, resp, msk, response_mu):
        if self.response_dist == 'bernoulli':
            return VF.BernoulliLogLik.apply(resp, msk, response_mu)
        # masked_gaussian_log_pdf with response_logvar = 2 log 0.1 (models.py:400-402, utils.py:52-56)
        mu = response_mu.reshape(resp.shape)
        sigma = 0.1
        lp = -((resp - mu) ** 2) / (2 * sigma * sigma) - math.log(sigma) - LOG_SQRT_2PI
        return (lp * (msk != 0).to(lp.dtype)).sum()

    def forward(self, response, mask):
        """reference models.py:337-354 -> 9-tuple (13-tuple with flows)."""
        ability, ability_mu, ability_logvar, item_feat, item_feat_mu, item_feat_logvar = \
            self.encode(response, mask)
        if self.n_norm_flows > 0:
            ability_k, ability_ldj = self.ability_norm_flows(ability)
            item_feat_k, item_ldj = self.item_norm_flows(item_feat)
            response_mu = self.decode(ability_k, item_feat_k)
            return (response, mask, response_mu, ability_k, ability, ability_mu, ability_logvar,
                    ability_ldj, item_feat_k, item_feat, item_feat_mu, item_feat_logvar, item_ldj)
        response_mu = self.decode(ability, item_feat)
        return (response, mask, response_mu, ability, ability_mu, ability_logvar,
                item_feat, item_feat_mu, item_feat_logvar)

    def elbo(self, response, mask, response_mu, ability, ability_mu, ability_logvar, item_feat,
             item_feat_mu, item_feat_logvar, annealing_factor=1, use_kl_divergence=True,
             ability_k=None, item_feat_k=None, ability_logabsdetjac=None, item_logabsdetjac=None):
        """reference models.py:380-443 on a materialised response_mu -> -ELBO (0-d)."""
        resp, msk = VF.prepare_rows(response, mask, binary=self.response_dist == 'bernoulli')
        ll = self._loglik_given_mu(resp, msk, response_mu)
        return -self._assemble_elbo(ll, ability, ability_mu, ability_logvar, item_feat, item_feat_mu,
                                    item_feat_logvar, annealing_factor, use_kl_divergence, ability_k,
                                    item_feat_k, ability_logabsdetjac, item_logabsdetjac)

    def _assemble_elbo(self, ll, ability, ability_mu, ability_logvar, item_feat, item_feat_mu,
                       item_feat_logvar, annealing_factor, use_kl_divergence, ability_k, item_feat_k,
                       ability_ldj, item_ldj):
        if self.n_norm_flows > 0:
            assert ability_ldj is not None and item_ldj is not None
            assert ability_k is not None and item_feat_k is not None
            log_q_u = normal_log_pdf(ability, ability_mu, ability_logvar).sum() - ability_ldj.sum()
            log_q_d = normal_log_pdf(item_feat, item_feat_mu, item_feat_logvar).sum() - item_ldj.sum()
            log_p = standard_normal_log_pdf(ability_k).sum() + standard_normal_log_pdf(item_feat_k).sum()
            return (ll + log_p) - (log_q_u + log_q_d)
        if use_kl_divergence:
            kl_u = kl_divergence_standard_normal_prior(ability_mu, ability_logvar).sum()
            kl_d = kl_divergence_standard_normal_prior(item_feat_mu, item_feat_logvar).sum()
            return ll - annealing_factor * kl_u - annealing_factor * kl_d
        log_p = standard_normal_log_pdf(ability).sum() + standard_normal_log_pdf(item_feat).sum()
        log_q = normal_log_pdf(ability, ability_mu, ability_logvar).sum() \
            + normal_log_pdf(item_feat, item_feat_mu, item_feat_logvar).sum()
        return (ll + log_p) - log_q

    def log_marginal(self, response, mask, num_samples=100, eps_item=None, eps_ability=None, seed=None):
        """reference models.py:445-504: logsumexp over num_samples batch-summed log-weights -
        log(num_samples) (0-d).  Unconditional product-of-experts posterior without flows on the
        GPU: ONE kernel with the sample loop inside (``vibo_log_marginal``: rows read once, scored
        against every sample from shared memory).  Otherwise each weight is one fused forward pass.
        eps_item (S, I, F) / eps_ability (S, P, D): optional pre-drawn noise."""
        with torch.no_grad():
            resp, msk = VF.prepare_rows(response, mask)
            if (resp.is_cuda and self.ability_merge == 'product' and not self.conditional_posterior
                    and self.n_norm_flows == 0 and self.generative_model == 'irt'
                    and self.response_dist == 'bernoulli'):
                if seed is None and (eps_item is None or eps_ability is None):
                    seed = int(torch.randint(0, 2 ** 62, (1,)).item())
                item_mu, item_lv = self.item_encoder()
                logp, _ = VF.K.log_marginal(resp, msk, self.ability_encoder.expert_table(), item_mu, item_lv,
                                            num_samples, irt_model=self.irt_num,
                                            missing_policy=self.ability_encoder.missing_policy,
                                            eps_item=eps_item, eps_ability=eps_ability,
                                            seed=0 if seed is None else seed)
                return logp.to(torch.float32)
            log_w = torch.stack([-self.fused_elbo(
                response, mask, use_kl_divergence=False,
                eps_item=None if eps_item is None else eps_item[s],
                eps_ability=None if eps_ability is None else eps_ability[s])
                for s in range(num_samples)])
            return torch.logsumexp(log_w, 0) - math.log(num_samples)

    def posterior_predictive_mean(self, response, mask, num_samples=100, seed=None):
        """mean over num_samples posterior draws of decode(ability_s, item_feat_s) -> (P, I, 1):
        sample_posterior_predictive + .mean(0) of the reference (vibo.py:349-390, :515) in one kernel
        (``vibo_predictive_mean``) instead of S decodes stacked on the host."""
        with torch.no_grad():
            _, a_mu, a_lv, _, i_mu, i_lv = self.encode(response, mask)
            if self.generative_model != 'irt':
                acc = 0.0
                for _ in range(num_samples):
                    acc = acc + self.decode(a_mu + torch.exp(0.5 * a_lv) * torch.randn_like(a_mu),
                                            i_mu + torch.exp(0.5 * i_lv) * torch.randn_like(i_mu))
                return acc / num_samples
            if seed is None:
                seed = int(torch.randint(0, 2 ** 62, (1,)).item())
            out = VF.K.predictive_mean(a_mu, a_lv, i_mu, i_lv, num_samples, irt_model=self.irt_num, seed=seed)
            return out.unsqueeze(2)

    # ------------------------------------------------------------ fused entry
    def fused_elbo(self, response, mask, annealing_factor=1, use_kl_divergence=True, eps_item=None,
                   eps_ability=None, seed=None, person_offset=0, item_term_scale=1.0,
                   return_outputs=False):
        """-ELBO of ``forward`` + ``elbo`` (reference vibo.py:264-266) without
        materialising response_mu.

        eps_item / eps_ability: optional pre-drawn noise (draw order of the
        reference: items (I, F) first, then abilities (P, D)).  With
        ``seed`` given and no ``eps_ability`` the ability noise is drawn inside
        the kernel (Philox keyed by ``person_offset`` + row).
        item_term_scale: weight of the item-side prior term; a person-sharded
        run passes 1/world_size so that the all-reduced sum counts it once.
        """
        resp, msk = VF.prepare_rows(response, mask, binary=self.response_dist == 'bernoulli')
        P = resp.shape[0]
        item_feat_mu, item_feat_logvar = self.item_encoder()
        host_rows = (not resp.is_cuda) and item_feat_mu.is_cuda
        if host_rows:
            if self.n_norm_flows > 0 or return_outputs or self.ability_merge == 'mean':
                raise NotImplementedError("host-resident rows: flows / mean merge / return_outputs need device rows")
            if eps_ability is None and seed is None:
                seed = int(torch.randint(0, 2 ** 62, (1,)).item())
        if eps_item is None:
            eps_item = torch.randn_like(item_feat_mu)
        kl_form = bool(use_kl_divergence)
        general = self.generative_model != 'irt' or self.response_dist != 'bernoulli'
        if general and host_rows:
            raise NotImplementedError("host-resident rows need generative_model='irt' and Bernoulli responses")
        # unconditional model on the GPU: the whole parameter-side chain is two small kernels
        if self.ability_merge == 'mean' or self.ability_encoder.per_cell:
            item_feat = eps_item * torch.exp(0.5 * item_feat_logvar) + item_feat_mu
            if eps_ability is None and seed is None:
                eps_ability = torch.randn(P, self.ability_dim, dtype=torch.float32, device=resp.device)
            return self._composed_elbo(resp, msk, None, item_feat, item_feat_mu, item_feat_logvar, eps_ability,
                                       seed, person_offset, item_term_scale, return_outputs,
                                       float(annealing_factor), kl_form)
        chain = (self.fuse_param_chain and item_feat_mu.is_cuda and not self.conditional_posterior
                 and self.hidden_dim <= 256 and not general)   # with flows: draw + expert table only
        if chain:
            mlp = self.ability_encoder.mlp
            item_feat, table, item_term_raw = VF.ParamChain.apply(
                item_feat_mu, item_feat_logvar, mlp[0].weight, mlp[0].bias, mlp[2].weight, mlp[2].bias,
                mlp[4].weight, mlp[4].bias, eps_item, self.irt_num,
                VF.ELBO_KL if kl_form else VF.ELBO_SAMPLE)
        else:
            item_feat = eps_item * torch.exp(0.5 * item_feat_logvar) + item_feat_mu
            table = self.ability_encoder.expert_table(item_feat if self.conditional_posterior else None)
        if eps_ability is None and seed is None:
            eps_ability = torch.randn(P, self.ability_dim, dtype=torch.float32, device=resp.device)
        beta = float(annealing_factor)

        if self.n_norm_flows > 0 or general:
            return self._composed_elbo(resp, msk, table, item_feat, item_feat_mu, item_feat_logvar,
                                       eps_ability, seed, person_offset, item_term_scale, return_outputs,
                                       beta, kl_form, eps_item=eps_item)

        cfg = dict(irt_model=self.irt_num, conditional=self.conditional_posterior,
                   missing_policy=self.ability_encoder.missing_policy,
                   elbo_form=VF.ELBO_KL if use_kl_divergence else VF.ELBO_SAMPLE, beta=beta,
                   seed=0 if seed is None else (seed if isinstance(seed, torch.Tensor) else int(seed)),
                   person_offset=int(person_offset),
                   want_person_outputs=return_outputs)
        if host_rows:
            cfg["chunk_person"] = getattr(self, "host_chunk_person", 65536)
            cfg["staging"] = getattr(self, "_host_staging", None)
            loss_k, scalars, scalars_host = VF.FusedElboHost.apply(resp, msk, table, item_feat,
                                                                   eps_ability, cfg)
            self._host_staging = cfg["staging"]
            a_mu = a_lv = ability = None
        else:
            loss_k, scalars, a_mu, a_lv, ability = VF.FusedElbo.apply(
                resp, msk, table, item_feat, eps_ability, cfg)
        if chain:
            item_term = beta * item_term_raw if use_kl_divergence else item_term_raw
        elif use_kl_divergence:
            item_term = beta * kl_divergence_standard_normal_prior(item_feat_mu, item_feat_logvar).sum()
        else:
            item_term = -(standard_normal_log_pdf(item_feat).sum()
                          - normal_log_pdf(item_feat, item_feat_mu, item_feat_logvar).sum())
        loss = loss_k + item_term_scale * item_term
        if return_outputs:
            return loss, dict(scalars=scalars, ability=ability, ability_mu=a_mu, ability_logvar=a_lv,
                              item_feat=item_feat, item_feat_mu=item_feat_mu,
                              item_feat_logvar=item_feat_logvar)
        return loss

    def _composed_elbo(self, resp, msk, table, item_feat, item_feat_mu, item_feat_logvar, eps_ability,
                       seed, person_offset, item_term_scale, return_outputs, beta, kl_form, eps_item=None):
        """-ELBO composed in autograd around the kernels: ability posterior (encode kernel for the
        product merge, PyTorch reductions / GEMMs on the collapsed table for the mean merge) ->
        draw [-> planar flows, fused per person on the GPU] -> link / log-likelihood kernel ->
        prior terms.  Used for --n-norm-flows (reference models.py:406-424, where
        annealing_factor is ignored exactly as in the reference) and --ability-merge mean."""
        enc = self.ability_encoder
        if enc.per_cell:
            a_mu, a_lv = enc._forward_per_cell(resp, msk, item_feat if self.conditional_posterior else None)
        elif self.ability_merge == 'mean':
            a_mu, a_lv = enc._forward_mean(resp, msk, item_feat if self.conditional_posterior else None)
        else:
            a_mu, a_lv = VF.EncodePosterior.apply(resp, msk, table, enc.conditional, enc.missing_policy)
        if eps_ability is None:
            # the noise the fused kernels draw in-kernel: Philox keyed by (seed, global person index)
            eps_ability = VF.K.philox_normal(a_mu.shape[0], a_mu.shape[1], seed, int(person_offset), resp.device)
        outputs = dict(ability_mu=a_mu, ability_logvar=a_lv, item_feat=item_feat, item_feat_mu=item_feat_mu,
                       item_feat_logvar=item_feat_logvar)
        if self.n_norm_flows > 0:
            fused_flows = (resp.is_cuda and self.n_norm_flows <= 8
                           and isinstance(self.ability_norm_flows.flows[0], PlanarFlow))
            item = None
            if fused_flows and eps_item is not None and self.item_feat_dim <= 8:
                # the item side is the same computation on I rows of width F: draw, K planar flows and
                # log N(z_K; 0, 1) - log N(z_0; mu, exp lv) + sum_k ldj_k in one kernel each way
                uhat, fw, fb = self.item_norm_flows.stacked_parameters()
                _, item_k, item = VF.FlowPerson.apply(item_feat_mu, item_feat_logvar, eps_item, uhat, fw, fb)
            else:
                item_k, i_ldj = self.item_norm_flows(item_feat)
            if fused_flows:
                # draw + K planar flows + person-side terms in one kernel each way
                uhat, fw, fb = self.ability_norm_flows.stacked_parameters()
                ability, ability_k, person = VF.FlowPerson.apply(a_mu, a_lv, eps_ability, uhat, fw, fb)
            else:
                ability = eps_ability * torch.exp(0.5 * a_lv) + a_mu
                ability_k, a_ldj = self.ability_norm_flows(ability)
                person = standard_normal_log_pdf(ability_k).sum() \
                    - (normal_log_pdf(ability, a_mu, a_lv).sum() - a_ldj.sum())
            ll = self._loglik(resp, msk, ability_k, item_k)
            if item is None:
                item = standard_normal_log_pdf(item_k).sum() \
                    - (normal_log_pdf(item_feat, item_feat_mu, item_feat_logvar).sum() - i_ldj.sum())
            loss = -(ll + person + item_term_scale * item)
            outputs.update(ability=ability, ability_k=ability_k, item_feat_k=item_k)
        else:
            ability = eps_ability * torch.exp(0.5 * a_lv) + a_mu
            ll = self._loglik(resp, msk, ability, item_feat)
            if kl_form:
                person = -beta * kl_divergence_standard_normal_prior(a_mu, a_lv).sum()
                item = -beta * kl_divergence_standard_normal_prior(item_feat_mu, item_feat_logvar).sum()
            else:
                person = standard_normal_log_pdf(ability).sum() - normal_log_pdf(ability, a_mu, a_lv).sum()
                item = standard_normal_log_pdf(item_feat).sum() \
                    - normal_log_pdf(item_feat, item_feat_mu, item_feat_logvar).sum()
            loss = -(ll + person + item_term_scale * item)
            outputs.update(ability=ability)
        if return_outputs:
            return loss, outputs
        return loss


class VIBO_2PL(VIBO_1PL):
    irt_num = 2


class VIBO_3PL(VIBO_2PL):
    irt_num = 3


# ---------------------------------------------------------------------------------------------
# Un-amortized variational IRT (reference models.py:89-243): per-person posterior tables instead
# of an inference network.  Same generative side, so the same link / log-likelihood kernel.
# ---------------------------------------------------------------------------------------------
class VI_1PL(nn.Module):
    irt_num = 1

    def __init__(self, latent_dim, num_person, num_item):
        super().__init__()
        self.latent_dim = latent_dim
        self.ability_dim = latent_dim
        self.response_dim = 1
        self.num_person = num_person
        self.num_item = num_item
        self.item_feat_dim = {1: 1, 2: latent_dim + 1, 3: latent_dim + 2}[self.irt_num]
        # construction order == reference (models.py:107-113): N(0, 1) embeddings, RNG-compatible
        self.ability_mu_lookup = nn.Embedding(num_person, self.ability_dim)
        self.ability_logvar_lookup = nn.Embedding(num_person, self.ability_dim)
        self.item_mu_lookup = nn.Embedding(num_item, self.item_feat_dim)
        self.item_logvar_lookup = nn.Embedding(num_item, self.item_feat_dim)

    reparameterize_gaussian = staticmethod(VIBO_1PL.reparameterize_gaussian)

    def encode(self, index, response, mask):
        """reference models.py:128-143 -> 6-tuple (item draw first, then abilities)."""
        item_feat_mu, item_feat_logvar = self.item_mu_lookup.weight, self.item_logvar_lookup.weight
        item_feat = self.reparameterize_gaussian(item_feat_mu, item_feat_logvar)
        idx = index.reshape(-1).long()
        ability_mu, ability_logvar = self.ability_mu_lookup(idx), self.ability_logvar_lookup(idx)
        ability = self.reparameterize_gaussian(ability_mu, ability_logvar)
        return ability, ability_mu, ability_logvar, item_feat, item_feat_mu, item_feat_logvar

    def decode(self, ability, item_feat):
        return VF.Decode.apply(ability, item_feat, self.irt_num)

    def forward(self, index, response, mask):
        ability, ability_mu, ability_logvar, item_feat, item_feat_mu, item_feat_logvar = \
            self.encode(index, response, mask)
        return (response, mask, self.decode(ability, item_feat), ability, ability_mu, ability_logvar,
                item_feat, item_feat_mu, item_feat_logvar)

    def _terms(self, ll, ability, ability_mu, ability_logvar, item_feat, item_feat_mu, item_feat_logvar,
               annealing_factor, use_kl_divergence):
        if use_kl_divergence:
            kl_u = kl_divergence_standard_normal_prior(ability_mu, ability_logvar).sum()
            kl_d = kl_divergence_standard_normal_prior(item_feat_mu, item_feat_logvar).sum()
            return ll - annealing_factor * kl_u - annealing_factor * kl_d
        log_p = standard_normal_log_pdf(ability).sum() + standard_normal_log_pdf(item_feat).sum()
        log_q = normal_log_pdf(ability, ability_mu, ability_logvar).sum() \
            + normal_log_pdf(item_feat, item_feat_mu, item_feat_logvar).sum()
        return (ll + log_p) - log_q

    def elbo(self, response, mask, response_mu, ability, ability_mu, ability_logvar, item_feat,
             item_feat_mu, item_feat_logvar, annealing_factor=1, use_kl_divergence=True):
        """reference models.py:148-177 on a materialised response_mu -> -ELBO (0-d)."""
        resp, msk = VF.prepare_rows(response, mask)
        ll = VF.BernoulliLogLik.apply(resp, msk, response_mu)
        return -self._terms(ll, ability, ability_mu, ability_logvar, item_feat, item_feat_mu, item_feat_logvar,
                            annealing_factor, use_kl_divergence)

    def fused_elbo(self, index, response, mask, annealing_factor=1, use_kl_divergence=True):
        """forward + elbo (reference vi.py training step) without materialising response_mu: the
        link + log-likelihood kernel (vibo_link_loglik) on the drawn abilities / item features."""
        resp, msk = VF.prepare_rows(response, mask)
        ability, ability_mu, ability_logvar, item_feat, item_feat_mu, item_feat_logvar = \
            self.encode(index, response, mask)
        ll = VF.LinkLogLik.apply(resp, msk, ability, item_feat, self.irt_num)
        return -self._terms(ll, ability, ability_mu, ability_logvar, item_feat, item_feat_mu, item_feat_logvar,
                            annealing_factor, use_kl_divergence)

    def log_marginal(self, index, response, mask, num_samples=100):
        """reference models.py:179-213 (which omits ``index`` in its forward call and cannot run as
        written): logsumexp over sample-form -ELBOs - log(num_samples)."""
        with torch.no_grad():
            log_w = torch.stack([-self.fused_elbo(index, response, mask, use_kl_divergence=False)
                                 for _ in range(num_samples)])
            return torch.logsumexp(log_w, 0) - math.log(num_samples)


class VI_2PL(VI_1PL):
    irt_num = 2


class VI_3PL(VI_2PL):
    irt_num = 3
