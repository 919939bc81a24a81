"""Oracle (test infrastructure): torch-CPU restatement of FOCOPS' learn
(/root/reference/fsrl/policy/focops.py:157-251) with autograd, on plain numpy batches.
Parity unpinned: the reference cannot be imported here (tianshou / gymnasium absent) and pins no
numbers for FOCOPS; this follows the cited lines."""
from __future__ import annotations

import numpy as np
import torch
from torch import nn
from torch.distributions import Independent, Normal, kl_divergence

from .cpo import process as cpo_process
from .ppo import split_indices


def process(actor, critics, batch, gamma, gae_lambda):
    """focops.py:135-155: GAE (no whole-batch normalisation) + old log-prob / mean / std."""
    return cpo_process(actor, critics, batch, gamma, gae_lambda, norm_adv=False)


def nu_step(nu, nu_lr, nu_max, cost_limit, ave_cost):
    """focops.py:157-162 (f32 tensor arithmetic)."""
    loss_nu = cost_limit - ave_cost
    nu_t = torch.tensor([nu], dtype=torch.float32)
    nu_t += -nu_lr * loss_nu
    nu_t = torch.clamp(nu_t, 0, nu_max)
    return float(nu_t.item()), loss_nu


def learn(actor, critics, actor_optim, critics_optim, batch, batch_size, repeat, nu, l2_reg=1e-3, delta=0.02,
          eta=0.02, tem_lambda=0.95, max_grad_norm=0.5, norm_adv=True):
    t = lambda k: torch.from_numpy(np.ascontiguousarray(batch[k]))
    obs_all, act_all, lpo_all, advs_all, rets = t("obs"), t("act"), t("logp_old"), t("advs"), t("rets")
    mo_all, so_all = t("mean_old"), t("std_old")
    n = obs_all.shape[0]
    stats = []
    for step in range(repeat):
        iters, approx_kl = 0, 0.0
        for idx in split_indices(n, batch_size):
            ix = torch.from_numpy(idx)
            obs, act, lpo = obs_all[ix], act_all[ix], lpo_all[ix]
            advs = advs_all[ix].clone()
            # critics (:164-180)
            loss = 0
            sc = {}
            for i, c in enumerate(critics):
                vf = (rets[ix, i] - c(obs).flatten()).pow(2).mean()
                for p in c.parameters():
                    vf = vf + p.pow(2).sum() * l2_reg
                loss = loss + vf
                sc["loss/vf" + str(i)] = vf.item()
            critics_optim.zero_grad(); loss.backward(); critics_optim.step()
            sc["loss/vf_total"] = loss.item()
            # actor (:182-215)
            mu, sigma = actor(obs)
            dist = Independent(Normal(mu, sigma), 1)
            ent = dist.entropy().mean()
            ratio = (dist.log_prob(act) - lpo).exp()
            kl_new_old = kl_divergence(dist, Independent(Normal(mo_all[ix], so_all[ix]), 1))
            if norm_adv:
                for i in range(advs.shape[1]):
                    a = advs[:, i]
                    advs[:, i] = (a - a.mean()) / a.std()
            loss_actor = ((kl_new_old - 1 / tem_lambda * ratio * (advs[:, 0] - nu * advs[:, 1])) *
                          (kl_new_old.detach() <= eta)).mean()
            actor_optim.zero_grad()
            loss_actor.backward()
            if max_grad_norm:
                nn.utils.clip_grad_norm_(actor.parameters(), max_norm=max_grad_norm)
            actor_optim.step()
            stats.append({"loss/actor_loss": loss_actor.item(), "loss/kl": kl_new_old.mean().item(),
                          "loss/entropy": ent.item(), **sc})
            approx_kl += stats[-1]["loss/kl"]
            iters += 1
        approx_kl /= iters + 1e-7
        if approx_kl > delta:
            break
    return stats
