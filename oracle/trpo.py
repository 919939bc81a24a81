"""Oracle (test infrastructure): torch-CPU restatement of TRPO-Lagrangian's learn
(/root/reference/fsrl/policy/trpo_lag.py:117-301) with autograd, on plain numpy batches."""
from __future__ import annotations

import numpy as np
import torch
from torch.distributions import Independent, Normal, kl_divergence

from .cpo import _flat_grad, _flat_params, _set_flat, process as cpo_process  # same process_fn shape
from .ppo import split_indices


def process(actor, critics, batch, gamma, gae_lambda, norm_adv=True):
    return cpo_process(actor, critics, batch, gamma, gae_lambda, norm_adv)


def learn(actor, critics, optim, batch, batch_size, repeat, lagrangian, rescaling=True, delta=0.001,
          backtrack=0.8, max_backtracks=10, optim_critic_iters=5, damping=0.1, use_lagrangian=True):
    t = lambda k: torch.from_numpy(np.ascontiguousarray(batch[k]))
    obs_all, act_all, lpo_all, advs, rets = t("obs"), t("act"), t("logp_old"), t("advs"), t("rets")
    n = obs_all.shape[0]
    resc = 1.0 / (lagrangian + 1.0) if (rescaling and use_lagrangian) else 1.0
    stats = []
    for _ in range(repeat):
        for idx in split_indices(n, batch_size):
            ix = torch.from_numpy(idx)
            obs, act, lpo, adv = obs_all[ix], act_all[ix], lpo_all[ix], advs[ix]

            def policy_loss(dist):
                ratio = (dist.log_prob(act) - lpo).exp().float()
                l_rew = -(ratio * adv[:, 0]).mean()
                l_saf = torch.mean(ratio * adv[:, 1] * lagrangian) if use_lagrangian else torch.zeros(())
                return resc * (l_rew + l_saf), l_rew, l_saf

            mu, sigma = actor(obs)
            dist = Independent(Normal(mu, sigma), 1)
            loss_actor, l_rew, l_saf = policy_loss(dist)
            flat_grads = _flat_grad(loss_actor, actor, retain_graph=True).detach()
            with torch.no_grad():
                mo, so = actor(obs)
                old_dist = Independent(Normal(mo, so), 1)
            kl = kl_divergence(old_dist, dist).mean()
            flat_kl_grad = _flat_grad(kl, actor, create_graph=True)

            def mvp(v):
                return _flat_grad((flat_kl_grad * v).sum(), actor, retain_graph=True).detach() + v * damping

            x = torch.zeros_like(flat_grads); r, p = flat_grads.clone(), flat_grads.clone()
            rdotr = r.dot(r)
            for _k in range(10):
                z = mvp(p)
                alpha = rdotr / p.dot(z)
                x += alpha * p; r -= alpha * z
                new = r.dot(r)
                if new < 1e-10:
                    break
                p = r + new / rdotr * p
                rdotr = new
            sd = -x
            step_size = torch.sqrt(2 * delta / (sd * mvp(sd)).sum(0, keepdim=True))
            with torch.no_grad():
                flat_params = _flat_params(actor).clone()
                for i in range(max_backtracks):
                    _set_flat(actor, flat_params + step_size * sd)
                    m2, s2 = actor(obs)
                    nd = Independent(Normal(m2, s2), 1)
                    loss_new, _, _ = policy_loss(nd)
                    kl = kl_divergence(old_dist, nd).mean()
                    if kl < delta and loss_new < loss_actor:
                        break
                    elif i < max_backtracks - 1:
                        step_size = step_size * backtrack
                    else:
                        step_size = torch.tensor([0.0])
            for _c in range(optim_critic_iters):
                loss = 0
                sc = {}
                for i, c in enumerate(critics):
                    vf = (rets[ix, i] - c(obs).flatten()).pow(2).mean()
                    loss = loss + vf
                    sc["loss/vf" + str(i)] = vf.item()
                optim.zero_grad(); loss.backward(); optim.step()
            stats.append({"loss/actor_rew": l_rew.item(), "loss/actor_safety": float(l_saf),
                          "loss/actor_total": loss_actor.item(), "loss/kl": kl.item(),
                          "loss/step_size": float(step_size), **sc})
    return stats
