"""Oracle (test infrastructure): torch-CPU restatement of the PPO-Lagrangian update,
following /root/reference/fsrl/policy/ppo_lag.py:134-257, lagrangian_base.py:145-166 and
base_policy.py:384-451 with autograd + torch.optim.Adam, on plain numpy batches.  The
minibatch permutation is drawn exactly like tianshou's Batch.split (np.random.permutation on
the global RNG, tail merged into the previous chunk -- SURVEY.md 2.3 [UNVERIFIED])."""
from __future__ import annotations

import numpy as np
import torch
from torch.distributions import Independent, Normal

from . import returns


def split_indices(n, size, shuffle=True, merge_last=True):
    idx = np.random.permutation(n) if shuffle else np.arange(n)
    merge_last = merge_last and n % size > 0
    out = []
    for i in range(0, n, size):
        if merge_last and i + size + size >= n:
            out.append(idx[i:])
            break
        out.append(idx[i:i + size])
    return out


def process(actor, critics, batch, gamma, gae_lambda, unfinished=None):
    """compute_gae_returns + logp_old (ppo_lag.py:134-150).  batch: dict of numpy arrays in
    batch order (obs, obs_next, act, rew, cost, terminated, truncated)."""
    obs = torch.from_numpy(batch["obs"]); obs_next = torch.from_numpy(batch["obs_next"])
    n = obs.shape[0]
    unf = np.zeros(n, bool) if unfinished is None else unfinished
    with torch.no_grad():
        v = np.stack([c(obs).flatten().numpy() for c in critics])
        vn = np.stack([c(obs_next).flatten().numpy() for c in critics])
        values, rets, advs = returns.dual_gae(v, vn, batch["rew"], batch["cost"], batch["terminated"],
                                              batch["truncated"], unf, gamma, gae_lambda)
        mu, sigma = actor(obs)
        logp_old = Independent(Normal(mu, sigma), 1).log_prob(torch.from_numpy(batch["act"])).numpy()
    return dict(batch, values=values, rets=rets, advs=advs, logp_old=logp_old)


def learn(actor, critics, optim, batch, batch_size, repeat, lagrangian, rescaling=True,
          eps_clip=0.2, vf_coef=0.25, max_grad_norm=None, target_kl=0.02, norm_adv=True,
          dual_clip=None, use_lagrangian=True, max_steps=None, value_clip=False, grads_out=None):
    """ppo_lag.py:214-257.  Returns a list of per-minibatch stat dicts (un-averaged)."""
    obs = torch.from_numpy(batch["obs"]); act = torch.from_numpy(batch["act"])
    logp_old_all = torch.from_numpy(batch["logp_old"])
    advs_all = torch.from_numpy(batch["advs"].copy()); rets_all = torch.from_numpy(batch["rets"])
    C = advs_all.shape[1]
    params = [p for m in [actor] + list(critics) for p in m.parameters()]
    stats = []
    n = obs.shape[0]
    if grads_out is not None:
        max_steps = 1
    resc = 1.0 / (lagrangian + 1.0) if (rescaling and use_lagrangian) else 1.0
    for step in range(repeat):
        kl_sum, iters = 0.0, 0
        for idx in split_indices(n, batch_size):
            idx_t = torch.from_numpy(idx)
            mu, sigma = actor(obs[idx_t])
            dist = Independent(Normal(mu, sigma), 1)
            log_p = dist.log_prob(act[idx_t])
            ratio = (log_p - logp_old_all[idx_t]).exp()
            if ratio.dtype != torch.float64:            # the fp64 twin (drift studies) keeps its precision
                ratio = ratio.float()                   # ppo_lag.py:176
            adv = advs_all[idx_t].clone()
            if norm_adv:                                                     # :178-182
                for i in range(C):
                    a = adv[:, i]
                    adv[:, i] = (a - a.mean()) / a.std()
                # the reference normalises batch.advs of the *minibatch copy* in place
            rew_adv = adv[:, 0]
            surr1 = ratio * rew_adv
            surr2 = ratio.clamp(1.0 - eps_clip, 1.0 + eps_clip) * rew_adv
            if dual_clip:
                clip1 = torch.min(surr1, surr2)
                clip2 = torch.max(clip1, dual_clip * rew_adv)
                loss_rew = -torch.where(rew_adv < 0, clip2, clip1).mean()
            else:
                loss_rew = -torch.min(surr1, surr2).mean()
            loss_saf = torch.zeros(())
            if use_lagrangian and C > 1:
                loss_saf = torch.mean(ratio * adv[:, 1] * lagrangian)          # lagrangian_base.py:161
            loss_actor = resc * (loss_rew + loss_saf)
            vf_losses = []
            for i, c in enumerate(critics):
                value = c(obs[idx_t]).flatten()
                if value_clip:                                                   # :156-163
                    v_old = torch.from_numpy(batch["values"])[idx_t, i]
                    v_clip = v_old + (value - v_old).clamp(-eps_clip, eps_clip)
                    vf1 = (rets_all[idx_t, i] - value).pow(2)
                    vf2 = (rets_all[idx_t, i] - v_clip).pow(2)
                    vf_losses.append(torch.max(vf1, vf2).mean())
                else:
                    vf_losses.append((rets_all[idx_t, i] - value).pow(2).mean())
            loss_vf = sum(vf_losses)
            loss = loss_actor + vf_coef * loss_vf
            optim.zero_grad()
            loss.backward()
            if grads_out is not None:                   # un-clipped gradients of the first minibatch (parity tests)
                grads_out.extend(p.grad.detach().clone() for p in params)
            gn = None
            if max_grad_norm:
                gn = torch.nn.utils.clip_grad_norm_(params, max_norm=max_grad_norm)
            else:
                gn = torch.sqrt(sum((p.grad ** 2).sum() for p in params))
            optim.step()
            kl = (logp_old_all[idx_t] - log_p).mean().item()
            kl_sum += kl; iters += 1
            stats.append({"loss/actor_rew": loss_rew.item(), "loss/actor_safety": float(loss_saf),
                          "loss/actor_total": loss_actor.item(), "loss/kl": kl,
                          "loss/vf_total": loss_vf.item(), "loss/total": loss.item(),
                          "loss/entropy": dist.entropy().mean().item(), "loss/grad_norm": float(gn),
                          **{f"loss/vf{i}": v.item() for i, v in enumerate(vf_losses)}})
            if max_steps is not None and len(stats) >= max_steps:
                return stats
        if kl_sum / (iters + 1e-7) > 1.5 * target_kl:                         # :251-255
            break
    return stats
