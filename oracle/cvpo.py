"""Oracle (test infrastructure): torch-CPU restatement of CVPO's update
(/root/reference/fsrl/policy/cvpo.py:248-430) on a batch with given n-step targets -- groundwork for SURVEY.md
8(f4); there is no CUDA path for CVPO yet.  Pinned against the reference's own ``CVPO.learn``
(oracle/make_golden_policies.py::golden_cvpo, tests/golden/policy_cvpo_golden.npz): the K action particles the
reference drew from the old policy are recorded and injected here.

One update = critics regression (:248-276), E-step (:320-361: Q of K sampled actions per state, one Adam step on
the dual (eta, lambda) of the temperature / cost constraints, softmax weights), M-step (:363-417: weighted
maximum likelihood with the mean / std contributions decoupled, Lagrange multipliers on KL_mu and KL_std updated
by their own Adam), Polyak update of the target critics (:201-202)."""
from __future__ import annotations

import numpy as np
import torch
from torch.distributions import Independent, Normal

EPS_CLAMP = float(np.finfo(np.float32).eps) * 10                      # cvpo.py:170


def qc_thresholds(cost_limit, gamma, max_episode_steps):
    """Per-step cost-value thresholds of the E-step (:133-136)."""
    limits = [cost_limit] if np.isscalar(cost_limit) else list(cost_limit)
    return [c * (1 - gamma ** max_episode_steps) / (1 - gamma) / max_episode_steps for c in limits]


def _q_min(critic, obs, act):
    """DoubleCritic.predict / SingleCritic.predict: min over the heads, (N, 1)."""
    heads = critic if isinstance(critic, (list, tuple)) else [critic]
    qs = [h(obs, act) for h in heads]
    q = qs[0]
    for other in qs[1:]:
        q = torch.min(q, other)
    return q, qs


def gaussian_kl(mu_old, std_old, mu, std):
    """Decoupled KL terms (:289-316): mean part under the OLD variance, variance part at the old mean."""
    var_old, var = torch.clamp_min(std_old ** 2, 1e-6), torch.clamp_min(std ** 2, 1e-6)
    kl_mu = torch.sum(0.5 * (mu_old - mu) ** 2 / var_old, dim=-1).mean()
    kl_std = torch.sum(0.5 * (torch.log(var / var_old) + var_old / var - 1), dim=-1).mean()
    return kl_mu, kl_std


def cvpo_update(actor, actor_old, critics, critics_old, actor_opt, critic_opt, estep_dual, estep_opt, mstep_duals,
                mstep_opt, obs, act, rets, sample_act, *, qc_thres, estep_kl=0.02, estep_dual_max=20.0,
                mstep_kl_mu=0.005, mstep_kl_std=0.0005, mstep_dual_max=0.5, estep_iters=1, mstep_iters=1, tau=0.05):
    """critics: list (one per return stream) of modules or (q1, q2) pairs; sample_act: (K, B, A) particles drawn
    from the old policy.  Returns the statistics the reference logs for this update."""
    C = len(critics)
    K, B = sample_act.shape[0], obs.shape[0]
    stats = {}
    # ---- critics (:248-276) ----------------------------------------------------------------------------
    loss_c = 0
    for i in range(C):
        _, qs = _q_min(critics[i], obs, act)
        li = 0
        for q in qs:
            li = li + (q.flatten() - rets[:, i]).pow(2).mean()
        loss_c = loss_c + li
        stats[f"loss/loss_q{i}"] = float(li)
        stats[f"estep/val_q{i}"] = float(rets[:, i].mean())
        if i >= 1:
            stats[f"estep/thres_q{i}"] = qc_thres[i - 1]
    critic_opt.zero_grad(); loss_c.backward(); critic_opt.step()
    stats["loss/q_total"] = float(loss_c)
    # ---- E-step (:320-361) -------------------------------------------------------------------------------
    with torch.no_grad():
        mu_old, std_old = actor_old(obs)
        flat_obs = obs[None].expand(K, -1, -1).reshape(K * B, -1)
        flat_act = sample_act.reshape(K * B, -1)
        q_values = [_q_min(critics[i], flat_obs, flat_act)[0].reshape(K, B).T for i in range(C)]      # (B, K)
    for _ in range(estep_iters):
        eta = estep_dual[0]
        loss = eta * estep_kl
        combined = q_values[0].detach()
        for i in range(1, C):
            combined -= estep_dual[i] * q_values[i].detach()       # in place, like the reference (:284)
            loss = loss + estep_dual[i] * qc_thres[i - 1]
        loss = loss + eta * torch.mean(torch.logsumexp(combined / eta, dim=1) - np.log(K))
        estep_opt.zero_grad(); loss.backward(); estep_opt.step()
        stats["loss/estep_loss"] = float(loss)
    estep_dual.data.clamp_(min=EPS_CLAMP, max=estep_dual_max)
    dual = [float(estep_dual[i].detach()) for i in range(C)]
    for i in range(C):
        stats[f"estep/dual{i}"] = dual[i]
    optimal_q = q_values[0].T                                       # (K, B) -- a view: the in-place updates above persist
    for i in range(1, C):
        optimal_q -= dual[i] * q_values[i].T
    weights = torch.softmax(optimal_q / dual[0], dim=0).detach()
    # ---- M-step (:363-417) ---------------------------------------------------------------------------------
    dual_mu_t, dual_std_t = mstep_duals
    for _ in range(mstep_iters):
        mu, std = actor(obs)
        d1, d2 = Independent(Normal(mu, std_old), 1), Independent(Normal(mu_old, std), 1)
        likelihood = d1.expand((K, B)).log_prob(sample_act) + d2.expand((K, B)).log_prob(sample_act)
        loss_mle = -torch.mean(weights * likelihood)
        kl_mu, kl_std = gaussian_kl(mu_old, std_old, mu, std)
        dual_loss = dual_mu_t * (mstep_kl_mu - kl_mu).detach() + dual_std_t * (mstep_kl_std - kl_std).detach()
        mstep_opt.zero_grad(); dual_loss.backward(); mstep_opt.step()
        dmu = float(np.clip(dual_mu_t.item(), 0.0, mstep_dual_max))
        dstd = float(np.clip(dual_std_t.item(), 0.0, mstep_dual_max))
        loss_kl = dmu * (kl_mu - mstep_kl_mu) + dstd * (kl_std - mstep_kl_std)
        loss_actor = loss_mle + loss_kl
        actor_opt.zero_grad(); loss_actor.backward(); actor_opt.step()
        stats.update({"mstep/mstep_kl_mu": float(kl_mu), "mstep/mstep_kl_std": float(kl_std),
                      "mstep/mstep_loss_kl": float(loss_kl), "mstep/mstep_loss_mle": float(loss_mle),
                      "mstep/mstep_loss_total": float(loss_actor), "mstep/mstep_dual_mu": dmu,
                      "mstep/mstep_dual_std": dstd,
                      "mstep/entropy": float(torch.mean(d1.entropy() + d2.entropy()))})
    # ---- target critics (:201-202) -------------------------------------------------------------------------
    with torch.no_grad():
        for i in range(C):
            new = critics[i] if isinstance(critics[i], (list, tuple)) else [critics[i]]
            old = critics_old[i] if isinstance(critics_old[i], (list, tuple)) else [critics_old[i]]
            for qo, qn in zip(old, new):
                for tp, sp in zip(qo.parameters(), qn.parameters()):
                    tp.copy_(tau * sp + (1 - tau) * tp)
    return stats
