"""Oracle (test infrastructure): torch-CPU restatement of CPO with autograd (incl. the double
backward of _MVP), following /root/reference/fsrl/policy/cpo.py:123-370 on plain numpy batches.
Returns the un-averaged per-step statistics the reference logs (:335-350)."""
from __future__ import annotations

import numpy as np
import torch
from torch import nn
from torch.distributions import Independent, Normal, kl_divergence

from . import returns
from .ppo import split_indices


def process(actor, critics, batch, gamma, gae_lambda, norm_adv=True):
    obs = torch.from_numpy(batch["obs"]); obs_next = torch.from_numpy(batch["obs_next"])
    with torch.no_grad():
        v = np.stack([c(obs).flatten().numpy() for c in critics])
        vn = np.stack([c(obs_next).flatten().numpy() for c in critics])
        n = obs.shape[0]
        values, rets, advs = returns.dual_gae(v, vn, batch["rew"], batch["cost"], batch["terminated"],
                                              batch["truncated"], np.zeros(n, bool), gamma, gae_lambda)
        advs = torch.from_numpy(advs.copy())
        if norm_adv:                                                       # :127-131
            for i in range(advs.shape[1]):
                a = advs[:, i]
                advs[:, i] = (a - a.mean()) / a.std()
        mu, sigma = actor(obs)
        logp_old = Independent(Normal(mu, sigma), 1).log_prob(torch.from_numpy(batch["act"]))
    return dict(batch, values=values, rets=rets, advs=advs.numpy(), logp_old=logp_old.numpy(),
                mean_old=mu.numpy(), std_old=sigma.numpy())


def _flat_grad(y, model, retain_graph=False, create_graph=False):
    retain_graph = True if create_graph else retain_graph
    grads = torch.autograd.grad(y, model.parameters(), retain_graph=retain_graph, create_graph=create_graph)
    return torch.cat([g.reshape(-1) for g in grads])


def _flat_params(model):
    return torch.cat([p.reshape(-1) for p in model.parameters()])


def _set_flat(model, new):
    n = 0
    for p in model.parameters():
        k = p.numel()
        p.data = new[n:n + k].view(p.size())
        n += k


def critics_step(critics, optim, obs, rets, l2_reg):
    loss = torch.zeros(1)
    stats = {}
    for i, c in enumerate(critics):
        value = c(obs).flatten()
        vf = (rets[:, i] - value).pow(2).mean()
        for p in c.parameters():
            vf = vf + p.pow(2).sum() * l2_reg
        loss = loss + vf
        stats["loss/vf" + str(i)] = vf.item()
    optim.zero_grad(); loss.backward(); optim.step()
    stats["loss/vf_total"] = loss.item()
    return stats


def policy_step(actor, mb, ave_cost_return, cost_limit, delta=0.01, damping=0.1, backtrack=0.8, max_backtracks=10):
    obs, act = mb["obs"], mb["act"]
    mu, sigma = actor(obs)
    dist = Independent(Normal(mu, sigma), 1)
    ent = dist.entropy().mean()
    logp = dist.log_prob(act)
    dist_old = Independent(Normal(mb["mean_old"], mb["std_old"]), 1)
    kl = kl_divergence(dist_old, dist).mean()
    objective = torch.mean(torch.exp(logp - mb["logp_old"]) * mb["adv_r"])
    cost_surrogate = ave_cost_return + torch.mean(torch.exp(logp - mb["logp_old"]) * mb["adv_c"]) - torch.mean(mb["adv_c"])
    grad_g = _flat_grad(objective, actor, retain_graph=True)
    grad_b = _flat_grad(-cost_surrogate, actor, retain_graph=True)
    flat_kl_grad = _flat_grad(kl, actor, create_graph=True)

    def mvp(v):
        kl_v = torch.dot(flat_kl_grad, v)
        return _flat_grad(kl_v, actor, retain_graph=True) + v * damping

    def cg(g, nsteps=10, tol=1e-8):
        x = torch.zeros_like(g); r, p = g.clone(), g.clone()
        rs_old = torch.sum(r * r)
        for _ in range(nsteps):
            z = mvp(p)
            alpha = rs_old / torch.sum(p * z)
            x += alpha * p; r -= alpha * z
            rs_new = torch.sum(r * r)
            if rs_new < tol:
                break
            p = r + (rs_new / rs_old) * p
            rs_old = rs_new
        return x

    H_inv_g = cg(grad_g)
    approx_g = mvp(H_inv_g)
    c_value = cost_surrogate - cost_limit
    EPS = 1e-8
    if torch.dot(grad_b, grad_b) <= EPS and c_value < 0:
        H_inv_b, scalar_r, scalar_s, A_value, B_value = [torch.zeros(1) for _ in range(5)]
        scalar_q = torch.dot(approx_g, H_inv_g)
        optim_case = 4
    else:
        H_inv_b = cg(grad_b)
        approx_b = mvp(H_inv_b)
        scalar_q = torch.dot(approx_g, H_inv_g)
        scalar_r = torch.dot(approx_g, H_inv_b)
        scalar_s = torch.dot(approx_b, H_inv_b)
        A_value = scalar_q - scalar_r ** 2 / scalar_s
        B_value = 2 * delta - c_value ** 2 / scalar_s
        if c_value < 0 and B_value < 0:
            optim_case = 3
        elif c_value < 0 and B_value >= 0:
            optim_case = 2
        elif c_value >= 0 and B_value >= 0:
            optim_case = 1
        else:
            optim_case = 0
    if optim_case in [3, 4]:
        lam = torch.sqrt(scalar_q / (2 * delta)); nu = torch.zeros_like(lam)
    elif optim_case in [1, 2]:
        LA, LB = [0, scalar_r / c_value], [scalar_r / c_value, np.inf]
        LA, LB = (LA, LB) if c_value < 0 else (LB, LA)
        proj = lambda x, L: max(L[0], min(L[1], x))
        lam_a = proj(torch.sqrt(A_value / B_value), LA)
        lam_b = proj(torch.sqrt(scalar_q / (2 * delta)), LB)
        f_a = lambda lam: -0.5 * (A_value / (lam + EPS) + B_value * lam) - scalar_r * c_value / (scalar_s + EPS)
        f_b = lambda lam: -0.5 * (scalar_q / (lam + EPS) + 2 * delta * lam)
        lam = lam_a if f_a(lam_a) >= f_b(lam_b) else lam_b
        lam = torch.as_tensor(lam).clone().detach()
        nu = max(0, (lam * c_value - scalar_r).item()) / (scalar_s + EPS)
    else:
        nu = torch.sqrt(2 * delta / (scalar_s + EPS)); lam = torch.zeros_like(nu)
    with torch.no_grad():
        delta_theta = (1. / (lam + EPS)) * (H_inv_g + nu * H_inv_b) if optim_case > 0 else nu * H_inv_b
        delta_theta /= torch.norm(delta_theta)
        beta = 1.0
        if not torch.isnan(lam):
            init_theta = _flat_params(actor).clone().detach()
            init_obj = objective.clone().detach(); init_cost = cost_surrogate.clone().detach()
            for _ in range(max_backtracks):
                _set_flat(actor, beta * delta_theta + init_theta)
                mu2, sigma2 = actor(obs)
                d2 = Independent(Normal(mu2, sigma2), 1)
                lp2 = d2.log_prob(act)
                new_kl = kl_divergence(dist_old, d2).mean().item()
                new_obj = torch.mean(torch.exp(lp2 - mb["logp_old"]) * mb["adv_r"])
                new_cost = ave_cost_return + torch.mean(torch.exp(lp2 - mb["logp_old"]) * mb["adv_c"]) - torch.mean(mb["adv_c"])
                if new_kl <= delta and (new_obj > init_obj if optim_case > 1 else True) and \
                        new_cost - init_cost <= max(-c_value.item(), 0):
                    break
                beta *= backtrack
    f = lambda t: float(torch.as_tensor(t).reshape(-1)[0])
    return {"loss/kl": kl.item(), "loss/entropy": ent.item(), "loss/rew_loss": objective.item(),
            "loss/cost_loss": cost_surrogate.item(), "loss/optim_A": f(A_value), "loss/optim_B": f(B_value),
            "loss/optim_C": f(c_value), "loss/optim_Q": f(scalar_q), "loss/optim_R": f(scalar_r),
            "loss/optim_S": f(scalar_s), "loss/optim_lam": f(lam), "loss/optim_nu": f(nu),
            "loss/optim_case": optim_case, "loss/step_size": beta,
            "_g": grad_g.detach().numpy(), "_b": grad_b.detach().numpy(), "_Hinv_g": H_inv_g.detach().numpy()}


def learn(actor, critics, optim, batch, batch_size, repeat, ave_cost_return, cost_limit, optim_critic_iters=10,
          l2_reg=1e-3, **kw):
    obs_all = torch.from_numpy(batch["obs"]); n = obs_all.shape[0]
    t = lambda k: torch.from_numpy(np.ascontiguousarray(batch[k]))
    act, lpo, mo, so = t("act"), t("logp_old"), t("mean_old"), t("std_old")
    advs, rets = t("advs"), t("rets")
    stats = []
    for _ in range(repeat):
        for idx in split_indices(n, batch_size):
            ix = torch.from_numpy(idx)
            for _ in range(optim_critic_iters):
                sc = critics_step(critics, optim, obs_all[ix], rets[ix], l2_reg)
            mb = dict(obs=obs_all[ix], act=act[ix], logp_old=lpo[ix], mean_old=mo[ix], std_old=so[ix],
                      adv_r=advs[ix, 0], adv_c=advs[ix, 1])
            sa = policy_step(actor, mb, ave_cost_return, cost_limit, **kw)
            stats.append({**sa, **sc})
    return stats
