"""Oracle (test infrastructure): torch-CPU restatement of the off-policy Lagrangian learners.

  compute_nstep_returns  <- /root/reference/fsrl/policy/base_policy.py:453-512 (+ :543-567)
  SAC-Lag                <- /root/reference/fsrl/policy/sac_lag.py:136-269
  DDPG-Lag               <- /root/reference/fsrl/policy/ddpg_lag.py:120-223
  safety_loss            <- /root/reference/fsrl/policy/lagrangian_base.py:145-166

Works on an oracle.collector.OracleBuffer-shaped object (numpy SoA, env-major rings).  The
reparameterisation noise is injected (``eps``) so that the device's documented Philox stream
can be replayed."""
from __future__ import annotations

import numpy as np
import torch

from . import returns


def buffer_next(buf, idx):
    """tianshou ReplayBufferManager.next: stay at a done transition or at the newest slot."""
    idx = np.asarray(idx, dtype=np.int64)
    env = idx // buf.cap
    newest = env * buf.cap + (buf.ptr[env] - 1) % buf.cap
    done = buf.terminated[idx] | buf.truncated[idx]
    nxt = env * buf.cap + (idx % buf.cap + 1) % buf.cap
    return np.where(done | (idx == newest), idx, nxt)


def unfinished_flags(buf):
    end = (buf.terminated | buf.truncated).copy()
    for e in range(buf.E):
        if buf.len[e] > 0:
            last = e * buf.cap + (buf.ptr[e] - 1) % buf.cap
            end[last] = True                                           # :493
    return end


def nstep_targets(buf, idx, target_q_list, gamma, n_step):
    """compute_nstep_returns (:481-509): target_q_list[i] is (bsz,) numpy evaluated at the terminal
    indices; returns rets (bsz, C) float32."""
    metrics = [buf.rew.astype(np.float64), buf.cost.astype(np.float64)]
    indices = [np.asarray(idx, np.int64)]
    for _ in range(n_step - 1):
        indices.append(buffer_next(buf, indices[-1]))
    indices = np.stack(indices)
    terminal = indices[-1]
    vmask = (~buf.terminated[terminal]).reshape(-1, 1)
    end_flag = unfinished_flags(buf)
    out = []
    for i, tq in enumerate(target_q_list):
        t = np.asarray(tq, np.float32).reshape(len(idx), -1) * vmask
        r = returns.nstep_return(metrics[i], end_flag, t, indices, gamma, n_step)
        out.append(r.astype(np.float32).reshape(-1))
    return np.stack(out, -1), terminal


def sac_forward(actor, obs, eps, deterministic=False):
    mu, sigma = actor(obs)
    dist = torch.distributions.Independent(torch.distributions.Normal(mu, sigma), 1)
    act = mu if deterministic else mu + sigma * eps                      # dist.rsample() with injected noise
    log_prob = dist.log_prob(act).unsqueeze(-1)
    sq = torch.tanh(act)
    log_prob = log_prob - torch.log((1 - sq.pow(2)) + np.finfo(np.float32).eps.item()).sum(-1, keepdim=True)
    return sq, log_prob


def sac_update(actor, critics, critics_old, actor_opt, critic_opt, obs, act, rets, eps_cur, *, alpha, tau,
               lagrangian, rescaling=True, use_lagrangian=True, auto_alpha=None):
    """SACLagrangian.learn (sac_lag.py:260-269) on a batch whose n-step targets `rets` [B, C] are given:
    critics_loss (:185-210), policy_loss (:212-258, reparameterisation noise eps_cur), sync_weight (:132-134)."""
    C = len(critics)
    stats = {}
    loss_c = 0
    for i in range(C):
        li = 0
        for j in range(2):
            td = critics[i][j](obs, act).flatten() - rets[:, i]
            li = li + td.pow(2).mean()
        loss_c = loss_c + li
        stats[f"loss/q{i}"] = float(li)
    critic_opt.zero_grad(); loss_c.backward(); critic_opt.step()
    stats["loss/q_total"] = float(loss_c)
    a, lp = sac_forward(actor, obs, eps_cur)
    q = torch.min(critics[0][0](obs, a), critics[0][1](obs, a)).flatten()
    loss_rew = (alpha * lp.flatten() - q).mean()
    loss_saf = torch.zeros(())
    if use_lagrangian and C > 1:
        qc = torch.min(critics[1][0](obs, a), critics[1][1](obs, a)).flatten()
        loss_saf = torch.mean(qc * lagrangian)
    resc = 1.0 / (lagrangian + 1.0) if (rescaling and use_lagrangian) else 1.0
    total = resc * (loss_rew + loss_saf)
    actor_opt.zero_grad(); total.backward(); actor_opt.step()
    stats.update({"loss/actor_rew": float(loss_rew), "loss/actor_safety": float(loss_saf),
                  "loss/actor_total": float(total)})
    new_alpha = alpha
    if auto_alpha is not None:
        target_entropy, log_alpha, alpha_opt = auto_alpha
        lpd = lp.detach() + target_entropy
        alpha_loss = -(log_alpha * lpd).mean()
        alpha_opt.zero_grad(); alpha_loss.backward(); alpha_opt.step()
        new_alpha = float(log_alpha.detach().exp())
        stats["loss/alpha_loss"] = float(alpha_loss); stats["loss/alpha_value"] = new_alpha
    with torch.no_grad():
        for i in range(C):
            for j in range(2):
                for tp, sp in zip(critics_old[i][j].parameters(), critics[i][j].parameters()):
                    tp.copy_(tau * sp + (1 - tau) * tp)
    return stats, new_alpha


def sac_step(actor, critics, critics_old, actor_opt, critic_opt, buf, idx, eps_next, eps_cur, *, alpha,
             gamma, n_step, tau, lagrangian, rescaling=True, use_lagrangian=True, auto_alpha=None):
    """One SACLagrangian.update: critics = list of (q1, q2) module pairs per return stream."""
    C = len(critics)
    with torch.no_grad():
        # process_fn -> _target_q at the n-step terminal indices (:136-145)
        _, terminal = nstep_targets(buf, idx, [np.zeros(len(idx))] * C, gamma, n_step)
        obs_next = torch.from_numpy(buf.obs_next[terminal])
        a_next, lp_next = sac_forward(actor, obs_next, eps_next)
        tq = []
        for i in range(C):
            q = torch.min(critics_old[i][0](obs_next, a_next), critics_old[i][1](obs_next, a_next))
            tq.append((q - alpha * lp_next).numpy())
    rets, _ = nstep_targets(buf, idx, tq, gamma, n_step)
    rets = torch.from_numpy(rets)
    obs = torch.from_numpy(buf.obs[idx]); act = torch.from_numpy(buf.act[idx])
    return sac_update(actor, critics, critics_old, actor_opt, critic_opt, obs, act, rets, eps_cur, alpha=alpha, tau=tau,
                      lagrangian=lagrangian, rescaling=rescaling, use_lagrangian=use_lagrangian, auto_alpha=auto_alpha)


def ddpg_update(actor, actor_old, critics, critics_old, actor_opt, critic_opt, obs, act, rets, *, tau, lagrangian,
                rescaling=True, use_lagrangian=True):
    """DDPGLagrangian.learn (ddpg_lag.py:215-223) on a batch with given n-step targets `rets` [B, C]:
    critics_loss (:165-189), policy_loss (:191-213), sync_weight (:120-123)."""
    C = len(critics)
    stats = {}
    loss_c = 0
    for i in range(C):
        td = critics[i](obs, act).flatten() - rets[:, i]
        li = td.pow(2).mean()
        loss_c = loss_c + li
        stats[f"loss/q{i}"] = float(li)
    critic_opt.zero_grad(); loss_c.backward(); critic_opt.step()
    stats["loss/q_total"] = float(loss_c)
    a = actor(obs)
    loss_rew = -critics[0](obs, a).mean()
    loss_saf = torch.zeros(())
    if use_lagrangian and C > 1:
        loss_saf = torch.mean(critics[1](obs, a).mean() * lagrangian)
    resc = 1.0 / (lagrangian + 1.0) if (rescaling and use_lagrangian) else 1.0
    total = resc * (loss_rew + loss_saf)
    actor_opt.zero_grad(); total.backward(); actor_opt.step()
    stats.update({"loss/actor_rew": float(loss_rew), "loss/actor_safety": float(loss_saf),
                  "loss/actor_total": float(total)})
    with torch.no_grad():
        for tp, sp in zip(actor_old.parameters(), actor.parameters()):
            tp.copy_(tau * sp + (1 - tau) * tp)
        for i in range(C):
            for tp, sp in zip(critics_old[i].parameters(), critics[i].parameters()):
                tp.copy_(tau * sp + (1 - tau) * tp)
    return stats


def ddpg_step(actor, actor_old, critics, critics_old, actor_opt, critic_opt, buf, idx, *, gamma, n_step,
              tau, lagrangian, rescaling=True, use_lagrangian=True):
    C = len(critics)
    with torch.no_grad():
        _, terminal = nstep_targets(buf, idx, [np.zeros(len(idx))] * C, gamma, n_step)
        obs_next = torch.from_numpy(buf.obs_next[terminal])
        a_next = actor_old(obs_next)
        tq = [critics_old[i](obs_next, a_next).numpy() for i in range(C)]
    rets, _ = nstep_targets(buf, idx, tq, gamma, n_step)
    rets = torch.from_numpy(rets)
    obs = torch.from_numpy(buf.obs[idx]); act = torch.from_numpy(buf.act[idx])
    return ddpg_update(actor, actor_old, critics, critics_old, actor_opt, critic_opt, obs, act, rets, tau=tau,
                       lagrangian=lagrangian, rescaling=rescaling, use_lagrangian=use_lagrangian)
