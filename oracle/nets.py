"""Oracle (test infrastructure): plain-torch CPU restatement of the tianshou networks the
reference builds (ActorProb / Critic / Actor over Net(MLP); SURVEY.md Appendix C [UNVERIFIED
-- tianshou 0.5 is absent], restated in-repo by the reference at
/root/reference/examples/customized/collect_dataset.py:189-215) and of the agent's
initialisation (/root/reference/fsrl/agent/ppo_lag_agent.py:136-162).  Independent of
fsrl_b200: it only consumes plain weight tensors."""
from __future__ import annotations

from typing import List, Sequence

import torch
from torch import nn


class Body(nn.Module):
    def __init__(self, d_in: int, hidden: Sequence[int]):
        super().__init__()
        dims = [d_in] + list(hidden)
        self.layers = nn.ModuleList(nn.Linear(i, o) for i, o in zip(dims[:-1], dims[1:]))

    def forward(self, x):
        for l in self.layers:
            x = torch.relu(l(x))
        return x


class GaussActor(nn.Module):
    """ActorProb: mu = max_action*tanh(Linear) (bounded), sigma = exp(sigma_param) or
    exp(clamp(Linear, -20, 2))."""

    def __init__(self, d_in, n_act, hidden, max_action=1.0, unbounded=False, conditioned_sigma=False):
        super().__init__()
        self.body = Body(d_in, hidden)
        self.mu = nn.Linear(hidden[-1], n_act)
        self.cond = conditioned_sigma
        if conditioned_sigma:
            self.sigma = nn.Linear(hidden[-1], n_act)
        else:
            self.sigma_param = nn.Parameter(torch.zeros(n_act, 1))
        self.max_action, self.unbounded = max_action, unbounded

    def forward(self, obs):
        h = self.body(obs)
        mu = self.mu(h)
        if not self.unbounded:
            mu = self.max_action * torch.tanh(mu)
        if self.cond:
            sigma = torch.clamp(self.sigma(h), min=-20, max=2).exp()
        else:
            sigma = (self.sigma_param.view(1, -1) + torch.zeros_like(mu)).exp()
        return mu, sigma


class DetActor(nn.Module):
    def __init__(self, d_in, n_act, hidden, max_action=1.0):
        super().__init__()
        self.body = Body(d_in, hidden)
        self.last = nn.Linear(hidden[-1], n_act)
        self.max_action = max_action

    def forward(self, obs):
        return self.max_action * torch.tanh(self.last(self.body(obs)))


class ValueNet(nn.Module):
    """Critic: V(s), or Q(s,a) on cat([obs, act])."""

    def __init__(self, d_in, hidden):
        super().__init__()
        self.body = Body(d_in, hidden)
        self.last = nn.Linear(hidden[-1], 1)

    def forward(self, obs, act=None):
        x = obs if act is None else torch.cat([obs, act], dim=1)
        return self.last(self.body(x))


def load_linear(lin: nn.Linear, weight, bias):
    with torch.no_grad():
        lin.weight.copy_(torch.as_tensor(weight, dtype=torch.float32).cpu())
        lin.bias.copy_(torch.as_tensor(bias, dtype=torch.float32).cpu())


def load_from_state_dict(net: nn.Module, sd: dict, prefix: str):
    """Copy weights out of a product-side state_dict (tianshou key names) into the plain
    oracle module: <prefix>preprocess.model.model.{0,2}.*, <prefix>{mu,sigma,last}.model.0.*,
    <prefix>sigma_param."""
    g = lambda k: sd[prefix + k].detach().cpu()
    load_linear(net.body.layers[0], g("preprocess.model.model.0.weight"), g("preprocess.model.model.0.bias"))
    load_linear(net.body.layers[1], g("preprocess.model.model.2.weight"), g("preprocess.model.model.2.bias"))
    if isinstance(net, GaussActor):
        load_linear(net.mu, g("mu.model.0.weight"), g("mu.model.0.bias"))
        if net.cond:
            load_linear(net.sigma, g("sigma.model.0.weight"), g("sigma.model.0.bias"))
        else:
            with torch.no_grad():
                net.sigma_param.copy_(g("sigma_param"))
    else:
        load_linear(net.last, g("last.model.0.weight"), g("last.model.0.bias"))
    return net
