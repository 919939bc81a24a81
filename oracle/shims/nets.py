"""tianshou 0.5's network classes as the reference constructs them (``Net(state_shape, hidden_sizes=...)``,
``ActorProb(preprocess_net, action_shape, max_action, unbounded, conditioned_sigma)``, ``Critic(preprocess_net)``,
``Actor(preprocess_net, action_shape, max_action)``), restated as plain torch modules from SURVEY.md Appendix C and the
reference's own in-repo restatement (/root/reference/examples/customized/collect_dataset.py:189-215).  Parameter
names follow tianshou's module tree (``preprocess.model.model.{0,2}``, ``mu.model.0``, ``last.model.0``) so that
reference checkpoints keep their keys.  CPU test / baseline infrastructure; tianshou itself is absent."""
from __future__ import annotations

import numpy as np
import torch
from torch import nn

SIGMA_MIN, SIGMA_MAX = -20.0, 2.0


class MLP(nn.Module):
    def __init__(self, input_dim, output_dim=0, hidden_sizes=(), norm_layer=None, activation=nn.ReLU, device=None,
                 linear_layer=nn.Linear, flatten_input=True):
        super().__init__()
        layers, d = [], int(input_dim)
        for h in hidden_sizes:
            layers += [linear_layer(d, int(h)), activation()]
            d = int(h)
        if output_dim > 0:
            layers.append(linear_layer(d, int(output_dim)))
            d = int(output_dim)
        self.output_dim = d
        self.model = nn.Sequential(*layers)
        self.device, self.flatten_input = device, flatten_input

    def forward(self, obs):
        dev = next(self.parameters()).device if len(self.model) else self.device
        obs = torch.as_tensor(obs, device=dev, dtype=torch.float32)
        if self.flatten_input:
            obs = obs.flatten(1)
        return self.model(obs)


class Net(nn.Module):
    def __init__(self, state_shape, action_shape=0, hidden_sizes=(), norm_layer=None, activation=nn.ReLU, device="cpu",
                 softmax=False, concat=False, num_atoms=1, dueling_param=None, linear_layer=nn.Linear):
        super().__init__()
        d_in = int(np.prod(state_shape))
        if concat:
            d_in += int(np.prod(action_shape))
        self.device = device
        self.model = MLP(d_in, 0, hidden_sizes, norm_layer, activation, device, linear_layer)
        self.output_dim = self.model.output_dim

    def forward(self, obs, state=None, info={}):
        return self.model(obs), state


def _last(d_in, d_out, hidden_sizes, device):
    return MLP(d_in, d_out, hidden_sizes, device=device)


class ActorProb(nn.Module):
    def __init__(self, preprocess_net, action_shape, hidden_sizes=(), max_action=1.0, device="cpu", unbounded=False,
                 conditioned_sigma=False, preprocess_net_output_dim=None):
        super().__init__()
        self.preprocess, self.device = preprocess_net, device
        self.output_dim = int(np.prod(action_shape))
        d = getattr(preprocess_net, "output_dim", preprocess_net_output_dim)
        self.mu = _last(d, self.output_dim, hidden_sizes, device)
        self._c_sigma = conditioned_sigma
        if conditioned_sigma:
            self.sigma = _last(d, self.output_dim, hidden_sizes, device)
        else:
            self.sigma_param = nn.Parameter(torch.zeros(self.output_dim, 1))
        self._max, self._unbounded = max_action, unbounded

    def forward(self, obs, state=None, info={}):
        logits, hidden = self.preprocess(obs, state)
        mu = self.mu(logits)
        if not self._unbounded:
            mu = self._max * torch.tanh(mu)
        if self._c_sigma:
            sigma = torch.clamp(self.sigma(logits), min=SIGMA_MIN, max=SIGMA_MAX).exp()
        else:
            shape = [1] * len(mu.shape)
            shape[1] = -1
            sigma = (self.sigma_param.view(shape) + torch.zeros_like(mu)).exp()
        return (mu, sigma), state


class Actor(nn.Module):
    def __init__(self, preprocess_net, action_shape, hidden_sizes=(), max_action=1.0, device="cpu",
                 preprocess_net_output_dim=None):
        super().__init__()
        self.preprocess, self.device = preprocess_net, device
        self.output_dim = int(np.prod(action_shape))
        d = getattr(preprocess_net, "output_dim", preprocess_net_output_dim)
        self.last = _last(d, self.output_dim, hidden_sizes, device)
        self._max = max_action

    def forward(self, obs, state=None, info={}):
        logits, hidden = self.preprocess(obs, state)
        return self._max * torch.tanh(self.last(logits)), hidden


class Critic(nn.Module):
    def __init__(self, preprocess_net, hidden_sizes=(), device="cpu", preprocess_net_output_dim=None,
                 linear_layer=nn.Linear, flatten_input=True):
        super().__init__()
        self.preprocess, self.device, self.output_dim = preprocess_net, device, 1
        d = getattr(preprocess_net, "output_dim", preprocess_net_output_dim)
        self.last = MLP(d, 1, hidden_sizes, device=device, linear_layer=linear_layer, flatten_input=flatten_input)

    def forward(self, obs, act=None, info={}):
        obs = torch.as_tensor(obs, device=self.device, dtype=torch.float32).flatten(1)
        if act is not None:
            act = torch.as_tensor(act, device=self.device, dtype=torch.float32).flatten(1)
            obs = torch.cat([obs, act], dim=1)
        logits, _ = self.preprocess(obs)
        return self.last(logits)
