"""Network containers with the module tree (and therefore the state_dict keys) of the
tianshou~=0.5 classes the reference agents build (fsrl/agent/ppo_lag_agent.py:136-145,
sac_lag_agent.py, ddpg_lag_agent.py; restated in examples/customized/collect_dataset.py:
189-215), plus the flat device arena all CUDA kernels read.

Every Linear is stored TRANSPOSED in the arena (Wt[in][out], include/fsrl_b200.h); the
torch-visible ``.weight`` is the strided view ``Wt.t()``, so ``state_dict()`` /
``load_state_dict()`` / ``torch.nn.init`` keep working while the kernels see the layout they
want.  The torch ``forward`` methods exist for API compatibility (and the parity tests use
them as an independent fp32 cross-check); the hot path never calls them.
"""
from __future__ import annotations

from copy import deepcopy

from typing import List, Optional, Sequence, Tuple, Union

import numpy as np
import torch
from torch import nn

from . import _lib

SIGMA_MIN, SIGMA_MAX = -20.0, 2.0     # tianshou ActorProb clamp [UNVERIFIED, SURVEY App. C]


class MLP(nn.Module):
    """[Linear -> ReLU] x len(hidden) then Linear(out) iff out > 0."""

    def __init__(self, input_dim: int, output_dim: int = 0, hidden_sizes: Sequence[int] = (),
                 device=None, linear_layer=nn.Linear, flatten_input: bool = True, **_):
        super().__init__()
        if linear_layer is not nn.Linear:
            raise NotImplementedError("only nn.Linear layers map onto the device engine")
        dims = [int(input_dim)] + [int(h) for h in hidden_sizes]
        layers: List[nn.Module] = []
        for i, o in zip(dims[:-1], dims[1:]):
            layers += [nn.Linear(i, o), nn.ReLU()]
        if output_dim > 0:
            layers.append(nn.Linear(dims[-1], int(output_dim)))
        self.output_dim = int(output_dim) if output_dim > 0 else dims[-1]
        self.model = nn.Sequential(*layers)
        self.device = device

    def forward(self, obs):
        dev = next(self.parameters()).device
        obs = torch.as_tensor(obs, device=dev, dtype=torch.float32)
        return self.model(obs.flatten(1))


class Net(nn.Module):
    """tianshou.utils.net.common.Net: an MLP body without output layer; ``concat=True``
    appends the action to the input (Q-networks)."""

    def __init__(self, state_shape, action_shape=0, hidden_sizes: Sequence[int] = (),
                 device=None, concat: bool = False, **_):
        super().__init__()
        in_dim = int(np.prod(state_shape))
        act_dim = int(np.prod(action_shape)) if action_shape else 0
        if concat:
            in_dim += act_dim
        self.input_dim = in_dim
        self.model = MLP(in_dim, 0, hidden_sizes, device)
        self.output_dim = self.model.output_dim
        self.device = device

    def forward(self, obs, state=None, info={}):
        return self.model(obs), state


class ActorProb(nn.Module):
    """Gaussian actor: mu = Linear(H, A) (bounded: max_action * tanh); sigma either the
    state-independent ``sigma_param`` (A, 1) or a clamped conditioned head."""

    def __init__(self, preprocess_net: Net, action_shape, hidden_sizes=(), max_action: float = 1.0,
                 device=None, unbounded: bool = False, conditioned_sigma: bool = False, **_):
        super().__init__()
        self.preprocess = preprocess_net
        self.output_dim = int(np.prod(action_shape))
        H = preprocess_net.output_dim
        self.mu = MLP(H, self.output_dim, ())
        self._c_sigma = conditioned_sigma
        if conditioned_sigma:
            self.sigma = MLP(H, self.output_dim, ())
        else:
            self.sigma_param = nn.Parameter(torch.zeros(self.output_dim, 1))
        self._max = float(max_action)
        self._unbounded = unbounded
        self.device = device

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
    """Deterministic actor: max_action * tanh(Linear(H, A))."""

    def __init__(self, preprocess_net: Net, action_shape, hidden_sizes=(), max_action: float = 1.0,
                 device=None, **_):
        super().__init__()
        self.preprocess = preprocess_net
        self.output_dim = int(np.prod(action_shape))
        self.last = MLP(preprocess_net.output_dim, self.output_dim, ())
        self._max = float(max_action)
        self.device = device

    def forward(self, obs, state=None, info={}):
        logits, hidden = self.preprocess(obs, state)
        return self._max * torch.tanh(self.last(logits)), hidden


class Critic(nn.Module):
    """V(s) or Q(s, a) (when ``act`` is given the input is cat([obs, act]))."""

    def __init__(self, preprocess_net: Net, hidden_sizes=(), device=None, preprocess_net_output_dim=None,
                 linear_layer=nn.Linear, flatten_input: bool = True, **_):
        # tianshou's positional order (subclasses such as the reference's SingleCritic pass all six positionally)
        super().__init__()
        if linear_layer is not nn.Linear:
            raise NotImplementedError("only nn.Linear layers map onto the device engine")
        self.preprocess = preprocess_net
        self.output_dim = 1
        self.last = MLP(preprocess_net.output_dim, 1, ())
        self.device = device

    def forward(self, obs, act=None, info={}):
        dev = next(self.parameters()).device
        obs = torch.as_tensor(obs, device=dev, dtype=torch.float32).flatten(1)
        if act is not None:
            act = torch.as_tensor(act, device=dev, dtype=torch.float32).flatten(1)
            obs = torch.cat([obs, act], dim=1)
        logits, _ = self.preprocess(obs)
        return self.last(logits)


# ---------------------------------------------------------------------------------------------
# flat arena
# ---------------------------------------------------------------------------------------------
def _linears(mlp: MLP) -> List[nn.Linear]:
    return [m for m in mlp.model if isinstance(m, nn.Linear)]


class NetSlot:
    """One 2-hidden-layer network inside the arena."""

    def __init__(self, name: str, module: nn.Module, lin1, lin2, heads: List[nn.Linear],
                 extra: Optional[nn.Parameter]):
        self.name, self.module = name, module
        self.lin1, self.lin2, self.heads, self.extra = lin1, lin2, heads, extra
        self.D, self.H = lin1.in_features, lin1.out_features
        if lin2.in_features != self.H or lin2.out_features != self.H:
            raise ValueError("fsrl_b200 kernels need two hidden layers of equal width "
                             f"(got {self.H} -> {lin2.out_features})")
        if self.H not in (64, 128, 256, 512):
            raise ValueError(f"hidden width {self.H} unsupported (64/128/256/512)")
        self.out = sum(h.out_features for h in heads)
        self.n_extra = 0 if extra is None else extra.numel()
        self.size = self.D * self.H + self.H + self.H * self.H + self.H + self.H * self.out + self.out + self.n_extra
        self.offset = -1

    def offsets(self):
        o = self.offset
        D, H, out = self.D, self.H, self.out
        w1 = o; b1 = w1 + D * H; w2 = b1 + H; b2 = w2 + H * H; w3 = b2 + H; b3 = w3 + H * out
        ex = b3 + out
        return w1, b1, w2, b2, w3, b3, ex


def slot_from_module(name: str, m: nn.Module) -> NetSlot:
    body = _linears(m.preprocess.model)
    if len(body) != 2:
        raise ValueError("fsrl_b200 kernels support exactly two hidden layers "
                         f"(hidden_sizes of length 2), got {len(body)}")
    if isinstance(m, ActorProb):
        heads = _linears(m.mu) + (_linears(m.sigma) if m._c_sigma else [])
        extra = None if m._c_sigma else m.sigma_param
    elif isinstance(m, (Actor, Critic)):
        heads, extra = _linears(m.last), None
    else:
        raise TypeError(f"unsupported network type {type(m)}")
    return NetSlot(name, m, body[0], body[1], heads, extra)


class Arena:
    """Flat fp32 device storage for a list of networks (+ grads, Adam moments, W2 mirror)."""

    def __init__(self, slots: List[NetSlot], device):
        self.slots = slots
        self.device = torch.device(device)
        off = 0
        for s in slots:
            s.offset = off
            off += s.size
            off = (off + 3) & ~3          # keep every network 16-byte aligned
        self.n_params = off
        self.theta = torch.zeros(off, dtype=torch.float32, device=self.device)
        self.grad = torch.zeros_like(self.theta)
        for s in slots:
            self._bind(s)

    def _bind(self, s: NetSlot):
        th = self.theta
        w1, b1, w2, b2, w3, b3, ex = s.offsets()
        D, H, out = s.D, s.H, s.out

        def adopt_linear(lin, woff, boff, in_f, out_total, col0):
            wt = th[woff:woff + in_f * out_total].view(in_f, out_total)[:, col0:col0 + lin.out_features]
            wt.copy_(lin.weight.detach().t().to(th.device))
            lin.weight.data = wt.t()
            bv = th[boff + col0: boff + col0 + lin.out_features]
            bv.copy_(lin.bias.detach().to(th.device))
            lin.bias.data = bv

        adopt_linear(s.lin1, w1, b1, D, H, 0)
        adopt_linear(s.lin2, w2, b2, H, H, 0)
        col = 0
        for h in s.heads:
            adopt_linear(h, w3, b3, H, out, col)
            col += h.out_features
        if s.extra is not None:
            ev = th[ex:ex + s.n_extra]
            ev.copy_(s.extra.detach().reshape(-1).to(th.device))
            s.extra.data = ev.view(s.extra.shape)

    def mlp3(self, s: NetSlot, out_cols: Optional[int] = None) -> _lib.Mlp3:
        w1, b1, w2, b2, w3, b3, _ = s.offsets()
        base = self.theta.data_ptr()
        m = _lib.Mlp3()
        m.w1t, m.b1, m.w2t, m.b2, m.w3t, m.b3 = (base + 4 * o for o in (w1, b1, w2, b2, w3, b3))
        m.in_, m.H, m.out = s.D, s.H, s.out
        return m

    def extra_ptr(self, s: NetSlot) -> Optional[int]:
        if s.extra is None:
            return None
        return self.theta.data_ptr() + 4 * s.offsets()[6]


# ---------------------------------------------------------------------------------------------
# Q-critics of the off-policy learners (reference: fsrl/utils/net/continuous.py:12-155)
# ---------------------------------------------------------------------------------------------
class DoubleCritic(nn.Module):
    """Two independent Q(s, a) heads; ``forward`` returns [q1, q2], ``predict`` (min, list)."""

    def __init__(self, preprocess_net1: Net, preprocess_net2: Net, hidden_sizes=(), device=None, **_):
        super().__init__()
        self.device = device
        self.preprocess1, self.preprocess2 = preprocess_net1, preprocess_net2
        self.output_dim = 1
        self.last1 = MLP(preprocess_net1.output_dim, 1, ())
        # a COPY of last1, not a fresh MLP: constructing one would draw from the torch RNG and every later
        # initialisation (the agents' orthogonal re-init) would leave the reference's seed-for-seed stream
        self.last2 = deepcopy(self.last1)

    def forward(self, obs, act=None, info={}):
        dev = next(self.parameters()).device
        obs = torch.as_tensor(obs, device=dev, dtype=torch.float32).flatten(1)
        if act is not None:
            act = torch.as_tensor(act, device=dev, dtype=torch.float32).flatten(1)
            obs = torch.cat([obs, act], dim=1)
        return [self.last1(self.preprocess1(obs)[0]), self.last2(self.preprocess2(obs)[0])]

    def predict(self, obs, act=None, info={}):
        q = self(obs, act, info)
        return torch.min(q[0], q[1]), q


class SingleCritic(Critic):
    """tianshou Critic with the list-valued API of DoubleCritic."""

    def forward(self, obs, act=None, info={}):
        return [super().forward(obs, act, info)]

    def predict(self, obs, act=None, info={}):
        q = self(obs, act, info)
        return q[0], q


def slots_from_module(name: str, m: nn.Module) -> List[NetSlot]:
    """1 slot for Actor/ActorProb/Critic, 2 for a DoubleCritic."""
    if isinstance(m, DoubleCritic):
        out = []
        for k, (pre, last) in enumerate(((m.preprocess1, m.last1), (m.preprocess2, m.last2))):
            body = _linears(pre.model)
            if len(body) != 2:
                raise ValueError("fsrl_b200 kernels support exactly two hidden layers")
            out.append(NetSlot(f"{name}.{k}", m, body[0], body[1], _linears(last), None))
        return out
    return [slot_from_module(name, m)]
