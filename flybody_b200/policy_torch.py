"""PyTorch ports of the reference's actor networks, for batched policy INFERENCE in the actor loop (the learner --
TF / Acme DMPO, Reverb, Ray -- is out of scope; BASELINE.json configs 4 / 5 only need the policy's forward pass on rank 0).

  LayerNormMLP          acme.tf.networks.LayerNormMLP as the reference uses it (`flybody/agents/network_factory.py:76-78`,
                        `network_factory_vis.py:296-341`): Linear -> LayerNorm -> tanh, then Linear (-> ELU) ...
  DMPOPolicy            `network_factory_dmpo` (`network_factory.py:66-109`): batch_concat -> LayerNormMLP(256, 256, 256, activate_final)
                        -> MultivariateNormalDiagHead (mean Linear; scale = softplus(Linear) * init_scale / log 2 + min_scale)
  VisNet                `network_factory_vis.py:141-220`: the two 32 x 32 eyes -> gray -> (x - 77) / 56 -> four VALID 3 x 3 convolutions
                        (2, 4, 8, 16 channels; strides 1, 1, 2, 2) -> Linear(vis_output_dim); output = [task_input, vis, other observations]
  TwoLevelController    `network_factory_vis.py:223-293`: high-level LayerNormMLP -> steering command (future reference displacements and
                        root quaternions) spliced into the low-level flight policy's observation at its `ref_displacement` slot

Weights are randomly initialised (there are no checkpoints here: no network); `load_state_dict` takes converted ones.  Small dense
layers and tiny convolutions: cuBLAS / cuDNN through torch, no custom kernel (this is not the hot path)."""
import numpy as np
import torch
from torch import nn


class LayerNormMLP(nn.Module):
    def __init__(self, n_in, layer_sizes, activate_final=False):
        super().__init__()
        sizes = list(layer_sizes)
        self.first, self.norm = nn.Linear(n_in, sizes[0]), nn.LayerNorm(sizes[0])
        self.rest = nn.ModuleList([nn.Linear(a, b) for a, b in zip(sizes[:-1], sizes[1:])])
        self.activate_final = activate_final

    def forward(self, x):
        h = torch.tanh(self.norm(self.first(x)))
        for i, layer in enumerate(self.rest):
            h = layer(h)
            if i < len(self.rest) - 1 or self.activate_final:
                h = nn.functional.elu(h)
        return h


class DMPOPolicy(nn.Module):
    """observation rows [N, n_obs] (the env's observables concatenated in spec order) -> action sample [N, n_act]"""

    def __init__(self, n_obs, n_act, layer_sizes=(256, 256, 256), init_scale=0.7, min_scale=1e-6):
        super().__init__()
        self.torso = LayerNormMLP(n_obs, layer_sizes, activate_final=True)
        self.mean, self.scale = nn.Linear(layer_sizes[-1], n_act), nn.Linear(layer_sizes[-1], n_act)
        self._k, self._min = init_scale / float(np.log(2.0)), min_scale

    def distribution(self, obs):
        h = self.torso(obs)
        return self.mean(h), nn.functional.softplus(self.scale(h)) * self._k + self._min

    def forward(self, obs, deterministic=False):
        mean, scale = self.distribution(obs)
        return mean if deterministic else mean + scale * torch.randn_like(mean)


class VisNet(nn.Module):
    def __init__(self, vis_output_dim=8, eye_size=32):
        super().__init__()
        c = nn.Conv2d
        self.conv = nn.Sequential(c(2, 2, 3), nn.ReLU(), c(2, 4, 3), nn.ReLU(), c(4, 8, 3, stride=2), nn.ReLU(), c(8, 16, 3, stride=2), nn.ReLU())
        s = ((eye_size - 2 - 2 - 3) // 2 + 1 - 3) // 2 + 1                 # 32 -> 30 -> 28 -> 13 -> 6
        self.head = nn.Linear(16 * s * s, vis_output_dim)
        self.mean, self.std = 77.0, 56.0                                   # "Mean and std from the trench task" (network_factory_vis.py:160-162)

    def forward(self, left_eye, right_eye, task_input, others):
        """eyes uint8 [N, S, S, 3] (or [N, S, S] gray), task_input [N, T] or None, others [N, K] -> [N, T + vis + K]"""
        gray = lambda e: (e.float().mean(-1) if e.dim() == 4 else e.float()).sub(self.mean).div(self.std)
        x = torch.stack((gray(left_eye), gray(right_eye)), 1)              # channels: left, right (tf.stack(..., axis=-1) in NHWC)
        x = self.head(self.conv(x).permute(0, 2, 3, 1).flatten(1))         # flatten in NHWC order, as snt.Flatten sees it
        parts = ([task_input] if task_input is not None else []) + [x, others]
        return torch.cat(parts, -1)


class TwoLevelController(nn.Module):
    """`x` = VisNet output [task_input, vis, low-level observation without its reference observables]"""

    def __init__(self, n_ll_obs_without_steering, n_act, steering_idx, hl_layer_sizes=(256, 256, 128), future_steps=5, task_input_dim=2,
                 vis_output_dim=8, ll_layer_sizes=(256, 256, 256)):
        super().__init__()
        n_rep = future_steps + 1
        self.steering_dim = 7 * n_rep                                      # ref_displacement [n_rep, 3] + ref_root_quat [n_rep, 4]
        self.offset, self.steering_idx = task_input_dim + vis_output_dim, int(steering_idx)
        self.hl = LayerNormMLP(self.offset + n_ll_obs_without_steering, list(hl_layer_sizes) + [self.steering_dim])
        with torch.no_grad():                                              # "scale=0.01 so the network output ... is close to the no-op steering command"
            for p in self.hl.parameters():
                if p.dim() == 2:
                    p.mul_(0.03)
        self.register_buffer('ballpark', torch.tensor(n_rep * [0.0, 0.0, 0.0] + n_rep * [1.0, 0.0, 0.0, 0.0]))
        self.ll = DMPOPolicy(n_ll_obs_without_steering + self.steering_dim, n_act, ll_layer_sizes)
        for p in self.ll.parameters():                                     # the pre-trained low-level controller stays frozen
            p.requires_grad_(False)

    def forward(self, x, deterministic=False):
        steering = self.hl(x) + self.ballpark
        ll_in = x[:, self.offset:]
        ll_in = torch.cat((ll_in[:, :self.steering_idx], steering, ll_in[:, self.steering_idx:]), -1)
        return self.ll(ll_in, deterministic=deterministic)


def vision_policy_for(env, device='cuda'):
    """(VisNet, TwoLevelController, column index of the non-visual observables) for a `vision_guided_flight` env: the low-level
    observation is the flight task's (accelerometer, gyro, joints_pos, joints_vel, velocimeter, world_zaxis + the steering command where
    flight_imitation has ref_displacement / ref_root_quat, i.e. after world_zaxis in spec order)."""
    spec = env.observation_spec()
    keys = [k for k in spec if not k.endswith(('_eye', 'task_input'))]
    widths = [int(np.prod(spec[k].shape[1 if env._batched else 0:])) for k in keys]
    n_others = int(sum(widths))
    vis = VisNet(eye_size=env._eye_size).to(device).eval()
    ctl = TwoLevelController(n_others, env.action_spec().shape[0], steering_idx=n_others).to(device).eval()
    return vis, ctl, keys
