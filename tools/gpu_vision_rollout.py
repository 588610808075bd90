"""BASELINE config 5 shape: vision_guided_flight + the vision policy of the reference (`agents/network_factory_vis.py:141-293`:
VisNet conv stack on the two 32 x 32 eyes -> high-level LayerNormMLP emitting a steering command (ref_displacement 6 x 3 +
ref_root_quat 6 x 4) -> low-level flight policy), random weights, PyTorch on the GPU; the env steps through its public API (host
task code).  Not yet run on a B200 when written; a measurement tool for the next round, not part of the product path.
    python tools/gpu_vision_rollout.py [n_envs] [steps]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.getcwd())
import torch
from flybody_b200 import fly_envs


class LayerNormMLP(torch.nn.Module):
    def __init__(self, n_in, sizes, activate_final):
        super().__init__()
        self.l0, self.ln = torch.nn.Linear(n_in, sizes[0]), torch.nn.LayerNorm(sizes[0])
        self.rest = torch.nn.ModuleList([torch.nn.Linear(a, b) for a, b in zip(sizes[:-1], sizes[1:])])
        self.activate_final = activate_final

    def forward(self, x):
        h = torch.tanh(self.ln(self.l0(x)))
        for i, l in enumerate(self.rest):
            h = l(h)
            if i < len(self.rest) - 1 or self.activate_final:
                h = torch.nn.functional.elu(h)
        return h


class VisNet(torch.nn.Module):
    def __init__(self, vis_output_dim=8):
        super().__init__()
        c = torch.nn.Conv2d
        self.conv = torch.nn.Sequential(c(2, 2, 3), torch.nn.ReLU(), c(2, 4, 3), torch.nn.ReLU(), c(4, 8, 3, stride=2), torch.nn.ReLU(),
                                        c(8, 16, 3, stride=2), torch.nn.ReLU(), torch.nn.Flatten(), torch.nn.Linear(16 * 6 * 6, vis_output_dim))

    def forward(self, left, right):                      # uint8 [N, 32, 32, 3]
        g = lambda e: (e.float().mean(-1) - 77.0) / 56.0
        return self.conv(torch.stack([g(left), g(right)], 1))


class TwoLevel(torch.nn.Module):
    def __init__(self, n_prop, n_act, vis_dim=8, task_dim=2, steer=42):
        super().__init__()
        self.vis = VisNet(vis_dim)
        self.hl = LayerNormMLP(task_dim + vis_dim + n_prop, (256, 256, 128, steer), activate_final=False)
        self.ball = torch.nn.Parameter(torch.tensor(6 * [0.0, 0, 0] + 6 * [1.0, 0, 0, 0]), requires_grad=False)
        self.ll = LayerNormMLP(n_prop + steer, (256, 256, 256), activate_final=True)
        self.mean = torch.nn.Linear(256, n_act)

    def forward(self, task, left, right, prop):
        x = torch.cat([task, self.vis(left, right), prop], -1)
        steering = self.hl(x) + self.ball
        return self.mean(self.ll(torch.cat([prop, steering], -1)))


N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
K = int(sys.argv[2]) if len(sys.argv) > 2 else 50
env = fly_envs.vision_guided_flight(n_envs=N, seed=1, terrain_bank=64)
ts = env.reset()
prop_keys = [k for k in ts.observation if not k.endswith(('_eye', 'task_input'))]
n_prop = sum(int(np.prod(ts.observation[k].shape[1:])) for k in prop_keys)
spec = env.action_spec()
policy = TwoLevel(n_prop, spec.shape[0]).cuda().eval()
lo, hi = torch.tensor(spec.minimum, device='cuda', dtype=torch.float32), torch.tensor(spec.maximum, device='cuda', dtype=torch.float32)


def act(ts):
    o = ts.observation
    with torch.no_grad():
        a = policy(torch.from_numpy(o['walker/task_input']).cuda(), torch.from_numpy(o['walker/left_eye']).cuda(),
                   torch.from_numpy(o['walker/right_eye']).cuda(), torch.from_numpy(np.concatenate([o[k].reshape(N, -1) for k in prop_keys], 1)).cuda())
        return torch.minimum(torch.maximum(a, lo), hi).cpu().numpy()


for k in range(5):
    ts = env.step(act(ts))
t0 = time.perf_counter(); n_last = 0
for k in range(K):
    ts = env.step(act(ts)); n_last += int((np.asarray(ts.step_type) == 2).sum())
dt = (time.perf_counter() - t0) / K
print(f'vision_guided_flight + VisNet / two-level policy, N={N}: {dt * 1e3:.3f} ms/step {N / dt:.0f} env-steps/s  terminations {n_last}  '
      f'mean reward {float(np.mean(ts.reward)):.3f}', flush=True)
