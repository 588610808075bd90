"""Device-resident actor rollout: a DMPO-shaped policy (reference `flybody/agents/network_factory.py:67-100`:
batch_concat -> LayerNormMLP(256, 256, 256) -> diagonal Gaussian head, random weights) in PyTorch on the stepper's stream,
fed by `env.step_device()` -- observations, actions, rewards never leave the GPU (BASELINE configs 2/4: "random policy" ->
"DMPO actor loop").  A measurement tool, not part of the product path:  python tools/gpu_actor_rollout.py [n_envs] [steps]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.getcwd())
import torch
from flybody_b200 import fly_envs


class DmpoPolicy(torch.nn.Module):
    """acme LayerNormMLP (Linear -> LayerNorm -> tanh, then Linear -> ELU ...) + MultivariateNormalDiagHead (init_scale 0.7)."""

    def __init__(self, n_obs, n_act, sizes=(256, 256, 256), init_scale=0.7, min_scale=1e-6):
        super().__init__()
        self.l0, self.ln = torch.nn.Linear(n_obs, sizes[0]), torch.nn.LayerNorm(sizes[0])
        self.rest = torch.nn.ModuleList([torch.nn.Linear(a, b) for a, b in zip(sizes[:-1], sizes[1:])])
        self.mean, self.scale = torch.nn.Linear(sizes[-1], n_act), torch.nn.Linear(sizes[-1], n_act)
        self.k, self.min_scale = init_scale / float(np.log(2.0)), min_scale

    def forward(self, obs):
        h = torch.tanh(self.ln(self.l0(obs)))
        for l in self.rest:
            h = torch.nn.functional.elu(l(h))
        mean, scale = self.mean(h), torch.nn.functional.softplus(self.scale(h)) * self.k + self.min_scale
        return mean + scale * torch.randn_like(mean)


N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
K = int(sys.argv[2]) if len(sys.argv) > 2 else 50
for variant, make in (('walk', lambda: fly_envs.walk_imitation(n_envs=N, device_task=True, reset_noise=0.05)),
                      ('flight', lambda: fly_envs.flight_imitation(n_envs=N, device_task=True))):
    env = make(); env.reset()
    spec = env.action_spec()
    lo, hi = torch.tensor(spec.minimum, device='cuda', dtype=torch.float32), torch.tensor(spec.maximum, device='cuda', dtype=torch.float32)
    cols = np.concatenate([np.arange(sl.start, sl.stop) for sl, _ in env.observation_layout().values()])
    cols_t = torch.tensor(cols, device='cuda')
    policy = DmpoPolicy(len(cols), spec.shape[0]).cuda()
    stream = torch.cuda.ExternalStream(env.physics.stepper.stream)
    with torch.cuda.stream(stream), torch.no_grad():
        obs, out = env.step_device(torch.zeros((N, spec.shape[0]), device='cuda'))
        ret, n_first = torch.zeros(N, device='cuda'), torch.zeros((), device='cuda')
        for k in range(K + 5):
            if k == 5:
                stream.synchronize(); t0 = time.perf_counter()
            act = torch.minimum(torch.maximum(policy(obs[:, cols_t]), lo), hi).contiguous()       # canonical spec clipping
            obs, out = env.step_device(act)
            ret += out[:, 0]; n_first += (out[:, 2] == 0).sum()
        stream.synchronize()
        dt = (time.perf_counter() - t0) / K
    print(f'{variant:7s} N={N} policy {len(cols)}->256x3->{spec.shape[0]}  {dt * 1e3:8.3f} ms/step {N / dt:12.0f} env-steps/s  '
          f'episodes restarted {int(n_first.item())}  mean return so far {float(ret.mean()):.2f}', flush=True)
    env.close()
