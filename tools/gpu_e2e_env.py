"""End-to-end env-steps/s through the public env API (host actions in, TimeStep out) with the task hooks in host numpy vs on
the device (fb_task_*):  python tools/gpu_e2e_env.py [n_envs] [steps]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.getcwd())
from flybody_b200 import fly_envs

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
K = int(sys.argv[2]) if len(sys.argv) > 2 else 30
for variant, make, na, scale in (('walk', lambda dev: fly_envs.walk_imitation(terminal_com_dist=float('inf'), n_envs=N, reset_noise=0.05, device_task=dev), 59, 0.5),
                                 ('flight', lambda dev: fly_envs.flight_imitation(n_envs=N, device_task=dev), 12, 0.2)):
    for dev in (False, True):
        env = make(dev)
        rs = np.random.RandomState(0)
        acts = rs.uniform(-scale, scale, (K + 5, N, na)).astype(np.float32)
        env.reset()
        for k in range(5):
            env.step(acts[k])
        t0 = time.perf_counter()
        n_last = 0
        for k in range(5, K + 5):
            ts = env.step(acts[k]); n_last += int((np.asarray(ts.step_type) == 2).sum())
        dt = (time.perf_counter() - t0) / K
        print(f'{variant:7s} N={N} device_task={dev!s:5s} {dt * 1e3:8.3f} ms/step {N / dt:12.0f} env-steps/s  terminations {n_last}  mean reward {float(np.mean(ts.reward)):.4f}', flush=True)
        env.close()

# device-resident rollout: actions from a CUDA tensor, observation / reward views left on the device (a GPU policy's loop)
import torch
for variant, make, na, scale in (('walk', lambda: fly_envs.walk_imitation(terminal_com_dist=float('inf'), n_envs=N, reset_noise=0.05, device_task=True), 59, 0.5),
                                 ('flight', lambda: fly_envs.flight_imitation(n_envs=N, device_task=True), 12, 0.2)):
    env = make(); env.reset()
    stream = torch.cuda.ExternalStream(env.physics.stepper.stream)
    with torch.cuda.stream(stream):
        acts = (torch.rand((K + 5, N, na), device='cuda') - 0.5) * 2 * scale
        for k in range(5):
            env.step_device(acts[k])
        stream.synchronize()
        t0 = time.perf_counter()
        for k in range(5, K + 5):
            obs, out = env.step_device(acts[k])
        stream.synchronize()
        dt = (time.perf_counter() - t0) / K
        print(f'{variant:7s} N={N} step_device (no host copies) {dt * 1e3:8.3f} ms/step {N / dt:12.0f} env-steps/s  mean reward {float(out[:, 0].mean()):.4f}', flush=True)
    env.close()
