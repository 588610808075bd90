"""A/B timing of stepper library variants on one GPU: ms per control step (CUDA-graph replay, device-resident actions) and the
per-stage kernel times (event pair around every launch), walk model, after a warm-up long enough for contacts to build up.
    python tools/gpu_ab.py libA.so libB.so ... [--envs 4096,16384] [--steps 30] [--warm 60]"""
import argparse, json, os, sys
import numpy as np
sys.path.insert(0, os.getcwd())
import torch
from flybody_b200 import stepper as st
from flybody_b200.flymodel import load_model

ap = argparse.ArgumentParser(); ap.add_argument('libs', nargs='+'); ap.add_argument('--envs', default='4096'); ap.add_argument('--steps', type=int, default=30)
ap.add_argument('--warm', type=int, default=60); ap.add_argument('--model', default='walk'); ap.add_argument('--rounds', type=int, default=2)
a = ap.parse_args()
m = load_model(a.model)
n_sub, scale = (10, 0.5) if a.model == 'walk' else (4, 0.2)
res = {}
for rnd in range(a.rounds):                      # interleaved rounds: box drift shows up as a difference between the rounds
    for spec in a.libs:                          # "path.so" or "path.so:VAR=val,VAR2=val2" (environment read by fb_create)
        lib, _, envs = spec.partition(':')
        sets = dict(kv.split('=') for kv in envs.split(',') if kv)
        for N in [int(x) for x in a.envs.split(',')]:
            old = {k: os.environ.get(k) for k in sets}
            os.environ.update(sets)
            sim = st.BatchedStepper(m, N, lib_path=lib)
            for k, v in old.items():
                if v is None: os.environ.pop(k, None)
                else: os.environ[k] = v
            rs = np.random.RandomState(1)
            q = np.tile(m.qpos0, (N, 1))
            if a.model == 'walk':
                for side in ('left', 'right'):
                    for dof, val in (('yaw', 1.5), ('roll', 0.7), ('pitch', -1.0)):
                        q[:, m.jnt_qposadr_of(f'walker/wing_{dof}_{side}')] = val
                q[:, 7:109] += rs.uniform(-0.05, 0.05, (N, 102))
            else:
                q[:, 2] = 1.0
            sim.reset(q)
            stream = torch.cuda.ExternalStream(sim.stream)
            gen = torch.Generator(device='cuda'); gen.manual_seed(7)
            acts = (torch.rand((16, N, m.nu), device='cuda', generator=gen) - 0.5) * (2 * scale)
            def step(k):
                sim.set_control_device(acts[k % 16].data_ptr()); sim.step(n_sub)
            with torch.cuda.stream(stream):
                for k in range(a.warm): step(k)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                for k in range(a.steps): step(k)
                e1.record(stream)
            sim.sync(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / a.steps
            sim.profile(True)
            with torch.cuda.stream(stream):
                for k in range(a.steps): step(k)
            sim.sync(); torch.cuda.synchronize()
            prof = {k: round(v[0] / a.steps, 3) for k, v in sim.profile_read().items() if v[1]}
            sim.profile(False)
            nefc = sim.get(st.NEFC)[:, 0]; nit = sim.get(st.SOLVER_NITER)[:, 0].astype(int)
            key = (os.path.basename(spec), N)
            res.setdefault(key, []).append(ms)
            print(json.dumps({'lib': key[0], 'envs': N, 'round': rnd, 'ms_per_step': round(ms, 3), 'env_steps_per_s': round(N / ms * 1e3), 'stages_ms': prof,
                              'nefc_mean': float(nefc.mean()), 'niter_hist': {str(b): int((nit == b).sum()) for b in range(0, 9)}, 'niter_gt8': int((nit > 8).sum()), 'niter_max': int(nit.max()), 'share_nefc_gt32': float((nefc > 32).mean()), 'flags': int((sim.get(st.FLAGS) != 0).sum())}), flush=True)
            sim.close()
print('SUMMARY', {f'{k[0]}@{k[1]}': [round(x, 3) for x in v] for k, v in res.items()})
