#!/usr/bin/env python
"""bench.py -- env-steps/sec of the batched walk_imitation physics step (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # our CUDA arm
    python bench.py --impl reference --gpus N --steps K ...  # CPU arm: the restated mj_step on the host cores

One "step" = one control step (10 physics substeps of 2e-4 s) of every environment of the batch:
workload `walk_imitation 4096 envs, random policy` per GPU (BASELINE.json configs[1]); N>1 shards
envs across ranks (weak scaling, 4096 envs per GPU) and gathers the packed observations + rewards to
rank 0 over NCCL every control step (configs[3]).

  value : env-steps/s, whole job, actions already resident in HBM, device-timed (CUDA events on the
          stepper's stream, max over ranks), gather included for N>1
  e2e   : the same metric through the public API `flybody_b200.fly_envs.walk_imitation(n_envs).step(a)`
          with pinned host actions in and the observation record out every step
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ENVS_PER_GPU = 4096
N_SUB = 10
BYTES_PER_ENV_SUBSTEP = 3628          # fp32 state read+write per physics substep (W model)
BYTES_PER_ENV_STEP = BYTES_PER_ENV_SUBSTEP * N_SUB
# --workload: the headline is walk (BASELINE.json configs[1]); flight is configs[2], same contract, not the driver's default
WORKLOADS = {'walk': dict(model='walk', n_sub=10, bytes_substep=3628, act_scale=0.5, n_action=59, dt='2e-4',
                          config='BASELINE.json configs[1]; N>1: configs[3] sharding + NCCL obs gather to rank 0'),
             'flight': dict(model='flight', n_sub=4, bytes_substep=1284, act_scale=0.2, n_action=12, dt='5e-5',
                            config='BASELINE.json configs[2]: ellipsoid fluid / wing forces, wing-beat pattern generator')}
WL = WORKLOADS['walk']


def set_workload(name):
    global WL, N_SUB, BYTES_PER_ENV_SUBSTEP, BYTES_PER_ENV_STEP
    WL = WORKLOADS[name]
    N_SUB, BYTES_PER_ENV_SUBSTEP = WL['n_sub'], WL['bytes_substep']
    BYTES_PER_ENV_STEP = BYTES_PER_ENV_SUBSTEP * N_SUB


def measured_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        return json.load(open(p)).get('hbm_gbs', 6650.0), 'measured (MEASURED_PEAKS.json)'
    return 6650.0, 'fallback (B200_PROFILING.md)'


class ClockSampler(threading.Thread):
    def __init__(self, device):
        super().__init__(daemon=True)
        self.device, self.stop_flag, self.samples = device, False, []

    def run(self):
        q = ('index,clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
             'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')
        while not self.stop_flag:
            try:
                out = subprocess.run(['nvidia-smi', f'--id={self.device}', f'--query-gpu={q}', '--format=csv,noheader,nounits'],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(',')])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.samples:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        sm = sorted(float(s[1]) for s in self.samples)
        reasons = []
        for i, name in ((3, 'hw_slowdown'), (4, 'hw_thermal_slowdown'), (5, 'sw_thermal_slowdown'), (6, 'sw_power_cap')):
            if any(s[i].lower().startswith('active') for s in self.samples):
                reasons.append(name)
        return {'sm_mhz': sm[len(sm) // 2], 'sm_max_mhz': float(self.samples[0][2]), 'reasons': reasons, 'samples': len(sm)}


def walk_reset_batch(m, n, rs):
    if WL['model'] == 'flight':                      # hovering start 1 cm above the floor (flight_imitation's synthetic trajectory height)
        qq = np.tile(m.qpos0, (n, 1)); qq[:, 2] = 1.0
        return qq
    q0 = m.qpos0.copy()
    for side in ('left', 'right'):
        for dof, val in (('yaw', 1.5), ('roll', 0.7), ('pitch', -1.0)):
            q0[m.jnt_qposadr_of(f'walker/wing_{dof}_{side}')] = val
    qq = np.tile(q0, (n, 1))
    # config 2: U(-0.05, 0.05) rad on the 48 actuated leg joints decorrelates the envs
    leg = [m.jnt_qposadr[m.actuator_trnid[i]] for i in range(m.nu)
           if m.actuator_trntype[i] == 0 and any(t in m.meta['actuator_names'][i] for t in ('T1', 'T2', 'T3'))]
    qq[:, leg] += rs.uniform(-0.05, 0.05, (n, len(leg)))
    return qq


# ------------------------------------------------------------------------------------------------
def cpu_oracle_worker(args):
    """One host core: restated mj_step (oracle), `steps` control steps of walk_imitation, random actions."""
    seed, budget_s, max_steps = args
    from flybody_b200.flymodel import load_model
    from oracle import fly_oracle as fo
    m = load_model(WL['model'])
    o = fo.Oracle(m)                       # MuJoCo default tolerance (1e-8)
    rs = np.random.RandomState(seed)
    o.reset(walk_reset_batch(m, 1, rs)[0])
    acts = rs.uniform(-WL['act_scale'], WL['act_scale'], (max_steps, m.nu))
    for k in range(3):
        o.set(fo.CTRL, acts[k]); o.control_step(N_SUB)
    t0 = time.perf_counter()
    n = 0
    while n < max_steps and time.perf_counter() - t0 < budget_s:
        o.set(fo.CTRL, acts[n]); o.control_step(N_SUB)
        n += 1
        if o.get(fo.FLAGS)[0] != 0:
            o.reset(walk_reset_batch(m, 1, rs)[0])
    return n, time.perf_counter() - t0


def cpu_baseline(budget_s=10.0, max_steps=4000, cores=None):
    import multiprocessing as mp
    from oracle import fly_oracle as fo
    fo.build()
    cores = cores or os.cpu_count()
    with mp.get_context('fork').Pool(cores) as pool:
        res = pool.map(cpu_oracle_worker, [(1000 + i, budget_s, max_steps) for i in range(cores)])
    rate = sum(n / t for n, t in res)
    return {'value': rate, 'unit': 'env-steps/s', 'cores': cores, 'kind': 'port',
            'sample': f'{cores} processes x ~{budget_s:.0f}s of {WL["model"]}_imitation control steps ({N_SUB} substeps, random actions), '
                      f'oracle/fly_oracle.c (restated mj_step, fp64, dense; NOT MuJoCo, which is not installable here and would be an estimated '
                      f'20-40x faster per core under load: DESIGN.md section 6), '
                      f'{sum(n for n, _ in res)} env-steps total',
            'per_core': rate / cores}


def cpu_kernel_source_worker(args):
    """One host core: the step kernels' own source compiled for the host (tests/_emu, g++ -O2, fp32, sparse factorisation, dual
    Newton) stepping a few envs -- a measured stand-in for what an optimised CPU engine does per core, next to the dense oracle."""
    seed, budget_s, emu = args
    from flybody_b200.flymodel import load_model
    from flybody_b200 import stepper as st
    m = load_model(WL['model'])
    n = 4
    sim = st.BatchedStepper(m, n, lib_path=emu)
    rs = np.random.RandomState(seed)
    sim.reset(walk_reset_batch(m, n, rs))
    acts = rs.uniform(-WL['act_scale'], WL['act_scale'], (64, n, m.nu)).astype(np.float32)
    for k in range(2):
        sim.set_control(acts[k]); sim.step(N_SUB)
    t0 = time.perf_counter(); steps = 0
    while time.perf_counter() - t0 < budget_s:
        sim.set_control(acts[steps % 64]); sim.step(N_SUB); steps += 1
    return steps * n, time.perf_counter() - t0


def cpu_kernel_source_baseline(budget_s=5.0, cores=None):
    """-> extra JSON block: env-steps/s of the host-emulation build of the kernel source on all host cores (None if it is absent)."""
    import multiprocessing as mp
    import __graft_entry__ as ge
    if not os.path.exists(ge.EMU):
        return None
    cores = cores or os.cpu_count()
    with mp.get_context('fork').Pool(cores) as pool:
        res = pool.map(cpu_kernel_source_worker, [(2000 + i, budget_s, ge.EMU) for i in range(cores)])
    rate = sum(n / t for n, t in res)
    return {'value': rate, 'unit': 'env-steps/s', 'cores': cores, 'per_core': rate / cores,
            'what': 'the step kernels\' source compiled for the host (tests/_emu/libfb_emu.so: g++ -O2, fp32, sparse L^T D L, dual Newton), one '
                    f'process per core x 4 envs x ~{budget_s:.0f}s; test artefact, never loaded by the product path -- reported as a measured, '
                    'stronger CPU number than the dense fp64 oracle of cpu_baseline'}


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    t0 = time.perf_counter()
    per_step_budget = 4.0
    cb = cpu_baseline(budget_s=per_step_budget * max(1, min(args.steps, 5)), cores=os.cpu_count())
    wall = time.perf_counter() - t0
    line = {'impl': 'reference', 'metric': f'env-steps/sec on {WL["model"]}_imitation (control steps of {N_SUB} substeps)', 'value': cb['value'],
            'unit': 'env-steps/s', 'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': 1e3 / cb['per_core'], 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f64', 'data': 'synthetic',
            'config': {'workload': f'{WL["model"]}_imitation, random policy, one env per host core (reference scaling model: '
                                   'one env per actor process, train_dmpo_ray.py:206-227)', 'cores': cb['cores']},
            'cpu_baseline': cb,
            'e2e': {'value': cb['value'], 'unit': 'env-steps/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0, 'wall_s': wall}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
class CudaView:
    def __init__(self, ptr, shape):
        self.__cuda_array_interface__ = {'shape': tuple(shape), 'typestr': '<f4', 'data': (int(ptr), False), 'version': 2}


def run_ours(args):
    import torch
    import torch.distributed as dist
    from flybody_b200.flymodel import load_model
    from flybody_b200 import stepper as st
    from flybody_b200 import fly_envs, sharding

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device; the stepper has no CPU path (use --impl reference for the CPU arm)')
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault('NCCL_DEBUG_FILE', '/dev/stderr')      # stdout carries exactly one JSON line (NCCL prints its version banner otherwise)
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    N = args.envs
    m = load_model(WL['model'])
    rs = np.random.RandomState(sharding.rank_seed(1234, rank))
    sim = st.BatchedStepper(m, N, device=local)
    Np = sim.n_envs_padded
    sim.reset(walk_reset_batch(m, N, rs))
    stream = torch.cuda.ExternalStream(sim.stream, device=torch.device('cuda', local))
    obs_ptr, obs_dim = sim.obs_ptr()
    obs = torch.as_tensor(CudaView(obs_ptr, (N, obs_dim)), device=f'cuda:{local}')
    K, W = args.steps, args.warmup
    gen = torch.Generator(device=f'cuda:{local}'); gen.manual_seed(1234 + rank)
    acts = (torch.rand((K + W, N, m.nu), device=f'cuda:{local}', generator=gen) - 0.5) * (2 * WL['act_scale'])      # ctrl rows [N][nu], resident in HBM
    # identity permutation between action and ctrl order is irrelevant for a random policy
    gather_list = [torch.empty((N, obs_dim), device=f'cuda:{local}') for _ in range(world)] if (world > 1 and rank == 0) else None
    bad_total = 0

    def one_step(k):
        with torch.cuda.stream(stream):
            sim.set_control_device(acts[k].data_ptr())
            sim.step(N_SUB)
            sim.pack_obs()
            if world > 1:
                sharding.gather_observations(obs, world, rank, gather_list)

    for k in range(W):
        one_step(k)
    sim.sync(); torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local); sampler.start()
    l0 = sim.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    with torch.cuda.stream(stream):
        ev0.record(stream)
        for k in range(W, W + K):
            one_step(k)
        ev1.record(stream)
    sim.sync(); torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1)
    launches = sim.launch_count - l0
    # the same K steps once more with a CUDA event pair around every kernel launch (this pass cannot replay the step graph,
    # so it is kept out of `value`): per-kernel durations for the roofline block
    sim.profile(True)
    pv0, pv1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        pv0.record(stream)
        for k in range(W, W + K):
            one_step(k)
        pv1.record(stream)
    sim.sync(); torch.cuda.synchronize()
    ms_profiled = pv0.elapsed_time(pv1)
    prof = sim.profile_read(); sim.profile(False)
    if world > 1:
        t = torch.tensor([ms], device=f'cuda:{local}'); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = float(t.item())
        dist.barrier()
    flags = sim.get(st.FLAGS)[:, 0]
    bad_total = int((flags != 0).sum())

    # ---- e2e through the public env API (host actions in pinned memory, observation record out)
    if WL['model'] == 'walk':
        env = fly_envs.walk_imitation(terminal_com_dist=float('inf'), n_envs=N, device=local, reset_noise=0.05, seed=1234 + rank,
                                      device_task=not args.host_task)
    else:
        env = fly_envs.flight_imitation(n_envs=N, device=local, seed=1234 + rank, device_task=not args.host_task)
    env.reset()
    na = WL['n_action']
    host_act = torch.empty((K + W, N, na), dtype=torch.float32).pin_memory()
    host_act.copy_(torch.from_numpy(rs.uniform(-WL['act_scale'], WL['act_scale'], (K + W, N, na)).astype(np.float32)))
    a_np = host_act.numpy()
    for k in range(W):
        env.step(a_np[k])
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for k in range(W, W + K):
        ts = env.step(a_np[k])
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device=f'cuda:{local}'); dist.all_reduce(t, op=dist.ReduceOp.MAX); e2e_s = float(t.item())
    sampler.stop_flag = True; sampler.join(timeout=2)
    clocks = sampler.summary()

    if rank == 0:
        peak, peak_src = measured_peaks()
        total_envs = N * world
        value = total_envs * K / (ms * 1e-3)
        # dominant kernel of the step
        dom = max(prof.items(), key=lambda kv: kv[1][0])
        dom_name, (dom_ms, dom_n) = dom
        step_ms_sum = sum(v[0] for v in prof.values())
        avg_launch_s = dom_ms / max(dom_n, 1) * 1e-3
        # one launch of a stage kernel = one substep over N envs; fused launch groupings (FB_FUSE) run 1 or N_SUB whole
        # substeps per launch: the substeps a launch of the dominant kernel covers = timed substeps / its launch count
        substeps_per_launch = max(1, round(K * N_SUB / max(dom_n, 1)))
        alg_bytes_per_launch = BYTES_PER_ENV_SUBSTEP * N * substeps_per_launch
        achieved = alg_bytes_per_launch / avg_launch_s / 1e9
        traffic = None
        tp = os.path.join(ROOT, 'profiles', 'dominant_kernel_traffic.json')
        if os.path.exists(tp) and WL['model'] == 'walk':        # the ncu capture is of the walk model's kernels
            traffic = json.load(open(tp)).get('dram_bytes_per_launch')
        line = {
            'metric': f'env-steps/sec on {WL["model"]}_imitation (control steps of {N_SUB} substeps)', 'value': value, 'unit': 'env-steps/s',
            'n_gpus': world, 'steps': K, 'warmup': W, 'ms_per_step': ms / K, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': f'{WL["model"]}_imitation {N} envs per GPU, random policy U(-{WL["act_scale"]},{WL["act_scale"]}), {N_SUB} substeps x {WL["dt"]} s '
                                   f'({WL["config"]})',
                       'envs_per_gpu': N, 'total_envs': total_envs, 'n_substeps': N_SUB,
                       'l2': f'inputs larger than L2: every launch streams the env records ({sim.record_bytes / 1e6:.2f} MB each, {sim.record_bytes * N / 1e9:.2f} GB per GPU vs 126 MB L2); no explicit flush',
                       'parallelism': f'env-sharded x{world}' + (', torch.distributed NCCL gather of packed obs per control step' if world > 1 else ''),
                       'unstable_envs_flagged': bad_total},
            'clocks': clocks, 'gpu_launches': int(launches),
            'e2e': {'value': total_envs * K / e2e_s, 'unit': 'env-steps/s', 'h2d_bytes_per_step': int(env.h2d_bytes_per_step),
                    'd2h_bytes_per_step': int(env.d2h_bytes_per_step), 'api': f'flybody_b200.fly_envs.{WL["model"]}_imitation(n_envs' + ('' if args.host_task else ', device_task=True') + ').step(action)'},
            'roofline': {'bound': 'hbm', 'kernel': dom_name, 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                         'traffic': traffic, 'peak_source': peak_src,
                         'how': f'algorithmic bytes = {BYTES_PER_ENV_SUBSTEP} B/env-substep x {N} envs x {substeps_per_launch} substep(s) per launch / mean CUDA-event duration of the '
                                f'"{dom_name}" kernel over a second pass of the same {K} steps with an event pair around every launch ({dom_n} launches, '
                                f'{ms_profiled / K:.3f} ms per step in that pass; `value` is the pass without per-launch events, which replays the step graph); whole-step algorithmic GB/s = '
                                f'{BYTES_PER_ENV_STEP * total_envs * K / (ms * 1e-3) / 1e9:.2f}',
                         'kernel_share': dom_ms / max(step_ms_sum, 1e-9),
                         'stage_ms_per_step': {k: v[0] / K for k, v in prof.items() if v[1] > 0}},
        }
        if world == 1 and not args.no_cpu:
            line['cpu_baseline'] = cpu_baseline(budget_s=args.cpu_seconds)
            line['cpu_kernel_source'] = cpu_kernel_source_baseline(budget_s=min(args.cpu_seconds, 5.0))
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--envs', type=int, default=ENVS_PER_GPU, help='envs per GPU')
    ap.add_argument('--cpu-seconds', type=float, default=10.0)
    ap.add_argument('--no-cpu', action='store_true')
    ap.add_argument('--workload', default='walk', choices=sorted(WORKLOADS), help='walk = the headline (BASELINE configs[1]); flight = configs[2]')
    ap.add_argument('--host-task', action='store_true', help='e2e leg: task hooks in host numpy instead of on the device (fb_task_*)')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    set_workload(args.workload)
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
