"""N>1 path on CPU: world_size-2 `gloo` ranks, each stepping its own env shard on the host-emulation build, actions
scattered from rank 0 and packed observations gathered to rank 0 (the exchange bench.py does over NCCL).  The result must
equal one process stepping all envs: env results do not depend on the rank that owns them."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import __graft_entry__ as ge
from flybody_b200 import sharding
from flybody_b200 import stepper as st
from flybody_b200.flymodel import load_model

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N_TOTAL, N_STEPS, N_SUB = 6, 3, 10


def _inputs(m):
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    from parity_common import reset_qpos
    rs = np.random.RandomState(5)
    q = np.tile(reset_qpos(m), (N_TOTAL, 1))
    q[:, 7:109] += rs.uniform(-0.05, 0.05, (N_TOTAL, 102))
    acts = rs.uniform(-0.5, 0.5, (N_STEPS, N_TOTAL, m.nu)).astype(np.float32)
    return q, acts


def _worker(rank, world, port, emu, out_path):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    m = load_model('walk')
    q, acts = _inputs(m)
    lo, hi = sharding.env_range(rank, world, N_TOTAL)
    sim = st.BatchedStepper(m, hi - lo, lib_path=emu)
    dim = sim.obs_ptr()[1]                      # default packed layout: qpos, qvel, act, sensor means, ...
    sim.reset(q[lo:hi])
    blocks = None
    for k in range(N_STEPS):
        a = sharding.scatter_actions(acts[k] if rank == 0 else None, hi - lo, m.nu, world, rank)
        sim.set_control(a.numpy())
        sim.step(N_SUB)
        obs = torch.from_numpy(sim.read_obs(np.empty((hi - lo, dim), np.float32)).copy())
        blocks = sharding.gather_observations(obs, world, rank)
    if rank == 0:
        np.save(out_path, torch.cat(blocks).numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_env_range_partitions():
    for total in (1, 7, 4096, 32768):
        for world in (1, 2, 3, 8):
            r = [sharding.env_range(k, world, total) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == total
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1


def test_two_rank_gloo_matches_single_process(tmp_path):
    ge.build()
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    out = str(tmp_path / 'gathered.npy')
    mp.spawn(_worker, args=(2, port, ge.EMU, out), nprocs=2, join=True)
    got = np.load(out)
    m = load_model('walk')
    q, acts = _inputs(m)
    sim = st.BatchedStepper(m, N_TOTAL, lib_path=ge.EMU)
    dim = sim.obs_ptr()[1]                      # default packed layout: qpos, qvel, act, sensor means, ...
    sim.reset(q)
    for k in range(N_STEPS):
        sim.set_control(acts[k])
        sim.step(N_SUB)
        want = sim.read_obs(np.empty((N_TOTAL, dim), np.float32))
    assert got.shape == want.shape
    assert np.array_equal(got, want)


def _task_worker(rank, world, port, emu, out_path):
    """device-side task logic under sharding: each rank runs `walk_imitation(n_envs=N/world, device_task=True)`; rank 0
    scatters the actions and gathers observation rows + (reward, discount, step_type)."""
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from flybody_b200 import fly_envs
    lo, hi = sharding.env_range(rank, world, N_TOTAL)
    env = fly_envs.walk_imitation(terminal_com_dist=0.05, n_envs=hi - lo, lib_path=emu, device_task=True)
    env.reset()
    rs = np.random.RandomState(9)
    acts = rs.uniform(-0.5, 0.5, (16, N_TOTAL, 59)).astype(np.float32)
    rows = []
    xch = sharding.ActorExchange(world, rank, hi - lo, 59)             # the exchange bench.py times at N > 1 (there over nccl)
    for k in range(16):
        a = xch.scatter_actions(torch.from_numpy(acts[k]).reshape(world, hi - lo, 59) if rank == 0 else None)
        env.step(a.numpy())
        obs_all, out_all = xch.gather(torch.from_numpy(env._rec.copy()), torch.from_numpy(env._out4.copy()))
        if rank == 0:
            rows.append(np.concatenate([torch.cat(obs_all).numpy(), torch.cat(out_all).numpy()], 1))
    if rank == 0:
        np.save(out_path, np.array(rows))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_device_task_rollout_matches_single_process(tmp_path):
    ge.build()
    from flybody_b200 import fly_envs
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    out = str(tmp_path / 'rollout.npy')
    mp.spawn(_task_worker, args=(2, port, ge.EMU, out), nprocs=2, join=True)
    got = np.load(out)
    env = fly_envs.walk_imitation(terminal_com_dist=0.05, n_envs=N_TOTAL, lib_path=ge.EMU, device_task=True)
    env.reset()
    rs = np.random.RandomState(9)
    acts = rs.uniform(-0.5, 0.5, (16, N_TOTAL, 59)).astype(np.float32)
    seen_last = False
    for k in range(16):
        env.step(acts[k])
        want = np.concatenate([env._rec, env._out4], 1)
        assert np.array_equal(got[k], want), k            # env results do not depend on the rank that owns them
        seen_last |= bool((env._out4[:, 2] == 2).any())
    assert seen_last                                      # the comparison ran through terminations and auto-resets
