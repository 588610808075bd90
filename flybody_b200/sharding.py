"""Multi-GPU plumbing for the env-sharded stepper (reference scaling model: one env per Ray actor,
`flybody/train_dmpo_ray.py:206-227`).

Environments are independent, so the data path has no collective: each rank owns a contiguous env range and its own
stepper handle.  The only exchange is per control step, outside the physics: the packed observation block
[N_local, obs_dim] of every rank is gathered to rank 0 (actor side), and actions are scattered back.
Backends: `nccl` on the GPUs (device tensors aliasing the stepper's observation buffer), `gloo` in the CPU tests.
"""
import numpy as np


def env_range(rank, world, total):
    """[lo, hi) of the envs rank `rank` owns; the first `total % world` ranks take one extra env."""
    base, extra = divmod(int(total), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def rank_seed(base_seed, rank):
    """per-rank RNG stream (BASELINE config 4: seeds 1234 + rank)."""
    return int(base_seed) + int(rank)


def gather_observations(obs, world, rank, gather_list=None, dst=0):
    """Gather each rank's [N_local, obs_dim] block to rank `dst`.  Returns the list of blocks on `dst`, None elsewhere.
    `obs` is a torch tensor (CUDA tensor aliasing fb_obs_ptr for nccl; CPU tensor for gloo); equal N_local per rank."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return [obs]
    if rank == dst and gather_list is None:
        gather_list = [torch.empty_like(obs) for _ in range(world)]
    dist.gather(obs, gather_list if rank == dst else None, dst=dst)
    return gather_list if rank == dst else None


def scatter_actions(actions_all, n_local, nu, world, rank, src=0, device=None):
    """Scatter rank `src`'s [world * n_local, nu] actions; every rank receives its [n_local, nu] rows."""
    import torch
    import torch.distributed as dist
    out = torch.empty((n_local, nu), dtype=torch.float32, device=device)
    if world == 1:
        out.copy_(torch.as_tensor(np.asarray(actions_all), dtype=torch.float32))
        return out
    chunks = None
    if rank == src:
        a = torch.as_tensor(np.asarray(actions_all), dtype=torch.float32, device=device).reshape(world, n_local, nu)
        chunks = [a[r].contiguous() for r in range(world)]
    dist.scatter(out, chunks, src=src)
    return out


class ActorExchange:
    """The per-control-step exchange of BASELINE.json configs[3] ("NCCL obs gather to rank 0 for the DMPO actor loop"): rank 0 is the
    actor.  `scatter_actions` sends every rank its [n_local, n_action] rows of rank 0's [world, n_local, n_action] action tensor;
    `gather` brings every rank's observation rows [n_local, obs_dim] and (reward, discount, step_type, 0) rows [n_local, 4] to rank 0.
    Tensors live on the device (nccl) or the host (gloo, CPU tests); the receive buffers are allocated once.  With world == 1 both calls
    are local no-ops.  bench.py times exactly these two calls inside its step loop; tests/test_multirank.py runs them under gloo."""

    def __init__(self, world, rank, n_local, n_action, device=None):
        import torch
        self.world, self.rank, self.n_local, self.n_action = int(world), int(rank), int(n_local), int(n_action)
        self.a_loc = torch.empty((n_local, n_action), dtype=torch.float32, device=device)
        self.obs_all = self.out_all = None

    def scatter_actions(self, actions_all):
        """actions_all: [world, n_local, n_action] float32 on rank 0 (ignored elsewhere) -> this rank's [n_local, n_action]"""
        import torch.distributed as dist
        if self.world == 1:
            return actions_all[0]
        dist.scatter(self.a_loc, [actions_all[r] for r in range(self.world)] if self.rank == 0 else None, src=0)
        return self.a_loc

    def gather(self, obs, out):
        """-> (list of [n_local, obs_dim], list of [n_local, 4]) on rank 0, (None, None) elsewhere"""
        import torch
        import torch.distributed as dist
        if self.world == 1:
            return [obs], [out]
        if self.rank == 0 and self.obs_all is None:
            self.obs_all = [torch.empty_like(obs) for _ in range(self.world)]
            self.out_all = [torch.empty_like(out) for _ in range(self.world)]
        dist.gather(obs, self.obs_all if self.rank == 0 else None, dst=0)
        dist.gather(out, self.out_all if self.rank == 0 else None, dst=0)
        return (self.obs_all, self.out_all) if self.rank == 0 else (None, None)
