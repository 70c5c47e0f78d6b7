"""Multi-GPU plumbing: the problem batch is sharded by contiguous blocks, one process per GPU, no traffic on the
hot path; the only collective is the final gather of counters / outcome codes.  Works with the
``nccl`` backend on GPUs and with ``gloo`` on CPU (tests)."""
from __future__ import annotations

import os


def env_world():
    return int(os.environ.get('RANK', 0)), int(os.environ.get('LOCAL_RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))


def shard_range(total: int, rank: int, world: int):
    """Contiguous block of problem indices owned by ``rank`` (sizes differ by at most one)."""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def init(backend: str):
    import torch.distributed as dist
    rank, local_rank, world = env_world()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        os.environ.setdefault('MASTER_PORT', '29511')
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, local_rank, world


def finalize():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        dist.destroy_process_group()


def barrier():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        dist.barrier()


def all_gather_vector(values, device='cpu'):
    """values: list of numbers on this rank -> [world][len] nested list on every rank."""
    import torch
    import torch.distributed as dist
    t = torch.tensor(values, dtype=torch.float64, device=device)
    if not (dist.is_available() and dist.is_initialized()):
        return [t.cpu().tolist()]
    out = [torch.empty_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return [o.cpu().tolist() for o in out]


def gather_outcomes(outcome, device='cpu'):
    """Concatenate per-rank outcome codes (int32 arrays of possibly different length) on every rank."""
    import numpy as np
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return np.asarray(outcome)
    n = all_gather_vector([len(outcome)], device)
    nmax = int(max(v[0] for v in n))
    pad = torch.full((nmax,), -1, dtype=torch.int32, device=device)
    pad[:len(outcome)] = torch.as_tensor(np.asarray(outcome), dtype=torch.int32, device=device)
    out = [torch.empty_like(pad) for _ in range(dist.get_world_size())]
    dist.all_gather(out, pad)
    return np.concatenate([o.cpu().numpy()[:int(k[0])] for o, k in zip(out, n)])


def outcome_counts(outcome):
    """The four numbers printed by scripts/mpc.py:287-291 from outcome bit codes."""
    import numpy as np
    from . import abi
    o = np.asarray(outcome)
    conv = (o & abi.OUT_CONVERGED) != 0
    coll = (o & abi.OUT_COLLIDED) != 0
    viable = ((o & abi.OUT_ABORTED) != 0) & ~conv & ~coll
    return {'completed': int(conv.sum()), 'collisions': int(coll.sum()), 'viable': int(viable.sum()),
            'not_converged': int(len(o) - (conv | coll | viable).sum())}
