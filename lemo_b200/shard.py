"""Sequence sharding over the GPUs of one box.

Independent motion sequences are the unit of data parallelism: sequence s goes to rank s % world_size, one
process per GPU, and there is NO collective on the data path (the reference has no cross-sequence state:
opt_amass_temp.py:251-458).  torch.distributed is used only to gather the [N,T,72] results on rank 0.
"""
import numpy as np
import torch


def assign(n_sequences, world_size, rank):
    """Round-robin: the global sequence ids this rank fits."""
    return list(range(rank, n_sequences, world_size))


def owner(seq_id, world_size):
    return seq_id % world_size


def gather_results(local_ids, local_params72, n_sequences, world_size, rank, group=None):
    """local_params72 [len(local_ids), T, 72] on any device -> on rank 0 the full [N,T,72] (numpy), else None.
    Uses gather_object so it works on gloo (CPU tests) and nccl alike; payload is ~34 KB per sequence."""
    import torch.distributed as dist
    payload = (list(local_ids), np.asarray(local_params72.detach().cpu().numpy() if torch.is_tensor(local_params72) else local_params72))
    if world_size == 1 or not dist.is_initialized():
        parts = [payload]
    else:
        parts = [None] * world_size if rank == 0 else None
        dist.gather_object(payload, parts, dst=0, group=group)
        if rank != 0:
            return None
    T = payload[1].shape[1] if payload[1].ndim == 3 else 0
    out = np.zeros((n_sequences, T, 72), np.float32)
    seen = np.zeros(n_sequences, bool)
    for ids, arr in parts:
        for i, s in enumerate(ids):
            assert not seen[s], 'sequence %d fitted twice' % s
            out[s] = arr[i]
            seen[s] = True
    assert seen.all(), 'sequences missing from the gather: %s' % np.nonzero(~seen)[0].tolist()
    return out
