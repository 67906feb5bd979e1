"""Multi-GPU plumbing for the only part of the path that shards: loop-closure candidate
batches.  One process per GPU; pairs are dealt to ranks with no data-path collective, and the fixed-size result
records are gathered once with torch.distributed (NCCL over NVLink on the GPU box, gloo in CPU tests)."""
import numpy as np


def partition_pairs(sizes, rank, world):
    """Indices of the pairs rank `rank` verifies: sort by size (descending, index as tie-break) and deal round-robin,
    so every rank receives ceil(P/W) or floor(P/W) pairs of similar total size."""
    order = sorted(range(len(sizes)), key=lambda i: (-int(sizes[i]), i))
    return sorted(order[rank::world])


def gather_records(local_records, n_total, pair_ids, rank, world):
    """All-gathers per-rank record tensors (n_local, R) into one (n_total, R) tensor ordered by pair id.
    local_records lives on the device the process group communicates on; ranks may hold different counts."""
    import torch
    import torch.distributed as dist
    if world == 1:
        out = torch.empty((n_total, local_records.shape[1]), dtype=local_records.dtype, device=local_records.device)
        out[pair_ids.to(local_records.device)] = local_records
        return out
    cap = (n_total + world - 1) // world
    R = local_records.shape[1]
    send = torch.zeros((cap, R + 1), dtype=local_records.dtype, device=local_records.device)
    n_local = local_records.shape[0]
    send[:n_local, :R] = local_records
    send[:, R] = -1
    send[:n_local, R] = pair_ids.to(device=local_records.device, dtype=local_records.dtype)
    recv = torch.empty((world * cap, R + 1), dtype=local_records.dtype, device=local_records.device)
    dist.all_gather_into_tensor(recv, send)
    ids = recv[:, R].round().long()
    valid = ids >= 0
    out = torch.empty((n_total, R), dtype=local_records.dtype, device=local_records.device)
    out[ids[valid]] = recv[valid][:, :R]
    return out


def records_to_array(records):
    """lgs_align_result ctypes records -> (n, 24) float64 array
    [T(16), fitness, trans_prob, iterations, converged, evaluations, line_search_trials, hessian_recomputes, pair_id]
    (float64: a fitness of DBL_MAX - no correspondence within range - stays finite, pair ids stay exact)."""
    out = np.zeros((len(records), 24), np.float64)
    for i, r in enumerate(records):
        out[i, :16] = np.array(r.T, np.float64)
        out[i, 16] = r.fitness
        out[i, 17] = r.trans_probability
        out[i, 18:24] = [r.iterations, r.converged, r.evaluations, r.line_search_trials, r.hessian_recomputes, r.pair_id]
    return out
