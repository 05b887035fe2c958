"""Multi-GPU partitioning: independent streams, replicas only.
No collective touches pixel data; `torch.distributed` is used for the barrier
and for the max-over-ranks of the elapsed time only."""


def streams_of_rank(n_streams, world_size, rank):
    """Static assignment stream i -> rank i mod world_size."""
    return [i for i in range(n_streams) if i % world_size == rank]


def job_throughput(units_per_rank, ms_per_rank):
    """Whole-job throughput: all units of all ranks over the slowest rank's time."""
    return sum(units_per_rank) / (max(ms_per_rank) * 1e-3)


def max_over_ranks(dist, value, device=None):
    """Reduce a python float with MAX over the default process group."""
    import torch
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
