"""Core binding of one-process-per-GPU runs (no torch import: meant to run before torch starts its thread pools)."""
import os


def bind_rank_cores(local_rank=None, local_world=None):
    """One process per GPU on one node: give this rank its own contiguous share of the cores the process may run
    on, so that the ranks' host threads (planner, copies' staging, JSON) do not migrate over each other.  Call it
    before the first ``import torch`` of the process where possible (thread pools inherit the mask).
    ``W2T_BIND_CORES=0`` disables it.  Returns the cores kept (or None)."""
    if os.environ.get("W2T_BIND_CORES", "1") == "0" or not hasattr(os, "sched_setaffinity"):
        return None
    local_rank = int(os.environ.get("LOCAL_RANK", "0")) if local_rank is None else int(local_rank)
    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", os.environ.get("WORLD_SIZE", "1"))) if local_world is None \
        else int(local_world)
    cores = sorted(os.sched_getaffinity(0))
    share = len(cores) // max(local_world, 1)
    if local_world <= 1 or share < 1:
        return None
    mine = cores[local_rank * share:(local_rank + 1) * share]
    os.sched_setaffinity(0, mine)
    return mine
