"""``get_num_workers`` of ``detnet/trainer/utils/__init__.py:33-58``.

The ensemble CLI keeps its ``-j/--jobs`` flag for drop-in compatibility and validates it the
way the reference does (same ``RuntimeError``), although the CUDA path merges every image in
one launch and never forks workers."""
import os


def get_num_workers(jobs, device='cpu'):
    """``jobs`` <= 0 means "all but ``-jobs``" of the available units; out-of-range raises."""
    kind = getattr(device, 'type', device)
    if kind == 'cpu':
        available = os.cpu_count()
    elif kind in ('gpu', 'cuda'):
        import torch
        available = torch.cuda.device_count()
    else:
        raise RuntimeError(f"unknown device {kind}")
    wanted = jobs if jobs > 0 else available + jobs
    if wanted < 0 or wanted > available:
        raise RuntimeError("System doesn't have so many {}: {} vs {}".format(kind, jobs, os.cpu_count()))
    return wanted
