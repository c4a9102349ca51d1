"""Batch-data-parallel training of the recurrent core: one process per GPU, each rank runs the
fused time loop on its own shard (per-replica BatchNorm statistics, exactly like the reference's
nn.DataParallel replicas, mainclean.py:132-134) and ONE fp32 all-reduce of the flat gradient vector
per step (NCCL over NVLink on GPUs; gloo in the CPU tests).  There is no data-path collective.
"""
import torch
import torch.distributed as dist


def shard_first_index(step, rank, world, per_rank_batch):
    """global sample index of this rank's first clip at `step` (disjoint, deterministic shards)."""
    return (step * world + rank) * per_rank_batch


def average_flat_gradient_(flat_grad, group=None):
    """in-place mean over ranks of the flat gradient bucket (equals DataParallel's gradient of the
    global-batch mean loss for equal shards).  Returns the tensor."""
    world = dist.get_world_size(group)
    if world == 1:
        return flat_grad
    if flat_grad.is_cuda and dist.get_backend(group) == "nccl":
        dist.all_reduce(flat_grad, op=dist.ReduceOp.AVG, group=group)      # one NCCL kernel, the mean is taken in it
    else:
        dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM, group=group)      # gloo (CPU tests) has no AVG
        flat_grad.div_(world)
    return flat_grad


def attach_data_parallel(model, group=None):
    """make `model` (any mirror module of this package) average its flat gradient across `group` during backward.

    The hook fires on the flat parameter vector right after the reverse-time kernels have been
    enqueued, i.e. before autograd scatters the bucket back into the 30 parameter tensors, so the
    collective is a single 0.43 MB all-reduce issued from the backward stream."""
    if not dist.is_initialized():
        raise RuntimeError("torch.distributed is not initialised")
    model._dp_group = group
    model._dp_enabled = True
    # parameters a mirror module keeps outside its flat vector (none today) would get their own hook
    for p in getattr(model, "dp_extra_parameters", lambda: [])():
        p.register_hook(lambda g, _grp=group: average_flat_gradient_(g.contiguous(), _grp))
    return model


def flat_gradient_hook(model, flat):
    """called by the mirror modules' forward on their flat parameter vector (ConvLSTM, kys.GRU; InT has it inline)"""
    if getattr(model, "_dp_enabled", False) and flat.requires_grad:
        group = model._dp_group
        flat.register_hook(lambda g: average_flat_gradient_(g.contiguous(), group))
    return flat


def broadcast_parameters(model, src=0, group=None):
    """replicate rank `src`'s parameters (DataParallel's replicate(), mainclean.py:133)."""
    for p in model.parameters():
        dist.broadcast(p.data, src=src, group=group)
