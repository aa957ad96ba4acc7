"""Data-parallel training of the hot path's parameters.

The reference wraps the whole model in torch's DistributedDataParallel when world_size > 1 (nerf/utils.py:439-441).  That works
unchanged on the drop-in modules (their parameters are ordinary nn.Parameters: `encoder.embeddings`, `sigma_net.weights`,
`color_net.weights`), and `ddp(model, local_rank)` is exactly that call.  `allreduce_gradients` is the explicit form for training
loops that do not use DDP: ONE all-reduce (NCCL over NVLink / NVSwitch on a B200 box) of all gradients packed into a flat buffer.
The table gradient is produced in fp16 by the backward kernel (atomicAdd on __half2, like the reference) and only widened to fp32 by
autograd, so exchanging it in fp16 loses nothing on the way in and halves the 48 MB all-reduce of the 12.2 M-entry table.
"""
import torch
import torch.distributed as dist


def ddp(model, local_rank=None):
    """what nerf/utils.py:439-441 does: SyncBatchNorm conversion (a no-op for these models) + DistributedDataParallel"""
    model = torch.nn.SyncBatchNorm.convert_sync_batchnorm(model)
    ids = None if local_rank is None else [local_rank]
    return torch.nn.parallel.DistributedDataParallel(model, device_ids=ids)


@torch.no_grad()
def allreduce_gradients(parameters, group=None, average=True, comm_dtype=None):
    """Sum (or average) `p.grad` of every parameter over the ranks of `group` with a single all-reduce.
    comm_dtype: dtype on the wire (default: fp16 on CUDA, where the table gradient is fp16-valued anyway; fp32 elsewhere).
    Parameters without a gradient contribute zeros (every rank must pass the same parameter list)."""
    params = [p for p in parameters]
    if not params or not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    dev = params[0].device
    if comm_dtype is None:
        comm_dtype = torch.float16 if dev.type == "cuda" else torch.float32
    sizes = [p.numel() for p in params]
    flat = torch.zeros(sum(sizes), dtype=comm_dtype, device=dev)
    off = 0
    for p, n in zip(params, sizes):
        if p.grad is not None:
            flat[off:off + n].copy_(p.grad.reshape(-1))
        off += n
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)          # the step's only collective
    if average:
        flat /= dist.get_world_size(group)
    off = 0
    for p, n in zip(params, sizes):
        g = flat[off:off + n].view_as(p)
        if p.grad is None:
            p.grad = g.to(p.dtype)
        else:
            p.grad.copy_(g)
        off += n
