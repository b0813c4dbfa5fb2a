"""Ray-sharded data parallelism for the render hot path: the grid is replicated, every rank renders its
slice of the ray batch into local gradient buffers, and the gradients + touched-voxel masks are summed across ranks
before the (identical, redundant) optimizer step.  The reference has no multi-GPU path; the contract is "equal to the
single-GPU result on the concatenated batch", which needs the fused losses normalised by the GLOBAL ray count
(svox2_csrc.set_loss_norm_rays).
"""
import torch
import torch.distributed as dist


def allreduce_grads(G, group=None):
    """Sum density / surface / SH gradients and OR the touched masks over all ranks (dense buckets)."""
    mask_u8 = G.mask.view(torch.uint8)
    dist.all_reduce(mask_u8, op=dist.ReduceOp.MAX, group=group)
    dist.all_reduce(G.density, op=dist.ReduceOp.SUM, group=group)
    if G.surface is not None:
        dist.all_reduce(G.surface, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(G.sh, op=dist.ReduceOp.SUM, group=group)
