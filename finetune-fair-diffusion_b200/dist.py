"""Collectives on the guidance path: one process per GPU, torch.distributed (NCCL on GPUs, gloo in
the CPU tests).  The path has exactly two exchanges:

  1. all-gather of the per-image class probabilities (the balanced assignment is global over the
     batch)  -- customized_all_gather E1:222-235, call sites E3:1978-1986;
  2. all-reduce(SUM) of the Monte-Carlo plan counts -- E3:1535.  The counts are exact int32 here,
     so the sum is order independent, every rank ends up bit-identical, and the reference's
     follow-up broadcasts (E3:2017-2020) are unnecessary.

Both messages are <= 256 KiB at the BASELINE configs: latency, not NVLink bandwidth, is the cost,
so each is a single NCCL call on a packed buffer.
"""
import torch
import torch.distributed as tdist


def _world(group=None):
    if not (tdist.is_available() and tdist.is_initialized()):
        return 1, 0
    return tdist.get_world_size(group), tdist.get_rank(group)


def customized_all_gather(tensor, accelerator=None, return_tensor_other_processes=False, group=None):
    """Drop-in for E1:222-235: concatenation over ranks along dim 0 (rank order), optionally also
    the concatenation of the OTHER ranks' tensors.  ``accelerator`` may be None (torch.distributed
    world) or any object with num_processes / local_process_index / device, as in the reference."""
    world, rank = _world(group)
    if accelerator is not None:
        world, rank = accelerator.num_processes, accelerator.local_process_index
    src = tensor.detach().contiguous()
    if world == 1:
        out = src.clone()
    else:
        wire = src.view(torch.uint8) if src.dtype == torch.bool else src
        out = torch.empty((world * wire.shape[0],) + tuple(wire.shape[1:]), dtype=wire.dtype, device=wire.device)
        tdist.all_gather_into_tensor(out, wire, group=group)
        if src.dtype == torch.bool:
            out = out.view(torch.bool)
    if not return_tensor_other_processes:
        return out
    n = src.shape[0]
    if world > 1:
        others = torch.cat([out[:n * rank], out[n * (rank + 1):]], dim=0)
    else:
        others = torch.empty([0] + list(out.shape[1:]), device=out.device, dtype=out.dtype)
    return out, others


def gather_probs(face_indicators, probs, group=None):
    """ONE collective for everything the assignment needs: packs {indicator, probs_*} of this
    rank's n images into an [n, 1+sum(widths)] buffer, all-gathers it, and unpacks.
    -> (face_indicators_all bool [N], [probs_all ...])."""
    world, _ = _world(group)
    widths = [p.shape[1] for p in probs]
    if world == 1:
        return face_indicators.clone(), [p.detach().clone() for p in probs]
    packed = pack_probs(face_indicators, probs)
    out = torch.empty((world * packed.shape[0], packed.shape[1]), dtype=packed.dtype, device=packed.device)
    all_gather_packed(out, packed, group)
    return unpack_probs(out, widths)


def pack_probs(face_indicators, probs):
    """[n, 1+sum(widths)] row block {indicator, probs_*} in the probs dtype: what one rank contributes to the all-gather."""
    dt = probs[0].dtype
    return torch.cat([face_indicators.to(dt).unsqueeze(1)] + [p.detach() for p in probs], dim=1).contiguous()


def all_gather_packed(out, packed, group=None):
    """The collective itself, on caller-owned buffers (static across CUDA-graph replays)."""
    tdist.all_gather_into_tensor(out, packed, group=group)
    return out


def unpack_probs(gathered, widths):
    """Inverse of pack_probs on the gathered [N, 1+sum(widths)] buffer -> (face_indicators_all bool [N], [probs_all ...])."""
    ind = gathered[:, 0] != 0
    cols, res = 1, []
    for w in widths:
        res.append(gathered[:, cols:cols + w].contiguous())
        cols += w
    return ind, res


def all_reduce_counts(counts, group=None):
    """Sum of the per-rank int32 plan counts, in place (E3:1535)."""
    world, _ = _world(group)
    if world > 1:
        tdist.all_reduce(counts, op=tdist.ReduceOp.SUM, group=group)
    return counts
