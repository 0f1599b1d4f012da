"""Consumer side of the image gradient: the adjusted-DFT gradient coefficients and the bucketed
gradient synchronisation that replaces the reference's per-tensor all-reduce loop.

    make_grad_hook                 E1:219-220
    adjusted-DFT coefficients      E1:1104-1109   (generate_image_w_gradient)
    manual gradient all-reduce     E1:1996-2011   (same loop at E3:2217-2229)
"""
import math

import numpy as np
import torch
import torch.distributed as tdist

from . import _lib, ops
from ._lib import check
from .ops import _p, _stream


def make_grad_hook(coef):
    """E1:219-220."""
    return lambda x: coef * x


def adjusted_dft_grad_coefs(alphas_cumprod, alphas, timesteps):
    """E1:1104-1109: per-timestep gradient coefficients sqrt(a_bar_t) * sqrt(1 - a_bar_t) / (1 - a_t), normalised by their
    geometric mean.  ``alphas_cumprod`` / ``alphas`` are the scheduler's tables (tensors or arrays), ``timesteps`` the
    scheduler's timestep sequence.  Host arithmetic on ~20 numbers, float64 like the reference's Python floats."""
    ac = torch.as_tensor(alphas_cumprod)
    al = torch.as_tensor(alphas)
    coefs = []
    for t in timesteps:
        t = int(t)
        coefs.append(ac[t].sqrt().item() * (1 - ac[t]).sqrt().item() / (1 - al[t].item()))
    coefs = np.array(coefs)
    coefs /= (math.prod(coefs) ** (1 / len(coefs)))
    return coefs


class GradBucket:
    """Flat fp32 bucket over the gradients of a fixed parameter list (one dtype).  ``layout`` is host logic; pack /
    unpack are one kernel each (csrc/fg_sync.cu); ``sync`` = pack -> ONE all-reduce -> unpack."""

    def __init__(self, params):
        self.params = [p for p in params]
        assert self.params, "GradBucket needs at least one parameter"
        self.offsets_host = self.layout([p.numel() for p in self.params])
        self.total = self.offsets_host[-1]
        dev = self.params[0].device
        self.dtype = self.params[0].dtype
        assert all(p.dtype == self.dtype and p.device == dev for p in self.params), "one dtype / device per bucket"
        if dev.type != "cuda":
            raise RuntimeError("fairguide ops run on CUDA tensors only (no CPU fallback)")
        self.offsets = torch.tensor(self.offsets_host, dtype=torch.int64, device=dev)
        self.bucket = torch.empty((self.total + 1,), dtype=torch.float32, device=dev)
        self.nonfinite = torch.zeros((1,), dtype=torch.int32, device=dev)
        self._ptrs, self._ptr_key = None, None

    @staticmethod
    def layout(numels):
        """Prefix sums of the element counts: tensor t owns bucket[offsets[t] : offsets[t+1]]."""
        off = [0]
        for n in numels:
            off.append(off[-1] + int(n))
        return off

    def _ptr_table(self):
        grads = [p.grad for p in self.params]
        assert all(g is not None and g.is_contiguous() for g in grads), "every parameter needs a contiguous .grad"
        key = tuple(g.data_ptr() for g in grads)
        if key != self._ptr_key:
            # .grad tensors that stay allocated between steps (``zero_grad(set_to_none=False)``) keep their addresses and
            # skip this upload; otherwise the table goes up from pinned memory without blocking the host
            if self._ptrs is None:
                self._ptrs = torch.empty((len(key),), dtype=torch.int64, device=self.bucket.device)
                self._ptrs_host = torch.empty((len(key),), dtype=torch.int64, pin_memory=True)
            else:
                self._ptr_event.synchronize()          # the previous upload has consumed the pinned buffer
            self._ptrs_host.copy_(torch.tensor(key, dtype=torch.int64))
            self._ptrs.copy_(self._ptrs_host, non_blocking=True)
            self._ptr_event = torch.cuda.Event()
            self._ptr_event.record()
            self._ptr_key = key
        return self._ptrs

    def pack(self):
        check(_lib.lib().fg_grad_bucket_pack(_p(self._ptr_table()), _p(self.offsets), len(self.params), self.total, _p(self.bucket),
                                             _p(self.nonfinite), ops._DT[self.dtype], _stream()), "fg_grad_bucket_pack")
        return self.bucket

    def unpack(self, divisor_a, divisor_b=1.0):
        check(_lib.lib().fg_grad_bucket_unpack(_p(self._ptr_table()), _p(self.offsets), len(self.params), self.total, _p(self.bucket),
                                               float(divisor_a), float(divisor_b), ops._DT[self.dtype], _stream()), "fg_grad_bucket_unpack")

    def sync(self, num_processes=None, n_backward=1, group=None):
        """E1:1996-2011 for all parameters at once: returns (grad_is_finite, nonfinite_total) where the first is THIS
        rank's flag like the reference's ``grad_is_finite`` and the second the count summed over all ranks (so every rank
        can take the same skip decision).  One host read instead of one per tensor."""
        world = tdist.get_world_size(group) if (tdist.is_available() and tdist.is_initialized()) else 1
        num_processes = world if num_processes is None else num_processes
        self.pack()
        local = self.nonfinite.clone()
        if world > 1:
            tdist.all_reduce(self.bucket, op=tdist.ReduceOp.SUM, group=group)
        self.unpack(num_processes, n_backward)
        flags = torch.stack([local[0].to(torch.float32), self.bucket[self.total]]).tolist()
        return flags[0] == 0, int(flags[1])


_BUCKETS = {}      # parameter-list identity -> GradBucket; bounded (the reference has at most two lists: UNet and text-encoder LoRA)
_MAX_BUCKETS = 4


def clear_gradient_buckets():
    """Drop the cached buckets (and the fp32 staging memory they hold), e.g. after rebuilding the models."""
    _BUCKETS.clear()


def allreduce_average_gradients(params, num_processes=None, n_backward=1, group=None):
    """Drop-in for the loop at E1:1999-2011: sums ``p.grad`` over the ranks and divides by num_processes and N_backward;
    returns ``grad_is_finite``.  The bucket of a parameter list is cached (least recently used of at most four lists is
    dropped, so changing lists cannot grow memory without bound); ``clear_gradient_buckets()`` frees them explicitly."""
    params = list(params)
    key = tuple(id(p) for p in params)
    bucket = _BUCKETS.pop(key, None)
    if bucket is None or any(a is not b for a, b in zip(bucket.params, params)):
        bucket = GradBucket(params)
    _BUCKETS[key] = bucket                               # re-inserted last: dict order is the LRU order
    while len(_BUCKETS) > _MAX_BUCKETS:
        _BUCKETS.pop(next(iter(_BUCKETS)))
    return bucket.sync(num_processes, n_backward, group)[0]
