"""Collectives on the guidance path: one process per GPU, torch.distributed (NCCL on GPUs, gloo in
the CPU tests).  The path has exactly two exchanges:

  1. all-gather of the per-image class probabilities (the balanced assignment is global over the
     batch)  -- customized_all_gather E1:222-235, call sites E3:1978-1986;
  2. all-reduce(SUM) of the Monte-Carlo plan counts -- E3:1535.  The counts are exact int32 here,
     so the sum is order independent, every rank ends up bit-identical, and the reference's
     follow-up broadcasts (E3:2017-2020) are unnecessary.

Both messages are <= 512 KiB at the BASELINE configs: latency, not NVLink bandwidth, is the cost.
Two transports: a single NCCL call per exchange on a packed buffer (gather_probs / all_reduce_counts;
also what the gloo CPU tests run), and -- on GPUs that can map each other's memory -- PeerExchange:
one-shot stores / loads over NVLink between symmetric buffers with epoch flags (csrc/fg_peer.cu), which
needs no host call between the kernels, so the whole multi-rank step is ONE CUDA graph.
"""
import ctypes
import os

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


class PeerExchange:
    """Symmetric-memory transport for the two exchanges of the path (csrc/fg_peer.cu; include/fairguide.h "multi-GPU
    exchanges").  One allocation of the same layout on every rank, mapped into every peer by
    torch.distributed._symmetric_memory (CUDA VMM handles over the NVLink fabric; no NVSHMEM, no NCCL on the data path):

        [flags 2 x 64 uint32][rows: 2 parities x world slots x slot_bytes][counts: 2 parities x counts_bytes]

    COLLECTIVE: every rank of ``group`` must construct it at the same point with the same sizes, and must run the same
    sequence of steps afterwards (a step = begin_step, push_rows, wait_rows, publish_counts, sum_counts, each at most
    once, in that order).  All methods are plain kernel launches on the current stream (graph-capturable)."""

    FLAGS_BYTES = 2 * 64 * 4

    @staticmethod
    def available(device=None):
        """True when the symmetric-memory allocator of this torch build can be used on the current GPUs.
        FG_PEER_EXCHANGE=0 turns the transport off (NCCL is used instead)."""
        if os.environ.get("FG_PEER_EXCHANGE", "1") == "0":
            return False
        try:
            import torch.distributed._symmetric_memory as symm  # noqa: F401
        except Exception:
            return False
        return torch.cuda.is_available() and tdist.is_available() and tdist.is_initialized() and tdist.get_backend() == "nccl"

    def __init__(self, slot_bytes, counts_bytes, device, group=None):
        import torch.distributed._symmetric_memory as symm
        from . import _lib
        self._lib = _lib
        self.group = group if group is not None else tdist.group.WORLD
        self.world, self.rank = tdist.get_world_size(self.group), tdist.get_rank(self.group)
        if slot_bytes % 16 or counts_bytes % 16:
            raise ValueError("PeerExchange: sizes must be multiples of 16 bytes")
        self.slot_bytes, self.counts_bytes = int(slot_bytes), int(counts_bytes)
        self.rows_off = self.FLAGS_BYTES
        self.rows_parity = self.world * self.slot_bytes
        self.counts_off = self.rows_off + 2 * self.rows_parity
        total = self.counts_off + 2 * self.counts_bytes
        self.buf = symm.empty(total, dtype=torch.uint8, device=device)
        self.buf.zero_()
        self.hdl = symm.rendezvous(self.buf, self.group)
        self.peer_base_dev = ctypes.c_void_p(int(self.hdl.buffer_ptrs_dev))
        self.epoch = torch.zeros(1, dtype=torch.int32, device=device)
        self.done = torch.zeros(1, dtype=torch.int32, device=device)
        self.status = torch.zeros(4, dtype=torch.int32, device=device)
        torch.cuda.synchronize(device)
        tdist.barrier(group=self.group)               # every rank's flags are zero before the first remote store

    def _call(self, name, *args):
        rc = getattr(self._lib.lib(), name)(*args)
        self._lib.CALLS[name] += 1
        if rc != 0:
            raise RuntimeError(f"{name} failed: {self._lib.lib().fg_error_string(rc).decode()}")

    @staticmethod
    def _p(t):
        return ctypes.c_void_p(t.data_ptr())

    @staticmethod
    def _st():
        return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

    def begin_step(self):
        self._call("fg_peer_epoch_advance", self._p(self.epoch), self._st())

    def push_rows(self, packed):
        """This rank's packed {indicator, probs} rows -> slot `rank` of every peer's row region, then the flag."""
        nbytes = packed.numel() * packed.element_size()
        assert packed.is_contiguous() and nbytes == self.slot_bytes
        self._call("fg_peer_push", self._p(packed), nbytes, self.peer_base_dev, self.rows_off, self.rows_parity,
                   self.rank * self.slot_bytes, 0, 0, 0, self.rank, self.world, self._p(self.epoch), self._p(self.done), self._st())

    def wait_rows(self, gathered):
        """Blocks the stream until every rank's rows of this step have arrived; ``gathered`` [world * n, w] receives them."""
        nbytes = gathered.numel() * gathered.element_size()
        assert gathered.is_contiguous() and nbytes == self.world * self.slot_bytes
        self._call("fg_peer_wait_copy", self.peer_base_dev, self.rank, self.world, 0, 0, self.rows_off, self.rows_parity,
                   self._p(self.epoch), self._p(gathered), nbytes, self._p(self.status), self._st())

    def push_rows_from(self, face_indicators, probs):
        """push_rows without the packed intermediate: the {indicator, probs} row block is assembled on the device from the head's
        outputs and stored into every peer by ONE kernel (fg_peer_push_rows)."""
        n = face_indicators.shape[0]
        ind = (face_indicators.view(torch.uint8) if face_indicators.dtype == torch.bool else face_indicators.to(torch.uint8)).contiguous()
        ps = [p.detach().contiguous() for p in probs]
        widths = (ctypes.c_int32 * len(ps))(*[p.shape[1] for p in ps])
        ptrs = (ctypes.c_void_p * len(ps))(*[p.data_ptr() for p in ps])
        dt = {torch.float32: 0, torch.bfloat16: 1, torch.float16: 2}[ps[0].dtype]
        assert n * (1 + sum(p.shape[1] for p in ps)) * ps[0].element_size() == self.slot_bytes
        self._call("fg_peer_push_rows", self._p(ind), ptrs, widths, len(ps), n, self.peer_base_dev, self.rows_off, self.rows_parity,
                   self.rank * self.slot_bytes, 0, self.rank, self.world, self._p(self.epoch), self._p(self.done), dt, self._st())

    def wait_rows_into(self, n_all, widths, dtype, device):
        """wait_rows + unpack_probs in one kernel (fg_peer_wait_unpack) -> (face_indicators_all bool [n_all], [probs_all ...])."""
        ind_all = torch.empty((n_all,), dtype=torch.uint8, device=device)
        outs = [torch.empty((n_all, w), dtype=dtype, device=device) for w in widths]
        wa = (ctypes.c_int32 * len(outs))(*widths)
        ptrs = (ctypes.c_void_p * len(outs))(*[o.data_ptr() for o in outs])
        dt = {torch.float32: 0, torch.bfloat16: 1, torch.float16: 2}[dtype]
        self._call("fg_peer_wait_unpack", self.peer_base_dev, self.rank, self.world, 0, self.rows_off, self.rows_parity, self._p(self.epoch),
                   self._p(ind_all), ptrs, wa, len(outs), n_all, self._p(self.status), dt, self._st())
        return ind_all.view(torch.bool), outs

    def sum_counts(self, counts):
        """All-reduce(SUM) of the int32 plan counts, in place: publish, wait for every rank, pull and add in rank order."""
        nbytes = counts.numel() * 4
        assert counts.is_contiguous() and counts.dtype == torch.int32 and nbytes <= self.counts_bytes and counts.numel() % 4 == 0
        self._call("fg_peer_push", self._p(counts), nbytes, self.peer_base_dev, self.counts_off, self.counts_bytes, 0, 1, 0, 1,
                   self.rank, self.world, self._p(self.epoch), self._p(self.done), self._st())
        self._call("fg_peer_wait_sum", self.peer_base_dev, self.rank, self.world, 0, 1, self.counts_off, self.counts_bytes,
                   self._p(self.epoch), self._p(counts), counts.numel(), self._p(self.status), self._st())
        return counts

    def check(self):
        """Raises if a wait of any earlier step timed out (one 16-byte read; synchronises)."""
        bits = int(self.status[0].item())
        if bits:
            raise RuntimeError(f"fairguide PeerExchange: a peer's flag did not arrive (status {bits:#x}); the results of that step are invalid")
