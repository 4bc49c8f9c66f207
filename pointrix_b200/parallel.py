"""View-sharded data parallelism for the render path.

One process per GPU, full replica of the Gaussian table; each rank renders its own
views.  The reference defines the semantics through ``batch_size=B`` on one GPU:
parameter gradients are the MEAN over the views of a batch (the loss is a mean over
the stacked batch, pointrix/model/loss.py:27-46), the densification statistic is the
SUM over views of each ``ndc.grad`` (pointrix/controller/gs.py:274-278), ``radii`` is
the MAX and ``visibility`` the ANY over views (pointrix/model/renderer/msplat.py:211-212).
``world_size`` ranks x 1 view therefore equals a reference batch of ``world_size``.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch
import torch.distributed as dist


def shard_views(num_views: int, rank: int, world: int, step: int = 0) -> List[int]:
    """Views of one step: view_id % world == rank, with a shared rotation by ``step``."""
    return [v for v in range(num_views) if (v + step) % world == rank]


def _coalesce(tensors: Sequence[torch.Tensor]) -> List[torch.Tensor]:
    """Tensors that tile one contiguous range of a shared storage (the fused backward's flat
    gradient buffer) are replaced by ONE flat view of that range."""
    groups = {}
    for t in tensors:
        groups.setdefault((t.untyped_storage().data_ptr(), t.dtype), []).append(t)
    out: List[torch.Tensor] = []
    for (_, _dt), ts in groups.items():
        ts = sorted(ts, key=lambda t: t.storage_offset())
        tiled = len(ts) > 1 and all(t.is_contiguous() for t in ts) and all(
            a.storage_offset() + a.numel() == b.storage_offset() for a, b in zip(ts, ts[1:]))
        if tiled:
            lo = ts[0].storage_offset()
            total = ts[-1].storage_offset() + ts[-1].numel() - lo
            out.append(ts[0].as_strided((total,), (1,), lo))
        else:
            out.extend(ts)
    return out


def begin_radii_reduce(radii: torch.Tensor, world: int, group=None):
    """``radii`` is final after the forward: start its MAX all-reduce then, underneath the backward.
    Pass the returned handle to :func:`allreduce_step`."""
    if world <= 1 or not dist.is_initialized():
        return None
    return dist.all_reduce(radii, op=dist.ReduceOp.MAX, group=group, async_op=True)


def allreduce_step(param_grads: Sequence[torch.Tensor], ndc_grad: Optional[torch.Tensor], radii: torch.Tensor,
                   world: int, group=None, average: bool = True, radii_work=None) -> torch.Tensor:
    """In-place exchange after a local backward.  ``radii`` is max-reduced; returns the batch visibility
    (``max radii > 0`` -- no collective of its own is needed).  Works on NCCL and gloo.

    ``average=True`` (each rank back-propagated its own un-scaled view loss): everything is scaled by
    1/world and summed -- the parameter gradients become the mean over views AND every view's
    ``ndc_grad`` carries 1/world before the sum, exactly as in the reference, where the batch loss is
    a mean (pointrix/model/loss.py:27-46), so each view's ``ndc.grad`` already holds 1/B when
    ``accumulate_viewspace_grad`` sums them (pointrix/controller/gs.py:274-278).

    ``average=False``: the caller's loss already carries the 1/world factor (the reference's mean
    over the stacked batch, pointrix/model/loss.py:27-46, applies it to every view's loss and
    therefore to every view's ``ndc.grad`` too): everything is summed, no scaling pass, and the
    gradients of the fused backward -- views of one flat buffer -- go out as ONE all-reduce."""
    if world <= 1 or not dist.is_initialized():
        return radii > 0
    works = []
    grads = [g for g in param_grads if g is not None]
    bufs = grads + ([ndc_grad] if ndc_grad is not None else [])
    if average:
        inv = 1.0 / world
        for g in _coalesce(bufs):
            g.mul_(inv)
    for t in _coalesce(bufs):
        works.append(dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group, async_op=True))
    if radii_work is None:
        radii_work = dist.all_reduce(radii, op=dist.ReduceOp.MAX, group=group, async_op=True)
    works.append(radii_work)
    for w in works:
        w.wait()
    return radii > 0


class NvlsGradExchange:
    """Per-step exchange of the fused render path over NVLink 5 / NVSwitch without NCCL on the data path.

    The fused backward writes every P-sized gradient (59 parameter floats + 2 ndc floats per Gaussian)
    straight into a *symmetric* buffer (``torch.distributed._symmetric_memory``: same allocation on
    every rank, mapped into all peers and into one multicast address); :meth:`exchange` then runs
    ``pxb_nvls_allreduce`` -- each rank pulls its 1/world slice reduced inside the switch
    (``multimem.ld_reduce``: SUM for the gradients, MAX for radii) and multicasts the result back
    (``multimem.st``) -- between two cross-rank barriers.  Afterwards the ``.grad`` tensors the backward
    handed to autograd (views of the buffer) and ``radii`` hold the batch values on every rank.

    Semantics are those of :func:`allreduce_step` with ``average=False``: the caller's loss carries the
    1/world factor.  ``nbuf`` buffers rotate, so the gradients of step k stay valid until the backward
    of step k + nbuf (an optimizer consumes them long before).  Use once per backward.
    """

    def __init__(self, P: int, device, group=None, nbuf: int = 2, mode: str = "auto"):
        import ctypes as C

        import torch.distributed._symmetric_memory as symm

        from . import _lib

        self._C, self._lib = C, _lib
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        self.P = int(P)
        self.device = torch.device(device)
        q = 4 * self.world
        self.n_f32 = (61 * self.P + q - 1) // q * q
        self.n_i32 = (self.P + q - 1) // q * q
        self.bufs, self.hdls = [], []
        for _ in range(nbuf):
            t = symm.empty(self.n_f32 + self.n_i32, dtype=torch.float32, device=self.device)
            h = symm.rendezvous(t, self.group.group_name)
            t.zero_()
            self.bufs.append(t)
            self.hdls.append(h)
        # multicast (in-switch reduction) moves (1 + 1/n) buffer sizes per GPU and direction, plain peer
        # loads/stores 2(n-1)/n: peer-to-peer wins for two GPUs, the switch from four on
        has_mc = all(bool(h.has_multicast_support) and bool(h.multicast_ptr) for h in self.hdls)
        if mode == "auto":
            mode = "nvls" if (has_mc and self.world > 2) else "p2p"
        if mode == "nvls" and not has_mc:
            raise RuntimeError("NVLS multicast is not available for this group")
        self.mode = mode
        self._peer_arrays = [(C.c_void_p * self.world)(*[int(x) for x in h.buffer_ptrs]) for h in self.hdls]
        self._next = 0
        self._cur = None
        torch.cuda.synchronize(self.device)
        dist.barrier(self.group)

    # called by the fused backward: where the flat gradient buffer of this step lives
    def next_buffer(self, numel: int) -> Optional[torch.Tensor]:
        if numel != 61 * self.P:
            return None
        if self._cur is not None:
            # the previous backward's gradients still alias buffer `_cur` (autograd adopted the views):
            # a second backward would rotate onto / overwrite memory `.grad` points into
            raise RuntimeError("NvlsGradExchange: one backward per exchange() -- the previous backward has not been "
                               "exchanged yet (render_batch with several views per rank, or gradient accumulation, "
                               "must use parallel.allreduce_step instead)")
        self._cur = self._next
        self._next = (self._next + 1) % len(self.bufs)
        return self.bufs[self._cur][:numel]

    def exchange(self, radii: torch.Tensor) -> torch.Tensor:
        """All-reduce the gradients written by the last backward (SUM) and ``radii`` (MAX, in place);
        returns the batch visibility."""
        if self._cur is None:
            raise RuntimeError("no backward has written into the exchange buffer since the last exchange")
        buf, h = self.bufs[self._cur], self.hdls[self._cur]
        self._last = self._cur
        self._cur = None
        ri = buf[self.n_f32:].view(torch.int32)
        ri[: self.P].copy_(radii.reshape(-1))
        stream = torch.cuda.current_stream(self.device)
        h.barrier(channel=0)  # every rank's replica is written
        if self.mode == "nvls":
            self._lib.launch("pxb_nvls_allreduce", self._C.c_void_p(h.multicast_ptr), self.n_f32, self.n_i32, self.rank,
                             self.world, self._C.c_void_p(stream.cuda_stream))
        else:
            self._lib.launch("pxb_p2p_allreduce", self._peer_arrays[self._last], self.n_f32, self.n_i32, self.rank,
                             self.world, self._C.c_void_p(stream.cuda_stream))
        h.barrier(channel=1)  # every slice has been multicast back
        radii.reshape(-1).copy_(ri[: self.P])
        return radii > 0


class ShFactoredExchange:
    """Per-step exchange of the fused render path that never moves the SH gradient (79 % of the bytes).

    dL/dshs of one view is the outer product ``basis(dir) (x) gated dL/drgb`` per Gaussian, so each rank
    publishes only its ``d_rgb[P,3]`` and its camera centre in symmetric memory; the 13 remaining floats per
    Gaussian (rotation 4, position 3, scaling 3, opacity 1, ndc 2) and ``radii`` are all-reduced in place as
    in :class:`NvlsGradExchange` (``pxb_nvls_allreduce`` from 4 GPUs on, ``pxb_p2p_allreduce`` for 2), and
    ``pxb_sh_grad_gather`` rebuilds ``sum_views basis (x) d_rgb`` on every rank, reading the peers' ``d_rgb``
    over NVLink inside the kernel.  Per Gaussian and GPU that is 52 B all-reduced + 12 (world-1) B gathered
    instead of 244 B all-reduced.  Same contract as :class:`NvlsGradExchange`: install with
    ``renderer.set_grad_sink``, ONE backward per :meth:`exchange`, the loss carries 1/world; after
    :meth:`exchange` the ``.grad`` tensors autograd received and ``radii`` hold the batch values on every rank.
    The ``shs`` gradient tensor handed to autograd is completed by :meth:`exchange` (it is a plain tensor,
    the others are views of the symmetric buffer).
    """

    def __init__(self, P: int, device, group=None, nbuf: int = 2, mode: str = "auto"):
        import ctypes as C

        import torch.distributed._symmetric_memory as symm

        from . import _lib

        self._C, self._lib = C, _lib
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        self.P = int(P)
        self.device = torch.device(device)
        q = 4 * self.world
        self.n_f32 = (13 * self.P + q - 1) // q * q      # all-reduced floats
        self.n_i32 = (self.P + q - 1) // q * q           # radii
        self.cam_off = self.n_f32 + self.n_i32           # camera centre (3 floats, padded to 4)
        self.rgb_off = self.cam_off + 4                  # d_rgb[P,3]
        total = self.rgb_off + 3 * self.P
        self.bufs, self.hdls = [], []
        for _ in range(nbuf):
            t = symm.empty(total, dtype=torch.float32, device=self.device)
            h = symm.rendezvous(t, self.group.group_name)
            t.zero_()
            self.bufs.append(t)
            self.hdls.append(h)
        has_mc = all(bool(h.has_multicast_support) and bool(h.multicast_ptr) for h in self.hdls)
        if mode == "auto":
            mode = "nvls" if (has_mc and self.world > 2) else "p2p"
        if mode == "nvls" and not has_mc:
            raise RuntimeError("NVLS multicast is not available for this group")
        self.mode = mode + "+sh-gather"
        self._nvls = mode == "nvls"
        self._peer_arrays = [(C.c_void_p * self.world)(*[int(x) for x in h.buffer_ptrs]) for h in self.hdls]
        self._next = 0
        self._cur = None
        self._pending = None
        # the all-reduce (NVLink bound) and the SH gather (latency / HBM bound) touch disjoint data: the
        # gather runs on a side stream between the two barriers
        self._side = torch.cuda.Stream(device=self.device)
        self._ev0, self._ev1 = torch.cuda.Event(), torch.cuda.Event()
        torch.cuda.synchronize(self.device)
        dist.barrier(self.group)

    def next_buffer(self, numel: int):  # the flat-buffer protocol is not used by this sink
        return None

    def plan(self, P: int, dev, cam_center: torch.Tensor, sh_degree: int):
        if P != self.P or torch.device(dev) != self.device:
            return None
        if self._cur is not None:
            raise RuntimeError("ShFactoredExchange: one backward per exchange() -- the previous backward has not been "
                               "exchanged yet (its d_rgb / camera centre would be lost and its gradient views "
                               "overwritten); use parallel.allreduce_step for several views per rank")
        self._cur = self._next
        self._next = (self._next + 1) % len(self.bufs)
        buf = self.bufs[self._cur]
        d_rot = buf[0:4 * P].view(P, 4)
        d_pos = buf[4 * P:7 * P].view(P, 3)
        d_sc = buf[7 * P:10 * P].view(P, 3)
        d_op = buf[10 * P:11 * P]
        d_ndc = buf[11 * P:13 * P].view(P, 2)
        buf[self.cam_off:self.cam_off + 3].copy_(cam_center.reshape(-1)[:3])
        d_rgb = buf[self.rgb_off:self.rgb_off + 3 * P].view(P, 3)
        # completed by exchange().  Autograd gets a view and this object keeps the base: AccumulateGrad adopts
        # an incoming gradient without copying only if nobody else holds that very tensor object
        base = torch.empty(48 * P, dtype=torch.float32, device=self.device)
        self._pending = (base, int(sh_degree))
        return base.view(P, 16, 3), d_rot, d_pos, d_sc, d_op, d_ndc, d_rgb

    def exchange(self, radii: torch.Tensor, position: torch.Tensor) -> torch.Tensor:
        """All-reduce the 13 non-SH gradient floats (SUM) and ``radii`` (MAX, in place) and rebuild the summed
        SH gradient from every rank's ``d_rgb``; ``position`` is the [P,3] tensor the views were rendered
        with.  Returns the batch visibility."""
        if self._cur is None or self._pending is None:
            raise RuntimeError("no backward has written into the exchange buffer since the last exchange")
        C, lib = self._C, self._lib
        cur, (d_sh, sh_degree) = self._cur, self._pending
        self._cur = self._pending = None
        buf, h = self.bufs[cur], self.hdls[cur]
        pos = position.detach()
        if pos.dtype != torch.float32 or not pos.is_contiguous():
            pos = pos.float().contiguous()
        ri = buf[self.n_f32:self.n_f32 + self.n_i32].view(torch.int32)
        ri[: self.P].copy_(radii.reshape(-1))
        main = torch.cuda.current_stream(self.device)
        stream = C.c_void_p(main.cuda_stream)
        h.barrier(channel=0)  # every rank's replica (gradients, d_rgb, camera centre) is written
        self._ev0.record(main)
        self._side.wait_event(self._ev0)
        # the side stream is made current around the launch so that a stage timer's events (recorded on the current
        # stream by _lib.launch) bracket the kernel where it runs -- on the main stream they measured nothing
        with torch.cuda.stream(self._side):
            lib.launch("pxb_sh_grad_gather", self._peer_arrays[cur], self.rgb_off, self.cam_off, self.world, self.P,
                       sh_degree, C.c_void_p(pos.data_ptr()), C.c_void_p(d_sh.data_ptr()), C.c_void_p(self._side.cuda_stream))
        self._ev1.record(self._side)
        if self._nvls:
            lib.launch("pxb_nvls_allreduce", C.c_void_p(h.multicast_ptr), self.n_f32, self.n_i32, self.rank, self.world,
                       stream)
        else:
            lib.launch("pxb_p2p_allreduce", self._peer_arrays[cur], self.n_f32, self.n_i32, self.rank, self.world, stream)
        main.wait_event(self._ev1)
        h.barrier(channel=1)  # every slice has landed and every peer has read this rank's d_rgb
        radii.reshape(-1).copy_(ri[: self.P])
        return radii > 0
