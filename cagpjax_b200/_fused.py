"""Differentiable torch entry points over the C ABI (custom VJPs of the kernel projection).

The reference differentiates this path with plain JAX reverse mode through ``lax.map`` +
``jax.checkpoint`` (operators/lazy_kernel.py:79-103); here the forward and the reverse pass are
hand-written sm_100a kernels and these ``torch.autograd.Function``s are the custom VJP glue
(a JAX host would wrap the same C entry points in ``jax.custom_vjp``).

Row sharding: every rank holds all of X / s / y (<= 40 MB at n = 1M) and owns the
rows ``[row_begin, row_end)`` of X1, hence of M' and of its cotangent.  The k x k partials are
summed with one all-reduce in the forward and the n-sized cotangents with one in the backward.
"""

from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Optional, Tuple

import torch

from . import _native as N

# Gram operators use the symmetric kernels (half the kernel evaluations); False forces the general
# rectangular kernels (kept for A/B measurements and as the cross-covariance path).
USE_SYMMETRIC = True


# --------------------------------------------------------------------------------------------
# Row partition / process-group plumbing
# --------------------------------------------------------------------------------------------
@dataclass(frozen=True)
class RowShard:
    """Rows [begin, end) of the n training points owned by this rank; ``group`` None => single GPU."""

    begin: int
    end: int
    n: int
    group: object = None  # torch.distributed process group (or None)
    # block-aligned shards (symmetric Gram path): action blocks [a_begin, a_end) of k, owned by `rank`
    a_begin: int = -1
    a_end: int = -1
    k: int = 0
    rank: int = 0
    world: int = 1

    @property
    def sharded(self) -> bool:
        return self.group is not None

    @staticmethod
    def full(n: int) -> "RowShard":
        return RowShard(0, n, n, None)

    @property
    def block_aligned(self) -> bool:
        return self.a_begin >= 0

    @staticmethod
    def block_range(k: int, rank: int, world: int) -> Tuple[int, int]:
        base, rem = divmod(k, world)
        a0 = rank * base + min(rank, rem)
        return a0, a0 + base + (1 if rank < rem else 0)

    @staticmethod
    def for_rank_blocks(n: int, k: int, rank: int, world: int, group=None) -> "RowShard":
        """Shard whose boundaries coincide with action-block boundaries (needs k >= world): rank r owns
        the action blocks ``block_range(k, r, world)`` and their rows.  The symmetric Gram kernels
        assign tiles by block, so a rank's tiles produce/consume whole blocks of rows."""
        if k < world:
            raise ValueError("block-aligned sharding needs n_actions >= world size")
        bs = n // k
        a0, a1 = RowShard.block_range(k, rank, world)
        begin = a0 * bs
        end = n if a1 == k else a1 * bs
        return RowShard(begin, end, n, group, a0, a1, k, rank, world)

    def peer_rows(self, r: int) -> Tuple[int, int]:
        """Row range of rank r under the same block-aligned partition."""
        bs = self.n // self.k
        a0, a1 = RowShard.block_range(self.k, r, self.world)
        return a0 * bs, (self.n if a1 == self.k else a1 * bs)

    @staticmethod
    def for_rank(n: int, rank: int, world: int, group=None) -> "RowShard":
        """Even split of rows; the remainder goes to the first ranks (any split is exact, F6)."""
        base, rem = divmod(n, world)
        begin = rank * base + min(rank, rem)
        end = begin + base + (1 if rank < rem else 0)
        return RowShard(begin, end, n, group)


def _all_reduce_sum(t: torch.Tensor, group) -> torch.Tensor:
    if group is None:
        return t
    import torch.distributed as dist

    with N._timed("host:all_reduce", t):
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


def _all_to_all(send: list, recv: list, group) -> None:
    """recv[r] <- rank r's send[this rank].  NCCL all_to_all; point-to-point on backends without it (gloo)."""
    import torch.distributed as dist

    if dist.get_backend(group) == "nccl":
        dist.all_to_all(recv, send, group=group)
        return
    me = dist.get_rank(group)
    recv[me].copy_(send[me])
    ops = []
    for r in range(dist.get_world_size(group)):
        if r != me:
            ops.append(dist.P2POp(dist.isend, send[r], dist.get_global_rank(group, r) if group is not dist.group.WORLD else r, group))
            ops.append(dist.P2POp(dist.irecv, recv[r], dist.get_global_rank(group, r) if group is not dist.group.WORLD else r, group))
    for req in dist.batch_isend_irecv(ops):
        req.wait()


def exchange_tiles_to_rows(Mfull: torch.Tensor, shard: RowShard) -> torch.Tensor:
    """Symmetric forward, ranks > 1.  Every rank holds ``Mfull`` (k, n): the entries its own tiles
    produced, zero elsewhere.  Entry (b, i) is produced by exactly one rank (the owner of blk(i) for
    row sums, the owner of b for column sums).  Returns the complete local panel Mt (k, rows of this
    rank): rank h sends block-rows own(h) x columns rows(g) to g, which adds them to its own part."""
    with N._timed("host:exchange_tiles_to_rows", Mfull):
        return _exchange_tiles_to_rows(Mfull, shard)


def _exchange_tiles_to_rows(Mfull: torch.Tensor, shard: RowShard) -> torch.Tensor:
    k, n = Mfull.shape
    world, me = shard.world, shard.rank
    a0, a1 = shard.a_begin, shard.a_end
    send, recv = [], []
    for r in range(world):
        r0, r1 = shard.peer_rows(r)
        send.append(Mfull[a0:a1, r0:r1].contiguous())
        b0, b1 = RowShard.block_range(k, r, world)
        recv.append(torch.empty((b1 - b0, shard.end - shard.begin), dtype=Mfull.dtype, device=Mfull.device))
    _all_to_all(send, recv, shard.group)
    Mt = Mfull[:, shard.begin:shard.end].contiguous()
    for r in range(world):
        if r != me:
            b0, b1 = RowShard.block_range(k, r, world)
            Mt[b0:b1] += recv[r]
    return Mt


def gather_cotangent_for_tiles(Gt_local: torch.Tensor, shard: RowShard, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Symmetric backward, ranks > 1.  The tiles (a in own, b) need G[i in own rows, b] (local) and
    G[j, a in own] for ALL j (remote rows).  Builds a (k, n) buffer holding both: rank h sends the
    block-rows own(g) of its local panel to g."""
    with N._timed("host:gather_cotangent_for_tiles", Gt_local):
        return _gather_cotangent_for_tiles(Gt_local, shard, out)


def _gather_cotangent_for_tiles(Gt_local: torch.Tensor, shard: RowShard, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    k = Gt_local.shape[0]
    world, me = shard.world, shard.rank
    a0, a1 = shard.a_begin, shard.a_end
    Gfull = out if out is not None else torch.empty((k, shard.n), dtype=Gt_local.dtype, device=Gt_local.device)
    Gfull[:, shard.begin:shard.end] = Gt_local
    send, recv = [], []
    for r in range(world):
        b0, b1 = RowShard.block_range(k, r, world)
        send.append(Gt_local[b0:b1].contiguous())
        r0, r1 = shard.peer_rows(r)
        recv.append(torch.empty((a1 - a0, r1 - r0), dtype=Gt_local.dtype, device=Gt_local.device))
    _all_to_all(send, recv, shard.group)
    for r in range(world):
        if r != me:
            r0, r1 = shard.peer_rows(r)
            Gfull[a0:a1, r0:r1] = recv[r]
    return Gfull


# --------------------------------------------------------------------------------------------
# Peer-memory forward: compute + exchange in one kernel (NVLink / NVSwitch, torch symmetric memory)
# --------------------------------------------------------------------------------------------
USE_PEER_MEMORY = True   # falls back to the all_to_all exchange if symmetric memory is unavailable on ANY rank
_peer_panels = {}        # (group id, k, ld, dtype, device) -> (symmetric tensor, handle) or None (unavailable)


def _peer_panel(shard: RowShard, k: int, ld: int, dtype, device):
    """Cached symmetric-memory scratch panel (k, ld) with its rendezvous handle (peer pointers, barrier), or None
    when symmetric memory cannot be set up.  Availability is decided COLLECTIVELY, once per (group, shape): every
    rank tries the allocation + rendezvous, the outcome is min-reduced over the group, and all ranks take the same
    branch - a rank-local fallback would leave the others waiting in the panel barrier.  Only the set-up is
    guarded; errors of the kernel itself propagate."""
    key = (id(shard.group), k, ld, dtype, str(device))
    if key in _peer_panels:
        return _peer_panels[key]
    import torch.distributed as dist

    ent, why = None, ""
    try:
        import torch.distributed._symmetric_memory as symm_mem

        t = symm_mem.empty((k, ld), dtype=dtype, device=device)
        hdl = symm_mem.rendezvous(t, shard.group)
        ent = (t, hdl)
    except (RuntimeError, ImportError, AttributeError) as exc:  # no symmetric memory on this system / build
        why = str(exc)
    ok = torch.tensor([1 if ent is not None else 0], dtype=torch.int32, device=device)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=shard.group)
    if int(ok.item()) == 0:
        if shard.rank == 0:
            import warnings

            warnings.warn(f"peer-memory forward unavailable on at least one rank ({why or 'another rank failed'}); "
                          "using the all_to_all exchange")
        ent = None
    _peer_panels[key] = ent
    return ent


def _peer_buffer(shard: RowShard, shape, dtype, device, tag: str):
    """Collectively agreed symmetric-memory buffer (see _peer_panel) for other shapes, e.g. the (k, n) cotangent buffer."""
    key = (id(shard.group), tag) + tuple(shape) + (dtype, str(device))
    if key in _peer_panels:
        return _peer_panels[key]
    import torch.distributed as dist

    ent = None
    try:
        import torch.distributed._symmetric_memory as symm_mem

        t = symm_mem.empty(tuple(shape), dtype=dtype, device=device)
        ent = (t, symm_mem.rendezvous(t, shard.group))
    except (RuntimeError, ImportError, AttributeError):
        ent = None
    ok = torch.tensor([1 if ent is not None else 0], dtype=torch.int32, device=device)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=shard.group)
    if int(ok.item()) == 0:
        ent = None
    _peer_panels[key] = ent
    return ent


def sym_forward_peer(family, xs, d, s, k, lengthscale, shard: RowShard, bbox, ent) -> torch.Tensor:
    """Symmetric Gram forward over the ranks of ``shard.group`` with the exchange fused into the kernel:
    column sums that belong to another rank's rows are stored straight into that rank's panel through
    peer-mapped memory.  Returns this rank's complete panel Mt (k, local rows)."""
    world, me = shard.world, shard.rank
    row_begin = [shard.peer_rows(r)[0] for r in range(world)] + [shard.n]
    panel, hdl = ent
    # column sums of a row block split over several CTAs are accumulated atomically: zero exactly those rows
    zero = N.sym_zero_rows(family, xs, d, k, lengthscale, bbox)
    if zero == 2:
        panel.zero_()
    elif zero == 1:
        panel[k - 1].zero_()
    hdl.barrier()         # every rank has zeroed its panel and finished reading it in the previous step
    N.ks_blocksparse_fwd_sym_peer(family, xs, d, s, k, lengthscale, shard.a_begin, shard.a_end,
                                  [int(p) for p in hdl.buffer_ptrs], row_begin, me, panel.shape[1], bbox=bbox)
    hdl.barrier()         # all remote stores have landed
    # a private copy: the panel is overwritten by the next forward, Mt lives until this step's backward (and longer
    # when a caller keeps the state for predict); 8 n k / G bytes of device copy, < 0.3 % of the step
    return panel[:, :shard.end - shard.begin].clone()


# --------------------------------------------------------------------------------------------
# Local compute (the part that runs native kernels).  Tests for the sharding logic replace this
# with an oracle-backed callable on CPU (gloo); the product always uses the native one.
# --------------------------------------------------------------------------------------------
def _box_centre(X: torch.Tensor) -> torch.Tensor:
    """Centre of the bounding box of the points (d,), one reduction kernel."""
    lo, hi = torch.aminmax(X.detach(), dim=0)
    return torch.lerp(lo, hi, 0.5)


import os as _os

# B' = M'^T M' on a side stream, concurrent with the A'-dependent k x k chain (see ProjectStats).  OFF by default:
# measured (r02) no gain at C3 on 1 GPU (1143.2 vs 1143.3 ms), -0.5 ms of 582 on 2 GPUs, and a REGRESSION on small
# steps (C2: 46.3 vs 37.1 ms per step when steps are enqueued back to back).  CAGP_OVERLAP_GRAM=1 enables it.
OVERLAP_GRAM = _os.environ.get("CAGP_OVERLAP_GRAM", "0") == "1"
_side_streams = {}
_pending_side = []   # (completion event, tensors the side stream reads): kept alive until the main stream is ordered after it


def _side_stream(device) -> "torch.cuda.Stream":
    key = str(device)
    if key not in _side_streams:
        _side_streams[key] = torch.cuda.Stream(device=device)
    return _side_streams[key]


def native_local_forward(family, X, lengthscale, s, yt, k, shard: RowShard, gram_stream=None):
    """Returns (A', B', c', Mt, xs, xs1, bbox): partial statistics over the local rows.  ``gram_stream``: CUDA stream
    on which B' is computed (after the pair kernel); the CALLER then owns the synchronisation of B'."""
    d = X.shape[1]
    # origin = centre of the bounding box: stationary kernels only see differences, and centred coordinates keep
    # |x|^2 small, which is what admits the expanded-distance fast path of the symmetric RBF kernels
    xs = N.scale_inputs(family, X, lengthscale, _box_centre(X))
    whole = shard.begin == 0 and shard.end == shard.n
    bbox = N.bounding_box(xs) if USE_SYMMETRIC and (whole or shard.block_aligned) else None
    if whole and USE_SYMMETRIC:
        # Gram case on one GPU: every kernel value is evaluated once per unordered block pair
        xs1 = xs
        Mt = N.ks_blocksparse_fwd_sym(family, xs, d, s, k, lengthscale, bbox=bbox)
    elif shard.block_aligned and USE_SYMMETRIC:
        xs1 = xs
        Mt = None
        if USE_PEER_MEMORY and shard.world <= 16:
            row_begin = [shard.peer_rows(r)[0] for r in range(shard.world)] + [shard.n]
            ld = max(row_begin[r + 1] - row_begin[r] for r in range(shard.world))
            ent = _peer_panel(shard, k, ld, xs.dtype, xs.device)
            if ent is not None:
                with N._timed("host:sym_forward_peer", xs):
                    Mt = sym_forward_peer(family, xs, d, s, k, lengthscale, shard, bbox, ent)
        if Mt is None:
            Mfull = N.ks_blocksparse_fwd_sym(family, xs, d, s, k, lengthscale, shard.a_begin, shard.a_end, bbox=bbox)
            Mt = exchange_tiles_to_rows(Mfull, shard)
            del Mfull
    else:
        xs1 = xs if whole else xs[:, shard.begin:shard.end].contiguous()
        Mt = N.ks_blocksparse_fwd(family, xs1, xs, d, s, k, lengthscale)
    sl = s[shard.begin:shard.end].contiguous()
    yl = yt[shard.begin:shard.end].contiguous()
    if gram_stream is not None:
        main = torch.cuda.current_stream(Mt.device)
        done = torch.cuda.Event()
        done.record(main)
        with torch.cuda.stream(gram_stream):
            gram_stream.wait_event(done)       # M' is complete
            B = N.gram_mtm(Mt)
        A, c = N.project_rows(Mt, shard.begin, shard.n, sl, yl)   # main stream, concurrently
    else:
        A, c = N.project_rows(Mt, shard.begin, shard.n, sl, yl)
        B = N.gram_mtm(Mt)
    return A, B, c, Mt, xs, xs1, bbox


def native_local_backward(family, X, lengthscale, s, yt, k, shard: RowShard, saved, Abar, Bbar, cbar):
    """Returns (s_bar_full_partial (n), yt_bar_local (m), ls_bar partial)."""
    Mt, xs, xs1, bbox = saved
    d = X.shape[1]
    sl = s[shard.begin:shard.end].contiguous()
    yl = yt[shard.begin:shard.end].contiguous()
    W = (Bbar + Bbar.T).contiguous()
    whole = shard.begin == 0 and shard.end == shard.n
    if (not whole and shard.block_aligned and USE_SYMMETRIC and USE_PEER_MEMORY and shard.world <= 16
            and Mt.dtype == torch.float64):
        # compute + exchange in one kernel: the cotangent GEMM stores its panel into this rank's (k, n) tile buffer and
        # every row b also into the buffer of the rank that owns action block b (remote stores from the epilogue)
        ent = _peer_buffer(shard, (k, shard.n), Mt.dtype, Mt.device, "cotangent")
        if ent is not None:
            Gfull, hdl = ent
            block_begin = [RowShard.block_range(k, r, shard.world)[0] for r in range(shard.world)] + [k]
            with N._timed("host:assemble_cotangent_peer", Mt):
                hdl.barrier()   # every rank is done reading its buffer in the previous backward
                N.assemble_cotangent_peer(Mt, shard.begin, shard.n, W, Abar.contiguous(), cbar.contiguous(), sl, yl,
                                          [int(p) for p in hdl.buffer_ptrs], block_begin, shard.rank, shard.n)
                hdl.barrier()   # all remote stores have landed
            s_bar, ls_bar = N.ks_blocksparse_bwd_sym(family, xs, d, s, k, Gfull, lengthscale, shard.a_begin, shard.a_end,
                                                     bbox=bbox)
            s_dir, yt_bar = N.project_rows_bwd(Mt, shard.begin, shard.n, Abar.contiguous(), cbar.contiguous())
            s_bar[shard.begin:shard.end] += s_dir
            return s_bar, yt_bar, ls_bar
    Gt = N.assemble_cotangent(Mt, shard.begin, shard.n, W, Abar.contiguous(), cbar.contiguous(), sl, yl)
    if whole and USE_SYMMETRIC:
        s_bar, ls_bar = N.ks_blocksparse_bwd_sym(family, xs, d, s, k, Gt, lengthscale, bbox=bbox)
    elif shard.block_aligned and USE_SYMMETRIC:
        Gfull = gather_cotangent_for_tiles(Gt, shard)
        s_bar, ls_bar = N.ks_blocksparse_bwd_sym(family, xs, d, s, k, Gfull, lengthscale, shard.a_begin, shard.a_end,
                                                 bbox=bbox)
        del Gfull
    else:
        s_bar, ls_bar = N.ks_blocksparse_bwd(family, xs1, xs, d, s, k, Gt, lengthscale)
    del Gt
    s_dir, yt_bar = N.project_rows_bwd(Mt, shard.begin, shard.n, Abar.contiguous(), cbar.contiguous())
    s_bar[shard.begin:shard.end] += s_dir
    return s_bar, yt_bar, ls_bar


class ProjectStats(torch.autograd.Function):
    """(X, lengthscale, s, yt) -> (A', B', c') with A' = S^T K' S, B' = (K'S)^T (K'S), c' = (K'S)^T yt.

    K' is the unit-variance kernel.  Non-differentiable w.r.t. X (training inputs are data in
    the reference path).  ``holder`` (a dict) receives the saved ``Mt`` so that predict-at-train
    can reuse it without a second n^2 pass.
    """

    @staticmethod
    def forward(ctx, X, lengthscale, s, yt, family, k, shard, holder, local_fwd, local_bwd):
        # B' = M'^T M' is the longest of the HBM-bound / GEMM kernels after the pair kernel, and nothing needs it before
        # the ELBO's closed form; the Cholesky factorisation, the solve and the KL chain only need A', c'.  With a
        # holder to carry the event, B' (and its all-reduce) runs on a side stream and the consumer waits for it
        # (KernelProjection.B).  Not under CUDA-graph capture (the side stream would have to re-join the capture).
        side = None
        if (local_fwd is None and OVERLAP_GRAM and holder is not None and X.is_cuda
                and not torch.cuda.is_current_stream_capturing()):
            side = _side_stream(X.device)
        if side is not None:
            # side-stream work of earlier steps: order the main stream after it before its operands can be recycled
            # (no Tensor.record_stream: that defers block reuse in the caching allocator and cost 60 % at C2)
            main = torch.cuda.current_stream(X.device)
            for ev, _keep in _pending_side:
                if not ev.query():
                    main.wait_event(ev)
            _pending_side.clear()
            A, B, c, *saved = native_local_forward(family, X, lengthscale, s, yt, k, shard, gram_stream=side)
            if shard.sharded:
                packed = torch.cat([A.reshape(-1), c.reshape(-1)])
                _all_reduce_sum(packed, shard.group)          # issued first: NCCL runs collectives in issue order
                A, c = packed[:k * k].view(k, k), packed[k * k:]
                with torch.cuda.stream(side):
                    _all_reduce_sum(B, shard.group)
            ready = torch.cuda.Event()
            ready.record(side)
            holder["B_event"] = ready
            _pending_side.append((ready, (saved[0], B)))   # M' and B' stay referenced until the side stream is done
        else:
            local_fwd = local_fwd or native_local_forward
            A, B, c, *saved = local_fwd(family, X, lengthscale, s, yt, k, shard)
            if shard.sharded:
                packed = torch.cat([A.reshape(-1), B.reshape(-1), c.reshape(-1)])
                _all_reduce_sum(packed, shard.group)
                kk = k * k
                A, B, c = packed[:kk].view(k, k), packed[kk:2 * kk].view(k, k), packed[2 * kk:]
        ctx.family, ctx.k, ctx.shard = family, k, shard
        ctx.local_bwd = local_bwd or native_local_backward
        ctx.saved_extra = saved
        ctx.save_for_backward(X, lengthscale, s, yt)
        if holder is not None:
            holder["Mt"] = saved[0]
            holder["shard"] = shard
        return A, B, c

    @staticmethod
    def backward(ctx, Abar, Bbar, cbar):
        X, lengthscale, s, yt = ctx.saved_tensors
        shard = ctx.shard
        s_bar, yt_bar_local, ls_bar = ctx.local_bwd(ctx.family, X, lengthscale, s, yt, ctx.k, shard,
                                                    ctx.saved_extra, Abar, Bbar, cbar)
        if shard.sharded or yt_bar_local.shape[0] != shard.n:
            yt_bar = torch.zeros_like(yt)
            yt_bar[shard.begin:shard.end] = yt_bar_local
        else:
            yt_bar = yt_bar_local
        if shard.sharded:
            n = shard.n
            packed = torch.cat([s_bar, yt_bar, ls_bar.reshape(-1)])
            _all_reduce_sum(packed, shard.group)
            s_bar, yt_bar, ls_bar = packed[:n], packed[n:2 * n], packed[2 * n:]
        return None, ls_bar.reshape(lengthscale.shape), s_bar, yt_bar, None, None, None, None, None, None


def _like(t: torch.Tensor, X: torch.Tensor) -> torch.Tensor:
    """``t`` in X's dtype on X's device, contiguous.  The kernels compute in the dtype of the inputs X (the reference
    relies on JAX type promotion; here every operand of a native call is cast explicitly - a differentiable no-op
    when it already matches).  Library defaults (gpx.Constant, BlockSparsePolicy.from_random) are float64, so
    float32 data with default hyper-parameters reaches this cast."""
    if t.dtype != X.dtype or t.device != X.device:
        t = t.to(device=X.device, dtype=X.dtype)
    return t.contiguous()


def project_stats(family: str, X, lengthscale, s, yt, k: int, shard: Optional[RowShard] = None, holder=None,
                  local_fwd: Optional[Callable] = None, local_bwd: Optional[Callable] = None):
    shard = shard or RowShard.full(X.shape[0])
    X = X.contiguous()
    return ProjectStats.apply(X, _like(lengthscale, X), _like(s, X), _like(yt, X), family, k, shard, holder,
                              local_fwd, local_bwd)


class KSMatrix(torch.autograd.Function):
    """Mt (k, m) = (K'(X1, X2) S)^T, differentiable w.r.t. lengthscale and s (not X1, X2)."""

    @staticmethod
    def forward(ctx, X1, X2, lengthscale, s, family, k):
        d = X2.shape[1]
        origin = _box_centre(X2)  # centred coordinates (see native_local_forward)
        xs2 = N.scale_inputs(family, X2, lengthscale, origin)
        same = X1.data_ptr() == X2.data_ptr() and X1.shape == X2.shape
        xs1 = xs2 if same else N.scale_inputs(family, X1, lengthscale, origin)
        Mt = N.ks_blocksparse_fwd(family, xs1, xs2, d, s, k, lengthscale, N.union_bounding_box(xs1, xs2))
        ctx.family, ctx.k, ctx.d = family, k, d
        ctx.save_for_backward(xs1, xs2, s, lengthscale)
        return Mt

    @staticmethod
    def backward(ctx, Gt):
        xs1, xs2, s, lengthscale = ctx.saved_tensors
        s_bar, ls_bar = N.ks_blocksparse_bwd(ctx.family, xs1, xs2, ctx.d, s, ctx.k, Gt.contiguous(), lengthscale)
        return None, None, ls_bar.reshape(lengthscale.shape), s_bar, None, None


def ks_matrix(family: str, X1, X2, lengthscale, s, k: int) -> torch.Tensor:
    X2 = X2.contiguous()
    X1 = X2 if X1 is X2 else _like(X1, X2)
    return KSMatrix.apply(X1, X2, _like(lengthscale, X2), _like(s, X2), family, k)


# --------------------------------------------------------------------------------------------
# Segment sums over the action blocks (n-sized, autograd-friendly torch ops)
# --------------------------------------------------------------------------------------------
def block_layout(n: int, n_blocks: int) -> Tuple[int, int, int]:
    """(block_size, n_blocks_main, n_main) - operators/block_diagonal_sparse.py:49-52."""
    block_size = n // n_blocks
    n_blocks_main = n_blocks if n % n_blocks == 0 else n_blocks - 1
    return block_size, n_blocks_main, n_blocks_main * block_size


def segment_sum(v: torch.Tensor, n_blocks: int) -> torch.Tensor:
    """Sum of ``v`` (n,) over each action block -> (n_blocks,)  (linalg/congruence.py:34-41)."""
    n = v.shape[0]
    block_size, n_blocks_main, n_main = block_layout(n, n_blocks)
    if n_blocks_main > 0:
        out = v[:n_main].reshape(n_blocks_main, block_size).sum(dim=1)
    else:
        out = v.new_zeros((0,))
    if n > n_main:
        out = torch.cat([out, v[n_main:].sum(dim=0, keepdim=True)])
    return out


# --------------------------------------------------------------------------------------------
# Dense right-hand side: Y = K'(X1, X2) V   (LazyKernel._matmat, operators/lazy_kernel.py:79-90)
# --------------------------------------------------------------------------------------------
class KernelMatmat(torch.autograd.Function):
    """Y = K'(X1, X2) V for dense V (n, p); differentiable w.r.t. lengthscale and V (not X1, X2)."""

    @staticmethod
    def forward(ctx, X1, X2, lengthscale, V, family):
        d = X2.shape[1]
        xs2 = N.scale_inputs(family, X2, lengthscale, X2[0])
        same = X1.data_ptr() == X2.data_ptr() and X1.shape == X2.shape
        xs1 = xs2 if same else N.scale_inputs(family, X1, lengthscale, X2[0])
        Vc = V.contiguous()
        Y = N.ks_dense_fwd(family, xs1, xs2, d, Vc, lengthscale)
        ctx.family, ctx.d = family, d
        ctx.save_for_backward(xs1, xs2, Vc, lengthscale)
        return Y

    @staticmethod
    def backward(ctx, Ybar):
        xs1, xs2, V, lengthscale = ctx.saved_tensors
        Ybar = Ybar.contiguous()
        V_bar = ls_bar = None
        if ctx.needs_input_grad[3]:
            V_bar = N.ks_dense_fwd(ctx.family, xs2, xs1, ctx.d, Ybar, lengthscale)  # K'^T Ybar, K'(x,y) = K'(y,x)
        if ctx.needs_input_grad[2]:
            ls_bar = N.ks_dense_bwd_ls(ctx.family, xs1, xs2, ctx.d, V, Ybar, lengthscale).reshape(lengthscale.shape)
        return None, None, ls_bar, V_bar, None


class CrossCovariance(torch.autograd.Function):
    """C (m, n) = K'(X1, X2) materialised; differentiable w.r.t. lengthscale and X2 (the trainable pseudo-inputs of
    PseudoInputPolicy, policies/pseudoinput.py:68-74), not X1 (the training inputs are non-trainable there)."""

    @staticmethod
    def forward(ctx, X1, X2, lengthscale, family):
        d = X2.shape[1]
        origin = X1[0].detach()
        xs1 = N.scale_inputs(family, X1, lengthscale, origin)
        xs2 = N.scale_inputs(family, X2, lengthscale, origin)
        C = N.cross_cov_fwd(family, xs1, xs2, d, lengthscale)
        ctx.family, ctx.d = family, d
        ctx.save_for_backward(xs1, xs2, lengthscale)
        return C

    @staticmethod
    def backward(ctx, Cbar):
        xs1, xs2, lengthscale = ctx.saved_tensors
        x2s_bar, ls_bar = N.cross_cov_bwd(ctx.family, xs1, xs2, ctx.d, Cbar, lengthscale)
        X2_bar = None
        if ctx.needs_input_grad[1]:
            scale = N.family_scale(ctx.family) / lengthscale.reshape(-1).to(x2s_bar.dtype)  # (1,) or (d,)
            X2_bar = x2s_bar[:ctx.d].T * scale
        ls_out = ls_bar.reshape(lengthscale.shape) if ctx.needs_input_grad[2] else None
        return None, X2_bar, ls_out, None


def cross_covariance(family: str, X1, X2, lengthscale) -> torch.Tensor:
    """Materialised unit-variance K'(X1, X2) (m, n) on the device."""
    X1 = X1.contiguous()
    return CrossCovariance.apply(X1, _like(X2, X1), _like(lengthscale, X1), family)


# Below this many columns a dense operand is cheaper as p one-block block-sparse passes (the pair
# kernels run at ~950 Gpairs/s; the GEMM-shaped dense kernel shares one kernel tile across up to 128
# output columns and only pays off when there are enough of them).
DENSE_MIN_COLUMNS = 12


def kernel_matmat(family: str, X1, X2, lengthscale, V: torch.Tensor) -> torch.Tensor:
    """K'(X1, X2) @ V for dense V (n, p), differentiable w.r.t. lengthscale and V."""
    X2 = X2.contiguous()
    X1 = X2 if X1 is X2 else _like(X1, X2)
    lengthscale, V = _like(lengthscale, X2), (V if V.dtype == X2.dtype and V.device == X2.device else V.to(device=X2.device, dtype=X2.dtype))
    if V.shape[1] < DENSE_MIN_COLUMNS:
        # every column of V is a one-block BlockDiagonalSparse (k = 1)
        cols = [ks_matrix(family, X1, X2, lengthscale, V[:, c].contiguous(), 1)[0] for c in range(V.shape[1])]
        return torch.stack(cols, dim=1)
    return KernelMatmat.apply(X1, X2, lengthscale, V, family)


# --------------------------------------------------------------------------------------------
# Dense action matrices: the same row-sum-reducible statistics with a dense S (n, k)
# --------------------------------------------------------------------------------------------
class DenseProjectStats(torch.autograd.Function):
    """(X, lengthscale, S, yt) -> (A', B', c') for a DENSE action matrix S (n, k):
    M' = K'(X, X) S,  A' = S^T M',  B' = M'^T M',  c' = M'^T yt   (unit-variance kernel K').

    Rows of X1 (hence of M') are sharded like in ``ProjectStats``: one dense kernel pass over the local
    rows, one packed all-reduce of [A'|B'|c'].  Reverse mode: Mbar = S Abar + M'(Bbar + Bbar^T) + yt cbar^T on the
    local rows, then S_bar = K'(X, X_local) Mbar (dense kernel, roles swapped) and the lengthscale pass;
    one all-reduce of [S_bar | yt_bar | ls_bar]."""

    @staticmethod
    def forward(ctx, X, lengthscale, S, yt, family, shard, holder, kernels=None):
        # ``kernels``: object providing scale_inputs / ks_dense_fwd / ks_dense_bwd_ls; the product always uses the
        # native module, the gloo tests of the sharding logic inject an oracle-backed stand-in (no GPU there)
        KN = N if kernels is None else kernels
        d = X.shape[1]
        xs = KN.scale_inputs(family, X, lengthscale, X[0])
        whole = shard.begin == 0 and shard.end == shard.n
        xs1 = xs if whole else xs[:, shard.begin:shard.end].contiguous()
        Sc = S.contiguous()
        M = KN.ks_dense_fwd(family, xs1, xs, d, Sc, lengthscale)  # (m_local, k)
        Sl, yl = Sc[shard.begin:shard.end], yt[shard.begin:shard.end]
        A, B, c = Sl.T @ M, M.T @ M, M.T @ yl
        if shard.sharded:
            k = S.shape[1]
            packed = torch.cat([A.reshape(-1), B.reshape(-1), c.reshape(-1)])
            _all_reduce_sum(packed, shard.group)
            kk = k * k
            A, B, c = packed[:kk].view(k, k), packed[kk:2 * kk].view(k, k), packed[2 * kk:]
        ctx.family, ctx.d, ctx.shard, ctx.kernels = family, d, shard, KN
        ctx.save_for_backward(xs, xs1, Sc, M, yt, lengthscale)
        if holder is not None:
            holder["M"] = M
        return A, B, c

    @staticmethod
    def backward(ctx, Abar, Bbar, cbar):
        xs, xs1, S, M, yt, lengthscale = ctx.saved_tensors
        shard, KN = ctx.shard, ctx.kernels
        Sl, yl = S[shard.begin:shard.end], yt[shard.begin:shard.end]
        Mbar = (Sl @ Abar + M @ (Bbar + Bbar.T) + torch.outer(yl, cbar)).contiguous()
        S_bar = KN.ks_dense_fwd(ctx.family, xs, xs1, ctx.d, Mbar, lengthscale)  # K'(X, X_local) Mbar, (n, k)
        ls_bar = KN.ks_dense_bwd_ls(ctx.family, xs1, xs, ctx.d, S, Mbar, lengthscale)
        S_bar[shard.begin:shard.end] += M @ Abar.T
        yt_bar = torch.zeros_like(yt)
        yt_bar[shard.begin:shard.end] = M @ cbar
        if shard.sharded:
            n, k = S.shape
            # S_bar (n x k: 4.1 GB at BASELINE config 5) is reduced in place; only the small vectors are packed
            _all_reduce_sum(S_bar, shard.group)
            packed = torch.cat([yt_bar, ls_bar.reshape(-1)])
            _all_reduce_sum(packed, shard.group)
            yt_bar, ls_bar = packed[:n], packed[n:]
        return None, ls_bar.reshape(lengthscale.shape), S_bar, yt_bar, None, None, None, None


def dense_project_stats(family: str, X, lengthscale, S, yt, shard: Optional[RowShard] = None, holder=None, kernels=None):
    shard = shard or RowShard.full(X.shape[0])
    X = X.contiguous()
    return DenseProjectStats.apply(X, _like(lengthscale, X), _like(S, X), _like(yt, X), family, shard, holder, kernels)
