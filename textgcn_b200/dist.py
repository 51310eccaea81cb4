"""Multi-GPU hot path on one NVSwitch box: one process per GPU, torch.distributed (NCCL) for the plumbing.

Propagation, default scheme (``GridPropagator``, ``bench.py --gpus N``): a G x R grid of feature slices x user
partitions.  E' = Â·E acts on every embedding column independently, so the FEATURE dimension splits with no exchange per
hop; inside a slice the users are partitioned by nnz, the (6x smaller) item-table slice is all-reduced once per hop
within the row group of R ranks, overlapped with the user-row SpMM, and the change of layout at the end rides in the
last passes as peer-memory stores (CUDA IPC over NVLink).  ``BipartitePropagator`` is the G = 1 hop loop it builds on,
``SlicedPropagator`` the G = P case as one C-ABI call per rank.

The north_star's wording — Â partitioned by ROW BLOCK with an all-gather of layer embeddings between hops — is
``DistPropagator``: rank p owns rows [starts[p], starts[p+1]) (balanced by nnz); one hop = all-gather of the P row blocks
(padded to a common ``max_rows`` so the collective is a single contiguous ``all_gather_into_tensor``) + the local SpMM,
whose column indices were relabelled once to index the padded gathered table.  It is comm-bound (5.1 GB per hop per
rank at the 200M-edge config) and kept as the comparison arm.

Evaluation: sharded by ITEM range (north_star).  Every rank ranks all requested users against its item shard
with the fused score+mask+top-k kernel, the (U, k) partial tables are exchanged with an all-to-all so that rank p
receives every shard's candidates for user slice p, and ``tgcn_topk_merge`` reduces P·k -> k under the same strict
order; the result is sharded by user slice (``gather=True`` all-gathers it).

The compute calls are injected (``spmm_fn`` / ``rank_fn`` / ``merge_fn``) so the partition / relabel / exchange
logic is unit-tested with world_size-2 gloo on CPU; the defaults are the CUDA kernels and nothing else.
"""
from __future__ import annotations

import os
from typing import Callable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def nnz_balanced_starts(rowptr: torch.Tensor, world_size: int) -> List[int]:
    """Row cut points s_0=0 < ... < s_P=N such that every block holds ~nnz/P non-zeros."""
    n = rowptr.numel() - 1
    nnz = int(rowptr[-1])
    targets = torch.tensor([nnz * p // world_size for p in range(1, world_size)], dtype=rowptr.dtype, device=rowptr.device)
    cuts = torch.searchsorted(rowptr.contiguous(), targets, right=False).tolist() if world_size > 1 else []
    starts = [0] + [min(max(int(c), 0), n) for c in cuts] + [n]
    for i in range(1, len(starts)):  # keep monotone (degenerate tiny graphs)
        starts[i] = max(starts[i], starts[i - 1])
    return starts


class RowPartition:
    """Row-block partition of the CSR of Â and the padded-table relabelling of its columns."""

    def __init__(self, rowptr: torch.Tensor, world_size: int):
        self.world_size = world_size
        self.starts = nnz_balanced_starts(rowptr, world_size)
        self.n_rows = rowptr.numel() - 1
        self.max_rows = max(self.starts[p + 1] - self.starts[p] for p in range(world_size))
        # 16-byte alignment of every block start for any d % 4 == 0 is automatic (rows are whole)

    def rows(self, rank: int) -> Tuple[int, int]:
        return self.starts[rank], self.starts[rank + 1]

    def owner(self, idx: torch.Tensor) -> torch.Tensor:
        bounds = torch.tensor(self.starts[1:], device=idx.device, dtype=idx.dtype)
        return torch.searchsorted(bounds, idx, right=True)

    def relabel(self, col: torch.Tensor) -> torch.Tensor:
        """Global row id -> row of the (P·max_rows, d) gathered table."""
        own = self.owner(col.to(torch.int64))
        starts = torch.tensor(self.starts[:-1], device=col.device, dtype=torch.int64)
        return (own * self.max_rows + (col.to(torch.int64) - starts[own])).to(torch.int32)

    def local_block(self, rank: int, rowptr: torch.Tensor, col: torch.Tensor, val: torch.Tensor):
        """(rowptr_local, col_relabelled, val_local) of rank's row block."""
        s, e = self.rows(rank)
        lo, hi = int(rowptr[s]), int(rowptr[e])
        rp = (rowptr[s:e + 1] - rowptr[s]).to(torch.int32).contiguous()
        return rp, self.relabel(col[lo:hi]).contiguous(), val[lo:hi].contiguous()


def _default_spmm(graph, x, y, addends, divisor):
    from . import ops
    return ops.spmm_ex(graph, x, y, addends=addends, divisor=divisor)


class DistPropagator:
    """K-layer propagation over a row-partitioned Â with an all-gather of layer embeddings between hops."""

    def __init__(self, part: RowPartition, rank: int, local_graph, d: int, n_layers: int, device,
                 group=None, spmm_fn: Callable = _default_spmm):
        self.part, self.rank, self.graph, self.d, self.n_layers = part, rank, local_graph, d, n_layers
        self.group, self.spmm_fn = group, spmm_fn
        self.n_local = part.starts[rank + 1] - part.starts[rank]
        P, mr = part.world_size, part.max_rows
        self.gathered = torch.zeros((P * mr, d), dtype=torch.float32, device=device)
        # layer buffers are max_rows tall so they can be handed to the collective without a pad copy
        self.bufs = [torch.zeros((mr, d), dtype=torch.float32, device=device) for _ in range(max(n_layers, 1))]
        self.comm_bytes_per_hop = (P - 1) * mr * d * 4  # received per rank

    def all_gather(self, block: torch.Tensor) -> torch.Tensor:
        dist.all_gather_into_tensor(self.gathered, block, group=self.group)
        return self.gathered

    def propagate(self, e0_local: torch.Tensor, out_local: Optional[torch.Tensor] = None, single: bool = False
                  ) -> torch.Tensor:
        """e0_local: this rank's rows of E0 (n_local, d).  Returns this rank's rows of the layer mean."""
        n, L = self.n_local, self.n_layers
        self.bufs[0][:n].copy_(e0_local)
        if out_local is None:
            out_local = torch.empty((n, self.d), dtype=torch.float32, device=e0_local.device)
        cur = self.bufs[0]
        for layer in range(1, L + 1):
            table = self.all_gather(cur)
            last = layer == L
            if last:
                addends = [] if single else [self.bufs[t][:n] for t in range(L)]
                self.spmm_fn(self.graph, table, out_local, addends, 1.0 if single else float(L + 1))
            else:
                y = self.bufs[layer]
                self.spmm_fn(self.graph, table, y[:n], [], 1.0)
                cur = y
        return out_local

    def gather_full(self, local: torch.Tensor) -> torch.Tensor:
        """All-gather local rows of a table and strip the padding -> (N, d) on every rank."""
        n = self.n_local
        self.bufs[0][:n].copy_(local)
        table = self.all_gather(self.bufs[0])
        mr = self.part.max_rows
        return torch.cat([table[p * mr: p * mr + (self.part.starts[p + 1] - self.part.starts[p])]
                          for p in range(self.part.world_size)])


# ---------------------------------------------------------------------------------------------------------
# Bipartite scheme: users partitioned, item table replicated, one all-reduce of the (I, d) table per hop
# ---------------------------------------------------------------------------------------------------------
def _default_mean(addends, out, divisor):
    from . import ops
    return ops.layer_mean(addends, out, divisor)


class BipartitePartition:
    """Users are split into P contiguous ranges balanced by nnz; every rank keeps the whole item table.

    Â is bipartite, so one hop is  E'_U = Â_UI·E_I  and  E'_I = Â_IU·E_U.  With users partitioned, the first product
    is local (owned user rows × replicated item table) and the second is a sum over users: every rank multiplies the
    columns of Â_IU that belong to ITS users by its local user rows and the (I, d) partial tables are summed with one
    all-reduce.  Per hop that moves I·d·4 bytes (1.02 GB at the 200M-edge config) instead of the (N, d) all-gather's
    6.1 GB, and the all-reduce overlaps the user-row SpMM of the same hop (it only feeds the next hop).
    """

    def __init__(self, rowptr: torch.Tensor, n_users: int, n_items: int, world_size: int):
        self.n_users, self.n_items, self.world_size = n_users, n_items, world_size
        self.starts = nnz_balanced_starts(rowptr[:n_users + 1], world_size)

    def users(self, rank: int) -> Tuple[int, int]:
        return self.starts[rank], self.starts[rank + 1]

    def user_block(self, rank: int, rowptr, col, val):
        """CSR of the owned user rows; columns = item ids (index into the replicated item table)."""
        u0, u1 = self.users(rank)
        lo, hi = int(rowptr[u0]), int(rowptr[u1])
        rp = (rowptr[u0:u1 + 1] - rowptr[u0]).to(torch.int32).contiguous()
        return rp, (col[lo:hi] - self.n_users).to(torch.int32).contiguous(), val[lo:hi].contiguous()

    def item_block(self, rank: int, rowptr, col, val):
        """CSR over ALL item rows restricted to the owned users' columns; columns = local user index."""
        u0, u1 = self.users(rank)
        nu, ni = self.n_users, self.n_items
        lo = int(rowptr[nu])
        icol, ival = col[lo:], val[lo:]
        keep = (icol >= u0) & (icol < u1)
        counts = (rowptr[nu + 1:] - rowptr[nu:-1]).to(torch.int64)
        rows = torch.repeat_interleave(torch.arange(ni, device=col.device), counts)[keep]
        rp = torch.zeros(ni + 1, dtype=torch.int64, device=col.device)
        rp[1:] = torch.cumsum(torch.bincount(rows, minlength=ni), 0)
        return rp.to(torch.int32).contiguous(), (icol[keep] - u0).to(torch.int32).contiguous(), ival[keep].contiguous()


class BipartitePropagator:
    """K-layer propagation with users partitioned and the item table all-reduced once per hop."""

    def __init__(self, part: BipartitePartition, rank: int, user_graph, item_graph, d: int, n_layers: int, device,
                 group=None, spmm_fn: Callable = _default_spmm, mean_fn: Callable = _default_mean, item_chunks=None):
        """``item_graph``: handle over all item rows restricted to the owned users.  ``item_chunks``: optional list of
        (row_begin, row_end, handle) covering the item rows in order; each chunk's partial rows are all-reduced as soon as
        that chunk's SpMM is enqueued, so the collective starts earlier and its exposed tail shrinks."""
        self.part, self.rank, self.ug, self.ig, self.d, self.n_layers = part, rank, user_graph, item_graph, d, n_layers
        self.item_chunks = item_chunks or [(0, part.n_items, item_graph)]
        self.group, self.spmm_fn, self.mean_fn = group, spmm_fn, mean_fn
        self.tail_fn = None  # optional hook (works, (addends, divisor), out_item, launch_last_user_pass) replacing the last layer's tail
        self.comm = None     # optional C-ABI communicator (CabiComm): the per-hop all-reduce through tgcn_allreduce_sum_f32
        u0, u1 = part.users(rank)
        self.n_local = u1 - u0
        self.ubufs = [torch.empty((self.n_local, d), dtype=torch.float32, device=device) for _ in range(max(n_layers - 1, 0))]
        self.ibufs = [torch.empty((part.n_items, d), dtype=torch.float32, device=device) for _ in range(n_layers)]
        self.comm_bytes_per_hop = part.n_items * d * 4  # all-reduce payload per rank
        self._probe = os.environ.get("TGCN_MG_PROBE", "")  # "nocomm" / "nocompute": timing probes for the overlap analysis, read once

    def propagate(self, e0_user_local: torch.Tensor, e0_item: torch.Tensor, out_user_local: torch.Tensor,
                  out_item: torch.Tensor, single: bool = False):
        """e0_user_local: owned rows of the user table; e0_item: the whole item table (identical on every rank).
        Writes this rank's rows of users_emb and the whole items_emb."""
        L = self.n_layers
        ws = self.part.world_size
        probe = self._probe
        spmm = (lambda *a: None) if probe == "nocompute" else self.spmm_fn

        def item_partials(src_u, dst):
            """B: partial item rows over the owned users, chunk by chunk, each chunk's all-reduce launched right behind it."""
            works = []
            for r0, r1, handle in self.item_chunks:
                spmm(handle, src_u, dst[r0:r1], [], 1.0)
                if ws > 1 and probe != "nocomm":
                    if self.comm is not None:
                        works.append(self.comm.all_reduce_async(dst[r0:r1]))
                    else:
                        works.append(dist.all_reduce(dst[r0:r1], group=self.group, async_op=True))
            return works

        # Software pipeline: the all-reduce of layer l's item table (AR_l) only feeds the USER rows of layer l+1, so it
        # runs behind two local SpMMs: A_l (user rows of layer l, needs AR_{l-1}) and B_{l+1} (item partials of layer
        # l+1, needs A_l's output).  Compute per hop = A + B back to back; the collective is hidden behind it.
        works = item_partials(e0_user_local, self.ibufs[0])                  # B_1 (+ AR_1)
        cur_i = e0_item
        for layer in range(1, L + 1):
            last = layer == L
            nxt = []
            if last and self.tail_fn is not None:
                # the item table's layer mean only needs AR_L, not the last user pass A_L: the owner's hook runs the mean on a
                # side stream (after AR_L) and A_L on this one, so that the two overlap (GridPropagator._tail)
                adds = [] if single else [e0_user_local] + self.ubufs
                self.tail_fn(works, ([self.ibufs[L - 1]], 1.0) if single else ([e0_item] + self.ibufs, float(L + 1)), out_item,
                             lambda: spmm(self.ug, cur_i, out_user_local, adds, 1.0 if single else float(L + 1)))
                return out_user_local, out_item
            if last:                                                        # A_l: user rows of layer l (local)
                adds = [] if single else [e0_user_local] + self.ubufs
                spmm(self.ug, cur_i, out_user_local, adds, 1.0 if single else float(L + 1))
            else:
                spmm(self.ug, cur_i, self.ubufs[layer - 1], [], 1.0)
                nxt = item_partials(self.ubufs[layer - 1], self.ibufs[layer])   # B_{l+1}; AR_{l+1} queues behind AR_l
            for wk in works:
                wk.wait()                                                   # item table of layer l is complete
            works = nxt
            cur_i = self.ibufs[layer - 1]
        if single:
            self.mean_fn([cur_i], out_item, 1.0)
        else:
            self.mean_fn([e0_item] + self.ibufs, out_item, float(L + 1))
        return out_user_local, out_item


class _ResultTables:
    """The row-sharded result of the sliced / grid schemes: (per, d) user rows + (n_items, d) item table per rank.
    ``p2p``: the tables live in CUDA-IPC peer memory and ``peer_u`` / ``peer_i`` hold every rank's mapping of every
    table, so SpMM epilogues can store into them over NVLink; otherwise plain torch tensors."""

    def __init__(self, per: int, n_items: int, d: int, world_size: int, rank: int, device, group, p2p: bool):
        self.world_size, self.group, self.p2p = world_size, group, p2p
        self._flag = torch.zeros(1, dtype=torch.float32, device=device)
        self._peer_flags, self._epoch, self._rank = None, 0, rank
        rows = max(per, 1)
        if p2p:
            from . import ops
            self._ops = ops
            self.buf_u = ops.PeerBuffer(rows * d * 4, device)
            self.buf_i = ops.PeerBuffer(n_items * d * 4, device)
            self.buf_f = ops.PeerBuffer(256, device)   # barrier flags: one int64 epoch slot per rank
            self.buf_f.tensor((64,)).zero_()
            torch.cuda.synchronize(device)
            handles = [None] * world_size
            if world_size > 1:
                dist.all_gather_object(handles, (self.buf_u.handle, self.buf_i.handle, self.buf_f.handle), group=group)
            self.peer_u = [self.buf_u.ptr if q == rank else self.buf_u.open_peer(handles[q][0]) for q in range(world_size)]
            self.peer_i = [self.buf_i.ptr if q == rank else self.buf_i.open_peer(handles[q][1]) for q in range(world_size)]
            # one process per GPU (NCCL): flag handshake in peer memory.  Processes SHARING a device (the gloo test) would
            # spin against each other's time slices: they keep the collective barrier.
            if world_size > 1 and dist.get_backend(group) == "nccl" and os.environ.get("TGCN_PEER_BARRIER", "1") != "0":
                self._peer_flags = [self.buf_f.ptr if q == rank else self.buf_f.open_peer(handles[q][2]) for q in range(world_size)]
            self.out_u = self.buf_u.tensor((rows, d))
            self.out_i = self.buf_i.tensor((n_items, d))
        else:
            self.out_u = torch.empty((rows, d), dtype=torch.float32, device=device)
            self.out_i = torch.empty((n_items, d), dtype=torch.float32, device=device)

    def barrier(self) -> None:
        """Stream-ordered barrier: returns (on the stream) once every rank's earlier work on its stream is complete.
        p2p tables: a flag handshake in peer memory (one tiny kernel: every rank stores the barrier's epoch into its slot of
        every peer's flag array and spins on its own array, tgcn_peer_barrier); otherwise a 1-float all-reduce."""
        if self.world_size <= 1:
            return
        if self.p2p and self._peer_flags is not None:
            self._epoch += 1
            self._ops.peer_barrier(self._peer_flags, self._rank, self._epoch)
        else:
            dist.all_reduce(self._flag, group=self.group)

    def close(self) -> None:
        if self.p2p:
            self.buf_u.close()
            self.buf_i.close()
            self.buf_f.close()


# ---------------------------------------------------------------------------------------------------------
# Feature-sliced scheme: every GPU runs all L hops on d/P columns of the tables; ONE exchange, fused into the last pass
# ---------------------------------------------------------------------------------------------------------
class FeatureSlicePartition:
    """E' = Â·E acts on each embedding column independently, so the hops need no exchange at all when the FEATURE
    dimension is what is split: rank p owns columns [p·d/P, (p+1)·d/P) of both tables for every layer and a replica
    of the CSR of Â (3.2 GB at the 200M-edge config; the tables, which dominate memory, shrink by P).  Only the
    RESULT has to change layout — users_emb row-sharded by user range, items_emb replicated, both full width, which is
    what eval sharding consumes — and that single exchange rides in the last SpMM's epilogue as peer-memory stores."""

    def __init__(self, n_users: int, n_items: int, d: int, world_size: int):
        if d % (4 * world_size) != 0:
            raise ValueError(f"embedding width {d} must be a multiple of 4·world_size = {4 * world_size}")
        self.n_users, self.n_items, self.d, self.world_size = n_users, n_items, d, world_size
        self.ds = d // world_size
        self.per = -(-n_users // world_size)  # users per rank (the last range may be short)

    def cols(self, rank: int) -> Tuple[int, int]:
        return rank * self.ds, (rank + 1) * self.ds

    def users(self, rank: int) -> Tuple[int, int]:
        return min(rank * self.per, self.n_users), min((rank + 1) * self.per, self.n_users)

    def slice_tables(self, rank: int, user_w: torch.Tensor, item_w: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        c0, c1 = self.cols(rank)
        return user_w[:, c0:c1].contiguous(), item_w[:, c0:c1].contiguous()


def _default_local_propagate(graph, user_slice, item_slice, n_layers, single, out):
    from . import ops
    return ops.propagate_fwd(graph, user_slice, item_slice, n_layers, single=single, out=out)


class SlicedPropagator:
    """K-layer propagation with the feature dimension sliced across ranks.

    ``exchange="p2p"`` (default, CUDA only): the last pass stores the layer mean directly into the peers' result
    tables (CUDA IPC peer memory, NVLink), bracketed by two stream-ordered barriers.  ``exchange="collective"``: the
    unfused comparison — local (N, d/P) result, then all-to-all (user rows) + all-gather (item rows) + an interleave
    copy; this is also the path the world_size-2 gloo test drives on CPU with the kernel stubbed."""

    def __init__(self, part: FeatureSlicePartition, rank: int, graph, n_layers: int, device, group=None,
                 exchange: str = "p2p", local_fn: Callable = _default_local_propagate):
        self.part, self.rank, self.graph, self.n_layers, self.group = part, rank, graph, n_layers, group
        self.exchange, self.local_fn = exchange, local_fn
        if exchange not in ("p2p", "collective"):
            raise ValueError(f"unknown exchange {exchange!r}")
        P, d, ds = part.world_size, part.d, part.ds
        u0, u1 = part.users(rank)
        self.n_local = u1 - u0
        self.tables = _ResultTables(part.per, part.n_items, d, P, rank, device, group, p2p=exchange == "p2p")
        self.out_u, self.out_i = self.tables.out_u, self.tables.out_i
        if exchange == "p2p":
            from . import ops
            self._ops = ops
        else:
            self.local = torch.empty((part.n_users + part.n_items, ds), dtype=torch.float32, device=device)
        # sent per rank: its column slice of every user row to the owner and of every item row to all peers
        self.comm_bytes_per_step = (part.n_users - self.n_local) * ds * 4 + (P - 1) * part.n_items * ds * 4

    def propagate(self, user_slice: torch.Tensor, item_slice: torch.Tensor, single: bool = False):
        """user_slice (n_users, d/P), item_slice (n_items, d/P): this rank's columns of E0.  Returns (users_emb rows of
        this rank's user range (n_local, d), items_emb (n_items, d)); the tensors are reused by the next call."""
        part, P = self.part, self.part.world_size
        nu, ni, ds = part.n_users, part.n_items, part.ds
        if self.exchange == "p2p":
            self.tables.barrier()  # every rank is done reading the previous result tables
            self._ops.propagate_sliced(self.graph, user_slice, item_slice, self.n_layers, part.d, part.cols(self.rank)[0],
                                       part.per, self.tables.peer_u, self.tables.peer_i, single=single)
            self.tables.barrier()  # every rank's stores have landed
        else:
            loc = self.local_fn(self.graph, user_slice, item_slice, self.n_layers, single, self.local)
            n_my = self.n_local
            if P > 1:
                sizes_in = [part.users(q)[1] - part.users(q)[0] for q in range(P)]
                recv = torch.empty((P * n_my, ds), dtype=loc.dtype, device=loc.device)
                dist.all_to_all_single(recv, loc[:nu], output_split_sizes=[n_my] * P, input_split_sizes=sizes_in, group=self.group)
                gath = torch.empty((P * ni, ds), dtype=loc.dtype, device=loc.device)
                dist.all_gather_into_tensor(gath, loc[nu:].contiguous(), group=self.group)
            else:
                recv, gath = loc[:nu], loc[nu:]
            self.out_u[:n_my].view(n_my, P, ds).copy_(recv.view(P, n_my, ds).transpose(0, 1))
            self.out_i.view(ni, P, ds).copy_(gath.view(P, ni, ds).transpose(0, 1))
        return self.out_u[:self.n_local], self.out_i

    def close(self) -> None:
        self.tables.close()


# ---------------------------------------------------------------------------------------------------------
# Grid scheme: G feature slices x R user partitions (P = G·R); all-reduce only inside a row group of R ranks
# ---------------------------------------------------------------------------------------------------------
class GridPartition:
    """rank = g·R + r.  Feature slice g owns columns [g·d/G, (g+1)·d/G) of every table; inside a slice the R ranks of a
    ROW GROUP split the users by nnz (BipartitePartition) and all-reduce the slice of the item table once per hop —
    (I, d/G) floats among R ranks instead of (I, d) among P.  G = P is the pure feature-sliced scheme (no exchange per
    hop, but d/P-wide rows: 64-byte gathers run HBM at ~60 % of its streaming rate), G = 1 the bipartite scheme (full
    rows, the largest all-reduce); in between both costs shrink.  The RESULT layout is the same for every G x R:
    users_emb rows [rank·per, (rank+1)·per) full width on each rank, items_emb replicated."""

    def __init__(self, rowptr: torch.Tensor, n_users: int, n_items: int, d: int, n_slices: int, n_row_parts: int):
        if d % (4 * n_slices) != 0:
            raise ValueError(f"embedding width {d} must be a multiple of 4·G = {4 * n_slices}")
        self.G, self.R = n_slices, n_row_parts
        self.world_size = n_slices * n_row_parts
        self.n_users, self.n_items, self.d, self.ds = n_users, n_items, d, d // n_slices
        self.rows = BipartitePartition(rowptr, n_users, n_items, n_row_parts)
        # result layout: the users of row partition r stay inside their row group — rank (g, r) ends up with the g-th of G
        # equal sub-ranges of partition r, full width.  Only the G feature-slice partners of a row group exchange user rows
        # (nothing moves at G = 1); round 1 re-cut the users into P equal ranges, which sent ~(P-1)/P of every rank's rows
        # across NVLink from inside the last SpMM pass.
        self.sub_per = [-(-(self.rows.users(r)[1] - self.rows.users(r)[0]) // n_slices) for r in range(n_row_parts)]
        self.per = max(max(self.sub_per), 1)  # rows of the largest result shard (table allocation)

    def coords(self, rank: int) -> Tuple[int, int]:
        return rank // self.R, rank % self.R

    def cols(self, g: int) -> Tuple[int, int]:
        return g * self.ds, (g + 1) * self.ds

    def final_users(self, rank: int) -> Tuple[int, int]:
        g, r = self.coords(rank)
        u0, u1 = self.rows.users(r)
        return min(u0 + g * self.sub_per[r], u1), min(u0 + (g + 1) * self.sub_per[r], u1)

    def slice_partners(self, r: int) -> List[int]:
        """The G ranks that hold the G feature slices of row partition r, in slice order."""
        return [g * self.R + r for g in range(self.G)]

    def row_group_ranks(self, g: int) -> List[int]:
        return [g * self.R + r for r in range(self.R)]


class GridPropagator:
    """K-layer propagation on a G x R grid.  The hops are BipartitePropagator's (on d/G-wide tables, all-reduce in
    ``row_group``); the change of layout rides in the last passes: ``exchange="p2p"`` — the last user-row SpMM stores
    each row into its final owner's table and the item table's layer mean is stored into every rank's replica (peer
    memory, NVLink), each row-group member broadcasting 1/R of the rows; ``exchange="collective"`` — the same through
    all-to-all / all-gather on ``group`` (the comparison arm, and what the gloo test drives on CPU)."""

    def __init__(self, part: GridPartition, rank: int, user_graph, item_graph, n_layers: int, device, row_group=None,
                 group=None, exchange: str = "p2p", spmm_fn: Callable = _default_spmm, mean_fn: Callable = _default_mean,
                 comm=None):
        self.part, self.rank, self.group, self.exchange = part, rank, group, exchange
        self.g, self.r = part.coords(rank)
        self._spmm_fn, self._mean_fn = spmm_fn, mean_fn
        self.inner = BipartitePropagator(part.rows, self.r, user_graph, item_graph, part.ds, n_layers, device, group=row_group,
                                         spmm_fn=self._spmm, mean_fn=self._mean)
        self.inner.comm = comm
        self._side = None
        if exchange == "p2p" and torch.device(device).type == "cuda" and os.environ.get("TGCN_GRID_SIDE_STREAM", "1") != "0":
            # the item table's layer mean + broadcast runs beside the last user pass instead of behind it
            self._side = torch.cuda.Stream(device=device)
            self._ev_main = torch.cuda.Event()
            self._ev_side = torch.cuda.Event()
            self.inner.tail_fn = self._tail
        self.ug = user_graph
        P, d = part.world_size, part.d
        f0, f1 = part.final_users(rank)
        self.n_final = f1 - f0
        u0, u1 = part.rows.users(self.r)
        self.n_local = u1 - u0
        if exchange not in ("p2p", "collective"):
            raise ValueError(f"unknown exchange {exchange!r}")
        self._events = [] if os.environ.get("TGCN_GRID_TIMING") else None
        self._tok_u = torch.empty(0, device=device)  # stand-ins for "the result tables" handed to the inner propagator
        self._tok_i = torch.empty(0, device=device)
        self.tables = _ResultTables(part.per, part.n_items, d, P, rank, device, group, p2p=exchange == "p2p")
        self.out_u, self.out_i = self.tables.out_u, self.tables.out_i
        if exchange == "p2p":
            from . import ops
            self._ops = ops
        else:
            self.loc_u = torch.empty((self.n_local, part.ds), dtype=torch.float32, device=device)
            self.loc_i = torch.empty((part.n_items, part.ds), dtype=torch.float32, device=device)
        self.comm_bytes_per_hop = part.n_items * part.ds * 4 if part.R > 1 else 0

    # -- hooks handed to the inner propagator: the two result tables are written by the exchange ----------------
    def _timed(self, label, fn):
        """TGCN_GRID_TIMING=1: CUDA events around every kernel call of a step (read back with timing_report)."""
        if self._events is None:
            return fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn()
        e1.record()
        self._events.append((label, e0, e1))
        return out

    def timing_report(self):
        torch.cuda.synchronize()
        rep = [(label, round(e0.elapsed_time(e1), 3)) for label, e0, e1 in (self._events or [])]
        self._events = [] if self._events is not None else None
        return rep

    def _spmm(self, graph, x, y, addends, divisor):
        label = ("A" if graph is self.ug else "B") + ("_scatter" if y is self._tok_u else "")
        return self._timed(label, lambda: self._spmm_impl(graph, x, y, addends, divisor))

    def _mean(self, addends, out, divisor):
        return self._timed("mean" + ("_scatter" if out is self._tok_i else ""), lambda: self._mean_impl(addends, out, divisor))

    def _spmm_impl(self, graph, x, y, addends, divisor):
        if y is not self._tok_u:
            return self._spmm_fn(graph, x, y, addends, divisor)
        if self.exchange == "p2p":
            part = self.part
            partners = part.slice_partners(self.r)   # user row (u0 + j) goes to partner j // sub_per, local row j % sub_per
            return self._ops.spmm_scatter(graph, x, addends, divisor, part.d, part.cols(self.g)[0], part.sub_per[self.r],
                                          [self.tables.peer_u[q] for q in partners], [self.tables.peer_i[q] for q in partners],
                                          user_row0=part.rows.users(self.r)[0])
        return self._spmm_fn(graph, x, self.loc_u, addends, divisor)

    def _mean_impl(self, addends, out, divisor):
        if out is not self._tok_i:
            return self._mean_fn(addends, out, divisor)
        if self.exchange == "p2p":
            part = self.part
            i0, i1 = item_shard(part.n_items, part.R, self.r)  # row-group members hold identical sums: each ships 1/R
            if i1 > i0:
                self._ops.layer_mean_scatter([a[i0:i1] for a in addends], divisor, part.d, part.cols(self.g)[0], i0, self.tables.peer_i)
            return None
        return self._mean_fn(addends, self.loc_i, divisor)

    def _tail(self, works, mean_args, out_item, launch_last_user_pass):
        """Last layer: the item table's layer mean (+ its broadcast into every rank's replica) needs the final all-reduce AR_L
        but not the last user pass A_L, and A_L does not need AR_L: the mean runs on the side stream, A_L on the main one."""
        main = torch.cuda.current_stream()
        self._ev_main.record(main)          # everything the mean reads except AR_L was produced before this point
        with torch.cuda.stream(self._side):
            self._side.wait_event(self._ev_main)
            for wk in works:
                wk.wait()                   # the side stream (not the main one) waits for the collective
            self._mean(mean_args[0], out_item, mean_args[1])
            self._ev_side.record(self._side)
        launch_last_user_pass()             # overlaps AR_L's tail and the mean / broadcast
        main.wait_event(self._ev_side)      # joins before the closing barrier

    def propagate(self, e0_user_slice_local: torch.Tensor, e0_item_slice: torch.Tensor, single: bool = False):
        """e0_user_slice_local: rows of this rank's row-group users, columns of its slice (n_local, d/G); e0_item_slice:
        (n_items, d/G).  Returns (users_emb rows [rank·per, ...) (n_final, d), items_emb (n_items, d))."""
        part, P = self.part, self.part.world_size
        if self.exchange == "p2p":
            self._timed("barrier", self.tables.barrier)  # every rank is done reading the previous result tables
            self.inner.propagate(e0_user_slice_local, e0_item_slice, self._tok_u, self._tok_i, single=single)
            self._timed("barrier", self.tables.barrier)  # every rank's stores have landed
            return self.out_u[:self.n_final], self.out_i
        self.inner.propagate(e0_user_slice_local, e0_item_slice, self._tok_u, self._tok_i, single=single)
        ds, ni = part.ds, part.n_items
        me0, me1 = part.final_users(self.rank)

        def overlap(a0, a1, b0, b1):
            lo, hi = max(a0, b0), min(a1, b1)
            return (lo, hi) if hi > lo else (lo, lo)

        u0, u1 = part.rows.users(self.r)
        send = [overlap(u0, u1, *part.final_users(q)) for q in range(P)]
        recv = [overlap(*part.rows.users(q % part.R), me0, me1) for q in range(P)]
        if P > 1:
            buf = torch.empty((sum(hi - lo for lo, hi in recv), ds), dtype=self.loc_u.dtype, device=self.loc_u.device)
            dist.all_to_all_single(buf, self.loc_u, output_split_sizes=[hi - lo for lo, hi in recv],
                                   input_split_sizes=[hi - lo for lo, hi in send], group=self.group)
            per_i = -(-ni // part.R)
            i0, i1 = item_shard(ni, part.R, self.r)
            mine = torch.zeros((per_i, ds), dtype=self.loc_i.dtype, device=self.loc_i.device)
            mine[:i1 - i0] = self.loc_i[i0:i1]
            gath = torch.empty((P * per_i, ds), dtype=mine.dtype, device=mine.device)
            dist.all_gather_into_tensor(gath, mine, group=self.group)
        else:
            buf, gath, per_i = self.loc_u, self.loc_i, ni
        off = 0
        for q in range(P):
            gq, rq = part.coords(q)
            lo, hi = recv[q]
            if hi > lo:
                self.out_u[lo - me0:hi - me0, gq * ds:(gq + 1) * ds] = buf[off:off + hi - lo]
                off += hi - lo
            j0, j1 = item_shard(ni, part.R, rq)
            if j1 > j0:
                self.out_i[j0:j1, gq * ds:(gq + 1) * ds] = gath[q * per_i:q * per_i + (j1 - j0)]
        return self.out_u[:self.n_final], self.out_i

    def close(self) -> None:
        self.tables.close()


class CabiComm:
    """An NCCL communicator owned through the C ABI (``tgcn_comm_*``): what a consumer of libtgcn_b200 that does not run
    torch.distributed uses for the path's two exchanges — the per-hop all-reduce of the item-table slice and the
    all-to-all of partial top-k tables.  The 128-byte unique id travels through any transport (here: a torch.distributed
    object broadcast inside ``group``).  Collectives are enqueued on this object's own high-priority stream and ordered
    with CUDA events, like torch's async NCCL work objects (``all_reduce_async(...).wait()``)."""

    class _Work:
        def __init__(self, event):
            self.event = event

        def wait(self):
            torch.cuda.current_stream().wait_event(self.event)

    def __init__(self, ranks: Sequence[int], rank: int, device, group=None):
        from . import ops
        self._ops = ops
        self.ranks, self.rank, self.device = list(ranks), rank, torch.device(device)
        me = self.ranks.index(rank)
        uid = [ops.comm_unique_id() if me == 0 else None]
        dist.broadcast_object_list(uid, src=self.ranks[0], group=group)
        with torch.cuda.device(self.device):
            self.handle = ops.comm_init_rank(len(self.ranks), me, uid[0])
            self.stream = torch.cuda.Stream(device=self.device, priority=-1)

    def all_reduce_async(self, t: torch.Tensor):
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream())
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(ready)
            self._ops.comm_allreduce_sum(self.handle, t, self.stream.cuda_stream)
            done = torch.cuda.Event()
            done.record(self.stream)
        return CabiComm._Work(done)

    def topk_exchange(self, part_ids: torch.Tensor, part_scores: torch.Tensor):
        """All-to-all of the partial (n_ranks·rows, k) top-k tables on the current stream (tgcn_topk_exchange)."""
        recv_ids, recv_sc = torch.empty_like(part_ids), torch.empty_like(part_scores)
        self._ops.comm_topk_exchange(self.handle, len(self.ranks), part_ids, part_scores, recv_ids, recv_sc,
                                     torch.cuda.current_stream().cuda_stream)
        return recv_ids, recv_sc

    def close(self):
        if self.handle:
            torch.cuda.synchronize(self.device)
            self._ops.comm_destroy(self.handle)
            self.handle = None


def item_shard(n_items: int, world_size: int, rank: int) -> Tuple[int, int]:
    per = (n_items + world_size - 1) // world_size
    return min(rank * per, n_items), min((rank + 1) * per, n_items)


def _default_rank(mask_graph, user_vecs, item_vecs, k, users, item_range, by_position=False):
    from . import ops
    return ops.eval_topk(mask_graph, user_vecs, item_vecs, k, users=users, item_range=item_range, finalize=False,
                         by_position=by_position)


def _default_merge(mask_graph, part_ids, part_scores, users):
    from . import ops
    return ops.topk_merge(mask_graph, part_ids, part_scores, users=users, finalize=True)


def sharded_eval_topk(mask_graph, user_vecs: torch.Tensor, item_vecs: torch.Tensor, users: torch.Tensor, k: int,
                      rank: int, world_size: int, group=None, gather: bool = False, by_position: bool = False,
                      rank_fn: Callable = _default_rank, merge_fn: Callable = _default_merge, comm: Optional["CabiComm"] = None):
    """Item-sharded full ranking with a cross-GPU top-k merge.

    ``users`` (int32 ids, identical on every rank; its length must be a multiple of world_size) are ranked by every
    rank against its own item range; partial (U, k) tables travel by all-to-all; rank p merges user slice p.
    Returns (ids, scores) for this rank's user slice, or for all users when ``gather``.
    """
    n_users_ranked = users.numel()
    assert n_users_ranked % world_size == 0, "pad the user list to a multiple of world_size"
    i0, i1 = item_shard(item_vecs.shape[0], world_size, rank)
    if by_position:  # user_vecs row m belongs to users[m] (a gathered sample) instead of being indexed by user id
        part_ids, part_sc = rank_fn(mask_graph, user_vecs, item_vecs, k, users, (i0, i1), by_position=True)
    else:
        part_ids, part_sc = rank_fn(mask_graph, user_vecs, item_vecs, k, users, (i0, i1))
    per = n_users_ranked // world_size
    recv_ids = torch.empty_like(part_ids)
    recv_sc = torch.empty_like(part_sc)
    if world_size > 1 and comm is not None:
        recv_ids, recv_sc = comm.topk_exchange(part_ids.contiguous(), part_sc.contiguous())
    elif world_size > 1:
        dist.all_to_all_single(recv_ids, part_ids.contiguous(), group=group)
        dist.all_to_all_single(recv_sc, part_sc.contiguous(), group=group)
    else:
        recv_ids, recv_sc = part_ids, part_sc
    my_users = users[rank * per:(rank + 1) * per].contiguous()
    ids, sc = merge_fn(mask_graph, recv_ids.view(world_size, per, k), recv_sc.view(world_size, per, k), my_users)
    if gather and world_size > 1:
        all_ids = torch.empty((n_users_ranked, k), dtype=ids.dtype, device=ids.device)
        all_sc = torch.empty((n_users_ranked, k), dtype=sc.dtype, device=sc.device)
        dist.all_gather_into_tensor(all_ids, ids.contiguous(), group=group)
        dist.all_gather_into_tensor(all_sc, sc.contiguous(), group=group)
        return all_ids, all_sc
    return ids, sc
