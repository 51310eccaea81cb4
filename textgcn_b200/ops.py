"""Torch-tensor front end of the C ABI: validates dtype / layout / device on the Python side (the
library validates sizes), hands raw pointers and the CURRENT torch stream to libtgcn_b200, and owns the
workspaces.  PyTorch here is plumbing (device memory, streams, autograd glue), not the compute path.
"""
from __future__ import annotations

import ctypes
from typing import Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import check


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _chk(t: torch.Tensor, dtype, name: str, dims: Optional[int] = None, align: int = 16) -> torch.Tensor:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise _lib.TgcnError(f"{name} must be a CUDA tensor (textgcn_b200 has no CPU path)")
    if t.dtype != dtype:
        raise _lib.TgcnError(f"{name} must be {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise _lib.TgcnError(f"{name} must be contiguous")
    if dims is not None and t.dim() != dims:
        raise _lib.TgcnError(f"{name} must have {dims} dims, got {t.dim()}")
    if t.data_ptr() % align != 0:
        raise _lib.TgcnError(f"{name} must be {align}-byte aligned")
    return t


def as_index(x, device) -> torch.Tensor:
    """int32 contiguous device copy of an index list / tensor."""
    if isinstance(x, torch.Tensor):
        return x.to(device=device, dtype=torch.int32).contiguous()
    return torch.as_tensor(np.asarray(x), device=device).to(torch.int32).contiguous()


class Graph:
    """Device CSR of Â plus the library handle (segment list, transpose permutation)."""

    def __init__(self, n_users: int, n_items: int, rowptr: torch.Tensor, col: torch.Tensor, val: torch.Tensor,
                 row_begin: int = 0, block: bool = False):
        self.lib = _lib.load()
        self.n_users, self.n_items = int(n_users), int(n_items)
        self.rowptr = _chk(rowptr, torch.int32, "rowptr", 1, align=4)  # CSR arrays are read with scalar loads
        self.col = _chk(col, torch.int32, "col", 1, align=4)
        self.val = _chk(val, torch.float32, "val", 1, align=4)
        self.nnz = int(col.numel())
        self.block = bool(block)
        self.row_begin = int(row_begin)
        self.n_rows = int(rowptr.numel()) - 1
        self.device = rowptr.device
        handle = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            if block:
                check(self.lib.tgcn_graph_create_block(ctypes.byref(handle), self.n_users, self.n_items, self.row_begin,
                                                       self.n_rows, self.nnz, _ptr(self.rowptr), _ptr(self.col),
                                                       _ptr(self.val), _stream()))
            else:
                if self.n_rows != self.n_users + self.n_items:
                    raise _lib.TgcnError("rowptr must have n_users + n_items + 1 entries")
                check(self.lib.tgcn_graph_create(ctypes.byref(handle), self.n_users, self.n_items, self.nnz,
                                                 _ptr(self.rowptr), _ptr(self.col), _ptr(self.val), _stream()))
        self.handle = handle
        self._ws = {}

    def __del__(self):
        h = getattr(self, "handle", None)
        if h:
            try:
                self.lib.tgcn_graph_destroy(h)
            except Exception:
                pass
            self.handle = None

    @property
    def n_nodes(self) -> int:
        return self.n_users + self.n_items

    @property
    def n_segments(self) -> int:
        return int(self.lib.tgcn_graph_num_segments(self.handle))

    @classmethod
    def from_norm_matrix(cls, norm_matrix: torch.Tensor, n_users: int, n_items: int, device=None) -> "Graph":
        """From the reference's ``dataset.norm_matrix`` (coalesced COO sorted by (row, col), int64 indices,
        fp32 values: dataset.py:138, :151-157).  COO order is CSR order, so only rowptr is computed."""
        device = torch.device(device) if device is not None else norm_matrix.device
        if device.type != "cuda":
            raise _lib.TgcnError("Graph needs a CUDA device (textgcn_b200 has no CPU path)")
        nm = norm_matrix.coalesce()
        idx = nm.indices()
        n = n_users + n_items
        counts = torch.bincount(idx[0], minlength=n)
        rowptr = torch.zeros(n + 1, dtype=torch.int64, device=idx.device)
        rowptr[1:] = torch.cumsum(counts, 0)
        return cls(n_users, n_items, rowptr.to(device=device, dtype=torch.int32).contiguous(),
                   idx[1].to(device=device, dtype=torch.int32).contiguous(),
                   nm.values().to(device=device, dtype=torch.float32).contiguous())

    def workspace(self, d: int, n_layers: int) -> torch.Tensor:
        """Caller-owned scratch for propagate / spmm (layer buffers + long-row partial sums); grown on demand."""
        nbytes = max(int(self.lib.tgcn_propagate_workspace_bytes(self.handle, d, n_layers)), 256)
        ws = self._ws.get("buf")
        if ws is None or ws.numel() < nbytes:
            if self._ws.get("pinned"):
                raise _lib.TgcnError("the workspace of this graph handle is referenced by a captured CUDA graph and cannot grow "
                                     f"(needs {nbytes} bytes for d={d}, L={n_layers}); use a second Graph handle for the wider call")
            ws = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
            self._ws["buf"] = ws
        return ws

    def pin_workspace(self) -> None:
        """Called after a CUDA graph has captured launches on this handle: the captured kernels hold the workspace's address,
        so it must neither move nor be freed; later calls that would need a larger one raise instead of reallocating.
        (One workspace per handle also means one stream per handle at a time — the library refuses a concurrent launch.)"""
        self._ws["pinned"] = True

    def set_mask_col_offset(self, off: int) -> None:
        check(self.lib.tgcn_graph_set_mask_col_offset(self.handle, int(off)))

    def build_transpose_perm(self) -> None:
        with torch.cuda.device(self.device):
            check(self.lib.tgcn_graph_build_transpose_perm(self.handle, _stream()))


def spmm(g: Graph, x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """layer_aggregation: Y = Â·X."""
    _chk(x, torch.float32, "x", 2)
    d = x.shape[1]
    out = torch.empty((g.n_rows, d), dtype=torch.float32, device=x.device) if out is None else _chk(out, torch.float32, "out", 2)
    ws = g.workspace(d, 1)
    with torch.cuda.device(x.device):
        check(g.lib.tgcn_spmm_fwd(g.handle, d, _ptr(x), _ptr(out), _ptr(ws), ws.numel(), _stream()))
    return out


def spmm_ex(g: Graph, x: torch.Tensor, y: torch.Tensor, addends: Sequence[torch.Tensor] = (), divisor: float = 1.0,
            keep: Optional[torch.Tensor] = None, dropout: float = 0.0, transposed: bool = False,
            accumulate: bool = False) -> torch.Tensor:
    """One SpMM pass with the fused epilogue  y = (Σ addends + Â·x) / divisor  (contiguous operands)."""
    _chk(x, torch.float32, "x", 2)
    _chk(y, torch.float32, "y", 2)
    d = x.shape[1]
    n = len(addends)
    arr = (ctypes.c_void_p * max(n, 1))(*[_chk(a, torch.float32, "addend", 2).data_ptr() for a in addends])
    nul = (ctypes.c_void_p * max(n, 1))()
    if keep is not None:
        keep = _keep_u8(keep)
    ws = g.workspace(d, 1)
    with torch.cuda.device(x.device):
        check(g.lib.tgcn_spmm_ex(g.handle, d, _ptr(x), None, _ptr(keep), float(dropout), int(transposed), n, arr, nul,
                                 float(divisor), int(accumulate), _ptr(y), _ptr(ws), ws.numel(), _stream()))
    return y


def layer_mean(addends: Sequence[torch.Tensor], out: torch.Tensor, divisor: Optional[float] = None) -> torch.Tensor:
    """out = (Σ addends) / divisor (default: their count) — layer_combination for a table no local SpMM produces."""
    lib = _lib.load()
    n = len(addends)
    arr = (ctypes.c_void_p * n)(*[_chk(a, torch.float32, "addend").data_ptr() for a in addends])
    _chk(out, torch.float32, "out")
    with torch.cuda.device(out.device):
        check(lib.tgcn_layer_mean(out.numel(), n, arr, float(n if divisor is None else divisor), _ptr(out), _stream()))
    return out


def dropout_mask(nnz: int, dropout: float, seed: int, device) -> torch.Tensor:
    """Bernoulli(1 - dropout) keep mask over nnz entries, drawn by one kernel on the device (uint8, 1 = keep)."""
    lib = _lib.load()
    keep = torch.empty(nnz, dtype=torch.uint8, device=device)
    with torch.cuda.device(device):
        check(lib.tgcn_dropout_mask(nnz, float(dropout), seed & (2 ** 64 - 1), _ptr(keep), _stream()))
    return keep


def dropout_mask_dev(nnz: int, dropout: float, base_seed: int, draws: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Same draw with the per-draw part of the seed read from the device counter ``draws`` (int64, 1 element), which is
    incremented first: no per-step host scalar, so the call can be replayed from a CUDA graph."""
    lib = _lib.load()
    if draws.dtype != torch.int64 or not draws.is_cuda or draws.numel() != 1:
        raise _lib.TgcnError("draws must be a 1-element int64 CUDA tensor")
    keep = torch.empty(nnz, dtype=torch.uint8, device=draws.device) if out is None else out
    with torch.cuda.device(draws.device):
        check(lib.tgcn_counter_inc(_ptr(draws), _stream()))
        check(lib.tgcn_dropout_mask_dev(nnz, float(dropout), base_seed & (2 ** 64 - 1), _ptr(draws), _ptr(keep), _stream()))
    return keep


def _keep_u8(keep: torch.Tensor) -> torch.Tensor:
    if keep.dtype == torch.bool:
        keep = keep.view(torch.uint8)
    if keep.dtype != torch.uint8 or not keep.is_cuda or not keep.is_contiguous():
        raise _lib.TgcnError("keep mask must be a contiguous CUDA bool/uint8 tensor of nnz entries")
    return keep


def propagate_fwd(g: Graph, user_w: torch.Tensor, item_w: torch.Tensor, n_layers: int, single: bool = False,
                  keep: Optional[torch.Tensor] = None, dropout: float = 0.0,
                  out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """representation: (N, d) result; rows [0, n_users) are users_emb."""
    _chk(user_w, torch.float32, "user_w", 2)
    _chk(item_w, torch.float32, "item_w", 2)
    d = user_w.shape[1]
    if user_w.shape[0] != g.n_users or item_w.shape != (g.n_items, d):
        raise _lib.TgcnError("embedding tables do not match the graph")
    if keep is not None:
        keep = _keep_u8(keep)
        if keep.numel() != g.nnz:
            raise _lib.TgcnError("keep mask must have nnz entries")
    out = torch.empty((g.n_nodes, d), dtype=torch.float32, device=user_w.device) if out is None else _chk(out, torch.float32, "out", 2)
    ws = g.workspace(d, n_layers)
    with torch.cuda.device(user_w.device):
        check(g.lib.tgcn_propagate_fwd(g.handle, d, n_layers, int(single), _ptr(user_w), _ptr(item_w), _ptr(keep),
                                       float(dropout), _ptr(out), _ptr(ws), ws.numel(), _stream()))
    return out


def propagate_host(g: Graph, h_user_w: torch.Tensor, h_item_w: torch.Tensor, h_out: torch.Tensor, n_layers: int,
                   d_stage: torch.Tensor, single: bool = False) -> torch.Tensor:
    """representation with HOST buffers (pinned): H2D of E0, L fused layers, D2H of the (N, d) result, all
    enqueued on the current stream by one C-ABI call.  ``d_stage`` is a (2N, d) device staging buffer."""
    for t, name in ((h_user_w, "h_user_w"), (h_item_w, "h_item_w"), (h_out, "h_out")):
        if t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous():
            raise _lib.TgcnError(f"{name} must be a contiguous fp32 host tensor")
    d = h_user_w.shape[1]
    _chk(d_stage, torch.float32, "d_stage", 2)
    if d_stage.numel() < 2 * g.n_nodes * d:
        raise _lib.TgcnError("d_stage must hold 2·N·d floats")
    ws = g.workspace(d, n_layers)
    with torch.cuda.device(g.device):
        check(g.lib.tgcn_propagate_host(g.handle, d, n_layers, int(single), h_user_w.data_ptr(), h_item_w.data_ptr(),
                                        h_out.data_ptr(), _ptr(d_stage), _ptr(ws), ws.numel(), _stream()))
    return h_out


class PeerBuffer:
    """Device memory other ranks can map (CUDA IPC): the row-sharded result tables of the feature-sliced
    multi-GPU propagation live here so that peers' SpMM epilogues can store into them over NVLink.
    ``handle`` (64 bytes) is what travels to the other ranks; ``tensor(shape)`` views the memory as fp32."""

    def __init__(self, nbytes: int, device):
        self.lib = _lib.load()
        self.device = torch.device(device)
        self.nbytes = int(nbytes)
        ptr = ctypes.c_void_p()
        hbuf = (ctypes.c_uint8 * 64)()
        with torch.cuda.device(self.device):
            check(self.lib.tgcn_peer_alloc(self.nbytes, ctypes.byref(ptr), ctypes.cast(hbuf, ctypes.c_void_p)))
        self.ptr = int(ptr.value)
        self.handle = bytes(hbuf)
        self._opened = []

    def tensor(self, shape) -> torch.Tensor:
        n = int(np.prod(shape))
        if n * 4 > self.nbytes:
            raise _lib.TgcnError("peer buffer too small for the requested view")
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (self.ptr, False), "version": 3}
        return torch.as_tensor(self, device=self.device).view(*shape)

    def open_peer(self, handle: bytes) -> int:
        """Map another rank's buffer into this process; returns the device pointer valid here."""
        ptr = ctypes.c_void_p()
        hbuf = (ctypes.c_uint8 * 64).from_buffer_copy(handle)
        with torch.cuda.device(self.device):
            check(self.lib.tgcn_peer_open(ctypes.cast(hbuf, ctypes.c_void_p), ctypes.byref(ptr)))
        self._opened.append(int(ptr.value))
        return int(ptr.value)

    def close(self) -> None:
        with torch.cuda.device(self.device):
            for p in self._opened:
                self.lib.tgcn_peer_close(p)
            self._opened = []
            if self.ptr:
                self.lib.tgcn_peer_free(self.ptr)
                self.ptr = 0


def peer_barrier(peer_flag_ptrs: Sequence[int], rank: int, epoch: int) -> None:
    """Stream-ordered barrier over peer-mapped flag arrays (tgcn_peer_barrier) on the current stream."""
    lib = _lib.load()
    n = len(peer_flag_ptrs)
    arr = (ctypes.c_void_p * n)(*peer_flag_ptrs)
    check(lib.tgcn_peer_barrier(n, int(rank), arr, int(epoch), _stream()))


def comm_unique_id() -> bytes:
    lib = _lib.load()
    buf = (ctypes.c_uint8 * 128)()
    check(lib.tgcn_comm_unique_id(ctypes.cast(buf, ctypes.c_void_p)))
    return bytes(buf)


def comm_init_rank(n_ranks: int, rank: int, uid: bytes) -> int:
    """ncclCommInitRank on the current device through the C ABI; returns the communicator as an integer handle."""
    lib = _lib.load()
    comm = ctypes.c_void_p()
    buf = (ctypes.c_uint8 * 128).from_buffer_copy(uid)
    check(lib.tgcn_comm_init_rank(ctypes.byref(comm), n_ranks, rank, ctypes.cast(buf, ctypes.c_void_p)))
    return int(comm.value)


def comm_destroy(comm: int) -> None:
    check(_lib.load().tgcn_comm_destroy(comm))


def comm_allreduce_sum(comm: int, t: torch.Tensor, stream: Optional[int] = None) -> None:
    """In-place sum of a contiguous fp32 tensor over the communicator, on ``stream`` (default: the current stream)."""
    _chk(t, torch.float32, "tensor", align=4)
    with torch.cuda.device(t.device):
        check(_lib.load().tgcn_allreduce_sum_f32(comm, t.data_ptr(), t.numel(), _stream() if stream is None else stream))


def comm_allgather(comm: int, send: torch.Tensor, recv: torch.Tensor, stream: Optional[int] = None) -> None:
    _chk(send, torch.float32, "send", align=4)
    _chk(recv, torch.float32, "recv", align=4)
    with torch.cuda.device(send.device):
        check(_lib.load().tgcn_allgather_f32(comm, send.data_ptr(), recv.data_ptr(), send.numel(), _stream() if stream is None else stream))


def comm_topk_exchange(comm: int, n_ranks: int, part_ids, part_scores, recv_ids, recv_scores, stream: Optional[int] = None) -> None:
    """All-to-all of partial top-k tables shaped (n_ranks · rows_per_rank, k) (tgcn_topk_exchange)."""
    for t, dt, n in ((part_ids, torch.int32, "part_ids"), (recv_ids, torch.int32, "recv_ids"), (part_scores, torch.float32, "part_scores"),
                     (recv_scores, torch.float32, "recv_scores")):
        _chk(t, dt, n, 2, align=4)
    rows, k = part_ids.shape
    if rows % n_ranks:
        raise _lib.TgcnError("the partial tables must hold a multiple of n_ranks rows")
    with torch.cuda.device(part_ids.device):
        check(_lib.load().tgcn_topk_exchange(comm, n_ranks, rows // n_ranks, k, part_ids.data_ptr(), part_scores.data_ptr(),
                                             recv_ids.data_ptr(), recv_scores.data_ptr(), _stream() if stream is None else stream))


def spmm_scatter(g: Graph, x: torch.Tensor, addends: Sequence[torch.Tensor], divisor: float, d_full: int, col_off: int,
                 users_per_rank: int, peer_user_ptrs: Sequence[int], peer_item_ptrs: Sequence[int], user_row0: int = 0) -> None:
    """One SpMM pass (row-block or whole-graph handle, contiguous operands) whose result rows are stored into the peers'
    full-width tables: user row u goes to peer (u - user_row0) // users_per_rank, local row (u - user_row0) % users_per_rank
    (see tgcn_spmm_scatter)."""
    _chk(x, torch.float32, "x", 2)
    ds = x.shape[1]
    n = len(addends)
    arr = (ctypes.c_void_p * max(n, 1))(*[_chk(t, torch.float32, "addend", 2).data_ptr() for t in addends])
    nul = (ctypes.c_void_p * max(n, 1))()
    np_ = len(peer_user_ptrs)
    pu = (ctypes.c_void_p * np_)(*peer_user_ptrs)
    pi = (ctypes.c_void_p * np_)(*peer_item_ptrs)
    ws = g.workspace(ds, 1)
    with torch.cuda.device(x.device):
        check(g.lib.tgcn_spmm_scatter(g.handle, ds, _ptr(x), None, n, arr, nul, float(divisor), int(d_full), int(col_off), np_,
                                      int(users_per_rank), int(user_row0), pu, pi, _ptr(ws), ws.numel(), _stream()))


def layer_mean_scatter(addends: Sequence[torch.Tensor], divisor: float, d_full: int, col_off: int, row0: int,
                       dst_ptrs: Sequence[int]) -> None:
    """(Σ addends) / divisor over (rows, ds) tables stored at column col_off, rows [row0, row0 + rows) of every
    destination table (d_full wide; peer-mapped pointers) — see tgcn_layer_mean_scatter."""
    lib = _lib.load()
    n = len(addends)
    rows, ds = addends[0].shape
    arr = (ctypes.c_void_p * n)(*[_chk(t, torch.float32, "addend", 2).data_ptr() for t in addends])
    dst = (ctypes.c_void_p * len(dst_ptrs))(*dst_ptrs)
    with torch.cuda.device(addends[0].device):
        check(lib.tgcn_layer_mean_scatter(rows, ds, n, arr, float(divisor), int(d_full), int(col_off), int(row0), len(dst_ptrs), dst,
                                          _stream()))


def propagate_sliced(g: Graph, user_slice: torch.Tensor, item_slice: torch.Tensor, n_layers: int, d_full: int, col_off: int,
                     users_per_rank: int, peer_user_ptrs: Sequence[int], peer_item_ptrs: Sequence[int], single: bool = False,
                     keep: Optional[torch.Tensor] = None, dropout: float = 0.0) -> None:
    """representation on one column slice of the tables, the layer mean stored straight into the peers' row-sharded
    full-width result tables (see tgcn_propagate_sliced).  The caller brackets it with barriers."""
    _chk(user_slice, torch.float32, "user_slice", 2)
    _chk(item_slice, torch.float32, "item_slice", 2)
    ds = user_slice.shape[1]
    if user_slice.shape[0] != g.n_users or item_slice.shape != (g.n_items, ds):
        raise _lib.TgcnError("embedding slices do not match the graph")
    if keep is not None:
        keep = _keep_u8(keep)
        if keep.numel() != g.nnz:
            raise _lib.TgcnError("keep mask must have nnz entries")
    n = len(peer_user_ptrs)
    pu = (ctypes.c_void_p * n)(*peer_user_ptrs)
    pi = (ctypes.c_void_p * n)(*peer_item_ptrs)
    ws = g.workspace(ds, n_layers)
    with torch.cuda.device(user_slice.device):
        check(g.lib.tgcn_propagate_sliced(g.handle, ds, n_layers, int(single), _ptr(user_slice), _ptr(item_slice), _ptr(keep),
                                          float(dropout), int(d_full), int(col_off), n, int(users_per_rank), pu, pi, _ptr(ws),
                                          ws.numel(), _stream()))


def propagate_bwd(g: Graph, grad_out: torch.Tensor, n_layers: int, single: bool = False,
                  keep: Optional[torch.Tensor] = None, dropout: float = 0.0,
                  grad_in: Optional[torch.Tensor] = None, accumulate: bool = False) -> torch.Tensor:
    _chk(grad_out, torch.float32, "grad_out", 2)
    d = grad_out.shape[1]
    if keep is not None:
        keep = _keep_u8(keep)
    if grad_in is None:
        grad_in = torch.empty_like(grad_out)
        accumulate = False
    _chk(grad_in, torch.float32, "grad_in", 2)
    ws = g.workspace(d, n_layers)
    with torch.cuda.device(grad_out.device):
        check(g.lib.tgcn_propagate_bwd(g.handle, d, n_layers, int(single), _ptr(grad_out), _ptr(keep), float(dropout),
                                       int(accumulate), _ptr(grad_in), _ptr(ws), ws.numel(), _stream()))
    return grad_in


def bpr_fwd_bwd(n_users: int, n_items: int, emb: torch.Tensor, user_w: torch.Tensor, item_w: torch.Tensor,
                users: torch.Tensor, pos: torch.Tensor, negs: torch.Tensor, reg_lambda: float,
                grad_emb: Optional[torch.Tensor], grad_w0: Optional[torch.Tensor]) -> torch.Tensor:
    """Fused BPR(SELU)+L2 step.  negs is (n_neg, batch) int32.  Returns losses = [bpr, reg] (device)."""
    lib = _lib.load()
    _chk(emb, torch.float32, "emb", 2)
    d = emb.shape[1]
    users, pos, negs = (_chk(t, torch.int32, n, align=4) for t, n in ((users, "users"), (pos, "pos"), (negs, "negs")))
    batch = users.numel()
    n_neg = negs.numel() // batch
    losses = torch.empty(2, dtype=torch.float32, device=emb.device)
    ws = torch.empty(int(lib.tgcn_bpr_workspace_bytes(batch)), dtype=torch.uint8, device=emb.device)
    with torch.cuda.device(emb.device):
        check(lib.tgcn_bpr_fwd_bwd(n_users, n_items, d, batch, n_neg, _ptr(users), _ptr(pos), _ptr(negs), _ptr(emb),
                                   _ptr(user_w), _ptr(item_w), float(reg_lambda), _ptr(losses), _ptr(grad_emb),
                                   _ptr(grad_w0), _ptr(ws), ws.numel(), _stream()))
    return losses


def eval_topk(mask_graph: Optional[Graph], user_vecs: torch.Tensor, item_vecs: torch.Tensor, k: int,
              users: Optional[torch.Tensor] = None, n_rank: Optional[int] = None,
              item_range: Optional[Tuple[int, int]] = None, user_bias: Optional[torch.Tensor] = None,
              item_bias: Optional[torch.Tensor] = None, finalize: bool = True, by_position: bool = False,
              precision: str = "auto", stats: Optional[dict] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Fused score + mask + top-k.  Returns (ids (n_rank, k) int32, scores (n_rank, k) fp32).

    ``users`` (int32 ids) selects the rows to rank and whose train items are masked; with ``by_position`` the
    user vectors / bias are already packed in list order (row m belongs to users[m]).
    ``precision``: "fp32" (exact FMA, SIMT kernel), "3xtf32" (tcgen05 tensor cores, three TF32 products per score),
    "screen" (one TF32 product per score to find the candidates, exact fp32 re-scoring of those, certificate per row, rows
    that cannot be certified ranked again in 3xTF32; k <= 24) or "auto" (screen where the item range is long enough for it to pay — 65 536 rows at K = 128, 131 072 at
    K = 64 —, else 3xTF32, else fp32).
    ``stats`` (diagnostics; synchronises): a dict that receives ``precision`` (what the call ran at) and ``second_pass_rows``,
    the number of rows a screened call had to rank again (None when the call did not take the screened path)."""
    lib = _lib.load()
    _chk(user_vecs, torch.float32, "user_vecs", 2)
    _chk(item_vecs, torch.float32, "item_vecs", 2)
    K = user_vecs.shape[1]
    if item_vecs.shape[1] != K:
        raise _lib.TgcnError("user and item vectors differ in width")
    if users is not None:
        users = _chk(users, torch.int32, "users", 1, align=4)
        n_rank = users.numel()
    elif n_rank is None:
        n_rank = user_vecs.shape[0]
    i0, i1 = item_range if item_range is not None else (0, item_vecs.shape[0])
    dev = user_vecs.device
    ids = torch.empty((n_rank, k), dtype=torch.int32, device=dev)
    scores = torch.empty((n_rank, k), dtype=torch.float32, device=dev)
    prec = _PRECISIONS[precision]
    nbytes = int(lib.tgcn_eval_workspace_bytes(n_rank, i1 - i0, K, k))
    if nbytes < 0:
        raise _lib.TgcnError("bad eval shape")
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        check(lib.tgcn_eval_topk(mask_graph.handle if mask_graph is not None else None, n_rank, _ptr(users),
                                 _ptr(user_vecs), user_vecs.stride(0), _ptr(item_vecs), item_vecs.stride(0), K, i0, i1,
                                 _ptr(user_bias), _ptr(item_bias), int(by_position), prec, k, int(finalize), _ptr(ids), _ptr(scores),
                                 _ptr(ws), ws.numel(), _stream()))
    if stats is not None:
        biased = int(user_bias is not None or item_bias is not None)
        ran = int(lib.tgcn_eval_resolve_precision(i1 - i0, K, k, biased, prec))
        stats["precision"] = {1: "fp32", 2: "3xtf32", 3: "screen"}[ran]
        off = int(lib.tgcn_eval_screen_queue_offset(n_rank, i1 - i0, K, k, biased)) if ran == 3 else -1
        stats["second_pass_rows"] = int(ws[off:off + 4].view(torch.int32).item()) if off >= 0 else None
    return ids, scores


_PRECISIONS = {"auto": 0, "fp32": 1, "3xtf32": 2, "screen": 3}


def eval_resolve_precision(n_items: int, K: int, k: int, has_bias: bool = False, precision: str = "auto") -> str:
    """The precision an ``eval_topk`` call of this shape runs at ("fp32", "3xtf32" or "screen")."""
    ran = int(_lib.load().tgcn_eval_resolve_precision(n_items, K, k, int(has_bias), _PRECISIONS[precision]))
    return {1: "fp32", 2: "3xtf32", 3: "screen"}[ran]


def topk_merge(mask_graph: Optional[Graph], part_ids: torch.Tensor, part_scores: torch.Tensor,
               users: Optional[torch.Tensor] = None, finalize: bool = True) -> Tuple[torch.Tensor, torch.Tensor]:
    """Merge (n_parts, n_rows, k) partial top-k tables."""
    lib = _lib.load()
    _chk(part_ids, torch.int32, "part_ids", 3)
    _chk(part_scores, torch.float32, "part_scores", 3)
    n_parts, n_rows, k = part_ids.shape
    ids = torch.empty((n_rows, k), dtype=torch.int32, device=part_ids.device)
    scores = torch.empty((n_rows, k), dtype=torch.float32, device=part_ids.device)
    with torch.cuda.device(part_ids.device):
        check(lib.tgcn_topk_merge(mask_graph.handle if mask_graph is not None else None, n_rows, _ptr(users), n_parts, k,
                                  _ptr(part_ids), _ptr(part_scores), int(finalize), _ptr(ids), _ptr(scores), _stream()))
    return ids, scores


def adv_select(g: Graph, emb: torch.Tensor, users: torch.Tensor, cands: torch.Tensor, kmax: int,
               want_scores: bool = False):
    """Hardest-negative selection.  Returns (negs (B, kmax) int32 -1 padded, counts (B,), scores or None)."""
    _chk(emb, torch.float32, "emb", 2)
    users = _chk(users, torch.int32, "users", 1, align=4)
    cands = _chk(cands, torch.int32, "cands", 2, align=4)
    b, c = cands.shape
    negs = torch.empty((b, kmax), dtype=torch.int32, device=emb.device)
    counts = torch.empty(b, dtype=torch.int32, device=emb.device)
    scores = torch.empty((b, c), dtype=torch.float32, device=emb.device) if want_scores else None
    with torch.cuda.device(emb.device):
        check(g.lib.tgcn_adv_select(g.handle, emb.shape[1], b, c, _ptr(users), _ptr(cands), _ptr(emb), kmax, _ptr(negs),
                                    _ptr(counts), _ptr(scores), _stream()))
    return negs, counts, scores


def ltr_pairwise_features(n_users: int, emb, users, items, users_rev, users_desc, items_rev, items_desc,
                          pop_users=None, pop_items=None) -> torch.Tensor:
    lib = _lib.load()
    n_feat = 7 if pop_users is not None else 5
    b = users.numel()
    out = torch.empty((b, n_feat), dtype=torch.float32, device=emb.device)
    with torch.cuda.device(emb.device):
        check(lib.tgcn_ltr_pairwise_features(n_users, emb.shape[1], users_rev.shape[1], b, n_feat,
                                             _ptr(_chk(users, torch.int32, "users", align=4)), _ptr(_chk(items, torch.int32, "items", align=4)),
                                             _ptr(_chk(emb, torch.float32, "emb")), _ptr(_chk(users_rev, torch.float32, "users_rev")),
                                             _ptr(_chk(users_desc, torch.float32, "users_desc")),
                                             _ptr(_chk(items_rev, torch.float32, "items_rev")),
                                             _ptr(_chk(items_desc, torch.float32, "items_desc")),
                                             _ptr(pop_users), _ptr(pop_items), _ptr(out), _stream()))
    return out


def ltr_pairwise_emb_bwd(n_users: int, emb, users, items, gf0, grad_emb) -> None:
    lib = _lib.load()
    with torch.cuda.device(emb.device):
        check(lib.tgcn_ltr_pairwise_emb_bwd(n_users, emb.shape[1], users.numel(), _ptr(users), _ptr(items), _ptr(emb),
                                            _ptr(_chk(gf0, torch.float32, "gf0")), _ptr(grad_emb), _stream()))


def ltr_pack_items(items_emb, items_rev, items_desc, w5: Sequence[float]) -> torch.Tensor:
    lib = _lib.load()
    n, d = items_emb.shape
    D = items_rev.shape[1]
    out = torch.empty((n, d + 2 * D), dtype=torch.float32, device=items_emb.device)
    w = (ctypes.c_float * 5)(*[float(x) for x in w5])
    with torch.cuda.device(items_emb.device):
        check(lib.tgcn_ltr_pack_items(n, d, D, _ptr(_chk(items_emb, torch.float32, "items_emb")),
                                      _ptr(_chk(items_rev, torch.float32, "items_rev")),
                                      _ptr(_chk(items_desc, torch.float32, "items_desc")), w, _ptr(out), _stream()))
    return out


def ltr_pack_users(users: Optional[torch.Tensor], users_emb, users_rev, users_desc) -> torch.Tensor:
    lib = _lib.load()
    d, D = users_emb.shape[1], users_rev.shape[1]
    n = users.numel() if users is not None else users_emb.shape[0]
    out = torch.empty((n, d + 2 * D), dtype=torch.float32, device=users_emb.device)
    with torch.cuda.device(users_emb.device):
        check(lib.tgcn_ltr_pack_users(n, _ptr(users), d, D, _ptr(_chk(users_emb, torch.float32, "users_emb")),
                                      _ptr(_chk(users_rev, torch.float32, "users_rev")),
                                      _ptr(_chk(users_desc, torch.float32, "users_desc")), _ptr(out), _stream()))
    return out


def _rows2d(t: torch.Tensor, name: str) -> torch.Tensor:
    """fp32 CUDA matrix whose rows are contiguous, 16-byte aligned and start every stride(0) % 4 == 0 floats (a column
    slice of a wider table qualifies: no copy)."""
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise _lib.TgcnError(f"{name} must be a CUDA tensor (textgcn_b200 has no CPU path)")
    if t.dtype != torch.float32 or t.dim() != 2:
        raise _lib.TgcnError(f"{name} must be a 2-D fp32 tensor")
    if t.shape[1] > 1 and t.stride(1) != 1 or t.stride(0) % 4 != 0 or t.stride(0) < t.shape[1] or t.data_ptr() % 16 != 0:
        t = t.contiguous()
    return t


def score_batchwise(a: torch.Tensor, b: torch.Tensor, row_bias: Optional[torch.Tensor] = None,
                    col_bias: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None,
                    plane: Optional[int] = None) -> torch.Tensor:
    """out[m, n] = <a[m], b[n]> (+ row_bias[m]) (+ col_bias[n]) in exact fp32 — base_model.py:173-179 as a kernel call.
    ``out`` (M, N) by default; with ``plane = f`` the result is written to out[:, :, f] of a contiguous (M, N, F) tensor
    (the feature planes of ltr_models.py:131-146)."""
    lib = _lib.load()
    a, b = _rows2d(a, "a"), _rows2d(b, "b")
    M, K = a.shape
    N = b.shape[0]
    if b.shape[1] != K or K % 4 != 0:
        raise _lib.TgcnError("operands must share a width that is a multiple of 4")
    if out is None:
        out = torch.empty((M, N), dtype=torch.float32, device=a.device)
    _chk(out, torch.float32, "out", align=4)
    if plane is None:
        if out.shape != (M, N):
            raise _lib.TgcnError("out must be (M, N)")
        ptr, ldr, ldc = out.data_ptr(), N, 1
    else:
        if out.dim() != 3 or out.shape[:2] != (M, N) or not 0 <= plane < out.shape[2]:
            raise _lib.TgcnError("out must be (M, N, F) with 0 <= plane < F")
        F = out.shape[2]
        ptr, ldr, ldc = out.data_ptr() + 4 * plane, N * F, F
    if M == 0 or N == 0:
        return out
    rb = None if row_bias is None else _chk(row_bias, torch.float32, "row_bias", align=4)
    cb = None if col_bias is None else _chk(col_bias, torch.float32, "col_bias", align=4)
    with torch.cuda.device(a.device):
        check(lib.tgcn_score_batchwise(M, N, K, a.data_ptr(), a.stride(0), b.data_ptr(), b.stride(0), _ptr(rb), _ptr(cb), ptr, ldr, ldc,
                                       _stream()))
    return out


def score_pairwise_adv(users_emb: torch.Tensor, items_emb: torch.Tensor) -> torch.Tensor:
    """(B, d) x (B, C, d) -> (B, C): advanced_sampling.py:37-44 as a kernel call (shape kept for B == 1, G14)."""
    lib = _lib.load()
    u = _rows2d(users_emb, "users_emb")
    it = _chk(items_emb.contiguous(), torch.float32, "items_emb", 3)
    B, C, d = it.shape
    if u.shape != (B, d) or d % 4 != 0:
        raise _lib.TgcnError("users_emb must be (B, d) for items_emb (B, C, d), d % 4 == 0")
    out = torch.empty((B, C), dtype=torch.float32, device=u.device)
    if B * C:
        with torch.cuda.device(u.device):
            check(lib.tgcn_score_pairwise_adv(B, C, d, u.data_ptr(), u.stride(0), it.data_ptr(), out.data_ptr(), _stream()))
    return out


def ltr_features_rows(ue, ie, ur, ud, ir, idesc) -> torch.Tensor:
    """(B, 5) features of row-aligned vectors (ltr_models.py:148-166)."""
    lib = _lib.load()
    ts = [_rows2d(t, n) for t, n in ((ue, "u emb"), (ie, "i emb"), (ur, "u reviews"), (ud, "u desc"), (ir, "i reviews"), (idesc, "i desc"))]
    B, d = ts[0].shape
    D = ts[2].shape[1]
    if ts[1].shape != (B, d) or any(t.shape != (B, D) for t in ts[2:]) or d % 4 or D % 4:
        raise _lib.TgcnError("feature vectors must be row-aligned: emb (B, d), text (B, D), widths multiples of 4")
    out = torch.empty((B, 5), dtype=torch.float32, device=ts[0].device)
    if B:
        args = []
        for t in ts:
            args += [t.data_ptr(), t.stride(0)]
        with torch.cuda.device(out.device):
            check(lib.tgcn_ltr_features_rows(B, d, D, *args, out.data_ptr(), 5, _stream()))
    return out


def topk_metrics(pred_ids: torch.Tensor, true_ptr: torch.Tensor, true_ids: torch.Tensor, ks: Sequence[int]) -> torch.Tensor:
    """(len(ks), 5) float64 means of [recall, precision, hit, ndcg, f1] (utils.py:36-63) from the (n, kmax) int32 id
    table and the CSR (ptr int64 (n+1), ids int32) of true test items; one kernel + a fixed-order reduction."""
    lib = _lib.load()
    pred = _chk(pred_ids, torch.int32, "pred_ids", 2, align=4)
    tp = _chk(true_ptr, torch.int64, "true_ptr", 1, align=8)
    ti = _chk(true_ids, torch.int32, "true_ids", 1, align=4)
    n, kmax = pred.shape
    if tp.numel() != n + 1:
        raise _lib.TgcnError("true_ptr must have n_rows + 1 entries")
    ks = [int(k) for k in ks]
    out = torch.empty((len(ks), 5), dtype=torch.float64, device=pred.device)
    ws = torch.empty(int(lib.tgcn_topk_metrics_workspace_bytes()), dtype=torch.uint8, device=pred.device)
    arr = (ctypes.c_int32 * len(ks))(*ks)
    with torch.cuda.device(pred.device):
        check(lib.tgcn_topk_metrics(n, kmax, pred.data_ptr(), tp.data_ptr(), ti.data_ptr(), len(ks), arr, out.data_ptr(), ws.data_ptr(),
                                    ws.numel(), _stream()))
    return out


def adam_prepare(step: torch.Tensor, bc: torch.Tensor, beta1: float, beta2: float) -> None:
    """++step (int64 device counter) and bc = [1 - beta1^step, sqrt(1 - beta2^step)] (2 floats on the device)."""
    lib = _lib.load()
    with torch.cuda.device(step.device):
        check(lib.tgcn_adam_prepare(_ptr(step), _ptr(bc), beta1, beta2, _stream()))


def adam_step_dev(p, g, m, v, lr, beta1, beta2, eps, bc: torch.Tensor) -> None:
    lib = _lib.load()
    with torch.cuda.device(p.device):
        check(lib.tgcn_adam_step_dev(p.numel(), _ptr(p), _ptr(_chk(g, torch.float32, "grad")), _ptr(m), _ptr(v), lr, beta1,
                                     beta2, eps, _ptr(bc), _stream()))


def adam_step(p, g, m, v, lr, beta1, beta2, eps, step) -> None:
    lib = _lib.load()
    with torch.cuda.device(p.device):
        check(lib.tgcn_adam_step(p.numel(), _ptr(p), _ptr(_chk(g, torch.float32, "grad")), _ptr(m), _ptr(v), lr, beta1,
                                 beta2, eps, step, _stream()))
