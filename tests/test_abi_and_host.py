"""CPU: the C-ABI library loads and exports every declared symbol; host-side logic that needs no GPU."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from conftest import ROOT, load_golden
from oracle import lightgcn_oracle as O


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "tgcn_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tgcn_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as entry
    entry.build()
    from textgcn_b200 import _lib
    return _lib.load()


def test_library_exports_every_declared_symbol(lib):
    from textgcn_b200 import _lib
    declared = _declared_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/tgcn_b200.h but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature"
    assert set(_lib.SIGNATURES) == set(declared)
    assert lib.tgcn_abi_version() == _lib.ABI_VERSION


def test_size_queries_and_argument_validation_without_gpu(lib):
    assert lib.tgcn_bpr_workspace_bytes(2048) >= 2048 * 8
    assert lib.tgcn_eval_workspace_bytes(2048, 63000, 64, 20) > 0
    assert lib.tgcn_eval_workspace_bytes(0, 63000, 64, 20) < 0
    # argument errors are reported through the return code + tgcn_last_error, never by aborting
    rc = lib.tgcn_topk_merge(None, 0, None, 1, 20, None, None, 1, None, None, None)
    assert rc != 0 and b"bad sizes" in lib.tgcn_last_error()
    rc = lib.tgcn_eval_topk(None, 8, None, None, 64, None, 64, 63, 0, 10, None, None, 0, 0, 20, 1, None, None, None, 0, None)
    assert rc != 0 and b"bad shapes" in lib.tgcn_last_error()
    handle = ctypes.c_void_p()
    rc = lib.tgcn_graph_create(ctypes.byref(handle), 5, 4, 26, None, None, None, None)
    assert rc != 0 and handle.value is None


def test_eval_precision_resolution_is_host_logic(lib, monkeypatch):
    """Which of the three eval paths a call takes is decided on the host from the shape alone (include/tgcn_b200.h, precision):
    the screened path for long item sweeps of narrow, bias-free contractions; 3xTF32 otherwise; fp32 where the tensor-core
    kernels do not apply.  An explicit precision is returned unchanged."""
    for var in ("TGCN_EVAL_SCREEN",):
        monkeypatch.delenv(var, raising=False)
    res = lambda n_items, K, k, bias=0, prec=0: lib.tgcn_eval_resolve_precision(n_items, K, k, bias, prec)  # noqa: E731
    assert res(2_000_000, 128, 20) == 3          # c5: screened
    assert res(63_000, 64, 20) == 2              # c2: short sweep stays on 3xTF32
    assert res(65_536, 128, 20) == 3 and res(65_535, 128, 20) == 2
    assert res(131_072, 64, 20) == 3 and res(131_071, 64, 20) == 2
    assert res(98_304, 96, 20) == 3 and res(98_303, 96, 20) == 2
    assert res(2_000_000, 128, 25) == 2          # k > 24: no margin in the 40-entry list
    assert res(2_000_000, 128, 20, bias=1) == 2  # bias terms / wide contractions: the streamed screened form is opt-in
    assert res(63_000, 1600, 20, bias=1) == 2
    assert res(63_000, 1600, 65) == 1            # k > 64: exact fp32 kernel
    assert res(63_000, 62, 20) == 1              # K % 4 != 0
    for prec in (1, 2, 3):
        assert res(63_000, 64, 20, prec=prec) == prec
    # the diagnostics offset exists exactly where the screened path can run, and lies inside the workspace
    for n_rank, n_items, K, k, bias in ((2048, 2_000_000, 128, 20, 0), (300, 20_000, 1600, 20, 1), (64, 500, 32, 1, 0)):
        off = lib.tgcn_eval_screen_queue_offset(n_rank, n_items, K, k, bias)
        assert 0 <= off < lib.tgcn_eval_workspace_bytes(n_rank, n_items, K, k) - 4
    assert lib.tgcn_eval_screen_queue_offset(2048, 2_000_000, 128, 30, 0) == -1


def test_product_refuses_cpu_tensors_and_never_imports_the_oracle():
    from textgcn_b200 import TgcnError, ops
    with pytest.raises(TgcnError):
        ops.propagate_fwd(None, torch.zeros(4, 4), torch.zeros(4, 4), 1)
    from textgcn_b200.models import BaseModel, make_params
    with pytest.raises(TgcnError):
        BaseModel(make_params(device=torch.device("cpu")), object())
    pkg = os.path.join(ROOT, "textgcn_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("parity oracle", ""), f"{f} references the oracle"


@pytest.mark.parametrize("case", ["dummy_lgcn", "small_lgcn_d64", "small_adv"])
def test_host_norm_adj_builder_is_bit_exact(case):
    from textgcn_b200.graph import csr_to_norm_matrix, norm_adj_csr
    g = load_golden(case)
    rowptr, col, val = norm_adj_csr(torch.from_numpy(g["train_u"]), torch.from_numpy(g["train_i"]), int(g["n_users"]), int(g["n_items"]))
    assert np.array_equal(col.numpy(), g["norm_col"])
    assert np.array_equal(val.numpy().view(np.uint32), g["norm_val"].view(np.uint32))
    nm = csr_to_norm_matrix(rowptr, col, val)
    assert np.array_equal(nm.indices()[0].numpy(), g["norm_row"])


def test_truth_csr_host_logic_and_metrics_refuse_cpu():
    from textgcn_b200 import TgcnError
    from textgcn_b200 import metrics as M
    t = M.TruthCSR.from_lists([[3], [], [2, 2, 7]], "cpu")
    assert t.ptr.tolist() == [0, 1, 1, 4] and t.ids.tolist() == [3, 2, 2, 7] and t.n_rows == 3
    p = M.TruthCSR.from_pairs(torch.tensor([2, 0, 2, 2]), torch.tensor([2, 3, 2, 7]), 3)   # order inside a row is kept
    assert p.ptr.tolist() == t.ptr.tolist() and p.ids.tolist() == t.ids.tolist()
    with pytest.raises(TgcnError):   # the metric pass is a kernel: no CPU path
        M.calculate_metrics(torch.tensor([[1, 2, 3]]), [[1]], [3])