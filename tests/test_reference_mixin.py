"""CPU, authoring container only: the mixins compose with the UNMODIFIED reference classes as INTEGRATION.md describes.
Skipped where /root/reference is absent (the GPU box); nothing here runs a kernel."""
import os
import sys

import pytest

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "TextGCN")), reason="reference not mounted")


def _ref():
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    import ref_shims
    ref_shims.install()
    import TextGCN
    from TextGCN.parser import parse_args
    return TextGCN, parse_args


def test_mixins_override_exactly_the_hot_path_methods(tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)
    TextGCN, parse_args = _ref()
    from textgcn_b200 import TgcnError
    from textgcn_b200.models import B200AdvSampl, B200HotPath, B200LTR

    class B200BaseModel(B200HotPath, TextGCN.BaseModel):
        pass

    class B200AdvSamplModel(B200AdvSampl, B200HotPath, TextGCN.AdvSamplModel):
        pass

    class B200LTRLinear(B200LTR, B200HotPath, TextGCN.LTRLinear):
        pass

    for name in ["representation", "layer_aggregation", "get_loss", "bpr_loss", "reg_loss", "predict", "evaluate", "score_batchwise"]:
        assert getattr(B200BaseModel, name) is getattr(B200HotPath, name), name
    for name in ["fit", "checkpoint", "load_model", "_copy_params", "_init_embeddings", "layer_combination"]:
        assert getattr(B200BaseModel, name) is getattr(TextGCN.BaseModel, name), name   # the shell stays the reference's
    assert B200AdvSamplModel.get_loss is B200AdvSampl.get_loss and B200AdvSamplModel.representation is B200HotPath.representation
    assert B200LTRLinear._rank is B200LTR._rank and B200LTRLinear.get_loss is B200LTR.get_loss

    args = parse_args(["--model", "lgcn", "-d", os.path.join(REF, "data", "dummy"), "-k", "2", "3", "--gpu", "", "--quiet", "--slurm", "--uid", "t"])
    ds = TextGCN.BaseDataset(args)
    model = B200BaseModel(args, ds)              # the reference constructor runs unchanged
    assert set(model.state_dict()) == {"embedding_user.weight", "embedding_item.weight"}
    with pytest.raises(TgcnError):               # CPU device: the kernel-backed path refuses instead of falling back
        model.representation
