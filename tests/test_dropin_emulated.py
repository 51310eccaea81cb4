"""CPU: the drop-in scenarios of tests/dropin_scenarios.py with the kernels EMULATED (tests/emul.py) — the host logic of the
mixins over the unmodified reference classes (method resolution, G18 load order, autograd wiring, triple construction,
checkpoint interchange) checked where no GPU exists.  The same scenarios run on the real kernels in tests/test_gpu_dropin.py.
Skipped where the reference is neither mounted nor staged."""
import pytest

import dropin_scenarios as S
import emul

pytestmark = S.needs_reference


@pytest.fixture(scope="module")
def env(tmp_path_factory):
    yield from S.make_env(tmp_path_factory, "cpu")


def test_lgcn_dropin_host_logic(env, monkeypatch):
    emul.install(monkeypatch)
    S.scenario_lgcn(env)


@pytest.mark.parametrize("model_name", ["ltr_linear", "ltr_pop"])
def test_ltr_dropin_host_logic(env, model_name, monkeypatch):
    emul.install(monkeypatch)
    S.scenario_ltr(env, model_name, monkeypatch)


def test_adv_sampling_dropin_host_logic(env, monkeypatch):
    emul.install(monkeypatch)
    S.scenario_adv(env)


def test_every_8b_method_resolves_to_a_mixin_and_the_shell_stays_the_references(env):
    """each hot-path method of the drop-in classes is the kernel-backed override; everything else (constructor,
    fit loop, checkpointing) is the reference's own code.  Without the emulation a CPU device refuses instead of falling back."""
    import pytest as _pytest
    from textgcn_b200 import TgcnError
    from textgcn_b200.models import B200AdvSampl, B200HotPath, B200LTR
    T, cls = env["T"], env["cls"]
    hot = ["representation", "layer_aggregation", "score_pairwise", "score_batchwise", "bpr_loss", "reg_loss", "get_loss", "predict", "evaluate"]
    for name in hot:
        assert getattr(cls["lgcn"], name) is getattr(B200HotPath, name), name
    for name in ["fit", "checkpoint", "load_model", "_copy_params", "_copy_dataset_params", "_init_embeddings", "layer_combination"]:
        assert getattr(cls["lgcn"], name) is getattr(T.BaseModel, name), name
    for name in ["score_pairwise_adv", "get_loss"]:
        assert getattr(cls["adv_sampling"], name) is getattr(B200AdvSampl, name), name
    assert cls["adv_sampling"].representation is B200HotPath.representation and cls["adv_sampling"].fit is T.BaseModel.fit
    ltr = ["get_user_vectors", "get_item_vectors", "get_features_batchwise", "get_features_pairwise", "score_batchwise_ltr",
           "score_pairwise_ltr", "evaluate_ltr", "bpr_loss", "get_loss", "_rank"]
    for key in ("ltr_linear", "ltr_pop"):
        for name in ltr:
            assert getattr(cls[key], name) is getattr(B200LTR, name), (key, name)
        assert cls[key].predict is B200HotPath.predict and cls[key].fit is T.BaseModel.fit
        assert cls[key].__init__ is (T.LTRLinear if key == "ltr_linear" else T.LTRLinearWPop).__init__ or cls[key].__init__ is T.LTRLinear.__init__
    a = env["args"]("lgcn", False, "structure")
    model = cls["lgcn"](a, T.BaseDataset(a))            # the reference constructor runs unchanged
    assert set(model.state_dict()) == {"embedding_user.weight", "embedding_item.weight"}
    with _pytest.raises(TgcnError):                     # CPU device, real ops: refuses instead of falling back
        model.representation
