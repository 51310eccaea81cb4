"""GPU: the REAL drop-in — ``textgcn_b200.dropin`` mixins in front of the UNMODIFIED reference classes —
against the same reference classes on the CPU: ``fit`` through the reference's own loop, ``predict``, ``evaluate``,
checkpoint round trips in both directions, ``LTRLinear --load_base --freeze`` (the G18 load order), AdvSampl steps, and a
call of every §8(b) method by its reference signature (scenarios in tests/dropin_scenarios.py).

The reference lives in ``/root/reference`` (authoring container) or in the git-ignored copy ``baseline/_ref`` that
``__graft_entry__.build()`` stages and the gpurun snapshot carries to the GPU box; skipped if neither is present.
"""
import pytest

import dropin_scenarios as S

pytestmark = [pytest.mark.gpu, S.needs_reference]


@pytest.fixture(scope="module")
def env(tmp_path_factory):
    yield from S.make_env(tmp_path_factory, "cuda")


def test_lgcn_dropin_fit_predict_evaluate_checkpoint(env):
    S.scenario_lgcn(env)


@pytest.mark.parametrize("model_name", ["ltr_linear", "ltr_pop"])
def test_ltr_dropin_with_load_base_and_freeze(env, model_name, monkeypatch):
    S.scenario_ltr(env, model_name, monkeypatch)


def test_adv_sampling_dropin_step_matches_reference(env):
    S.scenario_adv(env)
