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
