"""The drop-in itself: the kernel-backed mixins in front of the REFERENCE's own model classes.

The reference (sergey-volokhin/TextGCN) has no FFI; its plugin surface is method override on ``BaseModel``
(SURVEY.md §8b).  ``make_dropin_classes(TextGCN)`` takes the imported reference package and returns the four
subclasses a maintainer would put in ``TextGCN/b200.py`` and point ``main.get_class`` at (main.py:16-22).  Everything
that is not the hot path — the constructor, ``fit``'s loop, ``checkpoint`` / ``load_model``, logging, the DataLoader —
stays the reference's own code, unchanged:

    import TextGCN
    from textgcn_b200.dropin import make_dropin_classes
    cls = make_dropin_classes(TextGCN)
    model = cls["lgcn"](args, dataset)         # BaseDataset / AdvSamplDataset / LTRDataset as before
    model.fit(loader); model.predict(range(dataset.n_users), with_scores=True, save=True)

``tests/test_gpu_dropin.py`` drives exactly these classes on a B200 against the same reference classes on the CPU.
"""
from __future__ import annotations

from .models import B200AdvSampl, B200HotPath, B200LTR


def make_dropin_classes(ref) -> dict:
    """``ref``: the imported reference package (``import TextGCN``).  Returns {--model name: class} like main.get_class."""

    class B200BaseModel(B200HotPath, ref.BaseModel):
        pass

    class B200AdvSamplModel(B200AdvSampl, B200HotPath, ref.AdvSamplModel):
        pass

    # LTRLinear.__init__ re-binds evaluate / score_pairwise / score_batchwise to self.*_ltr by instance attribute
    # (ltr_models.py:177-179); B200LTR supplies those *_ltr methods, so the re-binding lands on the kernel-backed ones
    class B200LTRLinear(B200LTR, B200HotPath, ref.LTRLinear):
        pass

    class B200LTRLinearWPop(B200LTR, B200HotPath, ref.LTRLinearWPop):
        pass

    return {"lgcn": B200BaseModel, "adv_sampling": B200AdvSamplModel, "ltr_linear": B200LTRLinear, "ltr_pop": B200LTRLinearWPop}
