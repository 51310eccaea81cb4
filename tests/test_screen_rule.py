"""The certificate of the screened eval path as arithmetic on numbers (no GPU): a numpy restatement of what `screen_finalize`
(textgcn_b200/csrc/eval_tc.cu) decides from a row's candidate list, checked against brute force on adversarial inputs.

Claim under test ("Eval design, long sweeps"): let s~ be approximate scores with |s~ - s| <= eps for every item.
Keep the KL best items by s~ (sorted).  Re-score the k best of them exactly, E = the smallest of those exact scores; go on
re-scoring while s~_j + eps >= E; sort the re-scored entries on (exact score desc, id asc).  If the scan stopped before the end
of a FULL list (or the list is not full), the first k entries ARE the exact top-k in canonical order; otherwise the row must be
flagged for the second pass.  The test draws scores with ties, near-ties inside the band and errors that use the whole budget.
"""
import numpy as np
import pytest


def screen_rule(approx, exact_of, eps, k, kl):
    """-> (ids of the first k re-scored entries in canonical order, flagged).  `approx`: (n,) approximate scores, `exact_of(id)`."""
    n = approx.shape[0]
    order = np.lexsort((np.arange(n), -approx))[:kl]  # the list: best KL by approximate score (any tie order is allowed here)
    ms = approx[order]
    n_valid = order.shape[0]
    p1 = min(k, n_valid)
    me = {int(i): exact_of(int(i)) for i in order[:p1]}
    p = p1
    flagged = False
    if n_valid >= k:
        e_min = min(me.values())
        while p < n_valid and ms[p] + eps >= e_min:
            me[int(order[p])] = exact_of(int(order[p]))
            p += 1
        flagged = p == kl  # no entry of a full list could be ruled out: an item outside the list might belong to the k best
    ranked = sorted(me.items(), key=lambda t: (-t[1], t[0]))
    return [i for i, _ in ranked[:k]], flagged


def brute_force(exact, k):
    n = exact.shape[0]
    return list(np.lexsort((np.arange(n), -exact))[:k])


@pytest.mark.parametrize("seed", range(40))
def test_certified_rows_are_the_exact_topk(seed):
    rng = np.random.default_rng(seed)
    n = int(rng.integers(30, 400))
    k = int(rng.integers(1, 25))
    kl = 40
    style = seed % 4
    if style == 0:    # generic scores
        exact = rng.normal(size=n)
    elif style == 1:  # plateaus of exact ties
        exact = rng.integers(0, 6, size=n).astype(np.float64)
    elif style == 2:  # clusters of near-ties well inside the band
        exact = np.repeat(rng.normal(size=(n + 7) // 8), 8)[:n] + rng.normal(size=n) * 1e-4
    else:             # a dense top: many items within a few eps of each other
        exact = rng.normal(size=n) * 1e-3
    eps = float(10.0 ** rng.uniform(-4, -1.5))
    # errors that use the whole budget, signed adversarially for half of the draws
    err = rng.uniform(-eps, eps, size=n)
    if seed % 2:
        top = brute_force(exact, min(n, k + 5))
        err[:] = eps
        err[top] = -eps  # push the true top down and everything else up
    approx = exact + err
    ids, flagged = screen_rule(approx, lambda i: float(exact[i]), eps, k, kl)
    if not flagged:
        assert ids == [int(i) for i in brute_force(exact, k)], (seed, style, eps)
    else:
        # flagged rows are the ones the 3xTF32 second pass ranks again: only legitimate when the band really reaches the list's end
        order = np.lexsort((np.arange(n), -approx))[:kl]
        assert order.shape[0] == kl


def test_short_lists_are_never_flagged_and_exact():
    rng = np.random.default_rng(7)
    exact = rng.normal(size=25)  # fewer items than the list holds: nothing is ever excluded
    approx = exact + rng.uniform(-0.5, 0.5, size=25)
    ids, flagged = screen_rule(approx, lambda i: float(exact[i]), 0.5, 20, 40)
    assert not flagged and ids == [int(i) for i in brute_force(exact, 20)]
    ids, flagged = screen_rule(approx[:7], lambda i: float(exact[i]), 0.5, 20, 40)  # fewer items than k: all of them, in order
    assert not flagged and ids == [int(i) for i in brute_force(exact[:7], 7)]


def test_wide_band_flags_instead_of_guessing():
    exact = np.linspace(1.0, 0.0, 200)
    approx = exact.copy()
    ids, flagged = screen_rule(approx, lambda i: float(exact[i]), eps=1.0, k=20, kl=40)  # eps covers every gap: nothing can be ruled out
    assert flagged
    ids, flagged = screen_rule(approx, lambda i: float(exact[i]), eps=1e-4, k=20, kl=40)
    assert not flagged and ids == list(range(20))


# ---- the split-fastest rasterisation of the streamed eval variant (eval_tc.cu, `split_fastest`) ---------------------------------
def _raster(bx, by, gx, gy, ctas):
    """(blockIdx.x, blockIdx.y) -> (user tile, item split), restating the kernel's index arithmetic."""
    lin = by * gx + bx
    grp = lin // ctas
    return (grp // gy) * ctas + lin % ctas, grp % gy


@pytest.mark.parametrize("gx,gy,ctas", [(148, 8, 2), (6, 3, 2), (7, 4, 1), (2, 1, 2), (1, 5, 1), (150, 2, 2)])
def test_split_fastest_rasterisation_is_a_bijection_that_keeps_pairs_together(gx, gy, ctas):
    seen = {}
    for by in range(gy):
        for bx in range(gx):
            seen[(bx, by)] = _raster(bx, by, gx, gy, ctas)
    assert sorted(seen.values()) == [(t, s) for t in range(gx) for s in range(gy)]  # every (tile, split) exactly once
    if ctas == 2:  # the two CTAs of a cluster (adjacent blockIdx.x, same blockIdx.y) take adjacent user tiles of the SAME split
        for by in range(gy):
            for bx in range(0, gx, 2):
                (t0, s0), (t1, s1) = seen[(bx, by)], seen[(bx + 1, by)]
                assert s0 == s1 and t1 == t0 + 1 and t0 % 2 == 0
    # CTAs are scheduled in linear order: consecutive groups walk the splits of one tile group before moving to the next
    order = [seen[(lin % gx, lin // gx)] for lin in range(gx * gy)]
    tiles_in_order = [t // ctas for t, _ in order]
    assert tiles_in_order == sorted(tiles_in_order)
