"""CPU: the evaluation-metric oracle (oracle/evalmetrics.py) against hand-computed values and invariants."""
import numpy as np

from oracle import evalmetrics as em


def test_label_instances_semantics():
    a = np.zeros((8, 8), int)
    a[0:3, 0:3] = 5
    a[0:3, 4:7] = 5          # same value, joined only through the diagonal pixel below
    a[3, 3] = 5
    a[5:7, 1:4] = 2
    lab = em.label_instances(a)
    assert lab.max() == 2 and (lab[a == 5] == 1).all() and (lab[a == 2] == 2).all() and (lab[a == 0] == 0).all()
    b = np.zeros((6, 6), int)
    b[1:3, 1:3] = 7
    b[1:3, 3:5] = 9          # touching instances with different ids stay separate
    b[4, 0] = 7              # a second, disconnected part of id 7 becomes its own component
    lab = em.label_instances(b)
    assert lab.max() == 3 and lab[1, 1] == 1 and lab[1, 3] == 2 and lab[4, 0] == 3
    assert em.label_instances(np.zeros((4, 5), int)).max() == 0


def test_aji_plus_hand_computed():
    t = np.zeros((10, 10), int)
    t[1:4, 1:4] = 1
    t[6:9, 6:9] = 2
    p = np.zeros((10, 10), int)
    p[1:4, 2:5] = 1
    p[6:9, 6:9] = 2
    p[0, 9] = 3
    # pair (1,1): inter 6, union 12; pair (2,2): inter 9, union 9; unpaired prediction 3 adds 1 to the union
    assert abs(em.aji_plus(t, p) - (6 + 9) / (12 + 9 + 1)) < 1e-15
    assert em.aji_plus(t, t) == 1.0
    q = np.zeros((10, 10), int)
    q[0, 0] = 1
    assert em.aji_plus(t, q) == 0.0


def test_aji_plus_against_the_reference_function():
    """tests/golden/aji_plus_reference.npz: values computed by the reference's OWN get_fast_aji_plus
    (src/evaluation/stats_utils.py:98-179, imported by path in make_golden.py aji; numpy + scipy only) -- the oracle must
    reproduce them exactly (same pairwise sums, same linear_sum_assignment call)"""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "aji_plus_reference.npz"))
    for k in range(int(g["n"])):
        t, p = g[f"true{k}"], g[f"pred{k}"]
        assert em.aji_plus(t, p) == float(g[f"aji{k}"])
        assert em.aji_plus(p, t) == float(g[f"aji_swapped{k}"])


def test_aji_plus_is_invariant_to_id_permutation_and_symmetric_in_pairing():
    rng = np.random.default_rng(3)
    t = np.zeros((40, 40), int)
    p = np.zeros((40, 40), int)
    k = 1
    for y in range(0, 40, 10):
        for x in range(0, 40, 10):
            t[y + 1:y + 8, x + 1:x + 8] = k
            dy, dx = rng.integers(-1, 3, 2)
            p[max(0, y + 1 + dy):y + 8 + dy, max(0, x + 1 + dx):x + 8 + dx] = k
            k += 1
    base = em.aji_plus(t, p)
    perm = np.concatenate([[0], rng.permutation(np.arange(1, k))])
    assert abs(em.aji_plus(perm[t], p) - base) < 1e-15
    assert 0.0 < base < 1.0
