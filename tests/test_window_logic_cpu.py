"""CPU check of the window-consumption rule of k_wresolve (csrc/bsr_window.cuh, phase B).

The kernel decides in closed form how many slots of a speculative window a chain consumes; the rule it must equal is
the reference's sequential loop (codes/bsr_class.py:174-252): proposals are taken one at a time, an accept ends the
window (later slots were proposed from a stale state), and `total` consecutive rejections >= val stop the chain, but
only at a sweep boundary (the `while total < val` test of bsr_class.py:174 runs once per K proposals)."""
import numpy as np


def sequential_rule(valid, acc, total, val, K, p0):
    n_cons, a, done = 0, -1, False
    for i in range(len(valid)):
        if not valid[i]:
            break
        n_cons += 1
        total += 1
        if acc[i]:
            a = i
            break
        if val > 0 and total >= val and (p0 + i + 1) % K == 0:
            done = True
            break
    return n_cons, a, done, (0 if a >= 0 else total)


def closed_form(valid, acc, total, val, K, p0):
    """line-by-line mirror of the device code"""
    L = 64                                             # lanes per chain (two warps exchange their ballots)
    valid_mask = sum(1 << i for i in range(len(valid)) if valid[i])
    acc_mask = sum(1 << i for i in range(len(acc)) if acc[i])
    inval = (~valid_mask) & ((1 << L) - 1)
    n_valid = (inval & -inval).bit_length() - 1 if inval else L
    first_acc = acc_mask & (((1 << L) - 1) if n_valid >= L else ((1 << n_valid) - 1))
    n_cons = (first_acc & -first_acc).bit_length() if first_acc else n_valid
    a = n_cons - 1 if first_acc else -1
    done = False
    if val > 0:
        k0 = p0 % K
        need = max(1, val - total)
        stop = need + ((K - (k0 + need) % K) % K)
        if stop <= n_cons - (1 if a >= 0 else 0):
            n_cons, a, done = stop, -1, True
    total += n_cons
    return n_cons, a, done, (0 if a >= 0 else total)


def test_closed_form_matches_sequential_rule():
    rng = np.random.default_rng(1)
    for _ in range(20000):
        K = int(rng.integers(1, 7))
        W = int(rng.choice([1, 7, 32, 45, 64]))
        n_valid = int(rng.integers(0, W + 1))
        valid = [i < n_valid for i in range(64)]
        acc = list(rng.random(64) < rng.choice([0.0, 0.02, 0.3]))
        val = int(rng.choice([0, 1, 5, 25, 100]))
        total = int(rng.integers(0, max(val, 1)))
        p0 = int(rng.integers(0, 1000))
        assert closed_form(valid, acc, total, val, K, p0) == sequential_rule(valid, acc, total, val, K, p0), \
            (K, n_valid, val, total, p0)
