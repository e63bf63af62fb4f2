"""Model check of the cooperative search kernel's work schedule (csrc/gnss_kernels.cuh, search_kernel_coop: `tb`,
`part`, `count_parts`, the publish / merge hand-over).  The formulas are restated here line by line; the test
checks, over many (rows, blocks, groups) shapes, the invariants the kernel's correctness and liveness rest on:
every (row, block) unit is done exactly once, every cut row has exactly one finisher (the group holding block 0),
the planes the finisher waits for are exactly the ones the other groups publish, and a group never waits before it
has published everything it owes (no waiting on a waiter)."""
import random


def tb(g, n_tail, K, ngroups, row_granular):
    if row_granular:
        return min(g, n_tail) * K
    return (g * (n_tail * K)) // ngroups                     # 32-bit unsigned in the kernel; host checks the range


def parts_of(group, n_rows, K, ngroups, row_granular):
    """[(row, kb, ke)] in execution order: `full` strided whole rows, then the tail range cut at its row boundary."""
    full, n_tail = n_rows // ngroups, n_rows % ngroups
    tail_base = full * ngroups
    out = [(i * ngroups + group, 0, K) for i in range(full)]
    t0, t1 = tb(group, n_tail, K, ngroups, row_granular), tb(group + 1, n_tail, K, ngroups, row_granular)
    if t0 < t1:
        r0 = t0 // K
        e0 = min(t1, (r0 + 1) * K)
        out.append((tail_base + r0, t0 - r0 * K, e0 - r0 * K))
        if e0 < t1:
            out.append((tail_base + r0 + 1, 0, t1 - e0))
    return out


def finisher_of(group, row, n_rows, K, ngroups):
    n_tail, tail_base = n_rows % ngroups, (n_rows // ngroups) * ngroups
    gf = group - 1
    while tb(gf, n_tail, K, ngroups, False) > (row - tail_base) * K:
        gf -= 1
    return gf


def check(n_rows, K, ngroups, row_granular):
    done = {}
    owed = {}                                                 # finisher group -> blocks published to it
    expects = {}                                              # finisher group -> (row, blocks it waits for)
    for g in range(ngroups):
        ps = parts_of(g, n_rows, K, ngroups, row_granular)
        published_all = True
        for i, (row, kb, ke) in enumerate(ps):
            assert 0 <= row < n_rows and 0 <= kb < ke <= K
            for k in range(kb, ke):
                assert (row, k) not in done, "unit done twice"
                done[(row, k)] = g
            if kb > 0:                                        # later part: publish every block to the finisher
                assert not row_granular
                gf = finisher_of(g, row, n_rows, K, ngroups)
                assert gf < g
                owed[gf] = owed.get(gf, 0) + (ke - kb)
                assert i == n_rows // ngroups, "a group publishes only as the first thing of its tail (nothing waits before it)"
            elif ke < K:                                      # leading part of a cut row: this group finishes it
                assert not row_granular
                assert g not in expects, "one finisher role per group"
                expects[g] = (row, K - ke)
                assert i == len(ps) - 1, "a group waits only as the very last thing it does"
    assert len(done) == n_rows * K, "unit never done"
    assert {g: n for g, (row, n) in expects.items()} == owed
    for g, (row, n) in expects.items():                       # the finisher holds block 0, the publishers the rest
        assert done[(row, 0)] == g
        assert all(done[(row, k)] > g for k in range(K - n, K))


def test_schedule_invariants_on_the_reference_shapes():
    for n_prn in (1, 2, 4, 8, 16, 32):
        for groups in (74, 37, 148, 18):
            for K in (1, 2, 20):
                rows = n_prn * 41
                n = min(groups, rows * K)
                check(rows, K, n, False)
                check(rows, K, min(groups, rows), True)


def test_schedule_invariants_on_random_shapes():
    rng = random.Random(6102)
    for _ in range(3000):
        rows, K, groups = rng.randint(1, 400), rng.randint(1, 40), rng.randint(1, 160)
        check(rows, K, min(groups, rows * K), False)
        check(rows, K, min(groups, rows), True)
