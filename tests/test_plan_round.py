"""CPU: the layout of one chunk round in the anchor arena (rh_plan_round — the code the streaming scheduler runs before it
launches anything): order, arena groups, admission of waiting reads, heavy lane.  Checked as invariants over hand-made and
random rounds; no device involved."""
import numpy as np
import pytest

BPA, PAD = 144, 1024


def region(n):
    return ((int(n) * BPA + PAD + 255) // 256) * 256


def check(n_anchors, n_mand, max_opt, arena, heavy=True):
    from rawhash_b200 import api
    p = api.plan_round(n_anchors, n_mand, max_opt, arena, heavy)
    na = np.asarray(n_anchors, dtype=np.int64)
    ns = len(na)
    order, a_off, groups, n_run = p["order"], p["a_off"], p["groups"], p["n_run"]
    assert sorted(order.tolist()) == list(range(ns)), "order is a permutation"
    assert sorted(order[:n_mand].tolist()) == list(range(n_mand)), "mandatory chunks stay in front"
    assert order[n_mand:].tolist() == list(range(n_mand, ns)), "waiting reads keep their admission order"
    m = na[order[:n_mand]]
    assert np.all(m[:-1] >= m[1:]), "mandatory chunks heaviest first"
    assert n_mand <= n_run <= min(ns, n_mand + max_opt)
    covered = []
    heavy_groups = [g for g in groups if g[2]]
    assert len(heavy_groups) <= 1 and (not heavy_groups or groups[0][2] == 1)
    main_bytes = p["main_bytes"]
    assert main_bytes % 256 == 0 and main_bytes <= arena
    for first, count, hv in groups:
        assert count > 0
        covered += list(range(first, first + count))
        lo = main_bytes if hv else 0
        hi = arena if hv else main_bytes
        at = lo
        for q in range(first, first + count):   # regions are laid end to end inside the group's slice
            assert int(a_off[q]) == at and at % 256 == 0
            at += region(na[order[q]])
        assert at <= hi, "a group fits its slice of the arena"
    assert covered == list(range(n_run)), "every running slot is in exactly one group, in slot order"
    ordinary = [g for g in groups if not g[2]]
    for first, count, _ in ordinary[:-1]:       # greedy: the next chunk would not have fitted
        nxt = first + count
        used = sum(region(na[order[q]]) for q in range(first, nxt))
        assert used + region(na[order[nxt]]) > main_bytes
    if n_run > n_mand:                           # admitted reads sit in the last ordinary group only
        first, count, _ = ordinary[-1]
        assert first <= max(n_mand - 1, first) and first + count == n_run
        assert all(q >= first for q in range(n_mand, n_run))
    if n_run < min(ns, n_mand + max_opt) and ordinary:   # admission stopped because the next one did not fit
        first, count, _ = ordinary[-1]
        used = sum(region(na[order[q]]) for q in range(first, first + count))
        assert used + region(na[order[n_run]]) > main_bytes
    if heavy_groups:
        first, count, _ = heavy_groups[0]
        assert first == 0 and 2 <= count <= n_mand // 4 and n_mand >= 64
        assert np.all(m[:count] > 1.25 * m.mean())
        assert sum(region(x) for x in m[:count]) <= arena // 8
    return p


def test_everything_fits_one_group():
    p = check([1000, 5000, 3000], 3, 0, 64 << 20)
    assert p["groups"] == [(0, 3, 0)] and p["order"].tolist() == [1, 2, 0]


def test_mandatory_spans_groups_and_waiting_reads_fill_the_last():
    sizes = [40_000] * 30 + [10_000] * 50          # 30 in flight, 50 waiting; the arena holds 12 of the large ones
    arena = 12 * region(40_000) + 4096
    p = check(sizes, 30, 50, arena, heavy=False)
    assert [g[1] for g in p["groups"][:2]] == [12, 12]
    last = p["groups"][-1]
    assert last[0] == 24 and p["n_run"] > 30       # 6 mandatory + as many waiting reads as fit beside them
    assert p["n_run"] - 30 == (arena - 6 * region(40_000)) // region(10_000)


def test_nothing_mandatory_admits_one_arena():
    p = check([20_000] * 100, 0, 100, 10 * region(20_000), heavy=False)
    assert p["groups"] == [(0, 10, 0)] and p["n_run"] == 10
    assert check([20_000] * 100, 0, 3, 10 * region(20_000))["n_run"] == 3       # the in-flight cap binds first
    assert check([20_000] * 5, 5, 0, 10 * region(20_000))["groups"] == [(0, 5, 0)]
    assert check([], 0, 0, 1 << 20)["groups"] == []


def test_heavy_lane_takes_the_largest_chunks_to_the_top_of_the_arena():
    rng = np.random.default_rng(1)
    sizes = np.concatenate([rng.integers(300_000, 600_000, 400), rng.integers(900_000, 1_300_000, 20), rng.integers(350_000, 450_000, 300)])
    arena = 72 << 30
    p = check(sizes, 420, 300, arena)
    assert p["groups"][0][2] == 1 and p["groups"][0][1] >= 20
    assert p["main_bytes"] < arena and int(p["a_off"][0]) == p["main_bytes"]
    q = check(sizes, 420, 300, arena, heavy=False)
    assert all(g[2] == 0 for g in q["groups"]) and q["main_bytes"] == arena & ~255
    assert check(sizes[:40], 40, 0, arena)["groups"][0][2] == 0                 # fewer than 64 chunks in flight: one lane


def test_a_chunk_larger_than_the_arena_is_an_error():
    from rawhash_b200 import api
    with pytest.raises(api.RawHashError, match="anchor arena"):
        api.plan_round([10, 2_000_000, 10], 3, 0, 64 << 20)
    with pytest.raises(api.RawHashError, match="anchor arena"):
        api.plan_round([10, 2_000_000], 1, 1, 64 << 20)


def test_random_rounds_keep_the_invariants():
    rng = np.random.default_rng(7)
    for _ in range(300):
        ns = int(rng.integers(0, 400))
        n_mand = int(rng.integers(0, ns + 1))
        sizes = np.where(rng.random(ns) < 0.1, rng.integers(200_000, 1_500_000, ns), rng.integers(0, 300_000, ns))
        arena = int(rng.integers(region(1_500_000), 40 * region(1_500_000)))
        check(sizes, n_mand, int(rng.integers(0, ns + 1)), arena, heavy=bool(rng.integers(0, 2)))
