"""CPU: the host-side work decomposition of the matcher's tensor-core pass (oryon_match_plan, match.cu: build_tc_plan).

The units (pair, 256-row anchor block, query tile) of a batch are distributed over the workers of the persistent kernel -- the
CTAs of the default single-CTA kernel (128-column tiles) or the CTA pairs of the cta_group::2 kernel (ORYON_MATCH_PAIR=1;
256-column tiles, kind | 0x100): whole row blocks round robin for the full waves, the remaining row blocks cut into per-CTA tile quotas.  The
properties checked here are the ones the kernel and the refine pass rely on: every unit is covered exactly once, the
segments of a row block carry consecutive slots below the reported list count, no row block is shared by more than 8 CTAs,
and the per-CTA loads are level (to one unit when the pairs are uniform)."""
import ctypes
import random

import numpy as np
import pytest

from oryon_b200 import _lib

ROWS, COLS, MAX_LISTS = 256, 128, 8


HYBRID, CONTIGUOUS, WHOLE = 0, 1, 2


def plan(n_a, n_q, sm_count=148, kind=HYBRID):
    lib = _lib.load()
    B = len(n_a)
    na = (ctypes.c_int32 * B)(*n_a)
    nq = (ctypes.c_int32 * B)(*n_q)
    begin = (ctypes.c_int32 * (sm_count + 1))()
    info = (ctypes.c_int32 * 3)()
    _lib.check(lib.oryon_match_plan(na, nq, B, sm_count, kind, None, 0, begin, info))
    grid, lists, nseg = info[0], info[1], info[2]
    segs = (ctypes.c_int32 * (5 * max(nseg, 1)))()
    _lib.check(lib.oryon_match_plan(na, nq, B, sm_count, kind, segs, max(nseg, 1), begin, info))
    return grid, lists, np.frombuffer(segs, dtype=np.int32)[:5 * nseg].reshape(nseg, 5).copy(), np.array(begin[:grid + 1])


PAIR = 0x100


def check(n_a, n_q, sm_count=148, kind=HYBRID):
    grid, lists, segs, begin = plan(n_a, n_q, sm_count, kind)
    COLS = 256 if kind & PAIR else 128
    sm_count = sm_count // 2 if kind & PAIR else sm_count
    units = {}
    for b, (a, q) in enumerate(zip(n_a, n_q)):
        if a > 0 and q > 0:
            for rb in range(-(-a // ROWS)):
                units[(b, rb)] = -(-q // COLS)
    total = sum(units.values())
    if total == 0:
        assert grid == 0 and len(segs) == 0
        return grid, lists, []
    assert 1 <= grid <= sm_count and begin[0] == 0 and begin[-1] == len(segs)
    assert np.all(np.diff(begin) > 0), "every launched CTA has work"
    # coverage: in slot order the segments of a row block tile [0, tiles) without gap or overlap
    by_block = {}
    for b, rb, j0, j1, slot in segs.tolist():
        assert (b, rb) in units and 0 <= j0 < j1 <= units[(b, rb)]
        by_block.setdefault((b, rb), []).append((slot, j0, j1))
    assert set(by_block) == set(units)
    for k, lst in by_block.items():
        lst.sort()
        assert [x[0] for x in lst] == list(range(len(lst))), "slots of a row block are 0, 1, 2, ..."
        assert lst[0][1] == 0 and lst[-1][2] == units[k] and all(lst[i][2] == lst[i + 1][1] for i in range(len(lst) - 1))
    assert max(len(v) for v in by_block.values()) == lists <= MAX_LISTS
    per_cta = [int(sum(s[3] - s[2] for s in segs[begin[c]:begin[c + 1]].tolist())) for c in range(grid)]
    assert sum(per_cta) == total
    return grid, lists, per_cta


def test_config2_is_balanced_over_148_sms():
    grid, lists, per_cta = check([19200] * 32, [19200] * 32)
    assert grid == 148 and lists <= MAX_LISTS
    assert max(per_cta) - min(per_cta) <= 1          # 2 400 row blocks: 16 whole per CTA + 32.4 tiles of the last 32
    _, segs_lists, segs, begin = plan([19200] * 32, [19200] * 32)
    first = segs[begin[:-1]]                         # the CTAs start on 148 consecutive row blocks (shared query tiles in L2)
    assert [(int(s[0]), int(s[1])) for s in first] == [(c // 75, c % 75) for c in range(148)]
    grid_w, lists_w, per_cta_w = check([19200] * 32, [19200] * 32, kind=WHOLE)
    assert lists_w == 1 and max(per_cta_w) - min(per_cta_w) == 150   # the first version: a 17th, mostly idle wave
    grid_c, lists_c, per_cta_c = check([19200] * 32, [19200] * 32, kind=CONTIGUOUS)
    assert lists_c == 2 and max(per_cta_c) - min(per_cta_c) <= 1


def test_config5_and_reference_worst_case():
    grid, lists, per_cta = check([76800] * 8, [76800] * 8)
    assert grid == 148 and max(per_cta) - min(per_cta) <= 1
    grid, lists, per_cta = check([5000], [36864])   # reference sizes: 20 row blocks x 288 tiles must still fill the GPU
    assert grid >= 120 and lists <= MAX_LISTS and max(per_cta) <= 2 * (20 * 288 // grid + 1)


def test_small_and_degenerate_batches():
    assert check([1], [1])[:2] == (1, 1)
    assert check([0, 0], [5, 0])[0] == 0
    assert check([300, 0, 7], [0, 40, 129])[0] >= 1
    check([2304], [2304])
    check([257], [100000])                 # two row blocks, 782 tiles each: capped at 8 lists per row block


@pytest.mark.parametrize("kind", [HYBRID, CONTIGUOUS, WHOLE])
@pytest.mark.parametrize("seed", range(6))
def test_ragged_random_batches(seed, kind):
    rng = random.Random(seed)
    B = rng.randint(1, 40)
    n_a = [rng.choice([0, 1, 255, 256, 257, rng.randint(1, 20000)]) for _ in range(B)]
    n_q = [rng.choice([0, 1, 127, 128, 129, rng.randint(1, 40000)]) for _ in range(B)]
    sm = rng.choice([1, 7, 132, 148])
    grid, lists, per_cta = check(n_a, n_q, sm_count=sm, kind=kind)
    if grid and kind == HYBRID:
        tiles_max = max(-(-q // COLS) for a, q in zip(n_a, n_q) if a > 0 and q > 0)
        # level to within one task / one minimum quota of the ideal share
        assert max(per_cta) <= -(-sum(per_cta) // sm) + tiles_max + -(-tiles_max // 7)


def test_pair_kernel_plan_config2_config5_and_reference_shape():
    """The cta_group::2 kernel works in CTA pairs on 256-column tiles: 74 workers on a B200."""
    grid, lists, per = check([19200] * 32, [19200] * 32, kind=HYBRID | PAIR)
    assert grid == 74 and lists <= MAX_LISTS and max(per) - min(per) <= 1        # 2 400 row blocks x 75 tiles over 74 pairs
    grid, lists, per = check([76800] * 8, [76800] * 8, kind=HYBRID | PAIR)
    assert grid == 74 and max(per) - min(per) <= 1
    grid, lists, per = check([5000], [36864], kind=HYBRID | PAIR)                 # 20 row blocks x 144 tiles
    assert grid >= 60 and lists <= MAX_LISTS
    assert check([1], [1], kind=HYBRID | PAIR)[:2] == (1, 1)
    assert check([0, 3], [5, 0], kind=HYBRID | PAIR)[0] == 0


@pytest.mark.parametrize("seed", range(6))
def test_pair_kernel_plan_ragged(seed):
    rng = random.Random(100 + seed)
    B = rng.randint(1, 40)
    n_a = [rng.choice([0, 1, 255, 256, 257, rng.randint(1, 20000)]) for _ in range(B)]
    n_q = [rng.choice([0, 1, 255, 256, 257, rng.randint(1, 40000)]) for _ in range(B)]
    check(n_a, n_q, sm_count=rng.choice([2, 8, 132, 148]), kind=HYBRID | PAIR)
