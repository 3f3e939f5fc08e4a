"""The work partition of the persistent attention kernel, checked on the CPU through the host-only C-ABI entry
mmpl_attn_plan (attention_tcgen05.cu: plan_schedule / PieceIter / UnitPieces - the same code the kernel runs on the
device): for every schedule (uniform KV split, ranges, hybrid) and a spread of shapes every (query-tile pair, KV tile) must
be covered exactly once, partial pieces must own distinct workspace slots, the merge bookkeeping must find every piece of
a unit, and the CTAs must be balanced the way the schedule promises."""
import collections
import ctypes

import pytest

from mmpl_b200 import _lib

SCHED = {0: "uniform", 1: "ranges", 2: "hybrid"}


def plan(Lq, H, T, ctas, force_split=0, max_pieces=200000):
    lib = _lib.load()
    sched = (ctypes.c_int * 7)()
    pieces = (ctypes.c_int * (9 * max_pieces))()
    n = lib.mmpl_attn_plan(Lq, H, T, ctas, force_split, ctypes.cast(sched, ctypes.c_void_p), ctypes.cast(pieces, ctypes.c_void_p),
                           max_pieces)
    assert n >= 0, f"mmpl_attn_plan failed: {n}"
    rows = [tuple(pieces[9 * i + k] for k in range(9)) for i in range(n)]
    keys = ("schedule", "split", "hg", "u_base", "grid", "slots", "QP")
    return dict(zip(keys, sched)), rows


def check(Lq, H, T, ctas, force_split=0):
    info, rows = plan(Lq, H, T, ctas, force_split)
    QP = (Lq + 255) // 256
    assert info["QP"] == QP
    cover = collections.Counter()
    slots = {}
    per_cta = collections.Counter()
    unit_pieces = collections.defaultdict(list)
    for (cta, head, q_row0, t0, n, whole, slot, np_unit, merge_ok) in rows:
        assert 0 <= cta < info["grid"] and 0 <= head < H and q_row0 % 256 == 0 and 0 <= q_row0 < QP * 256
        assert n >= 1 and 0 <= t0 and t0 + n <= T
        u = head * QP + q_row0 // 256
        for t in range(t0, t0 + n):
            cover[(u, t)] += 1
        per_cta[cta] += n
        assert bool(whole) == (n == T) or info["schedule"] == 0  # uniform split: whole iff split == 1
        if not whole:
            assert 0 <= slot < info["slots"], (slot, info)
            assert slot not in slots, f"workspace slot {slot} used twice"
            slots[slot] = u
            assert merge_ok == 1, f"merge bookkeeping does not find the slot of piece {(cta, u, t0, n)}"
            unit_pieces[u].append((n, np_unit))
    assert len(cover) == QP * H * T and set(cover.values()) == {1}, "a (unit, KV tile) pair is missing or covered twice"
    for u, lst in unit_pieces.items():
        assert sum(n for n, _ in lst) == T and all(npu == len(lst) for _, npu in lst), (u, lst)
    return info, per_cta


@pytest.mark.parametrize("Lq,H,T,ctas", [
    (4680, 12, 37, 148), (4680, 12, 74, 148), (4680, 12, 147, 148), (4680, 12, 256, 148),   # cfg2: L_kv 4680 .. 32760
    (4680, 12, 4, 148),                                                                     # cross-attention
    (4680, 40, 147, 148), (10920, 40, 110, 148), (9360, 40, 183, 148), (3120, 40, 25, 148),  # Wan-14B / MMPL stages
    (1170, 12, 10, 148), (390, 2, 4, 148), (700, 3, 40, 4), (700, 3, 40, 7), (600, 2, 313, 5), (128, 1, 1, 148),
])
def test_cost_model_schedule_covers_everything_once(Lq, H, T, ctas):
    info, per_cta = check(Lq, H, T, ctas)
    if info["schedule"] in (1, 2):  # range / hybrid: every CTA gets the same number of KV tiles (+-1)
        total = ((Lq + 255) // 256) * H * T
        assert max(per_cta.values()) - min(per_cta.get(c, 0) for c in range(info["grid"])) <= 1
        assert sum(per_cta.values()) == total


def test_cfg2_picks_hybrid():
    """The cost model's choices at the cfg2 shapes: the hybrid schedule (148 whole units + 80 ranged)
    for self-attention at every KV length, whole units for cross-attention."""
    for T in (37, 74, 110, 147, 183, 220, 256):
        info, _ = plan(4680, 12, T, 148)
        assert SCHED[info["schedule"]] == "hybrid" and info["u_base"] == 148 and info["grid"] == 148, (T, info)
    info, _ = plan(4680, 12, 4, 148)
    assert SCHED[info["schedule"]] == "uniform" and info["split"] == 1
    info, _ = plan(4680, 40, 147, 148)  # Wan-14B: 760 units = 5 whole rounds + 20 ranged
    assert SCHED[info["schedule"]] == "hybrid" and info["u_base"] == 740


@pytest.mark.parametrize("force", [1, 2, 3, 5, -1, -2, -3, -12, -1000])
@pytest.mark.parametrize("Lq,H,T,ctas", [(700, 3, 40, 4), (4680, 12, 74, 148), (1300, 5, 17, 6), (3120, 8, 110, 148),
                                          (600, 2, 313, 5)])
def test_forced_schedules_cover_everything_once(Lq, H, T, ctas, force):
    check(Lq, H, T, ctas, force)


def test_hybrid_balance():
    """Hybrid: every CTA runs floor(U/G) whole units and an equal share (+-1 tile) of the remaining units' KV tiles."""
    info, per_cta = check(4680, 12, 256, 148, -1000)
    assert SCHED[info["schedule"]] == "hybrid"
    lo, hi = min(per_cta.values()), max(per_cta.values())
    assert lo >= 256 + 80 * 256 // 148 and hi <= 256 + 80 * 256 // 148 + 1, (lo, hi)


def test_bad_arguments():
    lib = _lib.load()
    buf = (ctypes.c_int * 90)()
    assert lib.mmpl_attn_plan(0, 12, 37, 148, 0, None, ctypes.cast(buf, ctypes.c_void_p), 10) == -4
    assert lib.mmpl_attn_plan(4680, 12, 37, 148, 0, None, ctypes.cast(buf, ctypes.c_void_p), 10) == -1  # too small
