"""CPU: the host-side launch plan of the tcgen05 scans (rnn_tc.cu::plan_tiles through sbr_plan_scan_tiles).

The scans cut a batch into cluster tiles of 8 or 16 rows; a B200 keeps 15 eight-CTA clusters co-resident, the hardware
hands clusters to free slots in launch order, so the launcher orders tiles longest-first, picks the tile height by
makespan, and folds the two shortest adjacent tiles into one 16-row tile when the batch is exactly one tile over."""
import ctypes as C

import numpy as np
import pytest


def plan(lens, B, t_max=200, slots=15, ratio8=0.71):
    from sbr_b200 import _capi
    lib = _capi.load_library()
    rows, n, extra = C.c_int(0), C.c_int(0), C.c_int(0)
    order = (C.c_ubyte * 64)()
    lp = None if lens is None else np.ascontiguousarray(lens, np.int32).ctypes.data_as(C.POINTER(C.c_int32))
    rc = lib.sbr_plan_scan_tiles(lp, B, t_max, slots, C.c_float(ratio8), C.byref(rows), C.byref(n), C.byref(extra), order)
    assert rc == 0
    return rows.value, n.value, extra.value, list(order)[:n.value]


def covered_rows(rows, order, extra, B):
    got = []
    for t in order:
        got += list(range(t * rows, min(B, (t + 1) * rows)))
    if extra >= 0:
        got += list(range(16 * extra, 16 * extra + 16))
    return got


@pytest.mark.parametrize("seed", range(6))
@pytest.mark.parametrize("B", [8, 16, 24, 64, 120, 128, 256])
def test_every_row_is_scheduled_exactly_once_longest_first(B, seed):
    rng = np.random.RandomState(seed)
    lens = np.sort(rng.randint(1, 201, B)) if seed % 2 else rng.randint(1, 201, B)
    rows, n, extra, order = plan(lens, B)
    assert rows in (8, 16)
    assert sorted(covered_rows(rows, order, extra, B)) == list(range(B))
    t_end = [int(lens[t * rows:(t + 1) * rows].max()) for t in order]
    assert t_end == sorted(t_end, reverse=True), "main launch must be ordered longest tile first"
    if extra >= 0:
        assert rows == 8 and B % 16 == 0
        grp = [int(lens[16 * g:16 * g + 16].max()) for g in range(B // 16)]
        assert grp[extra] == min(grp), "the folded 16-row group must be the shortest one"
        assert n + 1 <= 15


def test_fits_in_the_slots_means_plain_8_row_tiles():
    lens = np.full(64, 150, np.int32)
    assert plan(lens, 64)[:3] == (8, 8, -1)


def test_one_tile_over_folds_the_shortest_group_only_when_it_pays():
    # the 16th tile is short: it simply queues behind the first cluster that finishes -> plain 8-row tiles
    lens = np.concatenate([np.full(16, 20), np.full(112, 180)]).astype(np.int32)
    assert plan(lens, 128)[:3] == (8, 16, -1)
    # all tiles long: queueing would cost a second round -> the shortest group becomes one 16-row tile
    lens = np.concatenate([np.full(16, 170), np.full(112, 180)]).astype(np.int32)
    rows, n, extra, order = plan(lens, 128)
    assert (rows, n, extra) == (8, 14, 0)
    assert 0 not in order and 1 not in order


def test_uniform_long_batch_prefers_16_rows_or_mixed_over_queueing():
    """All rows equally long: 16 eight-row tiles on 15 slots would take two rounds (2 x 0.71 > 1)."""
    lens = np.full(128, 200, np.int32)
    rows, n, extra, _ = plan(lens, 128)
    assert (rows == 16 and extra == -1) or (rows == 8 and extra >= 0)


def test_ragged_batch_and_missing_lengths_use_the_static_rule():
    assert plan(np.arange(1, 14), 13)[0] == 16          # B % 8 != 0
    assert plan(None, 64)[:3] == (8, 8, -1)
    assert plan(None, 128)[:3] == (16, 8, -1)           # 16 tiles > 15 slots without lengths


def test_argument_errors():
    from sbr_b200 import _capi
    lib = _capi.load_library()
    assert lib.sbr_plan_scan_tiles(None, 0, 1, 15, C.c_float(0.7), None, None, None, None) < 0
