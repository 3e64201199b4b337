"""Host-side tiling logic (no GPU): the 1-D tile layout keeps the contract documented for
`cztile.AlmostEqualBorderFixedTotalAreaStrategy1D` (constant total tile size, full coverage,
zero border at the image edges, at least the minimum border on inner sides, borders differing by
at most one pixel), and the overlap region equals `calculate_overlap_rle`
(/root/reference/empanada/inference/tile.py:8-52) restated with plain counting."""
import numpy as np
import pytest


@pytest.mark.parametrize("length,total,border", [(5000, 2048, 128), (4096, 2048, 128), (2049, 2048, 128),
                                                 (3000, 1024, 100), (10000, 2048, 204), (1000, 2048, 128),
                                                 (777, 256, 25), (2048, 2048, 128), (6145, 2048, 128)])
def test_fixed_total_area_layout_contract(length, total, border):
    from empanada_napari_b200.tiling import fixed_total_area_tiles_1d
    tiles = fixed_total_area_tiles_1d(length, total, border)
    if total >= length:
        assert tiles == [(0, length)]
        return
    assert all(size == total for _, size in tiles)                       # constant total size
    assert tiles[0][0] == 0 and tiles[-1][0] + total == length           # zero border at the edges
    starts = [s for s, _ in tiles]
    assert starts == sorted(starts)
    overlaps = [starts[i] + total - starts[i + 1] for i in range(len(tiles) - 1)]
    assert min(overlaps) >= 2 * border                                   # >= min border on both inner sides
    assert max(overlaps) - min(overlaps) <= 2                            # borders differ by at most one pixel
    # as few tiles as possible: one tile fewer could not keep the minimum borders
    n = len(tiles)
    assert (n - 1) * total - 2 * (n - 2) * border < length


def test_overlap_region_is_rows_and_columns_covered_twice():
    from empanada_napari_b200.tiling import Tiler
    for shape, ts in [((300, 420), 128), ((260, 200), 96), ((1000, 700), 256)]:
        t = Tiler(shape, tile_size=ts, overlap_width=min(128, int(ts * 0.1)))
        cover = np.zeros(shape, dtype=np.int32)
        for (y0, y1), (x0, x1) in zip(t.yranges, t.xranges):
            cover[y0:y1, x0:x1] += 1
        assert cover.min() >= 1
        rows = np.zeros(shape[0], dtype=np.int32)
        for a, b in set(t.yranges):
            rows[a:b] += 1
        cols = np.zeros(shape[1], dtype=np.int32)
        for a, b in set(t.xranges):
            cols[a:b] += 1
        want = (rows >= 2)[:, None] | (cols >= 2)[None, :]
        assert np.array_equal(t.overlap_mask().astype(bool), want)
        assert np.array_equal(want, cover >= 2)                          # a grid: the same thing
