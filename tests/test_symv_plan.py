"""CPU: the host-side plan of the assembled-S symmetric matvec (vlg_ba_symv_plan, the code build_problem runs) replayed
in numpy with the kernel's semantics -- tiles aligned to the 256-row grid, row sums per fragment, column sums per run
of one strip, fold by the row/column lists -- against a dense product.  Covers a full triangle, a multi-GPU column
block, more CTAs than tiles, and a speed-weighted cut."""
import numpy as np
import pytest

from bundleadjustmentmatlab_b200 import capi

FIRST_STRIP, LAST_STRIP, FIRST_FRAG, LAST_FRAG = 1 << 16, 1 << 17, 1 << 18, 1 << 19


def replay(Np, G, J0, J1, speed=None, seed=0):
    pl = capi.symv_plan(Np, G, J0, J1, speed)
    BLK, SLAB = pl["blk_rows"], pl["slab"]
    rng = np.random.default_rng(seed)
    A = rng.normal(size=(Np, Np)); S = A + A.T
    x = rng.normal(size=Np)
    rowpart = np.full((pl["nfrag"], BLK), np.nan)
    colpart = np.full(pl["nfrag"] * 32 * SLAB, np.nan)
    tiles, tptr = pl["tiles"], pl["tile_ptr"]
    assert tptr[0] == 0 and tptr[-1] == len(tiles) and np.all(np.diff(tptr) >= 0)
    seen = np.zeros((Np, Np // 32), dtype=np.int32)             # every (row, strip) of the column block exactly once
    for g in range(G):
        yacc = np.full(BLK, np.nan); colacc = np.zeros(32)
        for t in range(tptr[g], tptr[g + 1]):
            J, r0, z, w = (int(v) for v in tiles[t])
            rows, frag, sl = z & 0xffff, w & 0xfffff, w >> 20
            assert rows > 0 and r0 >= 32 * J and (r0 // 256) == ((r0 + rows - 1) // 256), "a tile stays inside one 256-row slot"
            if z & FIRST_FRAG:
                yacc[:] = 0.0
            if z & FIRST_STRIP:
                assert np.all(colacc == 0.0)
            c0 = 32 * J
            r = np.arange(r0, r0 + rows)
            seen[r, J] += 1
            blk = S[r0:r0 + rows, c0:c0 + 32]
            xr = np.where(r < c0 + 32, 0.0, x[r])                  # the diagonal 32 x 32 block is used in full, no mirrored part
            yacc[r % BLK] += blk @ x[c0:c0 + 32]
            colacc += blk.T @ xr
            if z & LAST_STRIP:
                colpart[frag * 32 * SLAB + 32 * sl: frag * 32 * SLAB + 32 * sl + 32] = colacc
                colacc = np.zeros(32)
            if z & LAST_FRAG:
                rowpart[frag] = yacc
                yacc = np.full(BLK, np.nan)
    for J in range(J0, J1):
        assert np.all(seen[32 * J:, J] == 1) and np.all(seen[:32 * J, J] == 0)
    y = np.zeros(Np)
    for rb in range(Np // 32):
        b = rb // (BLK // 32)
        acc = np.zeros(32)
        for f in pl["row_list"][pl["row_ptr"][b]:pl["row_ptr"][b + 1]]:
            acc += rowpart[f, (rb % (BLK // 32)) * 32:(rb % (BLK // 32)) * 32 + 32]
        for off in pl["col_list"][pl["col_ptr"][rb]:pl["col_ptr"][rb + 1]]:
            acc += colpart[off:off + 32]
        y[32 * rb:32 * rb + 32] = acc
    ref = np.zeros(Np)
    for J in range(J0, J1):
        c0 = 32 * J
        ref[c0:] += S[c0:, c0:c0 + 32] @ x[c0:c0 + 32]
        ref[c0:c0 + 32] += S[c0 + 32:, c0:c0 + 32].T @ x[c0 + 32:]
    return y, ref, S @ x, pl


@pytest.mark.parametrize("Np,G", [(4608, 37), (2304, 148), (96, 148), (6400, 5)])
def test_full_triangle(Np, G):
    y, ref, full, pl = replay(Np, G, 0, Np // 32)
    assert np.all(np.isfinite(y))
    assert np.abs(y - ref).max() <= 1e-10 * np.abs(ref).max()
    assert np.abs(ref - full).max() <= 1e-10 * np.abs(full).max()      # all strips: the whole symmetric product
    assert pl["nfrag"] <= pl["ncell"] + G


def test_column_block_of_a_rank():
    Np = 4800
    nstrips = Np // 32
    parts = []
    for (J0, J1) in [(0, 44), (44, 97), (97, nstrips)]:
        y, ref, full, _ = replay(Np, 23, J0, J1)
        assert np.abs(y - ref).max() <= 1e-10 * np.abs(full).max()
        parts.append(y)
    assert np.abs(sum(parts) - full).max() <= 1e-10 * np.abs(full).max()  # the ranks' partial products add up


def test_speed_weighted_cut():
    Np, G = 4608, 16
    sp = np.linspace(0.8, 1.25, G)
    y, ref, full, pl = replay(Np, G, 0, Np // 32, speed=sp)
    assert np.abs(y - full).max() <= 1e-10 * np.abs(full).max()
    ntile = np.diff(pl["tile_ptr"]).astype(float)
    assert ntile[-1] > ntile[0]                                            # faster CTAs get longer pieces


def test_banded_matrix_skips_empty_slots():
    """Occupancy map: a banded S (cameras that share points only with their neighbours) keeps the tiles of the band; the
    replayed product over the kept tiles must still be the full product of the banded matrix."""
    Np, G, half = 4608, 37, 300
    nstrips = Np // 32
    rng = np.random.default_rng(3)
    A = rng.normal(size=(Np, Np)); S = A + A.T
    ii, jj = np.indices((Np, Np))
    S[np.abs(ii - jj) > half] = 0.0
    occ = np.zeros((Np // 256, nstrips), dtype=np.uint8)
    nzr, nzc = np.nonzero(np.tril(S))
    occ[nzr // 256, nzc // 32] = 1
    pl = capi.symv_plan(Np, G, 0, nstrips, occ=occ)
    dense = capi.symv_plan(Np, G, 0, nstrips)
    assert len(pl["tiles"]) < 0.35 * len(dense["tiles"])
    # replay with the kernel's semantics (see replay() above), on the banded matrix
    BLK, SLAB = pl["blk_rows"], pl["slab"]
    x = rng.normal(size=Np)
    rowpart = np.zeros((pl["nfrag"], BLK)); colpart = np.zeros(pl["nfrag"] * 32 * SLAB)
    covered = np.zeros((Np // 256, nstrips), dtype=bool)
    tiles, tptr = pl["tiles"], pl["tile_ptr"]
    for g in range(G):
        yacc = np.zeros(BLK); colacc = np.zeros(32)
        for t in range(tptr[g], tptr[g + 1]):
            J, r0, z, w = (int(v) for v in tiles[t])
            rows, frag, sl = z & 0xffff, w & 0xfffff, w >> 20
            covered[r0 // 256, J] = True
            if z & FIRST_FRAG:
                yacc[:] = 0.0
            c0 = 32 * J
            r = np.arange(r0, r0 + rows)
            blk = S[r0:r0 + rows, c0:c0 + 32]
            xr = np.where(r < c0 + 32, 0.0, x[r])
            yacc[r % BLK] += blk @ x[c0:c0 + 32]
            colacc += blk.T @ xr
            if z & LAST_STRIP:
                colpart[frag * 32 * SLAB + 32 * sl: frag * 32 * SLAB + 32 * sl + 32] = colacc
                colacc = np.zeros(32)
            if z & LAST_FRAG:
                rowpart[frag] = yacc
    assert np.array_equal(covered, occ.astype(bool) & (np.arange(Np // 256)[:, None] * 256 + 255 >= np.arange(nstrips)[None, :] * 32))
    y = np.zeros(Np)
    for rb in range(Np // 32):
        b = rb // (BLK // 32)
        acc = np.zeros(32)
        for f in pl["row_list"][pl["row_ptr"][b]:pl["row_ptr"][b + 1]]:
            acc += rowpart[f, (rb % (BLK // 32)) * 32:(rb % (BLK // 32)) * 32 + 32]
        for off in pl["col_list"][pl["col_ptr"][rb]:pl["col_ptr"][rb + 1]]:
            acc += colpart[off:off + 32]
        y[32 * rb:32 * rb + 32] = acc
    full = S @ x
    assert np.abs(y - full).max() <= 1e-10 * np.abs(full).max()
