"""Index-math model of the padded-pixel implicit GEMM (tools/halo_igemm_model.py, the round-2 igemm formulation of DESIGN.md §9)
against a direct convolution: the nine taps read one halo tile at constant row offsets, junk rows never reach the output, never-
loaded rows are only read by junk rows."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import halo_igemm_model as hm  # noqa: E402


@pytest.mark.parametrize("N,H,W,Cin,Cout,tile_m", [(2, 8, 8, 64, 16, 128), (1, 32, 32, 80, 8, 128), (3, 5, 7, 16, 4, 128),
                                                    (1, 32, 32, 64, 8, 256), (2, 16, 16, 128, 8, 64), (1, 4, 4, 8, 8, 128)])
def test_padded_pixel_schedule_equals_direct_convolution(N, H, W, Cin, Cout, tile_m):
    rng = np.random.default_rng(0)
    x = rng.standard_normal((N, H, W, Cin))
    w = rng.standard_normal((3, 3, Cout, Cin))
    got = hm.conv3x3_padded_pixel(x, w, tile_m)
    want = torch.nn.functional.conv2d(torch.from_numpy(x).permute(0, 3, 1, 2),
                                      torch.from_numpy(w).permute(2, 3, 0, 1).contiguous(), padding=1).permute(0, 2, 3, 1).numpy()
    assert np.abs(got).max() < 1e6                       # a never-loaded row (1e30 sentinel) reaching an output would show here
    np.testing.assert_allclose(got, want, rtol=1e-10, atol=1e-10)


def test_plan_and_buffer_sizes():
    assert hm.rows_per_tile(128, 32) == 3 and hm.rows_per_tile(256, 32) == 7
    tiles = hm.plan(2, 32, 32, 256)
    assert len(tiles) == 2 * 5 and tiles[4] == (0, 28, 4) and all(R >= 1 for _, _, R in tiles)
    assert sum(R for n, _, R in tiles if n == 0) == 32
    assert hm.halo_rows(128, 32) == 198                  # 25 KB of 128-byte rows per 64-channel chunk
    assert abs(hm.mma_row_efficiency(32, 32, 256) - 0.8) < 1e-9
    new, old = hm.operand_bytes_per_tile(32, 128, 64, 128)
    assert new < 0.6 * old
    with pytest.raises(ValueError):
        hm.plan(1, 4, 200, 128)


@pytest.mark.parametrize("N,H,W,Cin,Cout", [(2, 32, 32, 64, 8), (1, 16, 16, 128, 4), (1, 64, 64, 64, 4), (2, 9, 30, 16, 4), (1, 3, 126, 8, 2)])
def test_cta_pair_schedule_with_tma_store_clipping(N, H, W, Cin, Cout):
    """each CTA of a pair loads its own halo box (row-granular extra offset for CTA 1), the output box clips junk columns / rows"""
    rng = np.random.default_rng(1)
    x = rng.standard_normal((N, H, W, Cin))
    w = rng.standard_normal((3, 3, Cout, Cin))
    got = hm.conv3x3_padded_pixel_pair(x, w)
    want = torch.nn.functional.conv2d(torch.from_numpy(x).permute(0, 3, 1, 2),
                                      torch.from_numpy(w).permute(2, 3, 0, 1).contiguous(), padding=1).permute(0, 2, 3, 1).numpy()
    np.testing.assert_allclose(got, want, rtol=1e-10, atol=1e-10)


def test_pair_halo_box_sizes():
    assert hm.pair_halo_box(0, 32) == (0, 6, 0)                 # CTA 0: 6 padded rows (26 KB per 64-channel chunk)
    skip, rows, shift = hm.pair_halo_box(1, 32)
    assert (skip, shift) == (3, 26) and rows == 7              # CTA 1 starts 26 pixels into padded row 3
    assert rows * 34 * 128 <= 32 * 1024


@pytest.mark.parametrize("N,H,W,Cin,Cout", [(7, 4, 4, 16, 4), (128, 4, 4, 8, 2), (5, 8, 8, 16, 4), (3, 16, 16, 8, 2), (2, 5, 7, 8, 2), (1, 4, 4, 8, 2)])
def test_whole_batch_padded_pixel_schedule_equals_direct_convolution(N, H, W, Cin, Cout):
    """index math of indm_igemm_t.a_pp (igemm_halo_kernel<..., FLAT>): one box per tile of 128 padded pixels of the whole batch,
    nine row offsets, the shared zero row between images is the bottom padding of one image and the top padding of the next"""
    rng = np.random.default_rng(2)
    x = rng.standard_normal((N, H, W, Cin))
    w = rng.standard_normal((3, 3, Cout, Cin))
    got = hm.conv3x3_whole_batch_padded(x, w)
    want = torch.nn.functional.conv2d(torch.from_numpy(x).permute(0, 3, 1, 2),
                                      torch.from_numpy(w).permute(2, 3, 0, 1).contiguous(), padding=1).permute(0, 2, 3, 1).numpy()
    np.testing.assert_allclose(got, want, rtol=1e-10, atol=1e-10)


def test_whole_batch_padded_pixel_sizes():
    assert hm.to_padded_pixels(np.ones((128, 4, 4, 1))).shape[0] == (128 * 5 + 1) * 6
    assert hm.pp_box_rows(4) == 142 and hm.pp_box_rows(8) == 150 and hm.pp_box_rows(61) == 256      # the kernel's limit: 256 box rows
    pp = hm.to_padded_pixels(np.ones((3, 4, 4, 2)))
    assert pp.sum() == 3 * 16 * 2 and pp[:6].sum() == 0 and pp[5 * 6:5 * 6 + 6].sum() == 0
    # useful rows per tile: 16 / 30 at 4x4, 64 / 90 at 8x8, 256 / 306 at 16x16 (why only the 4x4 blocks use it)
    assert abs(16 / (5 * 6) - 0.533) < 1e-3
