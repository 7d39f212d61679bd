#!/usr/bin/env python
"""Index-math model (numpy, CPU) of the padded-pixel implicit GEMM planned for round 2: a 3x3 stride-1 pad-1
convolution over NHWC activations where ONE halo tile per K chunk serves all nine taps.

Geometry.  Wp = W + 2.  The padded image P[y', x'] (y' in [0, H + 2), x' in [0, Wp)) is X shifted by (1, 1) with a zero border;
its pixels are numbered q = y' * Wp + x'.  A tile owns R output rows y0 .. y0 + R - 1 of one image and issues MMAs of `tile_m` rows:
GEMM row m <-> output position (y0 + m // Wp, m % Wp); rows with m % Wp >= W or m >= R * Wp are junk (computed, never stored).
The shared-memory halo tile holds padded rows y0 .. y0 + R + 1 (what one TMA box [1][R + 2][Wp][64ch] at coordinates
(y0 - 1, -1) delivers, out-of-bounds zero-filled), i.e. local pixel index l = q - y0 * Wp, and tap (dy, dx) in {-1, 0, 1}^2 reads
the rows  l = (1 + dy) * Wp + (1 + dx) + m : the same tile at a constant row offset — a descriptor start-address offset of
((1 + dy) * Wp + 1 + dx) * 128 bytes on the device.  The largest row touched is 2 * Wp + 2 + tile_m - 1, so the buffer is
`halo_rows(tile_m, W)` rows long; rows past the loaded box are only ever read by junk GEMM rows.

`conv3x3_padded_pixel` executes exactly that schedule with numpy (fp64) and is checked against a direct convolution in
tests/test_halo_model_cpu.py; `plan` is the host-side tile list a launcher would build."""
import numpy as np


def rows_per_tile(tile_m, W):
    return max(tile_m // (W + 2), 0)


def halo_rows(tile_m, W):
    return 2 * (W + 2) + 2 + tile_m


def plan(N, H, W, tile_m=128):
    """[(image, y0, R)]: tiles never span images (TMA zero-fills only at tensor edges); the last tile of an image may be short"""
    R = rows_per_tile(tile_m, W)
    if R < 1:
        raise ValueError(f"image rows of {W} + 2 pixels do not fit a {tile_m}-row tile")
    return [(n, y0, min(R, H - y0)) for n in range(N) for y0 in range(0, H, R)]


def mma_row_efficiency(H, W, tile_m=128):
    """useful GEMM rows / issued GEMM rows over one image"""
    tiles = plan(1, H, W, tile_m)
    return H * W / (len(tiles) * tile_m)


def operand_bytes_per_tile(W, cin, cout_per_cta, tile_m=128, elem=2):
    """(padded-pixel, per-tap) shared-memory fill per tile: A halo once per 64-channel chunk + nine B taps, against nine A boxes +
    nine B taps"""
    chunks = (cin + 63) // 64
    R = rows_per_tile(tile_m, W)
    halo = (R + 2) * (W + 2) * 64 * elem
    b = cout_per_cta * 64 * elem
    return chunks * (halo + 9 * b), chunks * 9 * (tile_m * 64 * elem + b)


def conv3x3_padded_pixel(x, w, tile_m=128):
    """x [N, H, W, Cin], w [3, 3, Cout, Cin] (tap-major like the engine's weight pack) -> y [N, H, W, Cout], following the tile
    schedule above: per tile, per 64-channel chunk, one halo tile; per tap one GEMM of `tile_m` rows at a row offset."""
    N, H, W, Cin = x.shape
    Cout = w.shape[2]
    Wp = W + 2
    y = np.zeros((N, H, W, Cout), dtype=np.float64)
    for (n, y0, R) in plan(N, H, W, tile_m):
        acc = np.zeros((tile_m, Cout), dtype=np.float64)                 # the TMEM accumulator
        for c0 in range(0, Cin, 64):
            c1 = min(c0 + 64, Cin)
            halo = np.full((halo_rows(tile_m, W), c1 - c0), np.nan)      # NaN = never loaded: only junk rows may read it
            box = np.zeros((R + 2, Wp, c1 - c0))                          # TMA box, out-of-bounds zero-filled
            ylo, yhi = max(y0 - 1, 0), min(y0 + R + 1, H)
            box[ylo - (y0 - 1):yhi - (y0 - 1), 1:W + 1] = x[n, ylo:yhi, :, c0:c1]
            halo[:(R + 2) * Wp] = box.reshape(-1, c1 - c0)
            for dy in (-1, 0, 1):
                for dx in (-1, 0, 1):
                    off = (1 + dy) * Wp + (1 + dx)
                    a = halo[off:off + tile_m]                             # the shifted view: a start-address offset on the device
                    acc += np.nan_to_num(a, nan=1e30) @ w[dy + 1, dx + 1, :, c0:c1].T.astype(np.float64)
        for m in range(R * Wp):                                           # epilogue: drop junk rows
            if m % Wp < W:
                y[n, y0 + m // Wp, m % Wp] = acc[m]
    return y


def pair_halo_box(r, W, tile_m_cta=128):
    """CTA pair (cta_group::2, M = 256): CTA r owns GEMM rows [128 r, 128 r + 128) of the 256-row tile.  Its own halo box starts
    `skip` padded rows below the tile's first halo row and its taps carry an extra row offset `shift` (the position of row 128 r
    inside its first padded row); `rows` padded rows cover everything its 128 GEMM rows can touch.
    Returns (skip, rows, shift)."""
    Wp = W + 2
    first = r * tile_m_cta
    skip, shift = first // Wp, first % Wp
    last_touched = shift + 2 * Wp + 2 + tile_m_cta - 1            # largest local row index read by any tap
    rows = last_touched // Wp + 1
    return skip, rows, shift


def conv3x3_padded_pixel_pair(x, w):
    """Same schedule for CTA pairs: a 256-row tile owns R = 256 // Wp image rows; each CTA of the pair loads its OWN halo box
    (`pair_halo_box`) and contributes GEMM rows [128 r, 128 r + 128); the epilogue is the TMA-store view: the tile's rows are the
    box [Cout][Wp][R] at x = 0, whose columns x >= W and rows y >= H the hardware clips."""
    N, H, W, Cin = x.shape
    Cout = w.shape[2]
    Wp = W + 2
    y = np.zeros((N, H, W, Cout), dtype=np.float64)
    for (n, y0, R) in plan(N, H, W, 256):
        Rfull = rows_per_tile(256, W)
        acc = np.zeros((256, Cout), dtype=np.float64)
        for r in (0, 1):
            skip, rows, shift = pair_halo_box(r, W)
            for c0 in range(0, Cin, 64):
                c1 = min(c0 + 64, Cin)
                box = np.zeros((rows, Wp, c1 - c0))                                 # TMA box at (y0 - 1 + skip, -1), zero-filled
                for i in range(rows):
                    yy = y0 - 1 + skip + i
                    if 0 <= yy < H:
                        box[i, 1:W + 1] = x[n, yy, :, c0:c1]
                halo = box.reshape(-1, c1 - c0)
                for dy in (-1, 0, 1):
                    for dx in (-1, 0, 1):
                        off = shift + (1 + dy) * Wp + (1 + dx)
                        assert off + 128 <= halo.shape[0]                            # never reads past the loaded box
                        acc[128 * r:128 * r + 128] += halo[off:off + 128] @ w[dy + 1, dx + 1, :, c0:c1].T.astype(np.float64)
        # TMA store of the box [Cout][Wp][Rfull] at (x = 0, y = y0): out-of-bounds columns / rows are clipped by the hardware
        tile = acc[:Rfull * Wp].reshape(Rfull, Wp, Cout)
        for i in range(Rfull):
            if y0 + i < H:
                y[n, y0 + i] = tile[i, :W]
    return y


if __name__ == "__main__":
    for (H, W, cin, cout) in [(32, 32, 128, 128), (32, 32, 256, 128), (16, 16, 256, 256), (64, 64, 128, 128)]:
        for tile_m, half in ((128, cout), (256, cout // 2)):
            new, old = operand_bytes_per_tile(W, cin, half if tile_m == 256 else cout, 128 if tile_m == 128 else 128)
            print(f"{H}x{W} {cin}->{cout} tile_m={tile_m}: row efficiency {mma_row_efficiency(H, W, tile_m):.2f}, "
                  f"operand bytes per CTA tile {new / 1024:.0f} KB vs {old / 1024:.0f} KB per-tap")


# ---------------------------------------------------------------- whole-batch padded-pixel layout (indm_igemm_t.a_pp, FLAT kernel)
def to_padded_pixels(x):
    """[N,H,W,C] -> [(N (H + 1) + 1)(W + 2), C]: pixel (n, y, x) at row (n (H + 1) + y + 1)(W + 2) + x + 1, every other row zero
    (what indm_gn_apply_pp writes; consecutive images share one zero row)"""
    N, H, W, C = x.shape
    buf = np.zeros((N * (H + 1) + 1, W + 2, C), dtype=x.dtype)
    buf[:N * (H + 1)].reshape(N, H + 1, W + 2, C)[:, 1:, 1:W + 1] = x
    return buf.reshape(-1, C)


def pp_box_rows(W, tile_m=128):
    """rows of the one shared-memory box that serves the nine taps of a tile of `tile_m` padded pixels"""
    return tile_m + 2 * (W + 2) + 2


def conv3x3_whole_batch_padded(x, w, tile_m=128):
    """3x3 'same' convolution of x [N,H,W,Cin] with w [3,3,Cout,Cin] the way igemm_halo_kernel<..., FLAT> schedules it: tile k =
    padded pixels [k tile_m, (k + 1) tile_m) of the WHOLE BATCH; its box starts at row k tile_m - (W + 2) - 1 (rows outside the
    buffer are TMA zero fill); tap (ty, tx) reads box rows m + ty (W + 2) + tx; the epilogue maps padded pixel -> (n, y, x) and
    skips border rows / columns and the tail past the last image."""
    N, H, W, Cin = x.shape
    Cout = w.shape[2]
    Wp = W + 2
    pp = to_padded_pixels(x)
    T = pp.shape[0]
    out = np.full((N, H, W, Cout), 1e30)
    box_rows = pp_box_rows(W, tile_m)
    written = 0
    for k in range((T + tile_m - 1) // tile_m):
        first = k * tile_m - Wp - 1
        box = np.zeros((box_rows, Cin), dtype=x.dtype)
        lo, hi = max(first, 0), min(first + box_rows, T)
        if hi > lo:
            box[lo - first:hi - first] = pp[lo:hi]
        acc = np.zeros((tile_m, Cout))
        for ty in range(3):
            for tx in range(3):
                off = ty * Wp + tx
                assert off + tile_m <= box_rows
                acc += box[off:off + tile_m] @ w[ty, tx].T
        for m in range(tile_m):
            q = k * tile_m + m
            R, col = divmod(q, Wp)
            n, yr = divmod(R, H + 1)
            if yr >= 1 and 1 <= col <= W and n < N:
                out[n, yr - 1, col - 1] = acc[m]
                written += 1
    assert written == N * H * W
    return out
