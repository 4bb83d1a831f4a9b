// rt_tiles.cuh — which pixels a thread of the render kernels owns (2-D pixel blocks).
//
// A frame is cut into row tiles of `tile_rows` rows; inside a tile, CTA work items (blockIdx.x, 128 threads) are 2-D blocks and the
// 32 lanes of a warp are laid out 8 x 4: 8 threads side by side, 4 rows. A thread owns a SPAN of `ppt` adjacent pixels of one row
// (ppt = 4 on the tiny-scene path — one 128-bit store — and 1 on the heavy paths).
//   render_loop            CTA = 2 x 2 warps = 16 ppt columns x 8 rows; items per tile = (tile_rows / 8) * ceil(w / (16 ppt))
//   k_render_tiny_pack     CTA = 4 x 1 warps = 128 columns x 4 rows (ppt = 4, w % 128 == 0): the CTA owns whole flag bytes of the
//                          packed gather (one per row and 128-pixel group); items per tile = (tile_rows / 4) * (w / 128)
// Needs tile_rows % 8 == 0 (the host checks; other frames keep the linear pixel order). Rows at or below h and columns at or
// beyond w (ragged right edge on the heavy paths) have no pixels. Shared with tests/hostemu, which checks that the spans of all
// threads of all items cover every pixel of the frame exactly once.
#pragma once
#include "rt_math.cuh"

namespace rtb {

RT_HD int tile2d_items_per_tile(int ppt, int w, int tile_rows) { return (tile_rows / 8) * ((w + 16 * ppt - 1) / (16 * ppt)); }

// render_loop: item `block_x` of tile `tile`, thread `tid` (0..127). false: no pixels.
RT_HD bool tile2d_span(int ppt, int w, int h, int tile, int tile_rows, int block_x, int tid, int* x, int* y) {
    const int cols = (w + 16 * ppt - 1) / (16 * ppt);
    const int rg = block_x / cols, cb = block_x - rg * cols;          // CTA-uniform
    const int lane = tid & 31, wp = tid >> 5;
    *x = (cb * 16 + (wp & 1) * 8 + (lane & 7)) * ppt;
    *y = tile * tile_rows + rg * 8 + (wp >> 1) * 4 + (lane >> 3);
    return *x < w && *y < h;
}

// k_render_tiny_pack (ppt = 4, w % 128 == 0): item `chunk` of tile `tile`. *cb = the 128-pixel group, *y_first = first of the item's
// four rows. false: the row is below the frame (the lane still takes part in the warp's ballots and the CTA's barriers).
RT_HD bool pack2d_span(int w, int h, int tile, int tile_rows, int chunk, int tid, int* x, int* y, int* cb, int* y_first) {
    const int cols = w >> 7;
    const int rg = chunk / cols;
    const int lane = tid & 31, wp = tid >> 5;
    *cb = chunk - rg * cols;
    *x = *cb * 128 + wp * 32 + (lane & 7) * 4;
    *y_first = tile * tile_rows + rg * 4;
    *y = *y_first + (lane >> 3);
    return *y < h;
}

}  // namespace rtb
