// psim_tiled.h -- export-buffer layout shared by the tiled engine and the slab exchange.
#pragma once
#include <cstddef>

namespace psim {

// Byte layout of the exports of ONE tile row (all tiles of the row, section by section), so that a
// whole boundary row is a single contiguous message.  Per tile: 16 ints of list counts
// ([0..7] halo lists N S W E NW NE SW SE, [8] outbox), HL double2 halo entries, CO outbox records
// of 64 bytes (x y vx vy ax ay id).
struct ExportLayout {
    size_t off_cnt, off_hxy, off_obox, row_bytes;
};

}  // namespace psim
