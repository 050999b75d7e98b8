// host_tables.h -- host-side construction of the read-only tables (see host_tables.cpp).
#pragma once
#include "slot.cuh"

namespace caaa {
void default_map(CaaaMap* out);                              // the reference's compiled-in board (src/land.cpp, sea.cpp, canal.cpp, player.cpp)
bool valid_map(const CaaaMap& m);                            // indices in range, counts within the table capacities
void build_tables(const CaaaMap& m, CaaaTables* out);        // reference initialize_constants(), src/engine.cpp:165-513
void build_dev_tables(const CaaaMap& m, DevTables* out);     // + per-player team lists (refresh_allies, 716-725)
}  // namespace caaa
