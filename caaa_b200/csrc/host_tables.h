// host_tables.h -- host-side construction of the read-only tables (see host_tables.cpp).
#pragma once
#include "slot.cuh"

namespace caaa {
void build_tables(CaaaTables* out);      // reference initialize_constants(), src/engine.cpp:165-513
void build_dev_tables(DevTables* out);   // + per-player team lists (refresh_allies, 716-725)
}  // namespace caaa
