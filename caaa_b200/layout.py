"""numpy views of the C-ABI's plain-old-data layouts (include/caaa_types.h, include/caaa_gpu.h)."""
import numpy as np

STATE_BYTES = 670

# GameState: reference include/game_state.h:20-68
LAND_DTYPE = np.dtype([("owner_idx", "u1"), ("factory_hp", "i1"), ("factory_max", "u1"), ("bombard_max", "u1"),
                       ("fighters", "u1", (5,)), ("bombers", "u1", (7,)), ("infantry", "u1", (2,)),
                       ("artillery", "u1", (2,)), ("tanks", "u1", (3,)), ("aa_guns", "u1", (2,))])
SEA_DTYPE = np.dtype([("fighters", "u1", (5,)), ("trans_empty", "u1", (4,)), ("trans_1i", "u1", (5,)),
                      ("trans_1a", "u1", (5,)), ("trans_1t", "u1", (5,)), ("trans_2i", "u1", (4,)),
                      ("trans_1i_1a", "u1", (4,)), ("trans_1i_1t", "u1", (4,)), ("submarines", "u1", (3,)),
                      ("destroyers", "u1", (2,)), ("carriers", "u1", (2,)), ("cruisers", "u1", (3,)),
                      ("battleships", "u1", (3,)), ("bs_damaged", "u1", (3,)), ("bombers", "u1", (5,))])
STATE_DTYPE = np.dtype([("player_index", "u1"), ("money", "u1", (5,)), ("builds_left", "u1", (8,)),
                        ("land_state", LAND_DTYPE, (5,)), ("units_sea", SEA_DTYPE, (3,)),
                        ("other_land_units", "u1", (4, 5, 6)), ("other_sea_units", "u1", (4, 3, 14)),
                        ("flagged_for_combat", "u1", (8,)), ("skipped_moves", "u1", (8, 8))])
assert STATE_DTYPE.itemsize == STATE_BYTES

STATS_DTYPE = np.dtype([("result", "<f8"), ("turns", "<i4"), ("decisions", "<i4"), ("draws", "<i4"),
                        ("final_player", "<i4"), ("state_hash", "<u8")])
SUMMARY_DTYPE = np.dtype([("playouts", "<u8"), ("plies", "<u8"), ("decisions", "<u8"), ("draws", "<u8"),
                          ("flagged", "<u8"), ("sum_result", "<f8"), ("kernel_ms", "<f8")])
CACHE_DTYPE = np.dtype([
    ("income_per_turn", "<i4", (6,)), ("total_factory_count", "<i4", (6,)),
    ("factory_locations", "<i4", (6, 5)), ("total_player_land_units", "<i4", (6, 5)),
    ("total_player_sea_units", "<i4", (6, 3)), ("transports_with_large_cargo_space", "<i4", (3,)),
    ("transports_with_small_cargo_space", "<i4", (3,)), ("allied_carriers", "<i4", (3,)),
    ("enemy_units_count", "<i4", (8,)), ("answers_remaining", "<i4"),
    ("current_player_land_unit_types", "u1", (5, 6)), ("current_player_sea_unit_types", "u1", (3, 15)),
    ("enemy_blockade_total", "u1", (3,)), ("enemy_destroyers_total", "u1", (3,)),
    ("is_land_path_blocked", "u1", (25,)), ("is_sea_path_blocked", "u1", (9,)),
    ("is_sub_path_blocked", "u1", (9,)), ("enemies_0", "u1", (4,)), ("is_allied_0", "u1", (5,)),
    ("enemies_count_0", "u1"), ("canal_state", "u1"), ("valid_moves", "u1", (20,)),
    ("valid_moves_count", "u1"), ("pad_", "u1", (1,))])
TABLES_DTYPE = np.dtype([
    ("land_value", "u1", (5,)), ("original_owner", "u1", (5,)), ("capital", "u1", (5,)),
    ("is_allied", "u1", (5, 5)), ("canal_lands", "u1", (2, 2)), ("land_dist", "u1", (5, 8)),
    ("air_dist", "u1", (8, 8)), ("sea_dist", "u1", (4, 3, 3)), ("sea_path", "u1", (4, 3, 3)),
    ("sea_path_alt", "u1", (4, 3, 3)), ("land_path", "u1", (5, 8)), ("land_path_alt", "u1", (5, 8)),
    ("air_conn_count", "u1", (8,)), ("air_conn", "u1", (8, 7)), ("l2l_count", "u1", (5,)),
    ("l2l_conn", "u1", (5, 6)), ("l2s_count", "u1", (5,)), ("l2s_conn", "u1", (5, 4)),
    ("s2s_count", "u1", (3,)), ("s2s_conn", "u1", (3, 7)), ("s2l_count", "u1", (3,)),
    ("s2l_conn", "u1", (3, 6)), ("lands_within_2_count", "u1", (5,)), ("lands_within_2", "u1", (5, 4)),
    ("load_within_2_count", "u1", (5,)), ("load_within_2", "u1", (5, 3)),
    ("seas_within_1_count", "u1", (4, 3)), ("seas_within_1", "u1", (4, 3, 2)),
    ("seas_within_2_count", "u1", (4, 3)), ("seas_within_2", "u1", (4, 3, 2)),
    ("air_within_count", "u1", (6, 8)), ("air_within", "u1", (6, 8, 7)),
    ("air_to_land_within_count", "u1", (6, 8)), ("air_to_land_within", "u1", (6, 8, 5))])


MAP_DTYPE = np.dtype([
    ("land_owner", "u1", (5,)), ("land_value", "u1", (5,)), ("land_sea_count", "u1", (5,)), ("land_land_count", "u1", (5,)),
    ("land_sea", "u1", (5, 4)), ("land_land", "u1", (5, 6)), ("sea_sea_count", "u1", (3,)), ("sea_land_count", "u1", (3,)),
    ("sea_sea", "u1", (3, 7)), ("sea_land", "u1", (3, 6)), ("canal_seas", "u1", (2, 2)), ("canal_lands", "u1", (2, 2)),
    ("capital", "u1", (5,)), ("is_allied", "u1", (5, 5))])


def map_from_json(path):
    """A board description (tests/golden/alt_map.json shows the format) as a CaaaMap record."""
    import json
    with open(path) as f:
        d = json.load(f)
    m = np.zeros((), dtype=MAP_DTYPE)
    for i, land in enumerate(d["lands"]):
        m["land_owner"][i], m["land_value"][i] = land["owner"], land["value"]
        m["land_sea_count"][i], m["land_land_count"][i] = len(land["seas"]), len(land["lands"])
        m["land_sea"][i][:len(land["seas"])] = land["seas"]
        m["land_land"][i][:len(land["lands"])] = land["lands"]
    for i, sea in enumerate(d["seas"]):
        m["sea_sea_count"][i], m["sea_land_count"][i] = len(sea["seas"]), len(sea["lands"])
        m["sea_sea"][i][:len(sea["seas"])] = sea["seas"]
        m["sea_land"][i][:len(sea["lands"])] = sea["lands"]
    for i, c in enumerate(d["canals"]):
        m["canal_seas"][i], m["canal_lands"][i] = c["seas"], c["lands"]
    m["capital"] = d["capitals"]
    axis = np.asarray(d["axis"], dtype=bool)
    m["is_allied"] = (axis[:, None] == axis[None, :]).astype(np.uint8)
    return m


def start_state():
    """The reference's start position (reference game_data.json:1-42) as 670 bytes (uint8)."""
    s = np.zeros((), dtype=STATE_DTYPE)
    s["money"] = [10, 20, 2, 20, 2]
    s["builds_left"] = [0, 0, 0, 8, 0, 0, 0, 0]
    s["land_state"]["owner_idx"] = [4, 2, 1, 0, 3]
    s["land_state"]["factory_hp"] = [10, 8, 10, 8, 8]
    s["land_state"]["factory_max"] = [10, 8, 10, 8, 8]
    return np.frombuffer(s.tobytes(), dtype=np.uint8).copy()
