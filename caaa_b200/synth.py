"""Synthetic workloads for tests and bench.py (pure numpy; no game logic beyond laying out bytes).

* :func:`start_state` -- the reference's game_data.json start position.
* :func:`battle_states` -- BASELINE.json config 3: randomised single-territory land / sea battles.
"""
import numpy as np

from .layout import STATE_BYTES, STATE_DTYPE, start_state  # noqa: F401

# static map data needed to lay out plausible states (reference src/land.cpp:4-20, src/player.cpp:9-13)
ORIGINAL_OWNER = np.array([4, 2, 1, 0, 3], dtype=np.int64)
FACTORY_MAX = np.array([10, 8, 10, 8, 8], dtype=np.int64)
LAND_TO_SEA = [[0, 1], [1, 2], [2], [], [0]]


def battle_states(n, seed=0xBA77, max_count=12):
    """n randomised battle states (uint8 array (n, 670)); even indices are land battles, odd sea battles.

    Every state has the five lands owned by their original owners (all lands keep a factory, as on
    the reference map), a random player to move, and one territory flagged for a fresh battle
    (flagged_for_combat = 2) holding an attacking stack of the player to move and 1-3 defending
    enemy stacks with per-type counts uniform in 0..max_count.  Land battles may carry 0..3 pending
    shore bombardments with matching cruisers / battleships in the "can bombard" state on an
    adjacent sea.  The generator is a fixed function of (n, seed).
    """
    rng = np.random.default_rng(seed)
    st = np.zeros(n, dtype=STATE_DTYPE)
    pi = rng.integers(0, 5, n)
    st["player_index"] = pi
    st["money"] = rng.integers(0, 40, (n, 5))
    owner_rel = (ORIGINAL_OWNER[None, :] - pi[:, None]) % 5  # rotated ownership, relative to the mover
    st["land_state"]["owner_idx"] = owner_rel
    st["land_state"]["factory_max"] = FACTORY_MAX[None, :]
    st["land_state"]["factory_hp"] = FACTORY_MAX[None, :] - rng.integers(0, 6, (n, 5))
    is_land = (np.arange(n) % 2) == 0
    # enemies of the mover in relative numbering: teams alternate, so odd relative offsets ... depends on pi
    rel = np.arange(1, 5)[None, :]
    enemy = ((pi[:, None] + rel) % 5) % 2 != (pi[:, None] % 2)  # (n, 4) for relative players 1..4
    n_def = rng.integers(1, 4, n)

    def counts(shape, p_zero=0.35):
        c = rng.integers(0, max_count + 1, shape)
        return np.where(rng.random(shape) < p_zero, 0, c)

    idx = np.arange(n)
    # ---- land battles ----
    land = rng.integers(0, 5, n)
    # make the battle territory enemy-owned: pick the first enemy relative player as owner
    first_enemy = 1 + np.argmax(enemy, axis=1)
    att = counts((n, 6))
    L = np.nonzero(is_land)[0]
    ls = st["land_state"]
    ls["owner_idx"][L, land[L]] = first_enemy[L]
    ls["infantry"][L, land[L], 0] = att[L, 2]
    ls["artillery"][L, land[L], 0] = att[L, 3]
    ls["tanks"][L, land[L], 0] = att[L, 4]
    f_state = rng.integers(1, 4, n)
    b_state = rng.integers(1, 6, n)
    ls["fighters"][L, land[L], f_state[L]] = att[L, 0] // 2
    ls["bombers"][L, land[L], b_state[L]] = att[L, 1] // 3
    dcount = counts((n, 4, 6))
    rank = np.cumsum(enemy, axis=1)  # 1-based rank among the enemies
    use = enemy & (rank <= n_def[:, None])
    dcount = dcount * use[:, :, None]
    st["other_land_units"][L, :, land[L], :] = dcount[L]
    st["flagged_for_combat"][L, land[L]] = 2
    bombard = rng.integers(0, 4, n)
    for l, seas in enumerate(LAND_TO_SEA):
        if not seas:
            continue
        sel = L[(land[L] == l)]
        if sel.size == 0:
            continue
        sea = np.array(seas)[rng.integers(0, len(seas), sel.size)]
        ls["bombard_max"][sel, l] = bombard[sel]
        st["units_sea"]["cruisers"][sel, sea, 1] = rng.integers(0, 3, sel.size) * (bombard[sel] > 0)
        st["units_sea"]["battleships"][sel, sea, 1] = rng.integers(0, 2, sel.size) * (bombard[sel] > 0)
    # ---- sea battles ----
    S = np.nonzero(~is_land)[0]
    sea = rng.integers(0, 3, n)
    us = st["units_sea"]
    satt = counts((n, 15), p_zero=0.5)
    us["fighters"][S, sea[S], f_state[S]] = satt[S, 0] // 2
    us["trans_empty"][S, sea[S], 0] = satt[S, 1] // 3
    for k, name in ((2, "trans_1i"), (3, "trans_1a"), (4, "trans_1t"), (5, "trans_2i"), (6, "trans_1i_1a"), (7, "trans_1i_1t")):
        us[name][S, sea[S], 1] = satt[S, k] // 4
    us["submarines"][S, sea[S], 0] = satt[S, 8]
    us["destroyers"][S, sea[S], 0] = satt[S, 9] // 2
    us["carriers"][S, sea[S], 0] = satt[S, 10] // 3
    us["cruisers"][S, sea[S], 1] = satt[S, 11] // 2
    us["battleships"][S, sea[S], 1] = satt[S, 12] // 3
    us["bs_damaged"][S, sea[S], 1] = satt[S, 13] // 4
    sdef = counts((n, 4, 14), p_zero=0.5) * use[:, :, None]
    st["other_sea_units"][S, :, sea[S], :] = sdef[S]
    st["flagged_for_combat"][S, 5 + sea[S]] = 2
    del idx
    return np.frombuffer(st.tobytes(), dtype=np.uint8).reshape(n, STATE_BYTES).copy()
