// host_tables.cpp -- host side of the table upload: static map data and the topology tables the
// kernels read (product code; plain C++, no CUDA).
//
// The tables are computed exactly the way the reference's initialize_constants() does
// (reference src/engine.cpp:165-513) because several of its results only exist through uint8
// wrap-around in the Floyd-Warshall passes (255 + 1 -> 0); the kernels
// consume them as data (< 2 KB, staged once per CTA in shared memory).
#include <cstring>

#include "host_tables.h"
#include "game.cuh"

namespace caaa {
namespace {

// reference src/land.cpp:4-20, src/sea.cpp:5-8, src/canal.cpp:3-6, src/player.cpp:9-13: the default board
struct LandDef { uint8_t owner, value; uint8_t n_sea; uint8_t sea[4]; uint8_t n_land; uint8_t land[6]; };
struct SeaDef { uint8_t n_sea; uint8_t sea[7]; uint8_t n_land; uint8_t land[6]; };
constexpr LandDef kLands[NL] = {
    /* Washington */ {4, 10, 2, {0, 1}, 0, {}},
    /* London     */ {2, 8, 2, {1, 2}, 0, {}},
    /* Berlin     */ {1, 10, 1, {2}, 1, {3}},
    /* Moscow     */ {0, 8, 0, {}, 2, {2, 4}},
    /* Japan      */ {3, 8, 1, {0}, 1, {3}},
};
constexpr SeaDef kSeas[NS] = {
    /* Pacific  */ {1, {1}, 2, {0, 4}},
    /* Atlantic */ {2, {0, 2}, 2, {0, 1}},
    /* Baltic   */ {1, {1}, 2, {0, 4}},
};
constexpr uint8_t kCanalSeas[CAAA_CANALS][2] = {{0, 1}, {0, 2}};   // Panama, Suez
constexpr uint8_t kCanalLands[CAAA_CANALS][2] = {{0, 1}, {1, 2}};
constexpr uint8_t kCapital[NP] = {3, 2, 1, 4, 0};                   // Rus Ger Eng Jap USA
constexpr bool kAxis[NP] = {false, true, false, true, false};       // teams alternate

// all-pairs shortest paths over `rows` sources with the reference's uint8 arithmetic
template <int ROWS, int COLS>
void relax_u8(uint8_t (&d)[ROWS][COLS], int n_mid) {
  for (int k = 0; k < n_mid; k++)
    for (int s = 0; s < ROWS; s++)
      for (int e = 0; e < COLS; e++) {
        const uint8_t via = (uint8_t)(d[s][k] + d[k][e]);
        if (via < d[s][e]) d[s][e] = via;
      }
}

}  // namespace

void default_map(CaaaMap* out) {
  CaaaMap& m = *out;
  std::memset(&m, 0, sizeof(m));
  for (int l = 0; l < NL; l++) {
    const LandDef& L = kLands[l];
    m.land_owner[l] = L.owner;
    m.land_value[l] = L.value;
    m.land_sea_count[l] = L.n_sea;
    m.land_land_count[l] = L.n_land;
    std::memcpy(m.land_sea[l], L.sea, L.n_sea);
    std::memcpy(m.land_land[l], L.land, L.n_land);
  }
  for (int s = 0; s < NS; s++) {
    const SeaDef& S = kSeas[s];
    m.sea_sea_count[s] = S.n_sea;
    m.sea_land_count[s] = S.n_land;
    std::memcpy(m.sea_sea[s], S.sea, S.n_sea);
    std::memcpy(m.sea_land[s], S.land, S.n_land);
  }
  std::memcpy(m.canal_seas, kCanalSeas, sizeof(kCanalSeas));
  std::memcpy(m.canal_lands, kCanalLands, sizeof(kCanalLands));
  for (int p = 0; p < NP; p++) {
    m.capital[p] = kCapital[p];
    for (int q = 0; q < NP; q++) m.is_allied[p][q] = kAxis[p] == kAxis[q];
  }
}

bool valid_map(const CaaaMap& m) {
  for (int l = 0; l < NL; l++) {
    if (m.land_owner[l] >= NP || m.land_sea_count[l] > 4 || m.land_land_count[l] > 6) return false;
    // a land has at most NL - 1 land and NS sea neighbours, an air location at most 7 (air_conn[8][7])
    if (m.land_land_count[l] > NL - 1 || m.land_sea_count[l] > NS || m.land_land_count[l] + m.land_sea_count[l] > 7) return false;
    for (int i = 0; i < m.land_sea_count[l]; i++)
      if (m.land_sea[l][i] >= NS) return false;
    for (int i = 0; i < m.land_land_count[l]; i++)
      if (m.land_land[l][i] >= NL || m.land_land[l][i] == l) return false;
  }
  for (int s = 0; s < NS; s++) {
    if (m.sea_sea_count[s] > NS - 1 || m.sea_land_count[s] > NL || m.sea_sea_count[s] + m.sea_land_count[s] > 7) return false;
    for (int i = 0; i < m.sea_sea_count[s]; i++)
      if (m.sea_sea[s][i] >= NS || m.sea_sea[s][i] == s) return false;
    for (int i = 0; i < m.sea_land_count[s]; i++)
      if (m.sea_land[s][i] >= NL) return false;
  }
  for (int k = 0; k < CAAA_CANALS; k++)
    if (m.canal_seas[k][0] >= NS || m.canal_seas[k][1] >= NS || m.canal_lands[k][0] >= NL || m.canal_lands[k][1] >= NL) return false;
  for (int p = 0; p < NP; p++) {
    if (m.capital[p] >= NL || !m.is_allied[p][p]) return false;
    for (int q = 0; q < NP; q++)
      if (m.is_allied[p][q] > 1 || m.is_allied[p][q] != m.is_allied[q][p]) return false;
    int enemies = 0;
    for (int q = 0; q < NP; q++) enemies += !m.is_allied[p][q];
    if (enemies > 4) return false;
  }
  return true;
}

void build_tables(const CaaaMap& m, CaaaTables* out) {
  CaaaTables& t = *out;
  std::memset(&t, 0, sizeof(t));
  for (int l = 0; l < NL; l++) {
    t.land_value[l] = m.land_value[l];
    t.original_owner[l] = m.land_owner[l];
    t.l2l_count[l] = m.land_land_count[l];
    t.l2s_count[l] = m.land_sea_count[l];
    std::memcpy(t.l2l_conn[l], m.land_land[l], m.land_land_count[l]);
    std::memcpy(t.l2s_conn[l], m.land_sea[l], m.land_sea_count[l]);
  }
  for (int s = 0; s < NS; s++) {
    t.s2s_count[s] = m.sea_sea_count[s];
    t.s2l_count[s] = m.sea_land_count[s];
    std::memcpy(t.s2s_conn[s], m.sea_sea[s], m.sea_sea_count[s]);
    std::memcpy(t.s2l_conn[s], m.sea_land[s], m.sea_land_count[s]);
  }
  for (int p = 0; p < NP; p++) {
    t.capital[p] = m.capital[p];
    for (int q = 0; q < NP; q++) t.is_allied[p][q] = m.is_allied[p][q];
  }
  std::memcpy(t.canal_lands, m.canal_lands, sizeof(m.canal_lands));

  // distances: start at 255, own territory 0, neighbours 1, then relax (engine.cpp:179-377)
  std::memset(t.land_dist, 0xff, sizeof(t.land_dist));
  std::memset(t.sea_dist, 0xff, sizeof(t.sea_dist));
  std::memset(t.air_dist, 0xff, sizeof(t.air_dist));  // the air diagonal is NOT zeroed (engine.cpp:333-339)
  for (int l = 0; l < NL; l++) t.land_dist[l][l] = 0;
  for (int l = 0; l < NL; l++) {
    for (int i = 0; i < t.l2l_count[l]; i++) t.land_dist[l][t.l2l_conn[l][i]] = t.land_dist[t.l2l_conn[l][i]][l] = 1;
    for (int i = 0; i < t.l2s_count[l]; i++) t.land_dist[l][NL + t.l2s_conn[l][i]] = 1;
  }
  relax_u8(t.land_dist, NL);
  for (int c = 0; c < CAAA_CANAL_STATES; c++) {
    for (int s = 0; s < NS; s++) {
      t.sea_dist[c][s][s] = 0;
      for (int i = 0; i < t.s2s_count[s]; i++) t.sea_dist[c][s][t.s2s_conn[s][i]] = t.sea_dist[c][t.s2s_conn[s][i]][s] = 1;
    }
    for (int k = 0; k < CAAA_CANALS; k++)
      if (c & (1 << k)) t.sea_dist[c][m.canal_seas[k][0]][m.canal_seas[k][1]] = t.sea_dist[c][m.canal_seas[k][1]][m.canal_seas[k][0]] = 1;
    relax_u8(t.sea_dist[c], NS);
  }
  auto air_edge = [&t](int a, int b) {
    t.air_conn[a][t.air_conn_count[a]++] = (uint8_t)b;
    t.air_dist[a][b] = t.air_dist[b][a] = 1;
  };
  for (int l = 0; l < NL; l++) {
    for (int i = 0; i < t.l2l_count[l]; i++) air_edge(l, t.l2l_conn[l][i]);
    for (int i = 0; i < t.l2s_count[l]; i++) air_edge(l, NL + t.l2s_conn[l][i]);
  }
  for (int s = 0; s < NS; s++) {
    for (int i = 0; i < t.s2l_count[s]; i++) air_edge(NL + s, t.s2l_conn[s][i]);
    for (int i = 0; i < t.s2s_count[s]; i++) air_edge(NL + s, NL + t.s2s_conn[s][i]);
  }
  relax_u8(t.air_dist, NA);

  // intermediate territory of two-step moves (engine.cpp:41-42,66-70,378-429): zero-filled except
  // element [0][0]; the "alt" table walks the neighbour list backwards
  t.land_path[0][0] = t.land_path_alt[0][0] = 0xff;
  t.sea_path[0][0][0] = t.sea_path_alt[0][0][0] = 0xff;
  for (int s = 0; s < NL; s++)
    for (int i = 0; i < t.l2l_count[s]; i++)
      for (int alt = 0; alt < 2; alt++) {
        const int mid = t.l2l_conn[s][alt ? t.l2l_count[s] - 1 - i : i];
        uint8_t* row = alt ? t.land_path_alt[s] : t.land_path[s];
        for (int j = 0; j < t.l2l_count[mid]; j++)
          if (t.land_dist[s][t.l2l_conn[mid][j]] == 2) row[t.l2l_conn[mid][j]] = (uint8_t)mid;
        for (int j = 0; j < t.l2s_count[mid]; j++)
          if (t.land_dist[s][NL + t.l2s_conn[mid][j]] == 2) row[NL + t.l2s_conn[mid][j]] = (uint8_t)mid;
      }
  for (int c = 0; c < CAAA_CANAL_STATES; c++)
    for (int s = 0; s < NS; s++)
      for (int i = 0; i < t.s2s_count[s]; i++)
        for (int alt = 0; alt < 2; alt++) {
          const int mid = t.s2s_conn[s][alt ? t.s2s_count[s] - 1 - i : i];
          uint8_t* row = alt ? t.sea_path_alt[c][s] : t.sea_path[c][s];
          for (int j = 0; j < t.s2s_count[mid]; j++)
            if (t.sea_dist[c][s][t.s2s_conn[mid][j]] == 2) row[t.s2s_conn[mid][j]] = (uint8_t)mid;
        }

  // reachability lists (engine.cpp:430-513)
  for (int s = 0; s < NL; s++) {
    for (int d = 0; d < NL; d++)
      if (d != s && t.land_dist[s][d] <= 2) t.lands_within_2[s][t.lands_within_2_count[s]++] = (uint8_t)d;
    for (int d = 0; d < NS; d++)
      if (t.land_dist[s][NL + d] <= 2) t.load_within_2[s][t.load_within_2_count[s]++] = (uint8_t)d;
  }
  for (int c = 0; c < CAAA_CANAL_STATES; c++)
    for (int s = 0; s < NS; s++)
      for (int d = 0; d < NS; d++) {
        if (d == s || t.sea_dist[c][s][d] > 2) continue;
        t.seas_within_2[c][s][t.seas_within_2_count[c][s]++] = (uint8_t)d;
        if (t.sea_dist[c][s][d] <= 1) t.seas_within_1[c][s][t.seas_within_1_count[c][s]++] = (uint8_t)d;
      }
  for (int s = 0; s < NA; s++)
    for (int d = 0; d < NA; d++) {
      if (d == s) continue;
      for (int m = 1; m <= 5; m++)
        if (t.air_dist[s][d] <= m) t.air_within[m - 1][s][t.air_within_count[m - 1][s]++] = (uint8_t)d;
      if (d < NL && t.air_dist[s][d] <= 6) t.air_within[5][s][t.air_within_count[5][s]++] = (uint8_t)d;  // bombers end on land
      if (d < NL)
        for (int m = 1; m <= 6; m++)
          if (t.air_dist[s][d] <= m) t.air_to_land_within[m - 1][s][t.air_to_land_within_count[m - 1][s]++] = (uint8_t)d;
    }
}

void build_dev_tables(const CaaaMap& m, DevTables* out) {
  std::memset(out, 0, sizeof(*out));
  build_tables(m, &out->t);
  for (uint32_t n = 1; n < 32; n++) out->recip16[n] = 65536u / n + 1u;
  {  // cell lists of the move phases (reference loop orders, src/engine.cpp:1593-1599, 1734, 1998, 2088-2098, ...)
    using e2::Game;
    auto land_cell = [](int l, int ut, int st) { return (uint32_t)(offsetof(Game, lu) + l * e2::LU_ROW + e2::off_lu(ut) + st); };
    auto sea_cell = [](int sea, int ut, int st) { return (uint32_t)(offsetof(Game, su) + sea * e2::SU_ROW + off_sea(ut) + st); };
    auto pack = [](uint32_t off, int ut, int loc, int st) { return off | ((uint32_t)ut << 10) | ((uint32_t)loc << 14) | ((uint32_t)st << 17); };
    uint32_t* t = out->scan;
    for (int a = 0; a < NA; a++)
      t[SC_FIGHTERS + a] = pack(a < NL ? land_cell(a, FIGHTERS, 4) : sea_cell(a - NL, FIGHTERS, 4), FIGHTERS, a, 4);
    for (int l = 0; l < NL; l++) t[SC_BOMBERS + l] = pack(land_cell(l, BOMBERS_LAND_AIR, 6), BOMBERS_LAND_AIR, l, 6);
    for (int ut = TRANS_EMPTY; ut <= TRANS_1T; ut++)
      for (int sea = 0; sea < NS; sea++) {
        const int st = (int)states_sea(ut) - 1;
        t[SC_STAGE + (ut - TRANS_EMPTY) * NS + sea] = pack(sea_cell(sea, ut, st), ut, sea, st);
      }
    for (int l = 0; l < NL; l++) {
      t[SC_TANKS + l * 2 + 0] = pack(land_cell(l, TANKS, 2), TANKS, l, 2);
      t[SC_TANKS + l * 2 + 1] = pack(land_cell(l, TANKS, 1), TANKS, l, 1);
      t[SC_ARTILLERY + l] = pack(land_cell(l, ARTILLERY, 1), ARTILLERY, l, 1);
      t[SC_INFANTRY + l] = pack(land_cell(l, INFANTRY, 1), INFANTRY, l, 1);
      t[SC_AA_GUNS + l] = pack(land_cell(l, AA_GUNS, 1), AA_GUNS, l, 1);
    }
    for (int ut = TRANS_1I; ut <= TRANS_1I_1T; ut++)
      for (int sea = 0; sea < NS; sea++) {
        for (int i = 1; i <= 2; i++)
          t[SC_TRANSPORTS + ((ut - TRANS_1I) * NS + sea) * 2 + (i - 1)] = pack(sea_cell(sea, ut, 4 - i), ut, sea, 4 - i);
        t[SC_UNLOAD + (ut - TRANS_1I) * NS + sea] = pack(sea_cell(sea, ut, 1), ut, sea, 1);
      }
    for (int ut = DESTROYERS; ut <= BS_DAMAGED; ut++)
      for (int sea = 0; sea < NS; sea++) {
        const int st = (int)unmoved_sea(ut);
        t[SC_SHIPS + (ut - DESTROYERS) * NS + sea] = pack(sea_cell(sea, ut, st), ut, sea, st);
      }
    for (int st = 1; st <= 3; st++)
      for (int a = 0; a < NA; a++)
        t[SC_LAND_FIGHTERS + (st - 1) * NA + a] = pack(a < NL ? land_cell(a, FIGHTERS, st) : sea_cell(a - NL, FIGHTERS, st), FIGHTERS, a, st);
    for (int sea = 0; sea < NS; sea++) t[SC_SUBS + sea] = pack(sea_cell(sea, SUBMARINES, 2), SUBMARINES, sea, 2);
    for (int c1 = 0; c1 < 5; c1++)
      for (int a = 0; a < NA; a++) {
        const int st = a < NL ? c1 + 1 : c1;
        t[SC_LAND_BOMBERS + c1 * NA + a] =
            pack(a < NL ? land_cell(a, BOMBERS_LAND_AIR, st) : sea_cell(a - NL, BOMBERS_SEA, st), a < NL ? (int)BOMBERS_LAND_AIR : (int)BOMBERS_SEA, a, st);
      }
  }
  for (int pi = 0; pi < NP; pi++) {  // refresh_allies (engine.cpp:716-725) for every player to move
    for (int p = 0; p < NP; p++) {
      if (out->t.is_allied[pi][(pi + p) % NP]) out->allied_mask[pi] |= (uint8_t)(1u << p);
      else {
        out->enemies_abs[pi][out->n_enemies[pi]] = (uint8_t)((pi + p) % NP);
        out->enemies[pi][out->n_enemies[pi]++] = (uint8_t)p;
      }
      if (out->t.is_allied[pi][p]) out->allied_abs[pi] |= (uint8_t)(1u << p);
    }
  }
}

}  // namespace caaa
