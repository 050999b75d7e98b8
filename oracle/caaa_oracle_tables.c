/* caaa_oracle_tables.c -- TEST INFRASTRUCTURE.  CPU restatement (plain C) of the reference's
 * static game data and of initialize_constants() (reference src/engine.cpp:165-513), including
 * the uint8 wrap-around in its Floyd-Warshall passes.  Pinned against the patched reference's
 * own tables (oracle/_ref) and SURVEY.md Appendix D in tests/test_tables.py.
 * Never linked into the product library.
 */
#include "caaa_oracle.h"
#include <string.h>

/* ---- static map data: src/land.cpp:4-20, src/sea.cpp:5-8, src/canal.cpp:3-6, src/player.cpp:9-13 ---- */
typedef struct { uint8_t owner, value, nsea, nland, sea[4], land[6]; } OLand;
typedef struct { uint8_t nsea, nland, sea[7], land[6]; } OSea;
static const OLand O_LANDS[5] = {
    {4, 10, 2, 0, {0, 1, 0, 0}, {0, 0, 0, 0, 0, 0}}, /* Washington */
    {2, 8, 2, 0, {1, 2, 0, 0}, {0, 0, 0, 0, 0, 0}},  /* London */
    {1, 10, 1, 1, {2, 0, 0, 0}, {3, 0, 0, 0, 0, 0}}, /* Berlin */
    {0, 8, 0, 2, {0, 0, 0, 0}, {2, 4, 0, 0, 0, 0}},  /* Moscow */
    {3, 8, 1, 1, {0, 0, 0, 0}, {3, 0, 0, 0, 0, 0}},  /* Japan */
};
static const OSea O_SEAS[3] = {
    {1, 2, {1, 0, 0, 0, 0, 0, 0}, {0, 4, 0, 0, 0, 0}}, /* Pacific */
    {2, 2, {0, 2, 0, 0, 0, 0, 0}, {0, 1, 0, 0, 0, 0}}, /* Atlantic */
    {1, 2, {1, 0, 0, 0, 0, 0, 0}, {0, 4, 0, 0, 0, 0}}, /* Baltic */
};
static const uint8_t O_CANAL_SEAS[2][2] = {{0, 1}, {0, 2}};
static const uint8_t O_CANAL_LANDS[2][2] = {{0, 1}, {1, 2}};
static const uint8_t O_CAPITAL[5] = {3, 2, 1, 4, 0};
static const uint8_t O_ALLIED[5][5] = {
    {1, 0, 1, 0, 1}, {0, 1, 0, 1, 0}, {1, 0, 1, 0, 1}, {0, 1, 0, 1, 0}, {1, 0, 1, 0, 1}};

void caaa_oracle_build_tables(CaaaTables* t) {
  memset(t, 0, sizeof(*t));
  for (int l = 0; l < 5; l++) {
    t->land_value[l] = O_LANDS[l].value;
    t->original_owner[l] = O_LANDS[l].owner;
    t->l2l_count[l] = O_LANDS[l].nland; /* engine.cpp:205-216 */
    for (int i = 0; i < O_LANDS[l].nland; i++) t->l2l_conn[l][i] = O_LANDS[l].land[i];
    t->l2s_count[l] = O_LANDS[l].nsea;
    for (int i = 0; i < O_LANDS[l].nsea; i++) t->l2s_conn[l][i] = O_LANDS[l].sea[i];
  }
  for (int s = 0; s < 3; s++) { /* engine.cpp:263-274 */
    t->s2s_count[s] = O_SEAS[s].nsea;
    for (int i = 0; i < O_SEAS[s].nsea; i++) t->s2s_conn[s][i] = O_SEAS[s].sea[i];
    t->s2l_count[s] = O_SEAS[s].nland;
    for (int i = 0; i < O_SEAS[s].nland; i++) t->s2l_conn[s][i] = O_SEAS[s].land[i];
  }
  for (int p = 0; p < 5; p++) {
    t->capital[p] = O_CAPITAL[p];
    for (int q = 0; q < 5; q++) t->is_allied[p][q] = O_ALLIED[p][q];
  }
  memcpy(t->canal_lands, O_CANAL_LANDS, 4);

  /* land distances, engine.cpp:179-236: uint8 sums wrap (255+1 -> 0) */
  memset(t->land_dist, 255, sizeof(t->land_dist));
  for (int l = 0; l < 5; l++) t->land_dist[l][l] = 0;
  for (int l = 0; l < 5; l++) {
    for (int i = 0; i < t->l2l_count[l]; i++) {
      int d = t->l2l_conn[l][i];
      t->land_dist[l][d] = 1;
      t->land_dist[d][l] = 1;
    }
    for (int i = 0; i < t->l2s_count[l]; i++) t->land_dist[l][t->l2s_conn[l][i] + 5] = 1;
  }
  for (int k = 0; k < 5; k++)
    for (int s = 0; s < 5; s++)
      for (int d = 0; d < 8; d++) {
        uint8_t nd = (uint8_t)(t->land_dist[s][k] + t->land_dist[k][d]);
        if (nd < t->land_dist[s][d]) t->land_dist[s][d] = nd;
      }

  /* sea distances per canal state, engine.cpp:237-307 */
  memset(t->sea_dist, 255, sizeof(t->sea_dist));
  for (int c = 0; c < 4; c++) {
    for (int s = 0; s < 3; s++) t->sea_dist[c][s][s] = 0;
    for (int s = 0; s < 3; s++)
      for (int i = 0; i < t->s2s_count[s]; i++) {
        int d = t->s2s_conn[s][i];
        t->sea_dist[c][s][d] = 1;
        t->sea_dist[c][d][s] = 1;
      }
    for (int k = 0; k < 2; k++) {
      if ((c & (1 << k)) == 0) continue;
      t->sea_dist[c][O_CANAL_SEAS[k][0]][O_CANAL_SEAS[k][1]] = 1;
      t->sea_dist[c][O_CANAL_SEAS[k][1]][O_CANAL_SEAS[k][0]] = 1;
    }
    for (int k = 0; k < 3; k++)
      for (int s = 0; s < 3; s++)
        for (int d = 0; d < 3; d++) {
          uint8_t nd = (uint8_t)(t->sea_dist[c][s][k] + t->sea_dist[c][k][d]);
          if (nd < t->sea_dist[c][s][d]) t->sea_dist[c][s][d] = nd;
        }
  }

  /* air distances and air adjacency lists, engine.cpp:308-377 (diagonal starts at 255) */
  memset(t->air_dist, 255, sizeof(t->air_dist));
#define AIR_EDGE(a, b)                                  \
  do {                                                  \
    t->air_conn[a][t->air_conn_count[a]++] = (uint8_t)(b); \
    t->air_dist[a][b] = 1;                              \
    t->air_dist[b][a] = 1;                              \
  } while (0)
  for (int l = 0; l < 5; l++) {
    for (int i = 0; i < t->l2l_count[l]; i++) AIR_EDGE(l, t->l2l_conn[l][i]);
    for (int i = 0; i < t->l2s_count[l]; i++) AIR_EDGE(l, t->l2s_conn[l][i] + 5);
  }
  for (int s = 0; s < 3; s++) {
    for (int i = 0; i < t->s2l_count[s]; i++) AIR_EDGE(s + 5, t->s2l_conn[s][i]);
    for (int i = 0; i < t->s2s_count[s]; i++) AIR_EDGE(s + 5, t->s2s_conn[s][i] + 5);
  }
#undef AIR_EDGE
  for (int k = 0; k < 8; k++)
    for (int s = 0; s < 8; s++)
      for (int d = 0; d < 8; d++) {
        uint8_t nd = (uint8_t)(t->air_dist[s][k] + t->air_dist[k][d]);
        if (nd < t->air_dist[s][d]) t->air_dist[s][d] = nd;
      }

  /* two-hop land paths, engine.cpp:41-42,378-407: only element [0][0] starts at 255 */
  t->land_path[0][0] = 255;
  t->land_path_alt[0][0] = 255;
  for (int s = 0; s < 5; s++) {
    int n = t->l2l_count[s];
    for (int i = 0; i < n; i++) {
      for (int pass = 0; pass < 2; pass++) {
        int mid = pass == 0 ? t->l2l_conn[s][i] : t->l2l_conn[s][n - 1 - i];
        uint8_t(*path)[8] = pass == 0 ? t->land_path : t->land_path_alt;
        for (int j = 0; j < t->l2l_count[mid]; j++) {
          int d = t->l2l_conn[mid][j];
          if (t->land_dist[s][d] == 2) path[s][d] = (uint8_t)mid;
        }
        for (int j = 0; j < t->l2s_count[mid]; j++) {
          int d = t->l2s_conn[mid][j] + 5;
          if (t->land_dist[s][d] == 2) path[s][d] = (uint8_t)mid;
        }
      }
    }
  }
  /* two-hop sea paths, engine.cpp:66-70,408-429 */
  t->sea_path[0][0][0] = 255;
  t->sea_path_alt[0][0][0] = 255;
  for (int c = 0; c < 4; c++)
    for (int s = 0; s < 3; s++) {
      int n = t->s2s_count[s];
      for (int i = 0; i < n; i++) {
        int mid = t->s2s_conn[s][i];
        for (int j = 0; j < t->s2s_count[mid]; j++) {
          int d = t->s2s_conn[mid][j];
          if (t->sea_dist[c][s][d] == 2) t->sea_path[c][s][d] = (uint8_t)mid;
        }
        int mid_alt = t->s2s_conn[s][n - 1 - i];
        for (int j = 0; j < t->s2s_count[mid_alt]; j++) {
          int d = t->s2s_conn[mid_alt][j];
          if (t->sea_dist[c][s][d] == 2) t->sea_path_alt[c][s][d] = (uint8_t)mid_alt;
        }
      }
    }

  /* within-N lists, engine.cpp:430-513 */
  for (int s = 0; s < 5; s++) {
    for (int d = 0; d < 5; d++)
      if (s != d && t->land_dist[s][d] <= 2) t->lands_within_2[s][t->lands_within_2_count[s]++] = (uint8_t)d;
    for (int d = 0; d < 3; d++)
      if (t->land_dist[s][d + 5] <= 2) t->load_within_2[s][t->load_within_2_count[s]++] = (uint8_t)d;
  }
  for (int c = 0; c < 4; c++)
    for (int s = 0; s < 3; s++)
      for (int d = 0; d < 3; d++) {
        if (s == d) continue;
        if (t->sea_dist[c][s][d] <= 2) t->seas_within_2[c][s][t->seas_within_2_count[c][s]++] = (uint8_t)d;
        else continue;
        if (t->sea_dist[c][s][d] <= 1) t->seas_within_1[c][s][t->seas_within_1_count[c][s]++] = (uint8_t)d;
      }
  for (int s = 0; s < 8; s++)
    for (int d = 0; d < 8; d++) {
      if (s == d) continue;
      for (int i = 0; i < 5; i++)
        if (t->air_dist[s][d] <= i + 1) t->air_within[i][s][t->air_within_count[i][s]++] = (uint8_t)d;
      if (d >= 5) continue; /* bombers end on land */
      if (t->air_dist[s][d] <= 6) t->air_within[5][s][t->air_within_count[5][s]++] = (uint8_t)d;
    }
  for (int s = 0; s < 8; s++)
    for (int d = 0; d < 5; d++) {
      if (s == d) continue;
      for (int i = 0; i < 6; i++)
        if (t->air_dist[s][d] <= i + 1)
          t->air_to_land_within[i][s][t->air_to_land_within_count[i][s]++] = (uint8_t)d;
    }
}
