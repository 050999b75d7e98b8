// host_json.cpp -- game_data.json <-> GameState in the reference's wire format, without cJSON.
//
// Wire format: reference src/serialize_data.cpp:75-256 (read) and 258-404 (write).  Keys:
// player_index, money[5], builds_left[8], land_state[5]{owner_index, factory_hp, factory_max,
// fighters, bombers, infantry, artillery, tanks, aa_guns}, units_sea[3]{fighters, tran_empty, tran_1i,
// tran_1a, tran_1t, tran_2i, tran_1i_1a, tran_1i_1t, subs, destroyers, carriers, cruisers,
// battleship, damaged_bs, bombers}, flagged_for_combat[8], other_land_units[4][5][6],
// other_sea_units[4][3][15].  The last one is read and written with 15 entries per 14-entry row
// (src/serialize_data.cpp:198,313): entry 14 aliases the first byte of the next row, which is
// reproduced here through a flat index into the state.
#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "../../include/caaa_host.h"

namespace {

struct JVal {
  enum Kind { NUL, NUM, STR, ARR, OBJ } kind = NUL;
  double num = 0;
  std::string str;
  std::vector<JVal> arr;
  std::vector<std::pair<std::string, JVal>> obj;
  const JVal* get(const char* key) const {
    if (kind != OBJ) return nullptr;
    for (auto& kv : obj)
      if (kv.first == key) return &kv.second;
    return nullptr;
  }
  int as_int() const { return (int)num; }
};

struct Parser {
  const char* p;
  bool ok = true;
  explicit Parser(const char* s) : p(s) {}
  void ws() { while (*p && std::isspace((unsigned char)*p)) p++; }
  bool eat(char c) { ws(); if (*p == c) { p++; return true; } return false; }
  JVal value() {
    ws();
    JVal v;
    if (*p == '{') {
      p++;
      v.kind = JVal::OBJ;
      if (eat('}')) return v;
      do {
        ws();
        JVal k = string();
        if (!ok || !eat(':')) { ok = false; return v; }
        v.obj.emplace_back(k.str, value());
      } while (ok && eat(','));
      if (!eat('}')) ok = false;
    } else if (*p == '[') {
      p++;
      v.kind = JVal::ARR;
      if (eat(']')) return v;
      do v.arr.push_back(value()); while (ok && eat(','));
      if (!eat(']')) ok = false;
    } else if (*p == '"') {
      v = string();
    } else if (*p == '-' || std::isdigit((unsigned char)*p)) {
      char* end = nullptr;
      v.kind = JVal::NUM;
      v.num = std::strtod(p, &end);
      if (end == p) ok = false;
      p = end;
    } else if (!std::strncmp(p, "true", 4)) { v.kind = JVal::NUM; v.num = 1; p += 4; }
    else if (!std::strncmp(p, "false", 5)) { v.kind = JVal::NUM; v.num = 0; p += 5; }
    else if (!std::strncmp(p, "null", 4)) { p += 4; }
    else ok = false;
    return v;
  }
  JVal string() {
    JVal v;
    v.kind = JVal::STR;
    if (*p != '"') { ok = false; return v; }
    p++;
    while (*p && *p != '"') {
      if (*p == '\\' && p[1]) p++;
      v.str.push_back(*p++);
    }
    if (*p == '"') p++; else ok = false;
    return v;
  }
};

const char* kLandKeys[6] = {"fighters", "bombers", "infantry", "artillery", "tanks", "aa_guns"};
const int kLandOff[6] = {4, 9, 16, 18, 20, 23};
const int kLandStates[6] = {5, 7, 2, 2, 3, 2};
const char* kSeaKeys[15] = {"fighters", "tran_empty", "tran_1i", "tran_1a", "tran_1t", "tran_2i", "tran_1i_1a", "tran_1i_1t",
                            "subs", "destroyers", "carriers", "cruisers", "battleship", "damaged_bs", "bombers"};
const int kSeaOff[15] = {0, 5, 9, 14, 19, 24, 28, 32, 36, 39, 41, 43, 46, 49, 52};
const int kSeaStates[15] = {5, 4, 5, 5, 5, 4, 4, 4, 3, 2, 2, 3, 3, 3, 5};

void read_u8_array(const JVal* a, uint8_t* dst, int cap) {
  if (!a || a->kind != JVal::ARR) return;
  for (size_t i = 0; i < a->arr.size() && (int)i < cap; i++)
    if (a->arr[i].kind == JVal::NUM) dst[i] = (uint8_t)a->arr[i].as_int();
}

int from_dom(const JVal& j, CaaaGameState* st) {
  if (j.kind != JVal::OBJ) return -1;
  if (const JVal* v = j.get("player_index")) st->player_index = (uint8_t)v->as_int();
  read_u8_array(j.get("money"), st->money, CAAA_PLAYERS);
  read_u8_array(j.get("builds_left"), st->builds_left, CAAA_AIRS);
  if (const JVal* ls = j.get("land_state"))
    for (size_t i = 0; i < ls->arr.size() && i < (size_t)CAAA_LANDS; i++) {
      const JVal& o = ls->arr[i];
      CaaaLandState* l = &st->land_state[i];
      if (const JVal* v = o.get("owner_index")) l->owner_idx = (uint8_t)v->as_int();
      if (const JVal* v = o.get("factory_hp")) l->factory_hp = (int8_t)v->as_int();
      if (const JVal* v = o.get("factory_max")) l->factory_max = (uint8_t)v->as_int();
      for (int u = 0; u < 6; u++) read_u8_array(o.get(kLandKeys[u]), (uint8_t*)l + kLandOff[u], kLandStates[u]);
    }
  if (const JVal* us = j.get("units_sea"))
    for (size_t i = 0; i < us->arr.size() && i < (size_t)CAAA_SEAS; i++)
      for (int u = 0; u < 15; u++) read_u8_array(us->arr[i].get(kSeaKeys[u]), (uint8_t*)&st->units_sea[i] + kSeaOff[u], kSeaStates[u]);
  read_u8_array(j.get("flagged_for_combat"), st->flagged_for_combat, CAAA_AIRS);
  if (const JVal* a = j.get("other_land_units"))
    for (size_t p = 0; p < a->arr.size() && p < 4; p++)
      for (size_t l = 0; l < a->arr[p].arr.size() && l < (size_t)CAAA_LANDS; l++)
        read_u8_array(&a->arr[p].arr[l], st->other_land_units[p][l], CAAA_LAND_UNIT_TYPES);
  if (const JVal* a = j.get("other_sea_units")) {
    uint8_t* flat = &st->other_sea_units[0][0][0];
    const int limit = (int)((uint8_t*)st + sizeof(CaaaGameState) - flat);
    for (size_t p = 0; p < a->arr.size() && p < 4; p++)
      for (size_t s = 0; s < a->arr[p].arr.size() && s < (size_t)CAAA_SEAS; s++) {
        const JVal& row = a->arr[p].arr[s];
        for (size_t k = 0; k < row.arr.size() && k < 15; k++) {  // 15 entries into 14-entry rows, like the reference
          const int idx = (int)((p * 3 + s) * 14 + k);
          if (idx < limit && row.arr[k].kind == JVal::NUM) flat[idx] = (uint8_t)row.arr[k].as_int();
        }
      }
  }
  return 0;
}

void put_array(std::string& o, const char* key, const uint8_t* a, int n, const char* indent, bool comma = true) {
  o += indent;
  o += "\"";
  o += key;
  o += "\":\t[";
  for (int i = 0; i < n; i++) {
    o += std::to_string((int)a[i]);
    if (i + 1 < n) o += ", ";
  }
  o += comma ? "],\n" : "]\n";
}

std::string to_text(const CaaaGameState* st) {
  std::string o = "{\n";
  o += "\t\"player_index\":\t" + std::to_string((int)st->player_index) + ",\n";
  put_array(o, "money", st->money, CAAA_PLAYERS, "\t");
  put_array(o, "builds_left", st->builds_left, CAAA_AIRS, "\t");
  o += "\t\"land_state\":\t[";
  for (int l = 0; l < CAAA_LANDS; l++) {
    const CaaaLandState* s = &st->land_state[l];
    o += "{\n";
    o += "\t\t\t\"owner_index\":\t" + std::to_string((int)s->owner_idx) + ",\n";
    o += "\t\t\t\"factory_hp\":\t" + std::to_string((int)s->factory_hp) + ",\n";
    o += "\t\t\t\"factory_max\":\t" + std::to_string((int)s->factory_max) + ",\n";
    for (int u = 0; u < 6; u++) put_array(o, kLandKeys[u], (const uint8_t*)s + kLandOff[u], kLandStates[u], "\t\t\t", u < 5);
    o += l + 1 < CAAA_LANDS ? "\t\t}, " : "\t\t}";
  }
  o += "],\n\t\"units_sea\":\t[";
  for (int s = 0; s < CAAA_SEAS; s++) {
    o += "{\n";
    for (int u = 0; u < 15; u++) put_array(o, kSeaKeys[u], (const uint8_t*)&st->units_sea[s] + kSeaOff[u], kSeaStates[u], "\t\t\t", u < 14);
    o += s + 1 < CAAA_SEAS ? "\t\t}, " : "\t\t}";
  }
  o += "],\n";
  put_array(o, "flagged_for_combat", st->flagged_for_combat, CAAA_AIRS, "\t");
  o += "\t\"other_land_units\":\t[";
  for (int p = 0; p < 4; p++) {
    o += "[";
    for (int l = 0; l < CAAA_LANDS; l++) {
      o += "[";
      for (int u = 0; u < 6; u++) o += std::to_string((int)st->other_land_units[p][l][u]) + (u < 5 ? ", " : "");
      o += l + 1 < CAAA_LANDS ? "], " : "]";
    }
    o += p < 3 ? "], " : "]";
  }
  o += "],\n\t\"other_sea_units\":\t[";
  const uint8_t* flat = &st->other_sea_units[0][0][0];
  for (int p = 0; p < 4; p++) {
    o += "[";
    for (int s = 0; s < CAAA_SEAS; s++) {
      o += "[";
      for (int k = 0; k < 15; k++) o += std::to_string((int)flat[(p * 3 + s) * 14 + k]) + (k < 14 ? ", " : "");  // 15 per row
      o += s + 1 < CAAA_SEAS ? "], " : "]";
    }
    o += p < 3 ? "], " : "]";
  }
  o += "]\n}";
  return o;
}

}  // namespace

extern "C" {

int caaa_state_from_json_text(const char* text, CaaaGameState* state) {
  if (!text || !state) return -3;
  Parser ps(text);
  JVal root = ps.value();
  if (!ps.ok) return -5;
  return from_dom(root, state);
}

int caaa_state_from_json_file(const char* filename, CaaaGameState* state) {
  if (!filename || !state) return -3;
  FILE* f = std::fopen(filename, "rb");
  if (!f) return -6;  // the reference calls exit(1) here (src/serialize_data.cpp:38-41); a library reports instead
  std::string text;
  char buf[4096];
  size_t n;
  while ((n = std::fread(buf, 1, sizeof(buf), f)) > 0) text.append(buf, n);
  std::fclose(f);
  return caaa_state_from_json_text(text.c_str(), state);
}

int caaa_state_to_json_text(const CaaaGameState* state, char* out, int cap) {
  if (!state) return -3;
  const std::string t = to_text(state);
  if (out && cap > 0) {
    const int n = (int)t.size() < cap - 1 ? (int)t.size() : cap - 1;
    std::memcpy(out, t.data(), (size_t)n);
    out[n] = 0;
  }
  return (int)t.size() + 1;
}

int caaa_state_to_json_file(const char* filename, const CaaaGameState* state) {
  if (!filename || !state) return -3;
  FILE* f = std::fopen(filename, "wb");
  if (!f) return -6;
  const std::string t = to_text(state);
  std::fwrite(t.data(), 1, t.size(), f);
  std::fclose(f);
  return 0;
}

}  // extern "C"
