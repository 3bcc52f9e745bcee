// textnav_core.h — Text-Nav as an int32 state machine (host + device).
//
// The reference's TextNavEnv (llm_rl_scripts/text_nav/env/env.py:13-72) wraps a TextWorld game compiled from
// env/make_game.py:30-107: 8 rooms, 3 tables, a chest and a fridge (both closed), 7 objects, two quests (eating the stale food
// loses; the stale food inside the CLOSED fridge scores 1 and ends the game).  TextWorld (jxihong/TextWorld @ main, unpinned) is
// not vendored and not installable, and env.py itself cannot run (NameError `true` :71, `tuple(str)` :64), so the world model is
// restated from make_game.py and pinned on what the reference ships of its behaviour: the two played transcripts in
// text_nav/notebooks/play_game.ipynb (admissible-command lists after every command, score, done) — tests/golden/textnav_transcripts.json.
// Room prose is TextWorld-grammar random text and is NOT reproduced (the repo defines its own deterministic sentences, host side).
//
// State per environment: 12 int32 = {room, loc[7], open mask, score, done, moves}.
#pragma once
#include <stdint.h>
#include "env_core.h"

namespace lmrl {

constexpr int TN_STATE = 12;
constexpr int TN_ROOMS = 8, TN_OBJECTS = 7, TN_MAX_ADM = 32;
// rooms, in make_game.py:33-40 creation order
enum { TN_BEDROOM = 0, TN_OFFICE, TN_BATHROOM, TN_LIVING, TN_KITCHEN, TN_DINING, TN_GARDEN, TN_BACKYARD };
// objects (make_game.py:66-72)
enum { TN_STALE = 0, TN_FRESH, TN_BOWL, TN_CUP, TN_PLATE, TN_UTENSILS, TN_FRUIT };
// where an object is: 0..7 floor of that room, then
enum { TN_ON_TABLE0 = 8 /* bedroom table */, TN_ON_TABLE1 /* living room */, TN_ON_TABLE2 /* dining room */, TN_IN_CHEST = 11, TN_IN_FRIDGE = 12,
       TN_INVENTORY = 13, TN_GONE = 14 };
enum { TN_DIR_EAST = 0, TN_DIR_NORTH, TN_DIR_SOUTH, TN_DIR_WEST };      // alphabetical, the order admissible commands list them
// verbs, alphabetical
enum { TN_CLOSE = 0, TN_DROP, TN_EAT, TN_GO, TN_INSERT, TN_OPEN, TN_PUT, TN_TAKE, TN_NOOP /* look / inventory / examine: not listed, no effect */ };
// holders of a `take`: 0 floor, 1 chest, 2 fridge, 3 table (alphabetical: "take x" < "take x from chest" < "... fridge" < "... table")
enum { TN_H_FLOOR = 0, TN_H_CHEST, TN_H_FRIDGE, TN_H_TABLE };

LMRL_HD int tn_code(int verb, int a, int b) { return (verb << 8) | (a << 4) | b; }

// exits[room][dir] (make_game.py:42-49: M.connect(a.x, b.y) joins a's exit x with b's exit y)
LMRL_HD int tn_exit(int room, int dir) {
    switch (room * 4 + dir) {
        case TN_BEDROOM * 4 + TN_DIR_WEST: return TN_LIVING;     case TN_LIVING * 4 + TN_DIR_EAST: return TN_BEDROOM;
        case TN_LIVING * 4 + TN_DIR_WEST: return TN_DINING;      case TN_DINING * 4 + TN_DIR_EAST: return TN_LIVING;
        case TN_DINING * 4 + TN_DIR_WEST: return TN_KITCHEN;     case TN_KITCHEN * 4 + TN_DIR_EAST: return TN_DINING;
        case TN_KITCHEN * 4 + TN_DIR_NORTH: return TN_GARDEN;    case TN_GARDEN * 4 + TN_DIR_SOUTH: return TN_KITCHEN;
        case TN_LIVING * 4 + TN_DIR_SOUTH: return TN_OFFICE;     case TN_OFFICE * 4 + TN_DIR_NORTH: return TN_LIVING;
        case TN_OFFICE * 4 + TN_DIR_EAST: return TN_BATHROOM;    case TN_BATHROOM * 4 + TN_DIR_WEST: return TN_OFFICE;
        case TN_BEDROOM * 4 + TN_DIR_SOUTH: return TN_BATHROOM;  case TN_BATHROOM * 4 + TN_DIR_NORTH: return TN_BEDROOM;
        case TN_KITCHEN * 4 + TN_DIR_WEST: return TN_BACKYARD;   case TN_BACKYARD * 4 + TN_DIR_EAST: return TN_KITCHEN;
        default: return -1;
    }
}
// the table of a room (make_game.py:59-63), or -1;  chest in the office, fridge in the kitchen
LMRL_HD int tn_table_of(int room) { return room == TN_BEDROOM ? 0 : room == TN_LIVING ? 1 : room == TN_DINING ? 2 : -1; }
LMRL_HD int tn_container_room(int c) { return c == 0 ? TN_OFFICE : TN_KITCHEN; }          // c: 0 chest, 1 fridge
LMRL_HD bool tn_is_food(int o) { return o == TN_STALE || o == TN_FRESH || o == TN_FRUIT; }
// objects in alphabetical order of their names: bowl, coffee cup, fresh food, fruit, plate, stale food, utensils
LMRL_HD int tn_sorted_obj(int i) {
    switch (i) { case 0: return TN_BOWL; case 1: return TN_CUP; case 2: return TN_FRESH; case 3: return TN_FRUIT; case 4: return TN_PLATE; case 5: return TN_STALE; default: return TN_UTENSILS; }
}

LMRL_HD void tn_reset(int32_t* s, int start_room) {
    s[0] = start_room;
    s[1 + TN_STALE] = TN_ON_TABLE1; s[1 + TN_FRESH] = TN_ON_TABLE1; s[1 + TN_BOWL] = TN_ON_TABLE0; s[1 + TN_CUP] = TN_ON_TABLE2;
    s[1 + TN_PLATE] = TN_ON_TABLE2; s[1 + TN_UTENSILS] = TN_IN_CHEST; s[1 + TN_FRUIT] = TN_IN_FRIDGE;      // make_game.py:74-80
    s[8] = 0;                    // bit c set = container c is open (both start closed, :56-57)
    s[9] = 0; s[10] = 0; s[11] = 0;
}
// random.choice([bedroom, bathroom, office, livingroom, kitchen, diningroom, garden, backyard]) (make_game.py:82-83)
LMRL_HD int tn_choice_room(int k) {
    switch (k) { case 0: return TN_BEDROOM; case 1: return TN_BATHROOM; case 2: return TN_OFFICE; case 3: return TN_LIVING; case 4: return TN_KITCHEN;
                 case 5: return TN_DINING; case 6: return TN_GARDEN; default: return TN_BACKYARD; }
}

// is the command possible in this state?  (== membership of TextWorld's admissible list, plus the three unlisted no-ops)
LMRL_HD bool tn_possible(const int32_t* s, int verb, int a, int b) {
    const int room = s[0], open = s[8];
    const int table = tn_table_of(room);
    switch (verb) {
        case TN_GO: return tn_exit(room, a) >= 0;
        case TN_OPEN: return tn_container_room(a) == room && !((open >> a) & 1);
        case TN_CLOSE: return tn_container_room(a) == room && ((open >> a) & 1);
        case TN_DROP: return s[1 + a] == TN_INVENTORY;
        case TN_EAT: return s[1 + a] == TN_INVENTORY && tn_is_food(a);
        case TN_PUT: return s[1 + a] == TN_INVENTORY && table >= 0;
        case TN_INSERT: return s[1 + a] == TN_INVENTORY && tn_container_room(b) == room && ((open >> b) & 1);
        case TN_TAKE:
            if (b == TN_H_FLOOR) return s[1 + a] == room;
            if (b == TN_H_TABLE) return table >= 0 && s[1 + a] == TN_ON_TABLE0 + table;
            { const int c = b == TN_H_CHEST ? 0 : 1; return tn_container_room(c) == room && ((open >> c) & 1) && s[1 + a] == TN_IN_CHEST + c; }
        case TN_NOOP: return true;
        default: return false;
    }
}

// admissible commands in TextWorld's (alphabetical) order, as codes; returns the count
LMRL_HD int tn_admissible(const int32_t* s, int32_t* out) {
    int n = 0;
    for (int c = 0; c < 2; ++c) if (tn_possible(s, TN_CLOSE, c, 0)) out[n++] = tn_code(TN_CLOSE, c, 0);
    for (int i = 0; i < TN_OBJECTS; ++i) if (tn_possible(s, TN_DROP, tn_sorted_obj(i), 0)) out[n++] = tn_code(TN_DROP, tn_sorted_obj(i), 0);
    for (int i = 0; i < TN_OBJECTS; ++i) if (tn_possible(s, TN_EAT, tn_sorted_obj(i), 0)) out[n++] = tn_code(TN_EAT, tn_sorted_obj(i), 0);
    for (int d = 0; d < 4; ++d) if (tn_possible(s, TN_GO, d, 0)) out[n++] = tn_code(TN_GO, d, 0);
    for (int i = 0; i < TN_OBJECTS; ++i)
        for (int c = 0; c < 2; ++c) if (tn_possible(s, TN_INSERT, tn_sorted_obj(i), c)) out[n++] = tn_code(TN_INSERT, tn_sorted_obj(i), c);
    for (int c = 0; c < 2; ++c) if (tn_possible(s, TN_OPEN, c, 0)) out[n++] = tn_code(TN_OPEN, c, 0);
    for (int i = 0; i < TN_OBJECTS; ++i) if (tn_possible(s, TN_PUT, tn_sorted_obj(i), 0)) out[n++] = tn_code(TN_PUT, tn_sorted_obj(i), 0);
    for (int i = 0; i < TN_OBJECTS; ++i)
        for (int h = 0; h < 4; ++h) if (tn_possible(s, TN_TAKE, tn_sorted_obj(i), h)) out[n++] = tn_code(TN_TAKE, tn_sorted_obj(i), h);
    return n;
}

// apply a POSSIBLE command; then the quest events (make_game.py:90-100)
LMRL_HD void tn_apply(int32_t* s, int verb, int a, int b) {
    switch (verb) {
        case TN_GO: s[0] = tn_exit(s[0], a); break;
        case TN_OPEN: s[8] |= 1 << a; break;
        case TN_CLOSE: s[8] &= ~(1 << a); break;
        case TN_DROP: s[1 + a] = s[0]; break;
        case TN_EAT: s[1 + a] = TN_GONE; break;
        case TN_PUT: s[1 + a] = TN_ON_TABLE0 + tn_table_of(s[0]); break;
        case TN_INSERT: s[1 + a] = TN_IN_CHEST + b; break;
        case TN_TAKE: s[1 + a] = TN_INVENTORY; break;
        default: break;
    }
    if (s[1 + TN_STALE] == TN_GONE) s[10] = 1;                                                 // fail event: food eaten -> game lost
    if (s[1 + TN_STALE] == TN_IN_FRIDGE && !((s[8] >> 1) & 1)) { s[9] = 1; s[10] = 1; }         // win event: in the closed fridge
}

// ---- command text ----------------------------------------------------------------------------------------------------
LMRL_HD int tn_put_str(uint8_t* out, int n, const char* z) { while (*z) out[n++] = (uint8_t)*z++; return n; }
LMRL_HD const char* tn_obj_name(int o) {
    switch (o) { case TN_STALE: return "stale food"; case TN_FRESH: return "fresh food"; case TN_BOWL: return "bowl"; case TN_CUP: return "coffee cup";
                 case TN_PLATE: return "plate"; case TN_UTENSILS: return "utensils"; default: return "fruit"; }
}
LMRL_HD const char* tn_dir_name(int d) { return d == TN_DIR_EAST ? "east" : d == TN_DIR_NORTH ? "north" : d == TN_DIR_SOUTH ? "south" : "west"; }
LMRL_HD const char* tn_cont_name(int c) { return c == 0 ? "chest" : "fridge"; }
// the command string TextWorld lists for a code (<= 40 bytes)
LMRL_HD int tn_cmd_string(int verb, int a, int b, uint8_t* out) {
    int n = 0;
    switch (verb) {
        case TN_GO: n = tn_put_str(out, n, "go "); n = tn_put_str(out, n, tn_dir_name(a)); break;
        case TN_OPEN: n = tn_put_str(out, n, "open "); n = tn_put_str(out, n, tn_cont_name(a)); break;
        case TN_CLOSE: n = tn_put_str(out, n, "close "); n = tn_put_str(out, n, tn_cont_name(a)); break;
        case TN_DROP: n = tn_put_str(out, n, "drop "); n = tn_put_str(out, n, tn_obj_name(a)); break;
        case TN_EAT: n = tn_put_str(out, n, "eat "); n = tn_put_str(out, n, tn_obj_name(a)); break;
        case TN_PUT: n = tn_put_str(out, n, "put "); n = tn_put_str(out, n, tn_obj_name(a)); n = tn_put_str(out, n, " on table"); break;
        case TN_INSERT: n = tn_put_str(out, n, "insert "); n = tn_put_str(out, n, tn_obj_name(a)); n = tn_put_str(out, n, " into "); n = tn_put_str(out, n, tn_cont_name(b)); break;
        case TN_TAKE:
            n = tn_put_str(out, n, "take "); n = tn_put_str(out, n, tn_obj_name(a));
            if (b == TN_H_CHEST) n = tn_put_str(out, n, " from chest");
            else if (b == TN_H_FRIDGE) n = tn_put_str(out, n, " from fridge");
            else if (b == TN_H_TABLE) n = tn_put_str(out, n, " from table");
            break;
        default: break;
    }
    return n;
}
LMRL_HD bool tn_bytes_eq(const uint8_t* a, int na, const uint8_t* b, int nb) {
    if (na != nb) return false;
    for (int i = 0; i < na; ++i) if (a[i] != b[i]) return false;
    return true;
}
LMRL_HD bool tn_starts(const uint8_t* a, int na, const char* z) {
    int i = 0;
    for (; z[i]; ++i) if (i >= na || a[i] != (uint8_t)z[i]) return false;
    return true;
}

// `command.strip()` (env.py:38) then an exact match against the command strings; returns a code, or -1 when no command has that text
LMRL_HD int tn_parse(const uint8_t* s, int n) {
    while (n > 0 && (s[n - 1] == ' ' || (s[n - 1] >= 9 && s[n - 1] <= 13))) --n;
    while (n > 0 && (s[0] == ' ' || (s[0] >= 9 && s[0] <= 13))) { ++s; --n; }
    if (n <= 0 || n > 48) return -1;
    uint8_t buf[48];
    if (tn_bytes_eq(s, n, (const uint8_t*)"look", 4) || tn_bytes_eq(s, n, (const uint8_t*)"inventory", 9) || tn_starts(s, n, "examine ")) return tn_code(TN_NOOP, 0, 0);
    for (int d = 0; d < 4; ++d) if (tn_bytes_eq(s, n, buf, tn_cmd_string(TN_GO, d, 0, buf))) return tn_code(TN_GO, d, 0);
    for (int c = 0; c < 2; ++c) {
        if (tn_bytes_eq(s, n, buf, tn_cmd_string(TN_OPEN, c, 0, buf))) return tn_code(TN_OPEN, c, 0);
        if (tn_bytes_eq(s, n, buf, tn_cmd_string(TN_CLOSE, c, 0, buf))) return tn_code(TN_CLOSE, c, 0);
    }
    for (int o = 0; o < TN_OBJECTS; ++o) {
        if (tn_bytes_eq(s, n, buf, tn_cmd_string(TN_DROP, o, 0, buf))) return tn_code(TN_DROP, o, 0);
        if (tn_bytes_eq(s, n, buf, tn_cmd_string(TN_EAT, o, 0, buf))) return tn_code(TN_EAT, o, 0);
        if (tn_bytes_eq(s, n, buf, tn_cmd_string(TN_PUT, o, 0, buf))) return tn_code(TN_PUT, o, 0);
        for (int c = 0; c < 2; ++c) if (tn_bytes_eq(s, n, buf, tn_cmd_string(TN_INSERT, o, c, buf))) return tn_code(TN_INSERT, o, c);
        for (int h = 0; h < 4; ++h) if (tn_bytes_eq(s, n, buf, tn_cmd_string(TN_TAKE, o, h, buf))) return tn_code(TN_TAKE, o, h);
    }
    return -1;
}

// one env step.  Returns the command code applied, -1 = unknown text ("You can't see any such thing."), -2 = known command that is
// not possible here (state unchanged either way).  A finished game ignores further commands (-3).
LMRL_HD int tn_step(int32_t* s, const uint8_t* bytes, int n) {
    if (s[10]) return -3;
    s[11] += 1;
    const int code = tn_parse(bytes, n);
    if (code < 0) return -1;
    const int verb = code >> 8, a = (code >> 4) & 15, b = code & 15;
    if (!tn_possible(s, verb, a, b)) return -2;
    tn_apply(s, verb, a, b);
    return code;
}

}  // namespace lmrl
