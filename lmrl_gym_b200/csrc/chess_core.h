// chess_core.h — host-side chess rules for the Chess-endgames environment (BASELINE.json configs[3]).
//
// The reference env (llm_rl_scripts/chess/env/env.py:29-186) leans on python-chess (unpinned, requirements.txt:34; not installable
// here) for everything this file restates from the published rules of chess and python-chess's documented behaviour:
//   board.fen()                 en-passant square only when a LEGAL en-passant capture exists, cleaned castling rights
//   board.push_san / parse_san  SAN grammar ^([NBKRQ])?([a-h])?([1-8])?[-x]?([a-h][1-8])(=?[nbrqkNBRQK])?[+#]?$ , castling O-O / 0-0,
//                               null move "--"; a pawn move without a file may not capture; ambiguity and illegality are errors
//   board.san(move)             disambiguation by file, else rank, else both; "x", "=Q", "+", "#"
//   board.is_game_over()        checkmate, stalemate, insufficient material, seventy-five moves, fivefold repetition
// Known-answer anchor: perft counts of the standard test positions (tests/test_chess_host.py).  Host code: the environment is
// host-bound by construction (its opponent is a UCI engine process); the GPU runs the policy.
#pragma once
#include <stdint.h>
#include <string.h>
#include <string>
#include <vector>

namespace lmrl_chess_ns {

enum { EMPTY = 0, PAWN = 1, KNIGHT, BISHOP, ROOK, QUEEN, KING };
inline int make_piece(int color, int type) { return type | (color ? 8 : 0); }      // color 0 = white, 1 = black
inline int ptype(int p) { return p & 7; }
inline int pcolor(int p) { return (p >> 3) & 1; }
inline int file_of(int s) { return s & 7; }
inline int rank_of(int s) { return s >> 3; }

struct Move {
    int8_t from = 0, to = 0, promo = 0;      // promo: 0 or KNIGHT..QUEEN
    bool is_null() const { return from == 0 && to == 0 && promo == 0; }
    bool operator==(const Move& o) const { return from == o.from && to == o.to && promo == o.promo; }
};

struct Undo { Move m; int8_t captured, castling, ep; int halfmove; int8_t moved; };

struct Board {
    int8_t sq[64];
    int turn = 0;              // 0 white to move
    int castling = 0;          // bit 0 K, 1 Q, 2 k, 3 q
    int ep = -1;               // square behind a pawn that just advanced two ranks, else -1
    int halfmove = 0, fullmove = 1;
    std::vector<Undo> stack;
    std::vector<uint64_t> keys;            // repetition keys of every position reached (including the current one)

    // ---- attacks -----------------------------------------------------------------------------------------------------
    bool attacked(int s, int by) const {
        const int f = file_of(s), r = rank_of(s);
        const int pr = by == 0 ? r - 1 : r + 1;                       // pawns of `by` that attack s stand one rank behind (from their side)
        if (pr >= 0 && pr < 8) {
            if (f > 0 && sq[pr * 8 + f - 1] == make_piece(by, PAWN)) return true;
            if (f < 7 && sq[pr * 8 + f + 1] == make_piece(by, PAWN)) return true;
        }
        static const int kn[8][2] = {{1, 2}, {2, 1}, {2, -1}, {1, -2}, {-1, -2}, {-2, -1}, {-2, 1}, {-1, 2}};
        for (auto& d : kn) {
            const int ff = f + d[0], rr = r + d[1];
            if (ff >= 0 && ff < 8 && rr >= 0 && rr < 8 && sq[rr * 8 + ff] == make_piece(by, KNIGHT)) return true;
        }
        for (int df = -1; df <= 1; ++df)
            for (int dr = -1; dr <= 1; ++dr) {
                if (!df && !dr) continue;
                int ff = f + df, rr = r + dr, dist = 1;
                while (ff >= 0 && ff < 8 && rr >= 0 && rr < 8) {
                    const int p = sq[rr * 8 + ff];
                    if (p) {
                        if (pcolor(p) == by) {
                            const int t = ptype(p);
                            const bool diag = df != 0 && dr != 0;
                            if (t == QUEEN || (diag && t == BISHOP) || (!diag && t == ROOK) || (dist == 1 && t == KING)) return true;
                        }
                        break;
                    }
                    ff += df; rr += dr; ++dist;
                }
            }
        return false;
    }
    int king_sq(int color) const {
        for (int s = 0; s < 64; ++s) if (sq[s] == make_piece(color, KING)) return s;
        return -1;
    }
    bool in_check(int color) const { const int k = king_sq(color); return k >= 0 && attacked(k, color ^ 1); }

    // ---- make / unmake -----------------------------------------------------------------------------------------------
    void push(const Move& m) {
        Undo u{m, 0, (int8_t)castling, (int8_t)ep, halfmove, 0};
        if (!m.is_null()) {
            const int p = sq[m.from], t = ptype(p), c = pcolor(p);
            u.moved = (int8_t)p;
            u.captured = sq[m.to];
            int new_ep = -1;
            ++halfmove;
            if (t == PAWN || sq[m.to]) halfmove = 0;
            if (t == PAWN && m.to == ep && ep >= 0 && file_of(m.from) != file_of(m.to) && !sq[m.to]) {      // en passant
                const int cap = m.to + (c == 0 ? -8 : 8);
                u.captured = sq[cap];
                sq[cap] = EMPTY;
            }
            if (t == PAWN && (m.to - m.from == 16 || m.from - m.to == 16)) new_ep = (m.from + m.to) / 2;
            sq[m.to] = m.promo ? (int8_t)make_piece(c, m.promo) : (int8_t)p;
            sq[m.from] = EMPTY;
            if (t == KING && (m.to - m.from == 2 || m.from - m.to == 2)) {                                  // castling: move the rook too
                const int r = rank_of(m.from) * 8;
                if (m.to > m.from) { sq[r + 5] = sq[r + 7]; sq[r + 7] = EMPTY; }
                else { sq[r + 3] = sq[r + 0]; sq[r + 0] = EMPTY; }
            }
            auto touch = [&](int s) {
                if (s == 4) castling &= ~3; else if (s == 60) castling &= ~12;
                else if (s == 7) castling &= ~1; else if (s == 0) castling &= ~2; else if (s == 63) castling &= ~4; else if (s == 56) castling &= ~8;
            };
            touch(m.from); touch(m.to);
            ep = new_ep;
        } else {
            ep = -1;
            ++halfmove;
        }
        if (turn == 1) ++fullmove;
        turn ^= 1;
        stack.push_back(u);
        keys.push_back(key());
    }
    void pop() {
        const Undo u = stack.back();
        stack.pop_back();
        keys.pop_back();
        turn ^= 1;
        if (turn == 1) --fullmove;
        const Move& m = u.m;
        if (!m.is_null()) {
            const int c = pcolor(u.moved), t = ptype(u.moved);
            sq[m.from] = u.moved;
            sq[m.to] = EMPTY;
            if (t == PAWN && m.to == u.ep && u.ep >= 0 && file_of(m.from) != file_of(m.to)) sq[m.to + (c == 0 ? -8 : 8)] = u.captured;   // en passant
            else sq[m.to] = u.captured;
            if (t == KING && (m.to - m.from == 2 || m.from - m.to == 2)) {
                const int r = rank_of(m.from) * 8;
                if (m.to > m.from) { sq[r + 7] = sq[r + 5]; sq[r + 5] = EMPTY; }
                else { sq[r + 0] = sq[r + 3]; sq[r + 3] = EMPTY; }
            }
        }
        castling = u.castling; ep = u.ep; halfmove = u.halfmove;
    }

    // ---- move generation -------------------------------------------------------------------------------------------------
    void pseudo(std::vector<Move>& out) const {
        const int c = turn;
        for (int s = 0; s < 64; ++s) {
            const int p = sq[s];
            if (!p || pcolor(p) != c) continue;
            const int t = ptype(p), f = file_of(s), r = rank_of(s);
            auto add = [&](int to) {
                if (t == PAWN && (rank_of(to) == 7 || rank_of(to) == 0)) { for (int pr = QUEEN; pr >= KNIGHT; --pr) out.push_back(Move{(int8_t)s, (int8_t)to, (int8_t)pr}); }
                else out.push_back(Move{(int8_t)s, (int8_t)to, 0});
            };
            if (t == PAWN) {
                const int dir = c == 0 ? 8 : -8, start = c == 0 ? 1 : 6;
                const int one = s + dir;
                if (one >= 0 && one < 64 && !sq[one]) {
                    add(one);
                    if (r == start && !sq[one + dir]) add(one + dir);
                }
                for (int df = -1; df <= 1; df += 2) {
                    const int ff = f + df;
                    if (ff < 0 || ff > 7) continue;
                    const int to = one + df;
                    if (to < 0 || to >= 64) continue;
                    if ((sq[to] && pcolor(sq[to]) != c) || (to == ep && ep >= 0)) add(to);
                }
            } else if (t == KNIGHT || t == KING) {
                static const int kn[8][2] = {{1, 2}, {2, 1}, {2, -1}, {1, -2}, {-1, -2}, {-2, -1}, {-2, 1}, {-1, 2}};
                static const int kg[8][2] = {{1, 0}, {1, 1}, {0, 1}, {-1, 1}, {-1, 0}, {-1, -1}, {0, -1}, {1, -1}};
                const int (*d)[2] = t == KNIGHT ? kn : kg;
                for (int i = 0; i < 8; ++i) {
                    const int ff = f + d[i][0], rr = r + d[i][1];
                    if (ff < 0 || ff > 7 || rr < 0 || rr > 7) continue;
                    const int to = rr * 8 + ff;
                    if (!sq[to] || pcolor(sq[to]) != c) add(to);
                }
                if (t == KING) {
                    const int home = c == 0 ? 4 : 60;
                    if (s == home && !attacked(s, c ^ 1)) {
                        const int kbit = c == 0 ? 1 : 4, qbit = c == 0 ? 2 : 8;
                        if ((castling & kbit) && sq[home + 3] == make_piece(c, ROOK) && !sq[home + 1] && !sq[home + 2] && !attacked(home + 1, c ^ 1) && !attacked(home + 2, c ^ 1))
                            out.push_back(Move{(int8_t)s, (int8_t)(home + 2), 0});
                        if ((castling & qbit) && sq[home - 4] == make_piece(c, ROOK) && !sq[home - 1] && !sq[home - 2] && !sq[home - 3] && !attacked(home - 1, c ^ 1) &&
                            !attacked(home - 2, c ^ 1))
                            out.push_back(Move{(int8_t)s, (int8_t)(home - 2), 0});
                    }
                }
            } else {
                for (int df = -1; df <= 1; ++df)
                    for (int dr = -1; dr <= 1; ++dr) {
                        if (!df && !dr) continue;
                        const bool diag = df != 0 && dr != 0;
                        if ((t == BISHOP && !diag) || (t == ROOK && diag)) continue;
                        int ff = f + df, rr = r + dr;
                        while (ff >= 0 && ff < 8 && rr >= 0 && rr < 8) {
                            const int to = rr * 8 + ff;
                            if (!sq[to]) add(to);
                            else { if (pcolor(sq[to]) != c) add(to); break; }
                            ff += df; rr += dr;
                        }
                    }
            }
        }
    }
    void legal(std::vector<Move>& out) {
        std::vector<Move> ps;
        ps.reserve(64);
        pseudo(ps);
        const int c = turn;
        for (const Move& m : ps) {
            push(m);
            if (!in_check(c)) out.push_back(m);
            pop();
        }
    }
    bool has_legal_ep() {
        if (ep < 0) return false;
        std::vector<Move> ms;
        legal(ms);
        for (const Move& m : ms) if (m.to == ep && ptype(sq[m.from]) == PAWN && file_of(m.from) != file_of(m.to)) return true;
        return false;
    }
    int clean_castling() const {
        int c = castling;
        if (sq[4] != make_piece(0, KING)) c &= ~3;
        if (sq[60] != make_piece(1, KING)) c &= ~12;
        if (sq[7] != make_piece(0, ROOK)) c &= ~1;
        if (sq[0] != make_piece(0, ROOK)) c &= ~2;
        if (sq[63] != make_piece(1, ROOK)) c &= ~4;
        if (sq[56] != make_piece(1, ROOK)) c &= ~8;
        return c;
    }
    uint64_t key() {                                        // transposition key: pieces, turn, cleaned castling rights, legal ep square
        uint64_t h = 1469598103934665603ull;
        auto mix = [&](uint64_t v) { h ^= v; h *= 1099511628211ull; };
        for (int s = 0; s < 64; ++s) mix((uint64_t)(uint8_t)sq[s] + 31ull * s);
        mix(0x100 + turn); mix(0x200 + clean_castling());
        // the ep square only matters when a capture is possible (legality check would recurse into key(): pseudo-legal is the usual compromise)
        int e = -1;
        if (ep >= 0) {
            const int c = turn, from_rank = rank_of(ep) + (c == 0 ? -1 : 1);
            for (int df = -1; df <= 1; df += 2) {
                const int ff = file_of(ep) + df;
                if (ff >= 0 && ff < 8 && sq[from_rank * 8 + ff] == make_piece(c, PAWN)) e = ep;
            }
        }
        mix(0x300 + (uint64_t)(e + 1));
        return h;
    }

    // ---- FEN -----------------------------------------------------------------------------------------------------------------
    static const char* piece_chars() { return ".PNBRQK..pnbrqk"; }
    bool set_fen(const std::string& fen, std::string& err) {
        memset(sq, 0, sizeof sq);
        stack.clear(); keys.clear();
        std::vector<std::string> parts;
        size_t i = 0;
        while (i < fen.size()) {
            while (i < fen.size() && fen[i] == ' ') ++i;
            size_t j = i;
            while (j < fen.size() && fen[j] != ' ') ++j;
            if (j > i) parts.push_back(fen.substr(i, j - i));
            i = j;
        }
        if (parts.empty()) { err = "empty fen"; return false; }
        int r = 7, f = 0;
        for (char ch : parts[0]) {
            if (ch == '/') { if (f != 8) { err = "bad rank length in fen"; return false; } --r; f = 0; continue; }
            if (ch >= '1' && ch <= '8') { f += ch - '0'; continue; }
            const char* pc = strchr(piece_chars(), ch);
            if (!pc || ch == '.' || r < 0 || f > 7) { err = "bad piece placement in fen"; return false; }
            sq[r * 8 + f] = (int8_t)(pc - piece_chars());
            ++f;
        }
        if (r != 0 || f != 8) { err = "fen does not describe 8 ranks"; return false; }
        turn = parts.size() > 1 && parts[1] == "b" ? 1 : 0;
        if (parts.size() > 1 && parts[1] != "w" && parts[1] != "b") { err = "bad side to move in fen"; return false; }
        castling = 0;
        if (parts.size() > 2 && parts[2] != "-")
            for (char ch : parts[2]) {
                if (ch == 'K') castling |= 1; else if (ch == 'Q') castling |= 2; else if (ch == 'k') castling |= 4; else if (ch == 'q') castling |= 8;
                else { err = "bad castling field in fen"; return false; }
            }
        ep = -1;
        if (parts.size() > 3 && parts[3] != "-") {
            if (parts[3].size() != 2 || parts[3][0] < 'a' || parts[3][0] > 'h' || (parts[3][1] != '3' && parts[3][1] != '6')) { err = "bad en passant field in fen"; return false; }
            ep = (parts[3][1] - '1') * 8 + (parts[3][0] - 'a');
        }
        halfmove = parts.size() > 4 ? atoi(parts[4].c_str()) : 0;
        fullmove = parts.size() > 5 ? std::max(1, atoi(parts[5].c_str())) : 1;
        castling = clean_castling();
        keys.push_back(key());
        return true;
    }
    std::string fen() {
        std::string out;
        for (int r = 7; r >= 0; --r) {
            int empty = 0;
            for (int f = 0; f < 8; ++f) {
                const int p = sq[r * 8 + f];
                if (!p) { ++empty; continue; }
                if (empty) { out += (char)('0' + empty); empty = 0; }
                out += piece_chars()[p];
            }
            if (empty) out += (char)('0' + empty);
            if (r) out += '/';
        }
        out += turn ? " b " : " w ";
        const int c = clean_castling();
        if (!c) out += '-';
        else { if (c & 1) out += 'K'; if (c & 2) out += 'Q'; if (c & 4) out += 'k'; if (c & 8) out += 'q'; }
        out += ' ';
        if (has_legal_ep()) { out += (char)('a' + file_of(ep)); out += (char)('1' + rank_of(ep)); }
        else out += '-';
        out += ' ' + std::to_string(halfmove) + ' ' + std::to_string(fullmove);
        return out;
    }

    // ---- SAN -------------------------------------------------------------------------------------------------------------------
    static std::string uci(const Move& m) {
        if (m.is_null()) return "0000";
        std::string s;
        s += (char)('a' + file_of(m.from)); s += (char)('1' + rank_of(m.from)); s += (char)('a' + file_of(m.to)); s += (char)('1' + rank_of(m.to));
        if (m.promo) s += ".pnbrqk"[m.promo];
        return s;
    }
    bool is_castling(const Move& m) const { return ptype(sq[m.from]) == KING && (m.to - m.from == 2 || m.from - m.to == 2); }
    bool is_capture(const Move& m) const { return sq[m.to] != 0 || (ptype(sq[m.from]) == PAWN && m.to == ep && ep >= 0 && file_of(m.from) != file_of(m.to)); }
    std::string san(const Move& m) {
        std::string s;
        std::vector<Move> lm;
        legal(lm);
        if (is_castling(m)) s = m.to > m.from ? "O-O" : "O-O-O";
        else {
            const int t = ptype(sq[m.from]);
            const bool cap = is_capture(m);
            if (t != PAWN) {
                s += ".PNBRQK"[t];
                bool any = false, same_rank = false, same_file = false;
                for (const Move& o : lm)
                    if (o.to == m.to && o.from != m.from && ptype(sq[o.from]) == t) {
                        any = true;
                        if (rank_of(o.from) == rank_of(m.from)) same_rank = true;
                        if (file_of(o.from) == file_of(m.from)) same_file = true;
                    }
                if (any) {
                    bool row = false, column = false;
                    if (same_rank) column = true;
                    if (same_file) row = true; else column = true;
                    if (column) s += (char)('a' + file_of(m.from));
                    if (row) s += (char)('1' + rank_of(m.from));
                }
            } else if (cap) s += (char)('a' + file_of(m.from));
            if (cap) s += 'x';
            s += (char)('a' + file_of(m.to)); s += (char)('1' + rank_of(m.to));
            if (m.promo) { s += '='; s += ".PNBRQK"[m.promo]; }
        }
        push(m);
        std::vector<Move> reply;
        legal(reply);
        if (in_check(turn)) s += reply.empty() ? '#' : '+';
        pop();
        return s;
    }
    // 0 = ok (move in *out), 1 = null move, -1 = syntax / illegal / ambiguous
    int parse_san(const std::string& san_in, Move* out) {
        std::string s = san_in;
        if (s == "--" || s == "Z0" || s == "0000" || s == "@@@@") { *out = Move{}; return 1; }
        std::vector<Move> lm;
        legal(lm);
        auto strip_suffix = [](std::string x) { while (!x.empty() && (x.back() == '+' || x.back() == '#')) x.pop_back(); return x; };
        const std::string core = strip_suffix(s);
        if (core == "O-O" || core == "0-0" || core == "O-O-O" || core == "0-0-0") {
            const bool king_side = core.size() == 3;
            if (s.size() > core.size() + 1) return -1;
            for (const Move& m : lm) if (is_castling(m) && (m.to > m.from) == king_side) { *out = m; return 0; }
            return -1;
        }
        // ^([NBKRQ])?([a-h])?([1-8])?[-x]?([a-h][1-8])(=?[nbrqkNBRQK])?[+#]?$
        size_t i = 0;
        int piece = PAWN, ffile = -1, frank = -1, promo = 0;
        if (i < s.size() && strchr("NBKRQ", s[i])) { piece = (int)(strchr(".PNBRQK", s[i]) - ".PNBRQK"); ++i; }
        // the destination is the LAST square of the body: parse greedily from the right
        size_t end = s.size();
        if (end > i && (s[end - 1] == '+' || s[end - 1] == '#')) --end;
        if (end > i && strchr("nbrqkNBRQK", s[end - 1]) && end >= i + 3) {
            const char pc = s[end - 1];
            size_t e2 = end - 1;
            if (e2 > i && s[e2 - 1] == '=') --e2;
            if (e2 >= i + 2 && s[e2 - 1] >= '1' && s[e2 - 1] <= '8' && s[e2 - 2] >= 'a' && s[e2 - 2] <= 'h') {
                const char up = (char)(pc >= 'a' ? pc - 32 : pc);
                promo = (int)(strchr(".PNBRQK", up) - ".PNBRQK");
                end = e2;
            }
        }
        if (end < i + 2) return -1;
        const char tf = s[end - 2], tr = s[end - 1];
        if (tf < 'a' || tf > 'h' || tr < '1' || tr > '8') return -1;
        const int to = (tr - '1') * 8 + (tf - 'a');
        size_t j = i;
        const size_t body_end = end - 2;
        if (j < body_end && s[j] >= 'a' && s[j] <= 'h') { ffile = s[j] - 'a'; ++j; }
        if (j < body_end && s[j] >= '1' && s[j] <= '8') { frank = s[j] - '1'; ++j; }
        if (j < body_end && (s[j] == '-' || s[j] == 'x')) ++j;
        if (j != body_end) return -1;
        if (promo == KING) return -1;                        // the grammar admits =K, no legal move has it
        int found = 0;
        for (const Move& m : lm) {
            if (m.to != to || ptype(sq[m.from]) != piece || m.promo != promo) continue;
            if (ffile >= 0 && file_of(m.from) != ffile) continue;
            if (frank >= 0 && rank_of(m.from) != frank) continue;
            if (piece == PAWN && ffile < 0 && file_of(m.from) != file_of(to)) continue;      // a pawn move without a file may not capture
            if (is_castling(m)) continue;
            if (found) return -1;                            // ambiguous
            *out = m;
            ++found;
        }
        return found ? 0 : -1;
    }

    // ---- game state ------------------------------------------------------------------------------------------------------------
    bool insufficient_for(int color) const {
        int n_own = 0, own_knights = 0, own_bishops = 0;
        bool own_major = false, other_nonkq = false, any_pawn = false, any_knight = false, light = false, dark = false;
        for (int s = 0; s < 64; ++s) {
            const int p = sq[s];
            if (!p) continue;
            const int t = ptype(p);
            if (t == PAWN) any_pawn = true;
            if (t == KNIGHT) any_knight = true;
            if (t == BISHOP) { if ((file_of(s) + rank_of(s)) & 1) light = true; else dark = true; }
            if (pcolor(p) == color) {
                ++n_own;
                if (t == PAWN || t == ROOK || t == QUEEN) own_major = true;
                if (t == KNIGHT) ++own_knights;
                if (t == BISHOP) ++own_bishops;
            } else if (t != KING && t != QUEEN) other_nonkq = true;
        }
        if (own_major) return false;
        if (own_knights) return n_own <= 2 && !other_nonkq;
        if (own_bishops) return !(light && dark) && !any_pawn && !any_knight;
        return true;
    }
    // bit 0 checkmate, 1 stalemate, 2 insufficient material, 3 seventy-five moves, 4 fivefold repetition, 5 side to move is in check
    int status() {
        std::vector<Move> lm;
        legal(lm);
        const bool chk = in_check(turn);
        int st = chk ? 32 : 0;
        if (lm.empty()) st |= chk ? 1 : 2;
        if (insufficient_for(0) && insufficient_for(1)) st |= 4;
        if (halfmove >= 150 && !lm.empty()) st |= 8;
        int rep = 0;
        for (uint64_t k : keys) if (k == keys.back()) ++rep;
        if (rep >= 5) st |= 16;
        return st;
    }
    // python-chess Board.is_valid(), the checks that matter for generated endgames: one king each, no pawns on the back ranks, the side
    // that just moved is not in check, at most 16 men a side
    bool is_valid() const {
        int kings[2] = {0, 0}, men[2] = {0, 0};
        for (int s = 0; s < 64; ++s) {
            const int p = sq[s];
            if (!p) continue;
            ++men[pcolor(p)];
            if (ptype(p) == KING) ++kings[pcolor(p)];
            if (ptype(p) == PAWN && (rank_of(s) == 0 || rank_of(s) == 7)) return false;
        }
        if (kings[0] != 1 || kings[1] != 1 || men[0] > 16 || men[1] > 16) return false;
        return !in_check(turn ^ 1);
    }
    int64_t perft(int depth) {
        if (depth == 0) return 1;
        std::vector<Move> lm;
        legal(lm);
        if (depth == 1) return (int64_t)lm.size();
        int64_t n = 0;
        for (const Move& m : lm) { push(m); n += perft(depth - 1); pop(); }
        return n;
    }
};

}  // namespace lmrl_chess_ns
