// chess_api.inl — C ABI of the host-side chess rules (chess_core.h) behind lmrl_gym_b200/chess_env.py (C ABI part 6).
// Pure host code: no CUDA call is made here; errors go through lmrl_last_error() like everywhere else.
#include "chess_core.h"

struct lmrl_chess { lmrl_chess_ns::Board b; };

static int chess_copy(const std::string& s, char* out, int cap) {
    if (!out || cap <= (int)s.size()) return lmrl::fail("lmrl_chess: output buffer of %d bytes is too small for %zu", cap, s.size() + 1);
    memcpy(out, s.c_str(), s.size() + 1);
    return (int)s.size();
}

extern "C" {

lmrl_chess* lmrl_chess_new(const char* fen) {
    lmrl_chess* c = new lmrl_chess();
    std::string err;
    if (!c->b.set_fen(fen ? fen : "rnbqkbnr/pppppppp/8/8/8/8/PPPPPPPP/RNBQKBNR w KQkq - 0 1", err)) {
        lmrl::fail("lmrl_chess_new: %s", err.c_str());
        delete c;
        return nullptr;
    }
    return c;
}
void lmrl_chess_free(lmrl_chess* c) { delete c; }

int lmrl_chess_fen(lmrl_chess* c, char* out, int32_t cap) {
    LMRL_CHECK(c, "lmrl_chess_fen: null board");
    return chess_copy(c->b.fen(), out, cap);
}

int lmrl_chess_legal_moves(lmrl_chess* c, char* out_uci, int32_t cap) {
    LMRL_CHECK(c, "lmrl_chess_legal_moves: null board");
    std::vector<lmrl_chess_ns::Move> lm;
    c->b.legal(lm);
    std::string s;
    for (size_t i = 0; i < lm.size(); ++i) { if (i) s += ' '; s += lmrl_chess_ns::Board::uci(lm[i]); }
    if (chess_copy(s, out_uci, cap) < 0) return -1;
    return (int)lm.size();
}

int lmrl_chess_push_san(lmrl_chess* c, const char* san, char* uci_out, int32_t cap) {
    LMRL_CHECK(c && san, "lmrl_chess_push_san: null argument");
    lmrl_chess_ns::Move m;
    const int rc = c->b.parse_san(san, &m);
    if (rc < 0) return lmrl::fail("lmrl_chess_push_san: illegal, ambiguous or malformed san '%s'", san);
    if (uci_out && chess_copy(lmrl_chess_ns::Board::uci(m), uci_out, cap) < 0) return -1;
    if (rc == 0) c->b.push(m);           // a null move is reported (1) and NOT pushed: the reference pops it at once (chess/env/env.py:116-118)
    return rc;
}

int lmrl_chess_push_uci(lmrl_chess* c, const char* uci, char* san_out, int32_t cap) {
    LMRL_CHECK(c && uci, "lmrl_chess_push_uci: null argument");
    std::vector<lmrl_chess_ns::Move> lm;
    c->b.legal(lm);
    for (const auto& m : lm)
        if (lmrl_chess_ns::Board::uci(m) == uci) {
            if (san_out && chess_copy(c->b.san(m), san_out, cap) < 0) return -1;
            c->b.push(m);
            return 0;
        }
    return lmrl::fail("lmrl_chess_push_uci: '%s' is not a legal move here", uci);
}

int lmrl_chess_pop(lmrl_chess* c) {
    LMRL_CHECK(c && !c->b.stack.empty(), "lmrl_chess_pop: nothing to pop");
    c->b.pop();
    return 0;
}

int lmrl_chess_status(lmrl_chess* c) { return c ? c->b.status() : -1; }
int lmrl_chess_turn(lmrl_chess* c) { return c ? c->b.turn : -1; }
int lmrl_chess_is_valid(lmrl_chess* c) { return c ? (c->b.is_valid() ? 1 : 0) : -1; }
int64_t lmrl_chess_perft(lmrl_chess* c, int32_t depth) { return c ? c->b.perft(depth) : -1; }

}  // extern "C"
