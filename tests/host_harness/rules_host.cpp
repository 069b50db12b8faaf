// Host build of csrc/minitchess.cuh for CPU-only tests (no GPU in the build container).
// TEST INFRASTRUCTURE: lets `pytest -m "not gpu"` drive the very same bitboard rules the CUDA
// kernels compile, against the mailbox oracle.  The product never loads this library.
#include <cstring>
#include "minitchess.cuh"

static mc_rules rules_or_default(const mc_rules* r) {
    mc_rules d = {0, 1, 30, 1, 1};
    return r ? *r : d;
}

extern "C" {
void hh_legal_moves(const mc_state* s, int n, const mc_rules* rules, uint16_t* codes, int32_t* counts, int8_t* results) {
    mc_rules R = rules_or_default(rules);
    for (int i = 0; i < n; ++i) {
        int res;
        counts[i] = mc::generate(s[i], R, codes + (size_t)i * MC_MAX_MOVES, &res);
        results[i] = (int8_t)res;
    }
}
void hh_apply(const mc_state* s, const uint16_t* codes, int n, const mc_rules* rules, mc_state* out, int8_t* status) {
    mc_rules R = rules_or_default(rules);
    for (int i = 0; i < n; ++i) status[i] = (int8_t)mc::step(s[i], codes[i], R, &out[i]);
}
void hh_tokenize(const mc_state* s, int n, uint8_t* tokens, float* clocks) {
    for (int i = 0; i < n; ++i) mc::tokenize(s[i], tokens + (size_t)i * 60, clocks + i);
}
// Legal target sets of every own piece, by the per-position Guard (what generate/step use) and by the definition
// (one king-safety test per candidate move): out[i*30 + square].
void hh_targets_both(const mc_state* s, int n, const mc_rules* rules, uint32_t* by_guard, uint32_t* by_test) {
    mc_rules R = rules_or_default(rules);
    for (int i = 0; i < n; ++i) {
        mc::Sets t = mc::sets_of(s[i]);
        bool white = mc::white_to_move(s[i]);
        mc::Guard g = mc::make_guard(t, white);
        for (int sq = 0; sq < 30; ++sq) {
            uint32_t a = 0, b = 0;
            if ((t.own >> sq) & 1u) {
                int type = mc::piece_at(s[i], sq);
                a = mc::legal_targets(t, g, white, type, sq, R);
                b = mc::legal_targets_by_test(t, white, type, sq, R);
            }
            by_guard[(size_t)i * 30 + sq] = a;
            by_test[(size_t)i * 30 + sq] = b;
        }
    }
}
// The device tables (CODE_VIEW, MOVE_ORDER) are constexpr products of the arithmetic: the same objects built for the host.
static const mc::detail::CodeViewTable HH_CODE_VIEW = mc::detail::make_code_view();
static const mc::detail::MoveOrderTable HH_MOVE_ORDER = mc::detail::make_move_order();
int hh_code_view_table(int code) { return HH_CODE_VIEW.v[code]; }
// codes of the moves from view square fv to the (real-square) targets tg, by the walk and by the table form; returns the count of
// the walk, *n_table gets the count of the table form; out_table[place] = code
int hh_emit_both(int fv, int white, int knight, uint32_t tg, int promo_piece, int promo_rep, uint16_t* out_walk, uint16_t* out_table, int* n_table) {
    const int base = knight ? mc::nbase(fv) : mc::qbase(fv);
    int n = 0;
    mc::emit_square_codes(fv, white != 0, knight != 0, tg, base, promo_piece != 0, promo_rep, [&](uint16_t c) { out_walk[n++] = c; });
    *n_table = mc::emit_square_codes_table(HH_MOVE_ORDER, fv, white != 0, tg, base, promo_piece != 0, promo_rep,
                                           [&](int place, uint16_t c) { out_table[place] = c; });
    return n;
}
int hh_view_to_code(int fv, int tv) { return mc::view_to_code(fv, tv); }
int hh_code_to_view(int code, int* fv, int* tv) { return mc::code_to_view(code, *fv, *tv) ? 0 : -1; }
}
