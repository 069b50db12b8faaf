// minitchess.cuh -- bitboard MinitChess rules for sm_100a (also compiles for the host so the
// tests can drive exactly this code without a GPU; the product never runs it on the CPU).
//
// Replaces, on the self-play path, what the reference gets from the python-chess `minitchess`
// fork through exp/environment.py: Board(fen)/legal_moves/result()/push/fen()
// (exp/environment.py:25,36,39,48,76) and the action indexing of exp/generate_moves_list.py.
//
// Board: 5 files x 6 ranks, square s = 5*rank + file, one bit per square in a uint32_t.
// A position is the packed `mc_state` of include/mcaz.h (three piece-type bit planes, a
// colour plane and the turn/clock word): 20 bytes, everything the 4-field FEN holds.
#pragma once
#include <stdint.h>

#include "mcaz.h"

#if defined(__CUDACC__)
#define MC_HD __host__ __device__ __forceinline__
#else
#define MC_HD inline
#endif

namespace mc {

enum : int { EMPTY = 0, PAWN = 1, ROOK = 2, BISHOP = 3, KNIGHT = 4, QUEEN = 5, KING = 6 };

constexpr uint32_t FULL = (1u << 30) - 1u;
constexpr uint32_t FILE_A = 0x02108421u & FULL;  // bits 0,5,10,15,20,25
constexpr uint32_t FILE_B = FILE_A << 1;
constexpr uint32_t FILE_D = FILE_A << 3;
constexpr uint32_t FILE_E = FILE_A << 4;
constexpr uint32_t RANK_1 = 0x1fu;
constexpr uint32_t RANK_6 = 0x1fu << 25;
// squares with (file + rank) even
constexpr uint32_t SHADE_EVEN = 0x15555555u & FULL;

MC_HD int popc(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return __popc(x);
#else
    return __builtin_popcount(x);
#endif
}
MC_HD int lsb(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return __ffs((int)x) - 1;
#else
    return __builtin_ctz(x);
#endif
}

// a square set seen from the mover's side (black: rotated by 180 degrees, view square = 29 - real square)
MC_HD uint32_t view_of(uint32_t b, bool white) {
    if (white) return b;
#if defined(__CUDA_ARCH__)
    return __brev(b) >> 2;
#else
    b = ((b >> 1) & 0x55555555u) | ((b & 0x55555555u) << 1);
    b = ((b >> 2) & 0x33333333u) | ((b & 0x33333333u) << 2);
    b = ((b >> 4) & 0x0f0f0f0fu) | ((b & 0x0f0f0f0fu) << 4);
    b = ((b >> 8) & 0x00ff00ffu) | ((b & 0x00ff00ffu) << 8);
    return ((b >> 16) | (b << 16)) >> 2;
#endif
}

struct Sets {
    uint32_t occ, own, opp, pawns, rooks, bishops, knights, queens, kings;
};

MC_HD bool white_to_move(const mc_state& s) { return (s.meta & 1u) != 0; }
MC_HD int halfmove(const mc_state& s) { return (int)((s.meta >> 8) & 0xffu); }
MC_HD int fullmove(const mc_state& s) { return (int)((s.meta >> 16) & 0xffu); }

MC_HD Sets sets_of(const mc_state& s) {
    Sets t;
    t.occ = (s.pl0 | s.pl1 | s.pl2) & FULL;
    uint32_t w = s.white & t.occ;
    t.own = white_to_move(s) ? w : (t.occ & ~w);
    t.opp = t.occ & ~t.own;
    t.pawns = s.pl0 & ~s.pl1 & ~s.pl2;
    t.rooks = ~s.pl0 & s.pl1 & ~s.pl2;
    t.bishops = s.pl0 & s.pl1 & ~s.pl2;
    t.knights = ~s.pl0 & ~s.pl1 & s.pl2;
    t.queens = s.pl0 & ~s.pl1 & s.pl2;
    t.kings = ~s.pl0 & s.pl1 & s.pl2;
    return t;
}

MC_HD int piece_at(const mc_state& s, int sq) {
    return (int)(((s.pl0 >> sq) & 1u) | (((s.pl1 >> sq) & 1u) << 1) | (((s.pl2 >> sq) & 1u) << 2));
}

// ---- attack sets ---------------------------------------------------------------------------
MC_HD uint32_t king_attacks(uint32_t b) {
    uint32_t h = ((b & ~FILE_A) >> 1) | ((b & ~FILE_E) << 1) | b;
    return ((h | (h << 5) | (h >> 5)) & ~b) & FULL;
}
MC_HD uint32_t knight_attacks(uint32_t b) {
    uint32_t l1 = (b & ~FILE_A) >> 1, l2 = (b & ~(FILE_A | FILE_B)) >> 2;
    uint32_t r1 = (b & ~FILE_E) << 1, r2 = (b & ~(FILE_D | FILE_E)) << 2;
    uint32_t h1 = l1 | r1, h2 = l2 | r2;
    return ((h1 << 10) | (h1 >> 10) | (h2 << 5) | (h2 >> 5)) & FULL;
}
MC_HD uint32_t pawn_attacks(uint32_t p, bool white) {
    return white ? ((((p & ~FILE_A) << 4) | ((p & ~FILE_E) << 6)) & FULL)
                 : (((p & ~FILE_A) >> 6) | ((p & ~FILE_E) >> 4));
}
// Occluded ray fill of a whole piece set in one direction: every square reached, including the
// first blocker.  `up` selects << or >>, `keep` masks out the file the step would wrap from.
template <int SHIFT, bool UP>
MC_HD uint32_t ray(uint32_t src, uint32_t keep, uint32_t empty) {
    uint32_t reach = 0, cur = src;
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        cur &= keep;
        cur = UP ? (cur << SHIFT) : (cur >> SHIFT);
        cur &= FULL;
        reach |= cur;
        cur &= empty;
    }
    return reach;
}
MC_HD uint32_t rook_attacks(uint32_t src, uint32_t occ) {
    uint32_t e = ~occ;
    return ray<5, true>(src, FULL, e) | ray<5, false>(src, FULL, e) | ray<1, true>(src, ~FILE_E, e) |
           ray<1, false>(src, ~FILE_A, e);
}
MC_HD uint32_t bishop_attacks(uint32_t src, uint32_t occ) {
    uint32_t e = ~occ;
    return ray<6, true>(src, ~FILE_E, e) | ray<4, true>(src, ~FILE_A, e) | ray<4, false>(src, ~FILE_E, e) |
           ray<6, false>(src, ~FILE_A, e);
}

// All squares attacked by the side `by` (a bit set of that side's pieces) given piece sets.
MC_HD uint32_t attacked_by(uint32_t by, bool by_white, uint32_t occ, uint32_t pawns, uint32_t rooks,
                           uint32_t bishops, uint32_t knights, uint32_t queens, uint32_t kings) {
    return pawn_attacks(pawns & by, by_white) | knight_attacks(knights & by) | king_attacks(kings & by) |
           rook_attacks((rooks | queens) & by, occ) | bishop_attacks((bishops | queens) & by, occ);
}

MC_HD bool in_check(const mc_state& s) {
    Sets t = sets_of(s);
    uint32_t att = attacked_by(t.opp, !white_to_move(s), t.occ, t.pawns, t.rooks, t.bishops, t.knights,
                               t.queens, t.kings);
    return (att & t.kings & t.own) != 0;
}

// ---- pseudo-legal targets of the piece on `from` ---------------------------------------------
MC_HD uint32_t pseudo_targets(const Sets& t, bool white, int type, int from, const mc_rules& R) {
    uint32_t b = 1u << from;
    switch (type) {
        case PAWN: {
            uint32_t empty = ~t.occ & FULL;
            uint32_t one = (white ? (b << 5) : (b >> 5)) & empty;
            uint32_t tg = one | (pawn_attacks(b, white) & t.opp);
            if (R.pawn_double_step) {
                uint32_t start = white ? (RANK_1 << 5) : (RANK_6 >> 5);
                if ((b & start) && one) tg |= (white ? (one << 5) : (one >> 5)) & empty;
            }
            return tg;
        }
        case KNIGHT: return knight_attacks(b) & ~t.own;
        case KING: return king_attacks(b) & ~t.own;
        case ROOK: return rook_attacks(b, t.occ) & ~t.own;
        case BISHOP: return bishop_attacks(b, t.occ) & ~t.own;
        case QUEEN: return (rook_attacks(b, t.occ) | bishop_attacks(b, t.occ)) & ~t.own;
        default: return 0;
    }
}

// Would moving own piece `type` from->to leave the mover's king attacked?
MC_HD bool leaves_king_safe(const Sets& t, bool white, int type, int from, int to) {
    uint32_t fb = 1u << from, tb = 1u << to, clr = ~(fb | tb);
    uint32_t occ = (t.occ & ~fb) | tb;
    uint32_t opp = t.opp & ~tb;
    uint32_t kings = t.kings & clr;
    if (type == KING) kings |= tb;
    uint32_t own_king = kings & ~opp;
    uint32_t att = attacked_by(opp, !white, occ, t.pawns & clr, t.rooks & clr, t.bishops & clr, t.knights & clr,
                               t.queens & clr, kings);
    return (att & own_king) == 0;
}

// Legal targets by definition: one king-safety test (a full opponent attack set) per candidate move.  Kept as the
// statement of what `Guard` below must reproduce (tests/host_harness compares the two square by square) and for the
// warp-cooperative expansion, which deals these tests out one per lane.
MC_HD uint32_t legal_targets_by_test(const Sets& t, bool white, int type, int from, const mc_rules& R) {
    uint32_t tg = pseudo_targets(t, white, type, from, R), out = 0;
    while (tg) {
        int to = lsb(tg);
        tg &= tg - 1;
        if (leaves_king_safe(t, white, type, from, to)) out |= 1u << to;
    }
    return out;
}

// Everything the king-safety test depends on, worked out once per position: 17 ray fills and one attack set
// instead of one attack set (8 ray fills) per candidate move.  MinitChess has no castling and no en passant, so a
// move is legal iff
//   king:   the target is not attacked once the king is lifted off the board (`danger`);
//   others: the target is on `check_mask` (anywhere when not in check; the checker or a square between it and the
//           king in single check; nowhere in double check) and, if the piece is the first own piece on a line from
//           the king to an enemy slider of that line's kind, on that line (`pin[d]`: the squares from the king
//           up to and including the pinner).
struct Guard {
    uint32_t danger, check_mask, checkers;
    uint32_t pin[8];
};

template <int SHIFT, bool UP>
MC_HD void guard_line(uint32_t kb, uint32_t keep, uint32_t occ, uint32_t own, uint32_t sliders, uint32_t& checkers,
                      uint32_t& lines, uint32_t& pin) {
    const uint32_t e = ~occ;
    const uint32_t r1 = ray<SHIFT, UP>(kb, keep, e);            // up to and including the first piece
    const uint32_t b1 = r1 & occ;
    const uint32_t hit = b1 & sliders;
    checkers |= hit;
    lines |= hit ? r1 : 0u;
    const uint32_t r2 = ray<SHIFT, UP>(b1 & own, keep, e);      // from an own first piece on to the next one
    pin = (r2 & occ & sliders) ? (r1 | r2) : 0u;
}

MC_HD Guard make_guard(const Sets& t, bool white) {
    Guard g;
    const uint32_t kb = t.kings & t.own;
    g.danger = attacked_by(t.opp, !white, t.occ & ~kb, t.pawns, t.rooks, t.bishops, t.knights, t.queens, t.kings);
    const uint32_t straight = t.opp & (t.rooks | t.queens), diagonal = t.opp & (t.bishops | t.queens);
    uint32_t checkers = (knight_attacks(kb) & t.opp & t.knights) | (pawn_attacks(kb, white) & t.opp & t.pawns) |
                        (king_attacks(kb) & t.opp & t.kings);
    uint32_t lines = checkers;
    guard_line<5, true>(kb, FULL, t.occ, t.own, straight, checkers, lines, g.pin[0]);
    guard_line<5, false>(kb, FULL, t.occ, t.own, straight, checkers, lines, g.pin[1]);
    guard_line<1, true>(kb, ~FILE_E, t.occ, t.own, straight, checkers, lines, g.pin[2]);
    guard_line<1, false>(kb, ~FILE_A, t.occ, t.own, straight, checkers, lines, g.pin[3]);
    guard_line<6, true>(kb, ~FILE_E, t.occ, t.own, diagonal, checkers, lines, g.pin[4]);
    guard_line<4, true>(kb, ~FILE_A, t.occ, t.own, diagonal, checkers, lines, g.pin[5]);
    guard_line<4, false>(kb, ~FILE_E, t.occ, t.own, diagonal, checkers, lines, g.pin[6]);
    guard_line<6, false>(kb, ~FILE_A, t.occ, t.own, diagonal, checkers, lines, g.pin[7]);
    g.checkers = checkers;
    const int n = popc(checkers);
    g.check_mask = n == 0 ? FULL : (n == 1 ? lines : 0u);
    return g;
}

MC_HD uint32_t legal_targets(const Sets& t, const Guard& g, bool white, int type, int from, const mc_rules& R) {
    const uint32_t ps = pseudo_targets(t, white, type, from, R);
    if (type == KING) return ps & ~g.danger;
    const uint32_t fb = 1u << from;
    uint32_t allowed = g.check_mask;
#pragma unroll
    for (int d = 0; d < 8; ++d) allowed &= (g.pin[d] & fb) ? g.pin[d] : FULL;
    return ps & allowed;
}

// ---- action codes (exp/generate_moves_list.py:11-36) -----------------------------------------
// Directions as (d_rank, d_file) in the reference's order; codes are numbered over mover's-view
// squares (black sees the board rotated by 180 degrees: view square = 29 - real square).
// Queen block: code = QBASE[from] + rank of (dir, dist) among on-board targets; knight block
// likewise after 430.  Both are evaluated arithmetically: no table in memory.
MC_HD constexpr int qdir_dr(int d) { return d < 3 ? 1 : (d < 5 ? 0 : -1); }
MC_HD constexpr int qdir_df(int d) { return d == 0 || d == 3 || d == 5 ? 1 : (d == 1 || d == 6 ? 0 : -1); }
MC_HD constexpr int ndir_dr(int d) { return d < 2 ? 1 : (d < 4 ? -1 : (d < 6 ? 2 : -2)); }
MC_HD constexpr int ndir_df(int d) { return d < 4 ? ((d & 1) ? -2 : 2) : ((d & 1) ? -1 : 1); }
MC_HD constexpr int min_i(int a, int b) { return a < b ? a : b; }
// number of on-board squares from (r,f) along queen direction d
MC_HD constexpr int qreach(int r, int f, int d) {
    int dr = qdir_dr(d), df = qdir_df(d);
    int nr = dr > 0 ? 5 - r : (dr < 0 ? r : 5);
    int nf = df > 0 ? 4 - f : (df < 0 ? f : 5);
    return min_i(nr, nf);
}
MC_HD int qcount(int r, int f) {
    int n = 0;
#pragma unroll
    for (int d = 0; d < 8; ++d) n += qreach(r, f, d);
    return n;
}
MC_HD constexpr bool n_on(int r, int f, int d) {
    int rr = r + ndir_dr(d), ff = f + ndir_df(d);
    return rr >= 0 && rr < 6 && ff >= 0 && ff < 5;
}

// Emit the codes of all moves from view-square `fv` with legal real targets `tg`, ascending.
// `knight` selects the block.  Returns the number written to out[] (each repeated `rep` times
// for promotions when promo_multiplicity > 1).
template <typename Emit>
MC_HD int emit_square_codes(int fv, bool white, bool knight, uint32_t tg, int base, bool promo_piece, int promo_rep,
                            Emit&& emit) {
    int r = fv / 5, f = fv % 5, n = 0, code = base;
    const uint32_t tgv = view_of(tg, white);               // targets as view squares
    if (!knight) {
        for (int d = 0; d < 8; ++d) {
            const int reach = qreach(r, f, d), stride = 5 * qdir_dr(d) + qdir_df(d);
            int tv = fv;
            for (int k = 1; k <= reach; ++k, ++code) {
                tv += stride;
                if ((tgv >> tv) & 1u) {
                    int rep = (promo_piece && tv >= 25) ? promo_rep : 1;
                    for (int j = 0; j < rep; ++j) { emit((uint16_t)code); ++n; }
                }
            }
        }
    } else {
        for (int d = 0; d < 8; ++d) {
            if (!n_on(r, f, d)) continue;
            int tv = 5 * (r + ndir_dr(d)) + f + ndir_df(d);
            if ((tgv >> tv) & 1u) { emit((uint16_t)code); ++n; }
            ++code;
        }
    }
    return n;
}

// Base code of view-square fv in the queen block / knight block: prefix sums of the per-square widths, worked out by
// the compiler from the same direction lists and packed into 64-bit immediates (qbase - 13 fv in 6 bits, ten squares a
// word; nbase - 430 in 7 bits, nine squares a word) -- a shift and a mask at run time, still no table in memory.
namespace detail {
constexpr int c_abs(int x) { return x < 0 ? -x : x; }
constexpr int c_qwidth(int r, int f) {
    int n = 0;
    for (int dr = -1; dr <= 1; ++dr)
        for (int df = -1; df <= 1; ++df) {
            if (dr == 0 && df == 0) continue;
            int nr = dr > 0 ? 5 - r : (dr < 0 ? r : 5), nf = df > 0 ? 4 - f : (df < 0 ? f : 5);
            n += nr < nf ? nr : nf;
        }
    return n;
}
constexpr int c_nwidth(int r, int f) {
    int n = 0;
    for (int dr = -2; dr <= 2; ++dr)
        for (int df = -2; df <= 2; ++df)
            if (c_abs(dr) + c_abs(df) == 3 && dr != 0 && df != 0 && r + dr >= 0 && r + dr < 6 && f + df >= 0 && f + df < 5) ++n;
    return n;
}
constexpr int c_qbase(int fv) { int b = 0; for (int s = 0; s < fv; ++s) b += c_qwidth(s / 5, s % 5); return b; }
constexpr int c_nbase(int fv) { int b = 0; for (int s = 0; s < fv; ++s) b += c_nwidth(s / 5, s % 5); return b; }
constexpr uint64_t c_pack_q(int word) {
    uint64_t v = 0;
    for (int i = 0; i < 10; ++i) v |= (uint64_t)(c_qbase(word * 10 + i) - 13 * (word * 10 + i)) << (6 * i);
    return v;
}
constexpr uint64_t c_pack_n(int word) {
    uint64_t v = 0;
    for (int i = 0; i < 9 && word * 9 + i < 30; ++i) v |= (uint64_t)c_nbase(word * 9 + i) << (7 * i);
    return v;
}
static_assert(c_qbase(30) == 430 && c_nbase(30) == 124, "554 action codes (exp/generate_moves_list.py)");
static_assert(c_qbase(29) - 13 * 29 < 64 && c_nbase(29) < 128, "packing widths");
constexpr uint64_t QB0 = c_pack_q(0), QB1 = c_pack_q(1), QB2 = c_pack_q(2);
constexpr uint64_t NB0 = c_pack_n(0), NB1 = c_pack_n(1), NB2 = c_pack_n(2), NB3 = c_pack_n(3);
}  // namespace detail
MC_HD constexpr int qbase(int fv) {
    const int w = fv >= 20 ? 2 : (fv >= 10 ? 1 : 0);
    const uint64_t v = w == 2 ? detail::QB2 : (w == 1 ? detail::QB1 : detail::QB0);
    return 13 * fv + (int)((v >> (6 * (fv - 10 * w))) & 63u);
}
MC_HD constexpr int nbase(int fv) {
    const int w = fv >= 27 ? 3 : (fv >= 18 ? 2 : (fv >= 9 ? 1 : 0));
    const uint64_t v = w == 3 ? detail::NB3 : (w == 2 ? detail::NB2 : (w == 1 ? detail::NB1 : detail::NB0));
    return 430 + (int)((v >> (7 * (fv - 9 * w))) & 127u);
}

// code -> view squares.  Returns false for code >= 554.  The arithmetic definition (host, and the source of the device table below).
MC_HD constexpr bool code_to_view_arith(int code, int& fv, int& tv) {
    if (code < 0 || code >= MC_NUM_ACTIONS) return false;
    if (code < 430) {
        int s = 0;
        for (int q = 1; q < 30; ++q) s += qbase(q) <= code ? 1 : 0;      // the last square whose base is <= code
        const int r = s / 5, f = s % 5;
        int off = code - qbase(s);
        for (int d = 0; d < 8; ++d) {
            int reach = qreach(r, f, d);
            if (off < reach) { fv = s; tv = 5 * (r + (off + 1) * qdir_dr(d)) + f + (off + 1) * qdir_df(d); return true; }
            off -= reach;
        }
    } else {
        int s = 0;
        for (int q = 1; q < 30; ++q) s += nbase(q) <= code ? 1 : 0;
        int c = nbase(s);
        for (int d = 0; d < 8; ++d)
            if (n_on(s / 5, s % 5, d)) {
                if (c == code) { fv = s; tv = 5 * (s / 5 + ndir_dr(d)) + s % 5 + ndir_df(d); return true; }
                ++c;
            }
    }
    return false;
}

// ---- small tables for the one-warp-per-tree search, where instructions on the critical path count (a thread-per-position kernel is
// better off with the arithmetic above).  Both are worked out by the compiler from the arithmetic definitions:
//   CODE_VIEW[code]      = from | to << 8 (view squares): code_to_view in one load instead of ~100 instructions
//   MOVE_ORDER.rel[f][t] = offset of the move f -> t within f's queen block (or knight block; a pair of squares is never both), 0xFF: none
//   MOVE_ORDER.less[f][t] = the targets of f (same block) whose codes come before that of t: the place of a move in the sorted list of
//                          its square is a population count instead of a walk over 8 directions x reach
namespace detail {
struct CodeViewTable { uint16_t v[MC_NUM_ACTIONS]; };
constexpr CodeViewTable make_code_view() {
    CodeViewTable t{};
    for (int c = 0; c < MC_NUM_ACTIONS; ++c) {
        int fv = 0, tv = 0;
        code_to_view_arith(c, fv, tv);
        t.v[c] = (uint16_t)(fv | (tv << 8));
    }
    return t;
}
struct MoveOrderTable { uint8_t rel[30][32]; uint32_t less[30][32]; };
constexpr MoveOrderTable make_move_order() {
    MoveOrderTable t{};
    for (int fv = 0; fv < 30; ++fv) {
        for (int tv = 0; tv < 32; ++tv) { t.rel[fv][tv] = 0xFF; t.less[fv][tv] = 0u; }
        const int r = fv / 5, f = fv % 5;
        uint32_t seen = 0u;
        int off = 0;
        for (int d = 0; d < 8; ++d) {                      // queen block: (direction, distance) in the reference's order
            const int reach = qreach(r, f, d), stride = 5 * qdir_dr(d) + qdir_df(d);
            int tv = fv;
            for (int k = 1; k <= reach; ++k) {
                tv += stride;
                t.rel[fv][tv] = (uint8_t)off++;
                t.less[fv][tv] = seen;
                seen |= 1u << tv;
            }
        }
        seen = 0u; off = 0;
        for (int d = 0; d < 8; ++d) {                      // knight block
            if (!n_on(r, f, d)) continue;
            const int tv = 5 * (r + ndir_dr(d)) + f + ndir_df(d);
            t.rel[fv][tv] = (uint8_t)off++;
            t.less[fv][tv] = seen;
            seen |= 1u << tv;
        }
    }
    return t;
}
}  // namespace detail
#if defined(__CUDACC__)
static __device__ const detail::CodeViewTable CODE_VIEW_D = detail::make_code_view();
static __device__ const detail::MoveOrderTable MOVE_ORDER_D = detail::make_move_order();
#endif

MC_HD bool code_to_view(int code, int& fv, int& tv) {
#if defined(__CUDA_ARCH__)
    if (code < 0 || code >= MC_NUM_ACTIONS) return false;
    const uint32_t v = CODE_VIEW_D.v[code];
    fv = (int)(v & 0xffu); tv = (int)(v >> 8);
    return true;
#else
    return code_to_view_arith(code, fv, tv);
#endif
}

// emit_square_codes with the place of every code within its square's sorted list: emit(place, code), in no particular order.
// One table row per legal target instead of the walk.
template <typename Emit>
MC_HD int emit_square_codes_table(const detail::MoveOrderTable& T, int fv, bool white, uint32_t tg, int base, bool promo_piece, int promo_rep,
                                  Emit&& emit) {
    const uint32_t tgv = view_of(tg, white);
    const uint32_t promo = (promo_piece && promo_rep > 1) ? (tgv & 0x3E000000u) : 0u;      // view squares 25..29: promotions, repeated
    uint32_t rest = tgv;
    int n = 0;
    while (rest) {
        const int tv = lsb(rest);
        rest &= rest - 1;
        const uint32_t less = T.less[fv][tv];
        const int code = base + (int)T.rel[fv][tv];
        const int place = popc(tgv & less) + (promo_rep - 1) * popc(promo & less);
        const int rep = ((promo >> tv) & 1u) ? promo_rep : 1;
        for (int j = 0; j < rep; ++j) emit(place + j, (uint16_t)code);
        n += rep;
    }
    return n;
}
// Device: the table form; host: the walk (tests/host_harness checks the table form against it).
template <typename Emit>
MC_HD int emit_square_codes_indexed(int fv, bool white, bool knight, uint32_t tg, int base, bool promo_piece, int promo_rep, Emit&& emit) {
#if defined(__CUDA_ARCH__)
    (void)knight;
    return emit_square_codes_table(MOVE_ORDER_D, fv, white, tg, base, promo_piece, promo_rep, emit);
#else
    int place = 0;
    return emit_square_codes(fv, white, knight, tg, base, promo_piece, promo_rep, [&](uint16_t c) { emit(place, c); ++place; });
#endif
}

// (view from, view to) -> code or -1
MC_HD int view_to_code(int fv, int tv) {
    int r = fv / 5, f = fv % 5, dr = tv / 5 - r, df = tv % 5 - f;
    if (dr == 0 && df == 0) return -1;
    int adr = dr < 0 ? -dr : dr, adf = df < 0 ? -df : df;
    if (dr == 0 || df == 0 || adr == adf) {
        int k = adr > adf ? adr : adf, sr = (dr > 0) - (dr < 0), sf = (df > 0) - (df < 0), code = qbase(fv);
        for (int d = 0; d < 8; ++d) {
            if (qdir_dr(d) == sr && qdir_df(d) == sf) return code + k - 1;
            code += qreach(r, f, d);
        }
        return -1;
    }
    if ((adr == 1 && adf == 2) || (adr == 2 && adf == 1)) {
        int code = nbase(fv);
        for (int d = 0; d < 8; ++d) {
            if (!n_on(r, f, d)) continue;
            if (ndir_dr(d) == dr && ndir_df(d) == df) return code;
            ++code;
        }
    }
    return -1;
}

// ---- result of a position (board.result() as read by exp/environment.py:39-45) -------------
MC_HD bool side_insufficient(const Sets& t, uint32_t side) {
    if (side & (t.pawns | t.rooks | t.queens)) return false;
    if (side & t.knights) return popc(side) <= 2 && ((t.occ & ~side) & ~t.kings & ~t.queens) == 0;
    if (side & t.bishops) {
        bool same = (t.bishops & SHADE_EVEN) == 0 || (t.bishops & ~SHADE_EVEN) == 0;
        return same && t.pawns == 0 && t.knights == 0;
    }
    return true;
}

MC_HD int result_of(const mc_state& s, const Sets& t, int n_legal, const mc_rules& R) {
    if (n_legal == 0) {
        uint32_t att = attacked_by(t.opp, !white_to_move(s), t.occ, t.pawns, t.rooks, t.bishops, t.knights, t.queens,
                                   t.kings);
        if (att & t.kings & t.own) return white_to_move(s) ? MC_BLACK_WINS : MC_WHITE_WINS;
    }
    if (R.insufficient_material && side_insufficient(t, t.own) && side_insufficient(t, t.opp)) return MC_DRAW;
    if (n_legal == 0) return MC_DRAW;
    if (fullmove(s) > R.max_fullmoves) return MC_DRAW;
    return MC_ONGOING;
}

// ---- full move generation for one position, single thread ---------------------------------
// Writes the sorted legal codes to codes[] (at most MC_MAX_MOVES) and returns their number;
// *result receives the MC_* result.  Sorted by construction: queen block by ascending view
// square, then the knight block.
MC_HD int generate(const mc_state& s, const mc_rules& R, uint16_t* codes, int* result) {
    Sets t = sets_of(s);
    bool white = white_to_move(s);
    const Guard g = make_guard(t, white);
    int n = 0, n_moves = 0;
    const uint32_t own_knights = t.own & t.knights;
    for (int pass = 0; pass < 2; ++pass) {
        uint32_t left = view_of(pass ? own_knights : (t.own & ~own_knights), white);      // ascending view squares
        while (left) {
            const int fv = lsb(left);
            left &= left - 1;
            const int sq = white ? fv : 29 - fv;
            const int type = piece_at(s, sq);
            const uint32_t tg = legal_targets(t, g, white, type, sq, R);
            if (tg) {
                n_moves += popc(tg);
                int w = n;
                emit_square_codes(fv, white, pass == 1, tg, pass ? nbase(fv) : qbase(fv), type == PAWN, R.promo_multiplicity,
                                  [&](uint16_t c) { if (w < MC_MAX_MOVES) codes[w] = c; ++w; });
                n = w;
            }
        }
    }
    if (result) *result = result_of(s, t, n_moves, R);
    return n < MC_MAX_MOVES ? n : MC_MAX_MOVES;
}

// ---- apply a move given by real squares (caller has checked legality) -----------------------
MC_HD mc_state apply_move(const mc_state& s, int from, int to) {
    bool white = white_to_move(s);
    int type = piece_at(s, from);
    uint32_t fb = 1u << from, tb = 1u << to;
    bool capture = ((s.pl0 | s.pl1 | s.pl2) & tb) != 0;
    bool zeroing = capture || type == PAWN;
    if (type == PAWN && (to / 5 == (white ? 5 : 0))) type = QUEEN;  // exp/environment.py:72-74
    mc_state o;
    uint32_t clr = ~(fb | tb);
    o.pl0 = (s.pl0 & clr) | ((type & 1) ? tb : 0u);
    o.pl1 = (s.pl1 & clr) | ((type & 2) ? tb : 0u);
    o.pl2 = (s.pl2 & clr) | ((type & 4) ? tb : 0u);
    o.white = (s.white & clr) | (white ? tb : 0u);
    int hm = zeroing ? 0 : halfmove(s) + 1;
    int fm = fullmove(s) + (white ? 0 : 1);
    if (hm > 255) hm = 255;
    if (fm > 255) fm = 255;
    o.meta = MC_META(!white, hm, fm);
    return o;
}

// exp/environment.py:68-82 on a packed position.  status 0 ok, 1 illegal, 2 finished.
MC_HD int step(const mc_state& s, int code, const mc_rules& R, mc_state* out) {
    *out = s;
    Sets t = sets_of(s);
    bool white = white_to_move(s);
    // result needs the legal-move count
    const Guard g = make_guard(t, white);
    int n_moves = 0;
    uint32_t own = t.own;
    while (own) {
        int sq = lsb(own);
        own &= own - 1;
        n_moves += popc(legal_targets(t, g, white, piece_at(s, sq), sq, R));
    }
    if (result_of(s, t, n_moves, R) != MC_ONGOING) return 2;
    int fv, tv;
    if (!code_to_view(code, fv, tv)) return 1;
    int from = white ? fv : 29 - fv, to = white ? tv : 29 - tv;
    if (!((t.own >> from) & 1u)) return 1;
    int type = piece_at(s, from);
    // a knight-shaped code needs a knight, a queen-shaped code anything else (uci is the same)
    if (!((legal_targets(t, g, white, type, from, R) >> to) & 1u)) return 1;
    *out = apply_move(s, from, to);
    return 0;
}

// Network.process_observation (exp/policy.py:82-105).
MC_HD void tokenize(const mc_state& s, uint8_t* tokens, float* clock) {
    bool white = white_to_move(s);
    uint32_t occ = s.pl0 | s.pl1 | s.pl2;
    uint32_t mine = white ? (s.white & occ) : (occ & ~s.white);
    for (int i = 0; i < 30; ++i) {
        int sq = 5 * (5 - i / 5) + i % 5;
        if (!white) sq = 29 - sq;
        int type = piece_at(s, sq);
        bool m = (mine >> sq) & 1u;
        tokens[i] = (uint8_t)(m ? type : 0);
        tokens[30 + i] = (uint8_t)(m ? 0 : type);
    }
    // float(fullmove) (+0.5) / 30 computed in double then rounded to float32, like
    // torch.tensor([[clock / 30]]).float()
    double c = (double)fullmove(s) + (white ? 0.0 : 0.5);
    *clock = (float)(c / 30.0);
}

}  // namespace mc
