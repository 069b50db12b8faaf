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

MC_HD uint32_t legal_targets(const Sets& t, bool white, int type, int from, const mc_rules& R) {
    uint32_t tg = pseudo_targets(t, white, type, from, R), out = 0;
    while (tg) {
        int to = lsb(tg);
        tg &= tg - 1;
        if (leaves_king_safe(t, white, type, from, to)) out |= 1u << to;
    }
    return out;
}

// ---- action codes (exp/generate_moves_list.py:11-36) -----------------------------------------
// Directions as (d_rank, d_file) in the reference's order; codes are numbered over mover's-view
// squares (black sees the board rotated by 180 degrees: view square = 29 - real square).
// Queen block: code = QBASE[from] + rank of (dir, dist) among on-board targets; knight block
// likewise after 430.  Both are evaluated arithmetically: no table in memory.
MC_HD int qdir_dr(int d) { return d < 3 ? 1 : (d < 5 ? 0 : -1); }
MC_HD int qdir_df(int d) { return d == 0 || d == 3 || d == 5 ? 1 : (d == 1 || d == 6 ? 0 : -1); }
MC_HD int ndir_dr(int d) { return d < 2 ? 1 : (d < 4 ? -1 : (d < 6 ? 2 : -2)); }
MC_HD int ndir_df(int d) { return d < 4 ? ((d & 1) ? -2 : 2) : ((d & 1) ? -1 : 1); }
MC_HD int min_i(int a, int b) { return a < b ? a : b; }
// number of on-board squares from (r,f) along queen direction d
MC_HD int qreach(int r, int f, int d) {
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
MC_HD bool n_on(int r, int f, int d) {
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
    if (!knight) {
        for (int d = 0; d < 8; ++d) {
            int dr = qdir_dr(d), df = qdir_df(d), reach = qreach(r, f, d);
            for (int k = 1; k <= reach; ++k, ++code) {
                int tv = 5 * (r + k * dr) + f + k * df;
                int to = white ? tv : 29 - tv;
                if ((tg >> to) & 1u) {
                    int rep = (promo_piece && (tv / 5 == 5)) ? promo_rep : 1;
                    for (int j = 0; j < rep; ++j) { emit((uint16_t)code); ++n; }
                }
            }
        }
    } else {
        for (int d = 0; d < 8; ++d) {
            if (!n_on(r, f, d)) continue;
            int tv = 5 * (r + ndir_dr(d)) + f + ndir_df(d);
            int to = white ? tv : 29 - tv;
            if ((tg >> to) & 1u) { emit((uint16_t)code); ++n; }
            ++code;
        }
    }
    return n;
}

// Base code of view-square fv in the queen block / knight block (prefix sums, computed).
MC_HD int qbase(int fv) {
    int b = 0;
    for (int s = 0; s < fv; ++s) b += qcount(s / 5, s % 5);
    return b;
}
MC_HD int nbase(int fv) {
    int b = 430;
    for (int s = 0; s < fv; ++s)
        for (int d = 0; d < 8; ++d) b += n_on(s / 5, s % 5, d) ? 1 : 0;
    return b;
}

// code -> view squares.  Returns false for code >= 554.
MC_HD bool code_to_view(int code, int& fv, int& tv) {
    if (code < 0 || code >= MC_NUM_ACTIONS) return false;
    if (code < 430) {
        int base = 0;
        for (int s = 0; s < 30; ++s) {
            int r = s / 5, f = s % 5, c = qcount(r, f);
            if (code < base + c) {
                int off = code - base;
                for (int d = 0; d < 8; ++d) {
                    int reach = qreach(r, f, d);
                    if (off < reach) { fv = s; tv = 5 * (r + (off + 1) * qdir_dr(d)) + f + (off + 1) * qdir_df(d); return true; }
                    off -= reach;
                }
            }
            base += c;
        }
    } else {
        int c = 430;
        for (int s = 0; s < 30; ++s)
            for (int d = 0; d < 8; ++d)
                if (n_on(s / 5, s % 5, d)) {
                    if (c == code) { fv = s; tv = 5 * (s / 5 + ndir_dr(d)) + s % 5 + ndir_df(d); return true; }
                    ++c;
                }
    }
    return false;
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
    int n = 0, n_moves = 0;
    for (int pass = 0; pass < 2; ++pass) {
        int base = pass ? 430 : 0;
        for (int fv = 0; fv < 30; ++fv) {
            int sq = white ? fv : 29 - fv;
            int r = fv / 5, f = fv % 5;
            int width;
            if (pass == 0) width = qcount(r, f);
            else { width = 0; for (int d = 0; d < 8; ++d) width += n_on(r, f, d) ? 1 : 0; }
            if ((t.own >> sq) & 1u) {
                int type = piece_at(s, sq);
                if ((type == KNIGHT) == (pass == 1)) {
                    uint32_t tg = legal_targets(t, white, type, sq, R);
                    if (tg) {
                        n_moves += popc(tg);
                        int w = n;
                        emit_square_codes(fv, white, pass == 1, tg, base, type == PAWN, R.promo_multiplicity,
                                          [&](uint16_t c) { if (w < MC_MAX_MOVES) codes[w] = c; ++w; });
                        n = w;
                    }
                }
            }
            base += width;
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
    int n_moves = 0;
    uint32_t own = t.own;
    while (own) {
        int sq = lsb(own);
        own &= own - 1;
        n_moves += popc(legal_targets(t, white, piece_at(s, sq), sq, R));
    }
    if (result_of(s, t, n_moves, R) != MC_ONGOING) return 2;
    int fv, tv;
    if (!code_to_view(code, fv, tv)) return 1;
    int from = white ? fv : 29 - fv, to = white ? tv : 29 - tv;
    if (!((t.own >> from) & 1u)) return 1;
    int type = piece_at(s, from);
    // a knight-shaped code needs a knight, a queen-shaped code anything else (uci is the same)
    if (!((legal_targets(t, white, type, from, R) >> to) & 1u)) return 1;
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
