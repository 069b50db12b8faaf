/* minitchess_oracle.c -- CPU restatement of the MinitChess rules.  TEST INFRASTRUCTURE ONLY.
 *
 * PARITY UNPINNED for the rules themselves: the python-chess `minitchess` fork that the
 * reference calls (exp/environment.py:3,25,36,39,48,76) is absent from the reference tree, so
 * this file follows the artefacts the reference does pin -- the square numbering and the
 * 554-code action table of exp/generate_moves_list.py:5-57, the 4-field FEN of
 * exp/environment.py:6, the result strings of exp/environment.py:39-45, queen-only promotion
 * of exp/environment.py:72-74, the 30-move cap of exp/policy.py:11-12 -- and upstream
 * python-chess v1.x for the rest (SURVEY.md §8c ledger).  It is pinned against the pure-Python
 * mailbox shim `oracle/shims/chess` (on which the reference's own unmodified
 * exp/environment.py runs) by tests/test_oracle_rules.py.
 *
 * Deliberately a plain mailbox (int8 cell array, ray walking): the product is a bitboard
 * CUDA implementation, so the two share no technique.  Only the packed position struct and
 * the rule switches (include/mcaz.h) are shared, as the data format under test.
 *
 * Build: gcc -O2 -shared -fPIC -I../include minitchess_oracle.c -o _build/libmc_oracle.so
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "mcaz.h"

enum { EMPTY = 0, PAWN = 1, ROOK = 2, BISHOP = 3, KNIGHT = 4, QUEEN = 5, KING = 6 };

typedef struct {
    int8_t cell[30]; /* +type white, -type black */
    int white_to_move;
    int halfmove, fullmove;
} Board;

typedef struct { int8_t from, to; } Mv;

/* exp/generate_moves_list.py:11-12 and :26-27, (d_rank, d_file) */
static const int QDIR[8][2] = {{1, 1}, {1, 0}, {1, -1}, {0, 1}, {0, -1}, {-1, 1}, {-1, 0}, {-1, -1}};
static const int NDIR[8][2] = {{1, 2}, {1, -2}, {-1, 2}, {-1, -2}, {2, 1}, {2, -1}, {-2, 1}, {-2, -1}};

static int16_t g_code_of[30][30]; /* mover's-view (from,to) -> code, -1 if none */
static int8_t g_code_from[MC_NUM_ACTIONS], g_code_to[MC_NUM_ACTIONS];
static int g_tables_ready = 0;

static int on_board(int r, int f) { return r >= 0 && r < 6 && f >= 0 && f < 5; }

/* exp/generate_moves_list.py:13-36: queen-like block then knight block, rank-major squares */
static void build_tables(void) {
    if (g_tables_ready) return;
    memset(g_code_of, 0xff, sizeof(g_code_of));
    int code = 0;
    for (int r = 0; r < 6; ++r)
        for (int f = 0; f < 5; ++f)
            for (int d = 0; d < 8; ++d)
                for (int dist = 1; dist < 6; ++dist) {
                    int rr = r + dist * QDIR[d][0], ff = f + dist * QDIR[d][1];
                    if (!on_board(rr, ff)) continue;
                    g_code_of[5 * r + f][5 * rr + ff] = (int16_t)code;
                    g_code_from[code] = (int8_t)(5 * r + f);
                    g_code_to[code] = (int8_t)(5 * rr + ff);
                    ++code;
                }
    for (int r = 0; r < 6; ++r)
        for (int f = 0; f < 5; ++f)
            for (int d = 0; d < 8; ++d) {
                int rr = r + NDIR[d][0], ff = f + NDIR[d][1];
                if (!on_board(rr, ff)) continue;
                g_code_of[5 * r + f][5 * rr + ff] = (int16_t)code;
                g_code_from[code] = (int8_t)(5 * r + f);
                g_code_to[code] = (int8_t)(5 * rr + ff);
                ++code;
            }
    g_tables_ready = (code == MC_NUM_ACTIONS);
}

int orc_num_actions(void) { build_tables(); return g_tables_ready ? MC_NUM_ACTIONS : -1; }

/* exp/generate_moves_list.py:40-57: black keys are the 180-degree rotation 29 - sq */
int orc_code_of(int from_sq, int to_sq, int white_to_move) {
    build_tables();
    if (!white_to_move) { from_sq = 29 - from_sq; to_sq = 29 - to_sq; }
    return g_code_of[from_sq][to_sq];
}

void orc_code_squares(int code, int white_to_move, int* from_sq, int* to_sq) {
    build_tables();
    int f = g_code_from[code], t = g_code_to[code];
    if (!white_to_move) { f = 29 - f; t = 29 - t; }
    *from_sq = f; *to_sq = t;
}

static void unpack(const mc_state* s, Board* b) {
    for (int q = 0; q < 30; ++q) {
        int t = ((s->pl0 >> q) & 1) | (((s->pl1 >> q) & 1) << 1) | (((s->pl2 >> q) & 1) << 2);
        b->cell[q] = (int8_t)(((s->white >> q) & 1) ? t : -t);
    }
    b->white_to_move = (int)(s->meta & 1u);
    b->halfmove = (int)((s->meta >> 8) & 0xff);
    b->fullmove = (int)((s->meta >> 16) & 0xff);
}

static void pack(const Board* b, mc_state* s) {
    memset(s, 0, sizeof(*s));
    for (int q = 0; q < 30; ++q) {
        int c = b->cell[q], t = c < 0 ? -c : c;
        if (t & 1) s->pl0 |= 1u << q;
        if (t & 2) s->pl1 |= 1u << q;
        if (t & 4) s->pl2 |= 1u << q;
        if (c > 0) s->white |= 1u << q;
    }
    s->meta = MC_META(b->white_to_move, b->halfmove, b->fullmove);
}

static int own(int c, int white) { return white ? c > 0 : c < 0; }

static int attacked_by(const Board* b, int white, int sq) {
    int r = sq / 5, f = sq % 5, sgn = white ? 1 : -1;
    int dr = white ? -1 : 1; /* the attacking pawn sits one rank behind the target */
    for (int df = -1; df <= 1; df += 2)
        if (on_board(r + dr, f + df) && b->cell[5 * (r + dr) + f + df] == sgn * PAWN) return 1;
    for (int d = 0; d < 8; ++d) {
        int rr = r + NDIR[d][0], ff = f + NDIR[d][1];
        if (on_board(rr, ff) && b->cell[5 * rr + ff] == sgn * KNIGHT) return 1;
        rr = r + QDIR[d][0]; ff = f + QDIR[d][1];
        if (on_board(rr, ff) && b->cell[5 * rr + ff] == sgn * KING) return 1;
    }
    for (int d = 0; d < 8; ++d) {
        int diag = QDIR[d][0] != 0 && QDIR[d][1] != 0;
        int rr = r + QDIR[d][0], ff = f + QDIR[d][1];
        while (on_board(rr, ff)) {
            int c = b->cell[5 * rr + ff];
            if (c != EMPTY) {
                if (c == sgn * QUEEN || c == sgn * (diag ? BISHOP : ROOK)) return 1;
                break;
            }
            rr += QDIR[d][0]; ff += QDIR[d][1];
        }
    }
    return 0;
}

static int king_sq(const Board* b, int white) {
    for (int q = 0; q < 30; ++q)
        if (b->cell[q] == (white ? KING : -KING)) return q;
    return -1;
}

static int in_check(const Board* b) {
    int k = king_sq(b, b->white_to_move);
    return k >= 0 && attacked_by(b, !b->white_to_move, k);
}

static int pseudo_moves(const Board* b, const mc_rules* R, Mv* out) {
    int n = 0, w = b->white_to_move;
    for (int s = 0; s < 30; ++s) {
        int c = b->cell[s];
        if (!own(c, w)) continue;
        int t = c < 0 ? -c : c, r = s / 5, f = s % 5;
        if (t == PAWN) {
            int dr = w ? 1 : -1, start = w ? 1 : 4;
            if (on_board(r + dr, f) && b->cell[5 * (r + dr) + f] == EMPTY) {
                out[n].from = (int8_t)s; out[n++].to = (int8_t)(5 * (r + dr) + f);
                if (R->pawn_double_step && r == start && b->cell[5 * (r + 2 * dr) + f] == EMPTY) {
                    out[n].from = (int8_t)s; out[n++].to = (int8_t)(5 * (r + 2 * dr) + f);
                }
            }
            for (int df = -1; df <= 1; df += 2)
                if (on_board(r + dr, f + df)) {
                    int x = b->cell[5 * (r + dr) + f + df];
                    if (x != EMPTY && !own(x, w)) { out[n].from = (int8_t)s; out[n++].to = (int8_t)(5 * (r + dr) + f + df); }
                }
        } else if (t == KNIGHT || t == KING) {
            const int(*D)[2] = (t == KNIGHT) ? NDIR : QDIR;
            for (int d = 0; d < 8; ++d) {
                int rr = r + D[d][0], ff = f + D[d][1];
                if (on_board(rr, ff) && !own(b->cell[5 * rr + ff], w)) { out[n].from = (int8_t)s; out[n++].to = (int8_t)(5 * rr + ff); }
            }
        } else {
            for (int d = 0; d < 8; ++d) {
                int diag = QDIR[d][0] != 0 && QDIR[d][1] != 0;
                if (t == BISHOP && !diag) continue;
                if (t == ROOK && diag) continue;
                int rr = r + QDIR[d][0], ff = f + QDIR[d][1];
                while (on_board(rr, ff)) {
                    int x = b->cell[5 * rr + ff];
                    if (own(x, w)) break;
                    out[n].from = (int8_t)s; out[n++].to = (int8_t)(5 * rr + ff);
                    if (x != EMPTY) break;
                    rr += QDIR[d][0]; ff += QDIR[d][1];
                }
            }
        }
    }
    return n;
}

static void make_move(Board* b, Mv m) {
    int c = b->cell[m.from], t = c < 0 ? -c : c;
    int zeroing = (t == PAWN) || b->cell[m.to] != EMPTY;
    if (t == PAWN && (m.to / 5 == (b->white_to_move ? 5 : 0))) c = b->white_to_move ? QUEEN : -QUEEN;
    b->cell[m.to] = (int8_t)c;
    b->cell[m.from] = EMPTY;
    b->halfmove = zeroing ? 0 : b->halfmove + 1;
    if (!b->white_to_move) b->fullmove += 1;
    b->white_to_move = !b->white_to_move;
}

static int legal_moves(const Board* b, const mc_rules* R, Mv* out) {
    Mv ps[256];
    int np = pseudo_moves(b, R, ps), n = 0;
    for (int i = 0; i < np; ++i) {
        Board c = *b;
        c.cell[ps[i].to] = c.cell[ps[i].from];
        c.cell[ps[i].from] = EMPTY;
        int k = king_sq(&c, b->white_to_move);
        if (k < 0 || !attacked_by(&c, !b->white_to_move, k)) out[n++] = ps[i];
    }
    return n;
}

static int side_insufficient(const Board* b, int white) {
    int cnt[7] = {0}, ocnt[7] = {0}, total = 0;
    for (int q = 0; q < 30; ++q) {
        int c = b->cell[q];
        if (c == EMPTY) continue;
        int t = c < 0 ? -c : c;
        if (own(c, white)) { cnt[t]++; total++; } else ocnt[t]++;
    }
    if (cnt[PAWN] || cnt[ROOK] || cnt[QUEEN]) return 0;
    if (cnt[KNIGHT]) return total <= 2 && !(ocnt[PAWN] || ocnt[KNIGHT] || ocnt[BISHOP] || ocnt[ROOK]);
    if (cnt[BISHOP]) {
        int light = 0, dark = 0;
        for (int q = 0; q < 30; ++q) {
            int c = b->cell[q];
            if (c == BISHOP || c == -BISHOP) { if (((q % 5) + (q / 5)) & 1) light = 1; else dark = 1; }
        }
        return !(light && dark) && !(cnt[PAWN] + ocnt[PAWN]) && !(cnt[KNIGHT] + ocnt[KNIGHT]);
    }
    return 1;
}

/* result() without move-stack rules: checkmate first, then every kind of draw */
static int result_of(const Board* b, const mc_rules* R, int n_legal) {
    if (n_legal == 0 && in_check(b)) return b->white_to_move ? MC_BLACK_WINS : MC_WHITE_WINS;
    if (R->insufficient_material && side_insufficient(b, 1) && side_insufficient(b, 0)) return MC_DRAW;
    if (n_legal == 0) return MC_DRAW;
    if (b->fullmove > R->max_fullmoves) return MC_DRAW;
    return MC_ONGOING;
}

static const mc_rules DEFAULT_RULES = {0, 1, 30, 1, 1};
static const mc_rules* rules_or_default(const mc_rules* r) { return r ? r : &DEFAULT_RULES; }

static int cmp_u16(const void* a, const void* b) { return (int)*(const uint16_t*)a - (int)*(const uint16_t*)b; }

/* exp/environment.py:47-50: sorted codes of board.legal_moves; promo multiplicity 4 repeats a code */
int orc_legal_moves(const mc_state* s, const mc_rules* rules, uint16_t* codes, int8_t* result) {
    build_tables();
    const mc_rules* R = rules_or_default(rules);
    Board b; unpack(s, &b);
    Mv mv[256];
    int n = legal_moves(&b, R, mv), k = 0;
    for (int i = 0; i < n; ++i) {
        int code = orc_code_of(mv[i].from, mv[i].to, b.white_to_move);
        int t = b.cell[mv[i].from]; t = t < 0 ? -t : t;
        int promo = (t == PAWN) && (mv[i].to / 5 == (b.white_to_move ? 5 : 0));
        int reps = promo ? R->promo_multiplicity : 1;
        for (int j = 0; j < reps; ++j) codes[k++] = (uint16_t)code;
    }
    qsort(codes, (size_t)k, sizeof(uint16_t), cmp_u16);
    if (result) *result = (int8_t)result_of(&b, R, n);
    return k;
}

void orc_legal_moves_bulk(const mc_state* s, int n, const mc_rules* rules, uint16_t* codes, int32_t* counts, int8_t* results) {
    for (int i = 0; i < n; ++i) counts[i] = orc_legal_moves(&s[i], rules, codes + (size_t)i * MC_MAX_MOVES, &results[i]);
}

/* exp/environment.py:68-82.  status: 0 ok, 1 illegal, 2 finished */
int orc_apply(const mc_state* s, uint16_t code, const mc_rules* rules, mc_state* out) {
    build_tables();
    const mc_rules* R = rules_or_default(rules);
    Board b; unpack(s, &b);
    Mv mv[256];
    int n = legal_moves(&b, R, mv);
    *out = *s;
    if (result_of(&b, R, n) != MC_ONGOING) return 2;
    if (code >= MC_NUM_ACTIONS) return 1;
    int from, to; orc_code_squares(code, b.white_to_move, &from, &to);
    for (int i = 0; i < n; ++i)
        if (mv[i].from == from && mv[i].to == to) { make_move(&b, mv[i]); pack(&b, out); return 0; }
    return 1;
}

void orc_apply_bulk(const mc_state* s, const uint16_t* codes, int n, const mc_rules* rules, mc_state* out, int8_t* status) {
    for (int i = 0; i < n; ++i) status[i] = (int8_t)orc_apply(&s[i], codes[i], rules, &out[i]);
}

static uint64_t perft_rec(const Board* b, const mc_rules* R, int depth) {
    if (depth == 0) return 1;
    Mv mv[256];
    int n = legal_moves(b, R, mv);
    if (result_of(b, R, n) != MC_ONGOING) return 0;
    if (depth == 1) return (uint64_t)n;
    uint64_t total = 0;
    for (int i = 0; i < n; ++i) { Board c = *b; make_move(&c, mv[i]); total += perft_rec(&c, R, depth - 1); }
    return total;
}

uint64_t orc_perft(const mc_state* s, int depth, const mc_rules* rules) {
    build_tables();
    Board b; unpack(s, &b);
    return perft_rec(&b, rules_or_default(rules), depth);
}

/* Network.process_observation (exp/policy.py:82-105): mover's view, FEN order (rank 6 first) */
void orc_tokenize(const mc_state* s, uint8_t* tokens, float* clock) {
    Board b; unpack(s, &b);
    for (int i = 0; i < 30; ++i) {
        int rank = 5 - i / 5, file = i % 5, sq = 5 * rank + file;
        if (!b.white_to_move) sq = 29 - sq; /* bfen[::-1].swapcase() */
        int c = b.cell[sq];
        if (!b.white_to_move) c = -c;
        tokens[i] = (uint8_t)(c > 0 ? c : 0);
        tokens[30 + i] = (uint8_t)(c < 0 ? -c : 0);
    }
    double clk = (double)b.fullmove + (b.white_to_move ? 0.0 : 0.5);
    *clock = (float)(clk / 30.0);
}

/* xorshift64* for reproducible random playouts */
static uint64_t rng_next(uint64_t* st) {
    uint64_t x = *st; x ^= x >> 12; x ^= x << 25; x ^= x >> 27; *st = x;
    return x * 0x2545F4914F6CDD1DULL;
}

static const char START_ROWS[6][6] = {"KBN..", "PPP..", ".....", ".....", "..ppp", "..nbk"};

void orc_start_state(mc_state* out) {
    Board b; memset(&b, 0, sizeof(b));
    for (int r = 0; r < 6; ++r)
        for (int f = 0; f < 5; ++f) {
            char ch = START_ROWS[r][f];
            int t = 0, w = ch >= 'A' && ch <= 'Z';
            switch (ch | 0x20) { case 'p': t = PAWN; break; case 'r': t = ROOK; break; case 'b': t = BISHOP; break;
                                 case 'n': t = KNIGHT; break; case 'q': t = QUEEN; break; case 'k': t = KING; break; default: t = 0; }
            b.cell[5 * r + f] = (int8_t)(w ? t : -t);
        }
    b.white_to_move = 1; b.halfmove = 0; b.fullmove = 1;
    pack(&b, out);
}

/* Uniform-random legal playouts from the start position; every position visited (including
 * the finished one) is a sample.  Repetition is not tracked (stateless positions).  Returns
 * the number written (== n).  SURVEY.md §8d config 2. */
int orc_random_positions(uint64_t seed, int n, const mc_rules* rules, mc_state* out) {
    build_tables();
    const mc_rules* R = rules_or_default(rules);
    uint64_t st = seed * 0x9E3779B97F4A7C15ULL + 0x1234567ULL;
    if (!st) st = 1;
    int k = 0;
    while (k < n) {
        mc_state s; orc_start_state(&s);
        Board b; unpack(&s, &b);
        for (;;) {
            pack(&b, &out[k++]);
            if (k >= n) break;
            Mv mv[256];
            int nl = legal_moves(&b, R, mv);
            if (result_of(&b, R, nl) != MC_ONGOING) break;
            make_move(&b, mv[rng_next(&st) % (uint64_t)nl]);
        }
    }
    return k;
}
