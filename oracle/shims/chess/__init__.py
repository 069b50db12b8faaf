"""CPU MinitChess rules: a `chess`-compatible stand-in for the python-chess `minitchess` fork.

TEST INFRASTRUCTURE ONLY -- never imported by the product path.

The reference's rules live in github.com/schouhy/python-chess (branch `minitchess`, pinned
`1f7c95a8...` in `Dockerfile:14-15`, `a4cf4ddd...` in `README.md:17`), which is NOT in the
reference tree and cannot be fetched here.  **Parity with that fork is therefore UNPINNED.**
This module restates the rules from the artefacts the reference does pin:

* board 5 files x 6 ranks, ``sq = 5*rank + file``        (`exp/generate_moves_list.py:5-9,14-22`)
* 4-field FEN ``<board> <w|b> <halfmove> <fullmove>``     (`exp/environment.py:6`, `exp/policy.py:98`)
* result vocabulary '1-0' / '0-1' / '1/2-1/2' / '*'        (`exp/environment.py:39-45`)
* promotion exists, the env always queens                  (`exp/environment.py:49,72-74`)
* a 30-move cap exists                                     (`exp/policy.py:11-12`)

and upstream python-chess v1.x semantics for the rest.  Every unpinned choice is a named
switch in ``RULES`` with the same default as the CUDA kernels (`csrc/minitchess.cuh`).

It is written as a plain mailbox (list of 30 chars) on purpose: the GPU path is a bitboard
implementation, so agreement between the two is a meaningful cross-check.

Only the API surface the reference touches is provided: ``Board(fen)``, ``.fen()``,
``.result()``, ``.legal_moves``, ``.push``, ``.turn``, ``Move(from_square, to_square,
promotion)``, ``Move.from_uci``, ``Move.uci()``, ``Move.__eq__``
(call sites: `exp/environment.py:25,36,39,48-49,66,72-76`, `exp/generate_moves_list.py:45,55`).
"""

WHITE, BLACK = True, False
NUM_FILES, NUM_RANKS, NUM_SQUARES = 5, 6, 30
FILE_NAMES = 'abcde'
STARTING_FEN = '2nbk/2ppp/5/5/PPP2/KBN2 w 0 1'

PAWN, KNIGHT, BISHOP, ROOK, QUEEN, KING = 1, 2, 3, 4, 5, 6
PIECE_SYMBOLS = [None, 'p', 'n', 'b', 'r', 'q', 'k']

# Unpinned-rule switches (SURVEY.md §8c ledger).  Same names/defaults as mc_rules in include/mcaz.h.
RULES = {
    'pawn_double_step': False,      # no en-passant field in the FEN -> assume no double step
    'promo_multiplicity': 1,        # 1: queen only; 4: q,r,b,n (duplicate 4-char codes)
    'max_fullmoves': 30,            # draw once fullmove_number > max_fullmoves
    'insufficient_material': True,  # upstream has_insufficient_material, true square colours
    'fivefold_repetition': True,    # upstream is_fivefold_repetition over the move stack
}

KNIGHT_STEPS = [(1, 2), (1, -2), (-1, 2), (-1, -2), (2, 1), (2, -1), (-2, 1), (-2, -1)]
KING_STEPS = [(1, 1), (1, 0), (1, -1), (0, 1), (0, -1), (-1, 1), (-1, 0), (-1, -1)]
BISHOP_DIRS = [(1, 1), (1, -1), (-1, 1), (-1, -1)]
ROOK_DIRS = [(1, 0), (-1, 0), (0, 1), (0, -1)]


def square(file_index, rank_index):
    return 5 * rank_index + file_index


def square_name(sq):
    return FILE_NAMES[sq % 5] + str(sq // 5 + 1)


def parse_square(name):
    return 5 * (int(name[1]) - 1) + FILE_NAMES.index(name[0])


SQUARE_NAMES = [square_name(s) for s in range(NUM_SQUARES)]


class Move:
    __slots__ = ('from_square', 'to_square', 'promotion')

    def __init__(self, from_square, to_square, promotion=None):
        self.from_square = int(from_square)
        self.to_square = int(to_square)
        self.promotion = promotion

    def uci(self):
        s = square_name(self.from_square) + square_name(self.to_square)
        if self.promotion:
            s += PIECE_SYMBOLS[self.promotion]
        return s

    @classmethod
    def from_uci(cls, uci):
        promo = PIECE_SYMBOLS.index(uci[4]) if len(uci) == 5 else None
        return cls(parse_square(uci[0:2]), parse_square(uci[2:4]), promo)

    def __eq__(self, other):
        return (isinstance(other, Move) and self.from_square == other.from_square
                and self.to_square == other.to_square and self.promotion == other.promotion)

    def __hash__(self):
        return hash((self.from_square, self.to_square, self.promotion))

    def __repr__(self):
        return 'Move.from_uci(%r)' % self.uci()


def _on_board(r, f):
    return 0 <= r < NUM_RANKS and 0 <= f < NUM_FILES


class Board:
    def __init__(self, fen=STARTING_FEN):
        self.cells = ['.'] * NUM_SQUARES       # upper = white, lower = black, '.' = empty
        self.turn = WHITE
        self.halfmove_clock = 0
        self.fullmove_number = 1
        self.move_stack = []
        self._history = []                     # (cells, turn) keys since the last irreversible move
        self.set_fen(fen)

    # ------------------------------------------------------------------ FEN
    def set_fen(self, fen):
        rows, turn, half, full = fen.split()
        rows = rows.split('/')
        assert len(rows) == NUM_RANKS, fen
        for i, row in enumerate(rows):
            rank = NUM_RANKS - 1 - i
            f = 0
            for ch in row:
                if ch.isdigit():
                    f += int(ch)
                else:
                    self.cells[square(f, rank)] = ch
                    f += 1
            assert f == NUM_FILES, fen
        self.turn = (turn == 'w')
        self.halfmove_clock = int(half)
        self.fullmove_number = int(full)
        self.move_stack = []
        self._history = [self._key()]

    def board_fen(self):
        out = []
        for rank in range(NUM_RANKS - 1, -1, -1):
            run, row = 0, ''
            for f in range(NUM_FILES):
                c = self.cells[square(f, rank)]
                if c == '.':
                    run += 1
                else:
                    if run:
                        row += str(run)
                        run = 0
                    row += c
            if run:
                row += str(run)
            out.append(row)
        return '/'.join(out)

    def fen(self):
        return '%s %s %d %d' % (self.board_fen(), 'w' if self.turn else 'b',
                                self.halfmove_clock, self.fullmove_number)

    def _key(self):
        return (''.join(self.cells), self.turn)

    # ------------------------------------------------------------ attacks
    def _is_own(self, c, color):
        return c != '.' and (c.isupper() == color)

    def is_attacked_by(self, color, sq):
        """True if a piece of `color` attacks square `sq`."""
        r, f = divmod(sq, 5)
        pawn = 'P' if color else 'p'
        dr = -1 if color else 1                      # a white pawn attacks upward: it sits one rank below
        for df in (-1, 1):
            rr, ff = r + dr, f + df
            if _on_board(rr, ff) and self.cells[square(ff, rr)] == pawn:
                return True
        knight = 'N' if color else 'n'
        for a, b in KNIGHT_STEPS:
            rr, ff = r + a, f + b
            if _on_board(rr, ff) and self.cells[square(ff, rr)] == knight:
                return True
        king = 'K' if color else 'k'
        for a, b in KING_STEPS:
            rr, ff = r + a, f + b
            if _on_board(rr, ff) and self.cells[square(ff, rr)] == king:
                return True
        for dirs, sliders in ((BISHOP_DIRS, 'BQ' if color else 'bq'), (ROOK_DIRS, 'RQ' if color else 'rq')):
            for a, b in dirs:
                rr, ff = r + a, f + b
                while _on_board(rr, ff):
                    c = self.cells[square(ff, rr)]
                    if c != '.':
                        if c in sliders:
                            return True
                        break
                    rr, ff = rr + a, ff + b
        return False

    def king_square(self, color):
        k = 'K' if color else 'k'
        for s in range(NUM_SQUARES):
            if self.cells[s] == k:
                return s
        return None

    def is_check(self):
        ks = self.king_square(self.turn)
        return ks is not None and self.is_attacked_by(not self.turn, ks)

    # ------------------------------------------------------------ move gen
    def _pseudo_legal(self):
        color = self.turn
        promos = [QUEEN, ROOK, BISHOP, KNIGHT][:RULES['promo_multiplicity']]
        for s in range(NUM_SQUARES):
            c = self.cells[s]
            if not self._is_own(c, color):
                continue
            r, f = divmod(s, 5)
            kind = c.lower()
            if kind == 'p':
                dr = 1 if color else -1
                last = NUM_RANKS - 1 if color else 0
                start = 1 if color else NUM_RANKS - 2
                targets = []
                rr = r + dr
                if _on_board(rr, f) and self.cells[square(f, rr)] == '.':
                    targets.append(square(f, rr))
                    if RULES['pawn_double_step'] and r == start and self.cells[square(f, rr + dr)] == '.':
                        targets.append(square(f, rr + dr))
                for df in (-1, 1):
                    ff = f + df
                    if _on_board(rr, ff):
                        t = self.cells[square(ff, rr)]
                        if t != '.' and not self._is_own(t, color):
                            targets.append(square(ff, rr))
                for t in targets:
                    if t // 5 == last:
                        for p in promos:
                            yield Move(s, t, p)
                    else:
                        yield Move(s, t)
            elif kind in 'nk':
                for a, b in (KNIGHT_STEPS if kind == 'n' else KING_STEPS):
                    rr, ff = r + a, f + b
                    if _on_board(rr, ff) and not self._is_own(self.cells[square(ff, rr)], color):
                        yield Move(s, square(ff, rr))
            else:
                dirs = {'b': BISHOP_DIRS, 'r': ROOK_DIRS, 'q': BISHOP_DIRS + ROOK_DIRS}[kind]
                for a, b in dirs:
                    rr, ff = r + a, f + b
                    while _on_board(rr, ff):
                        t = self.cells[square(ff, rr)]
                        if self._is_own(t, color):
                            break
                        yield Move(s, square(ff, rr))
                        if t != '.':
                            break
                        rr, ff = rr + a, ff + b

    def _leaves_king_safe(self, move):
        color = self.turn
        saved_from, saved_to = self.cells[move.from_square], self.cells[move.to_square]
        self.cells[move.to_square] = saved_from
        self.cells[move.from_square] = '.'
        ks = self.king_square(color)
        ok = ks is None or not self.is_attacked_by(not color, ks)
        self.cells[move.from_square], self.cells[move.to_square] = saved_from, saved_to
        return ok

    def generate_legal_moves(self):
        for m in self._pseudo_legal():
            if self._leaves_king_safe(m):
                yield m

    @property
    def legal_moves(self):
        return list(self.generate_legal_moves())

    # ------------------------------------------------------------ make move
    def push(self, move):
        c = self.cells[move.from_square]
        captured = self.cells[move.to_square]
        zeroing = (c.lower() == 'p') or captured != '.'
        if move.promotion:
            sym = PIECE_SYMBOLS[move.promotion]
            c = sym.upper() if self.turn else sym
        self.cells[move.to_square] = c
        self.cells[move.from_square] = '.'
        self.move_stack.append(move)
        self.halfmove_clock = 0 if zeroing else self.halfmove_clock + 1
        if not self.turn:
            self.fullmove_number += 1
        self.turn = not self.turn
        if zeroing:
            self._history = []
        self._history.append(self._key())

    # ------------------------------------------------------------ results
    def is_checkmate(self):
        return self.is_check() and not any(self.generate_legal_moves())

    def is_stalemate(self):
        return not self.is_check() and not any(self.generate_legal_moves())

    def has_insufficient_material(self, color):
        own = [c.lower() for c in self.cells if self._is_own(c, color)]
        if any(k in 'prq' for k in own):
            return False
        if 'n' in own:
            others = [c.lower() for c in self.cells if c != '.' and not self._is_own(c, color)]
            return len(own) <= 2 and all(k in 'kq' for k in others)
        if 'b' in own:
            shades = set()
            for s in range(NUM_SQUARES):
                if self.cells[s].lower() == 'b':
                    shades.add((s % 5 + s // 5) & 1)
            everything = [c.lower() for c in self.cells if c != '.']
            return len(shades) == 1 and 'p' not in everything and 'n' not in everything
        return True

    def is_insufficient_material(self):
        return self.has_insufficient_material(WHITE) and self.has_insufficient_material(BLACK)

    def is_fivefold_repetition(self):
        return self._history.count(self._key()) >= 5

    def is_max_moves(self):
        return self.fullmove_number > RULES['max_fullmoves']

    def result(self):
        if self.is_checkmate():
            return '0-1' if self.turn else '1-0'
        if RULES['insufficient_material'] and self.is_insufficient_material():
            return '1/2-1/2'
        if not any(self.generate_legal_moves()):
            return '1/2-1/2'
        if self.is_max_moves():
            return '1/2-1/2'
        if RULES['fivefold_repetition'] and self.is_fivefold_repetition():
            return '1/2-1/2'
        return '*'
