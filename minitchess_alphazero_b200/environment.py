"""Drop-in for exp/environment.py: the same Environment / Episode surface, with the rules computed
by the CUDA kernels behind mc_legal_moves / mc_apply instead of the python-chess fork.

A single episode makes one tiny GPU call per step; the fast path for many games is `engine.Engine`.
"""
import numpy as np

from . import rules
from .erlyx_compat import BaseEnvironment, Episode, EpisodeStatus
from .moves import MOVES_DICT, MOVES_DICT_INV, NUM_ACTIONS  # noqa: F401  (re-exported like the reference)
from .rules import STARTING_FEN  # noqa: F401


class TerminatedEpisodeStepException(BaseException):       # exp/environment.py:8-9
    pass


class IlegalMoveException(BaseException):                  # exp/environment.py:12-13 (sic)
    pass


class MinitChessEpisode(Episode):
    """exp/environment.py:23-85."""

    def __init__(self, fen, rules_switches=None):
        self._rules = rules_switches
        self._state = rules.state_from_fen(fen)
        self._history = [self._rep_key()]
        self._update_attributes()

    def _rep_key(self):
        s = self._state
        return (int(s['pl0']), int(s['pl1']), int(s['pl2']), int(s['white']), int(s['meta']) & 1)

    def _update_attributes(self):                          # :34-50
        self._observation = rules.state_to_fen(self._state)
        codes, counts, results = rules.legal_moves(self._state, self._rules)
        self._legal_moves = codes[0, :counts[0]].astype(int).tolist()
        result = int(results[0])
        fivefold = self._rules.fivefold_repetition if self._rules is not None else 1
        if result == 0 and fivefold and self._history.count(self._history[-1]) >= 5:
            result = 3
        self._result = result
        if result in (1, 2):
            self._reward, self._done = 1., True
        elif result == 3:
            self._reward, self._done = 0., True
        else:
            self._reward, self._done = None, False

    def get_observation(self):
        return self._observation

    def get_reward(self):
        return self._reward

    def is_done(self):
        return self._done

    def get_legal_moves(self):
        return self._legal_moves

    def get_result(self):
        return rules.result_string(self._result)

    @property
    def turn(self):
        return bool(int(self._state['meta']) & 1)

    def step(self, action, return_status=True):            # :68-82
        if self.is_done():
            raise TerminatedEpisodeStepException
        out, status = rules.apply(self._state, np.uint16(action) if 0 <= int(action) < NUM_ACTIONS else np.uint16(0xffff),
                                  self._rules)
        if status[0] != 0:
            raise IlegalMoveException
        self._state = out[0]
        if ((int(self._state['meta']) >> 8) & 0xff) == 0:  # pawn move or capture: earlier positions cannot recur
            self._history = []
        self._history.append(self._rep_key())
        self._update_attributes()
        if return_status:
            return self.get_status()

    def get_status(self):
        return EpisodeStatus(self.get_observation(), self.get_reward(), self.is_done())


class MinitChessEnvironment(BaseEnvironment):              # exp/environment.py:88-91
    def __init__(self, rules_switches=None):
        self._rules = rules_switches

    def new_episode(self, fen=None):
        episode = MinitChessEpisode(fen or STARTING_FEN, self._rules)
        return episode, episode.get_observation()
