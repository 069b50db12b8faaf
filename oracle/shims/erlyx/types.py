from collections import namedtuple

ActionData = namedtuple('ActionData', ['action', 'info'])
EpisodeStatus = namedtuple('EpisodeStatus', ['observation', 'reward', 'done'])
