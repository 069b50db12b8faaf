"""Minimal stand-in for `erlyx` (github.com/schouhy/erlyx@0f6409c3, NOT in the reference tree).

TEST INFRASTRUCTURE ONLY.  The reference imports base classes and `run_episodes` from erlyx
(`exp/agent.py:1-2`, `exp/environment.py:2,4`, `exp/policy.py:1`, `exp/callbacks.py:3`,
`exp/learner.py:8-9`, `app/base.py:14`).  The real package is unavailable offline, so the
episode-loop contract below is INFERRED from how the reference uses it:

    env.new_episode() -> cb.on_episode_begin(obs) -> loop{ agent.select_action(obs);
    episode.step(action.action); cb.on_step_end(action, obs', reward, done) } -> cb.on_episode_end()

A truthy return value from `on_episode_end` aborts the remaining episodes
(`app/base.py:57,62` returns True from `MQTTDataset.push`, forwarded by `exp/callbacks.py:54`).
"""
from .types import ActionData, EpisodeStatus  # noqa: F401


def run_episodes(environment, agent, n_episodes, callbacks=None, use_tqdm=False):
    callbacks = list(callbacks or [])
    for _ in range(n_episodes):
        episode, observation = environment.new_episode()
        for cb in callbacks:
            cb.on_episode_begin(observation)
        done = False
        while not done:
            action = agent.select_action(observation)
            observation, reward, done = episode.step(action.action)
            for cb in callbacks:
                cb.on_step_end(action, observation, reward, done)
        stop = False
        for cb in callbacks:
            if cb.on_episode_end():
                stop = True
        if stop:
            break
