"""Run the reference's UNMODIFIED exp/*.py on the oracle shims (build container only).

`/root/reference` does not exist on the GPU box, so nothing imported at run time by the
`-m gpu` tests, smoke() or bench.py may call this; it is used by `tests/golden/make_golden.py`
to produce committed fixtures and by CPU tests that skip when the reference is absent.
"""
import contextlib
import importlib
import os
import sys

REFERENCE_ROOT = os.environ.get('MCAZ_REFERENCE_ROOT', '/root/reference')
SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'shims')


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, 'exp', 'agent.py'))


@contextlib.contextmanager
def _cwd(path):
    old = os.getcwd()
    os.chdir(path)
    try:
        yield
    finally:
        os.chdir(old)


def load_reference():
    """Returns (exp.agent, exp.environment, exp.policy) imported from /root/reference."""
    if not reference_available():
        raise RuntimeError('reference tree not found at %s' % REFERENCE_ROOT)
    for p in (SHIMS, REFERENCE_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    # exp/environment.py:16 opens 'moves_dict.json' relative to the cwd.
    with _cwd(os.path.join(REFERENCE_ROOT, 'exp')):
        env = importlib.import_module('exp.environment')
    agent = importlib.import_module('exp.agent')
    policy = importlib.import_module('exp.policy')
    return agent, env, policy
