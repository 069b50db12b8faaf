"""CPU: the logic the CUDA kernels share with the host (csrc/mcts_core.cuh, csrc/minitchess.cuh) under AddressSanitizer
and UndefinedBehaviorSanitizer.  compute-sanitizer is closed on the GPU pool (profiles/README.md), so this is the
memory-safety net under the tree code: whole self-play games, virtual loss, arenas that overflow (must raise the
capacity flag, never overrun) and the bitboard rules on dense synthetic positions -- tests/host_harness/sanitize_main.cpp."""
import os
import shutil
import subprocess

import pytest

from conftest import REPO


@pytest.mark.skipif(shutil.which('g++') is None, reason='needs g++')
def test_shared_tree_and_rules_code_is_clean_under_asan_ubsan():
    hh = os.path.join(REPO, 'tests', 'host_harness')
    exe = os.path.join(hh, '_build', 'sanitize_main')
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    build = subprocess.run(['g++', '-O1', '-g', '-std=c++17', '-fsanitize=address,undefined', '-fno-sanitize-recover=all',
                            '-fno-omit-frame-pointer', '-ffp-contract=off', '-I', os.path.join(REPO, 'include'),
                            '-I', os.path.join(REPO, 'minitchess_alphazero_b200', 'csrc'), os.path.join(hh, 'sanitize_main.cpp'), '-o', exe],
                           capture_output=True, text=True, timeout=600)
    if build.returncode != 0 and 'sanitize' in build.stderr.lower() and 'cannot find' in build.stderr.lower():
        pytest.skip('libasan / libubsan are not installed')
    assert build.returncode == 0, build.stderr[-3000:]
    env = dict(os.environ, ASAN_OPTIONS='detect_leaks=1:abort_on_error=0', UBSAN_OPTIONS='print_stacktrace=1')
    run = subprocess.run([exe], capture_output=True, text=True, timeout=600, env=env)
    assert run.returncode == 0, run.stdout[-2000:] + run.stderr[-4000:]
    assert 'sanitize ok' in run.stdout and 'runtime error' not in run.stderr and 'AddressSanitizer' not in run.stderr
