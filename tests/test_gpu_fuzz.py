"""Randomised differential test of the fused CUDA path against the CPU oracle (tests/fuzz_vs_oracle.py): random shapes
(1 .. 700 px, degenerate ones included), all seven colour spaces, block ranges 2 .. 256, random quality ranges, synthetic /
noise / flat inputs, both stream layouts.  30 cases per run; the command-line tool runs longer campaigns."""
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_fuzz_vs_oracle(seed):
    import fuzz_vs_oracle as F
    problems = F.run_cases(10, seed=seed, verbose=False)
    assert not problems, problems
