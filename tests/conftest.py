import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def native():
    """Builds (if stale) and loads the native libraries; the tests fail when they cannot be built."""
    import __graft_entry__ as entry
    entry.build(stage_models=False)
    import optimal_control_problem_b200 as ocp
    return ocp


_problem_cache = {}


@pytest.fixture(scope="session")
def problems(native):
    """name -> (optimal_control_problem_b200.Problem, _oracle.OracleProblem), built lazily."""
    import _oracle

    def get(name, horizon=0):
        key = (name, horizon)
        if key not in _problem_cache:
            _problem_cache[key] = (native.Problem(name, horizon=horizon), _oracle.OracleProblem(name, horizon=horizon))
        return _problem_cache[key]

    return get
