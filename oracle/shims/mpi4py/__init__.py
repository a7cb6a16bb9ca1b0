"""TEST INFRASTRUCTURE ONLY -- single-rank `mpi4py` stand-in (algos/multiagent/rl_tools/mpi_tools.py:1)."""
from . import MPI  # noqa: F401
