"""TEST INFRASTRUCTURE ONLY -- one-rank COMM_WORLD: Allreduce copies, Bcast is a no-op
(algos/multiagent/rl_tools/mpi_tools.py:38-95)."""
import numpy as np

SUM, MIN, MAX = "sum", "min", "max"


class _Comm:
    def Get_rank(self):
        return 0

    def Get_size(self):
        return 1

    def Allreduce(self, x, buff, op=SUM):
        np.copyto(buff, x)

    def Bcast(self, x, root=0):
        return None


COMM_WORLD = _Comm()
