"""Single-process mpi4py stand-in (oracle/shims/README.md)."""


class _Op:
    pass


class _Comm:
    def Get_size(self):
        return 1

    def Get_rank(self):
        return 0

    def Bcast(self, a, src):
        pass

    def Allreduce(self, a, b, op):
        b[...] = a

    def Barrier(self):
        pass


class MPI:
    COMM_WORLD = _Comm()
    MAX = _Op()
    SUM = _Op()
