"""mpi4py stand-in (oracle/shims/README.md).

`theforce.distributed` falls back to `theforce._mpi4py` when torch has no MPI backend
(distributed.py:4-10); that module needs only `COMM_WORLD.{Get_size, Get_rank, Bcast,
Allreduce, Barrier}` and `MPI.{SUM, MAX}` (_mpi4py.py:5-60).

* default: a single-process world (size 1, rank 0);
* with SGPR_SHIM_WORLD / SGPR_SHIM_RANK / SGPR_SHIM_PORT in the environment (set by
  oracle/ref_bench.py for its worker processes): a real multi-process world whose
  collectives run over torch.distributed's gloo backend on 127.0.0.1 -- this is how the
  reference's own MPI decomposition over atoms (descriptor/atoms.py:228-259,321-341;
  calculator/active.py:562,601-602) is exercised on the host cores without MPI.
"""
import os


class _Op:
    def __init__(self, name):
        self.name = name


class _Comm:
    def __init__(self):
        self._world = int(os.environ.get("SGPR_SHIM_WORLD", "1"))
        self._rank = int(os.environ.get("SGPR_SHIM_RANK", "0"))
        self._pg = None

    def _group(self):
        if self._pg is None:
            import datetime

            import torch.distributed as dist

            if not dist.is_initialized():
                dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{}".format(os.environ["SGPR_SHIM_PORT"]),
                                        rank=self._rank, world_size=self._world, timeout=datetime.timedelta(seconds=1800))
            self._pg = dist
        return self._pg

    def Get_size(self):
        return self._world

    def Get_rank(self):
        return self._rank

    def Bcast(self, a, src):
        if self._world == 1:
            return
        import torch

        t = torch.from_numpy(a) if a.flags.writeable and a.flags.c_contiguous else torch.tensor(a)
        self._group().broadcast(t, src)
        if t.data_ptr() != a.ctypes.data:
            a[...] = t.numpy()

    def Allreduce(self, a, b, op):
        if self._world == 1:
            b[...] = a
            return
        import torch

        dist = self._group()
        t = torch.tensor(a)
        dist.all_reduce(t, op=dist.ReduceOp.MAX if getattr(op, "name", "SUM") == "MAX" else dist.ReduceOp.SUM)
        b[...] = t.numpy()

    def Barrier(self):
        if self._world > 1:
            self._group().barrier()


class MPI:
    COMM_WORLD = _Comm()
    MAX = _Op("MAX")
    SUM = _Op("SUM")
