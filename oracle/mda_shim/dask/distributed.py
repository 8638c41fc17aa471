"""TEST INFRASTRUCTURE ONLY -- inline stand-in for `dask.distributed.Client`.

Implements the calls `/root/reference/waterEntropy/recipes/interfacial_solvent.py:219-240`
makes: Client(processes=, threads_per_worker=, silence_logs=), scheduler_info(),
map(), gather(), restart(), close().  Tasks run in a multiprocessing pool (one
process per core, like the reference's LocalCluster) or inline when
WE_SHIM_INLINE=1.
"""
import multiprocessing as mp
import os


class Client:
    def __init__(self, *args, n_workers=None, **kwargs):
        self._n = n_workers or max(1, (os.cpu_count() or 2) // 1)
        self._inline = os.environ.get("WE_SHIM_INLINE", "0") == "1"
        self._pool = None

    def scheduler_info(self):
        return {"n_workers": self._n, "workers": {i: {} for i in range(self._n)}}

    def map(self, fn, args):
        args = list(args)
        if self._inline:
            return [fn(a) for a in args]
        if self._pool is None:
            self._pool = mp.get_context("fork").Pool(self._n)
        return self._pool.map_async(fn, args)

    def gather(self, futures):
        if isinstance(futures, list):
            return futures
        return futures.get()

    def restart(self):
        if self._pool is not None:
            self._pool.terminate()
            self._pool = None

    def close(self):
        self.restart()
