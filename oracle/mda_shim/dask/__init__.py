"""TEST INFRASTRUCTURE ONLY -- stub so `from dask.distributed import Client` resolves."""
