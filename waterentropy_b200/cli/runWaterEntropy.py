"""`waterEntropy` command line on the B200 path.

Keeps the analysis flags of the reference's console script
(`cli/runWaterEntropy.py:102-160`: -top/--file-topology, -crd/--file-coords, -s/--start,
-e/--end, -dt/--step, -temp/--temperature, -p/--parallel) and its printed output
(`Number of frames analysed`, the orientation table, the vibration table).  The
SLURM / dask-jobqueue flags (--hpc*, --conda-*) are accepted and ignored with a
notice: frame parallelism here is one process per GPU under torchrun, and the
parser does not hard-exit outside a conda environment like upstream does
(`cli/runWaterEntropy.py:57-86`).
"""
from __future__ import annotations

import argparse
import logging
import sys
from datetime import datetime

from waterentropy_b200.entropy import orientations as OR
from waterentropy_b200.entropy import vibrations as VIB
from waterentropy_b200.recipes import interfacial_solvent as GetSolvent


def build_parser():
    p = argparse.ArgumentParser(
        prog="waterEntropy",
        description="Interfacial water orientational + vibrational entropy (B200-native hot path).")
    p.add_argument("-top", "--file-topology", metavar="file", default=None,
                   help="name of file containing system topology.")
    p.add_argument("-crd", "--file-coords", metavar="file", default=None,
                   help="name of file containing positions and forces in a single file.")
    p.add_argument("-s", "--start", metavar="int", type=int, default=0, help="frame number to start analysis from.")
    p.add_argument("-e", "--end", metavar="int", type=int, default=1, help="frame number to end analysis at.")
    p.add_argument("-dt", "--step", metavar="int", type=int, default=1,
                   help="steps to take between start and end frame selections.")
    p.add_argument("-temp", "--temperature", metavar="float", type=float, default=298,
                   help="Target temperature the simulation was performed at in Kelvin.")
    p.add_argument("-p", "--parallel", action="store_true",
                   help="shard frames over the GPUs of a torchrun job (one process per GPU).")
    p.add_argument("--bulk", action="store_true", help="run the bulk-water recipe instead of the interfacial one.")
    # accepted for command-line compatibility with upstream's SLURM mode
    for flag in ("--conda-exec", "--conda-env", "--conda-path", "--hpc-account", "--hpc-constraint", "--hpc-cores",
                 "--hpc-memory", "--hpc-nodes", "--hpc-processes", "--hpc-queue", "--hpc-qos", "--hpc-walltime"):
        p.add_argument(flag, default=None, help=argparse.SUPPRESS)
    p.add_argument("--hpc", action="store_true", help=argparse.SUPPRESS)
    p.add_argument("--submit", action="store_true", help=argparse.SUPPRESS)
    return p


def load_universe(topology, coords):
    try:
        from MDAnalysis import Universe  # the reference's reader, if the user has it
    except Exception:
        from waterentropy_b200.io import Universe
    return Universe(topology, coords)


def run_waterEntropy(args):
    start_time = datetime.now()
    print(start_time)
    rank = 0
    if args.parallel:
        import os

        if "RANK" in os.environ:
            import torch
            import torch.distributed as dist

            torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
            dist.init_process_group("nccl")
            rank = dist.get_rank()
    u = load_universe(args.file_topology, args.file_coords)
    if args.bulk:
        from waterentropy_b200.recipes.bulk_water import get_bulk_water_orient_entropy

        Sorient_dict, covariances, vibrations = get_bulk_water_orient_entropy(
            u, args.start, args.end, args.step, args.temperature)
        n_frames = len(range(args.start, args.end, args.step))
    else:
        Sorient_dict, covariances, vibrations, _fsi, n_frames = GetSolvent.get_interfacial_water_orient_entropy(
            u, args.start, args.end, args.step, args.temperature, args.parallel)
    if rank == 0:
        print(f"Number of frames analysed: {n_frames}")
        OR.print_Sorient_dicts(Sorient_dict)
        VIB.print_Svib_data(vibrations, covariances)
        print(datetime.now() - start_time)


def main():
    args = build_parser().parse_args()
    if args.file_topology is None or args.file_coords is None:
        logging.error("a topology (-top) and a coordinates+forces file (-crd) are required")
        sys.exit(1)
    if args.hpc or args.submit:
        print("note: --hpc/--submit (SLURM + dask-jobqueue orchestration) is out of scope of the B200 path; "
              "running on the local GPU(s) instead")
    run_waterEntropy(args)


if __name__ == "__main__":
    main()
