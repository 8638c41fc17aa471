"""`waterShells` command line: print the RAD shells of interfacial waters per frame.

Reference: `cli/run_find_Nc.py:16-141` - same function names, the same -top/-crd/-s/-e/-dt flags and
the same printed lines (`frame atom_idx [shell indices] n`); the shells come from the CUDA path
(`get_interfacial_shells`, which needs positions only: a coordinate file without forces is fine).
"""
from __future__ import annotations

import argparse
import logging
import sys
from datetime import datetime

from waterentropy_b200.cli.runWaterEntropy import load_universe
from waterentropy_b200.recipes import interfacial_solvent as GetSolvent


def run_waterEntropy(file_topology="file_topology", file_coords="file_coords", file_forces="file_forces",
                     file_energies="file_energies", list_files="list_files", start="start", end="end", step="step"):
    """Reference: `cli/run_find_Nc.py:16-45` (keyword names kept, the unused ones included)."""
    start_time = datetime.now()
    print(start_time)
    u = load_universe(file_topology, file_coords)
    print(u.trajectory)
    frame_solvent_shells = GetSolvent.get_interfacial_shells(u, start, end, step)
    GetSolvent.print_frame_solvent_shells(frame_solvent_shells)
    sys.stdout.flush()
    print(datetime.now() - start_time)


def build_parser():
    p = argparse.ArgumentParser(
        prog="waterShells",
        description="Coordination shells (RAD) of the water molecules at solute interfaces (B200-native hot path).",
        formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    p.add_argument("-top", "--file-topology", metavar="file", default=None, help="name of file containing system topology.")
    p.add_argument("-crd", "--file-coords", metavar="file", default=None, help="name of file containing positions.")
    p.add_argument("-s", "--start", metavar="int", type=int, default=0, help="frame number to start analysis from.")
    p.add_argument("-e", "--end", metavar="int", type=int, default=1, help="frame number to end analysis at.")
    p.add_argument("-dt", "--step", metavar="int", type=int, default=1,
                   help="steps to take between start and end frame selections.")
    return p


def main(argv=None):
    args = build_parser().parse_args(argv)
    if args.file_topology is None or args.file_coords is None:
        logging.error("a topology (-top) and a coordinates file (-crd) are required")
        sys.exit(1)
    run_waterEntropy(file_topology=args.file_topology, file_coords=args.file_coords, start=args.start,
                     end=args.end, step=args.step)


if __name__ == "__main__":
    main()
