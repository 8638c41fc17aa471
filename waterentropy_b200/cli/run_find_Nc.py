"""`waterShells` command line: print the RAD shells of interfacial waters per frame
(reference `cli/run_find_Nc.py:16-141`, same -top/-crd/-s/-e/-dt flags)."""
from __future__ import annotations

import argparse
import logging
import sys
from datetime import datetime

from waterentropy_b200.cli.runWaterEntropy import load_universe
from waterentropy_b200.recipes import interfacial_solvent as GetSolvent


def main():
    p = argparse.ArgumentParser(prog="waterShells")
    p.add_argument("-top", "--file-topology", metavar="file", default=None)
    p.add_argument("-crd", "--file-coords", metavar="file", default=None)
    p.add_argument("-s", "--start", metavar="int", type=int, default=0)
    p.add_argument("-e", "--end", metavar="int", type=int, default=1)
    p.add_argument("-dt", "--step", metavar="int", type=int, default=1)
    args = p.parse_args()
    if args.file_topology is None or args.file_coords is None:
        logging.error("a topology (-top) and a coordinates file (-crd) are required")
        sys.exit(1)
    start_time = datetime.now()
    print(start_time)
    u = load_universe(args.file_topology, args.file_coords)
    shells = GetSolvent.get_interfacial_shells(u, args.start, args.end, args.step)
    GetSolvent.print_frame_solvent_shells(shells)
    print(datetime.now() - start_time)


if __name__ == "__main__":
    main()
