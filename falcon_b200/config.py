"""Command-line and INI configuration -- mirror of ``falcon.config``
(/root/reference/falcon/config.py:38-183; the published 0.1.x flag set of SURVEY A.6 where the
snapshot replaced it).  ``configargparse`` is not installed here: plain ``argparse`` plus a small
reader for ``config.ini`` / ``-c FILE`` (``key = value`` lines, command line wins).
"""
from __future__ import annotations

import argparse
import os
from typing import List, Optional

from . import __version__  # noqa: E402  (one version string for the package and the CSV header)


def _ini_args(path: str) -> List[str]:
    args: List[str] = []
    with open(path) as fh:
        for line in fh:
            line = line.split("#", 1)[0].split(";", 1)[0].strip()
            if not line or line.startswith("["):
                continue
            key, _, value = line.partition("=")
            key, value = key.strip(), value.strip()
            if value.lower() in ("true", "yes", "on"):
                args.append(f"--{key}")
            elif value.lower() in ("false", "no", "off") and key in ("overwrite", "export_representatives"):
                continue
            else:
                args.append(f"--{key}")
                args.extend(value.split())
    return args


class Config:
    """Settings as attributes (``config.eps``) after ``parse``; same names and defaults as falcon."""

    def __init__(self) -> None:
        p = argparse.ArgumentParser(
            prog="falcon", description="falcon: fast spectrum clustering using nearest neighbor searching "
                                       f"(B200 build {__version__})")
        p.add_argument("-c", "--config", default=None, help="INI file with settings (default: ./config.ini if present)")
        # IO
        p.add_argument("input_filenames", nargs="+", help="Input peak files (this build reads .mgf).")
        p.add_argument("output_filename", help="Output file name (without extension).")
        p.add_argument("--work_dir", default=None, help="Working directory (unused: nothing is spilled to disk).")
        p.add_argument("--overwrite", action="store_true", help="Overwrite existing results.")
        p.add_argument("--export_representatives", action="store_true",
                       help="Export cluster representatives to an MGF file.")
        p.add_argument("--usi_pxd", default="USI000000", help="ProteomeXchange dataset identifier for USIs.")
        # NN index / clustering (published 0.1.x flags)
        p.add_argument("--precursor_tol", nargs=2, default=[20, "ppm"], help='Precursor tolerance mass and mode.')
        p.add_argument("--rt_tol", type=float, default=None, help="Retention time tolerance (default: none).")
        p.add_argument("--fragment_tol", type=float, default=0.05, help="Fragment mass tolerance in m/z.")
        p.add_argument("--eps", type=float, default=0.1, help="The eps parameter (cosine distance) for DBSCAN.")
        p.add_argument("--distance_threshold", type=float, default=None,
                       help="Alias of --eps (the snapshot's name, config.py:98-105).")
        p.add_argument("--mz_interval", type=int, default=1, help="Precursor m/z interval of the buckets.")
        p.add_argument("--low_dim", type=int, default=400, help="Low-dimensional vector length.")
        p.add_argument("--n_neighbors", type=int, default=64, help="Neighbors in the pairwise distance matrix.")
        p.add_argument("--n_neighbors_ann", type=int, default=128, help="Neighbors retrieved from the index.")
        p.add_argument("--batch_size", type=int, default=2 ** 16, help="Batch size (kept for compatibility).")
        p.add_argument("--n_probe", type=int, default=32, help="Maximum number of lists to probe per query.")
        p.add_argument("--exhaustive", action="store_true", help="Probe every list (n_probe = n_list).")
        p.add_argument("--singletons_as_clusters", action="store_true",
                       help="Give every unclustered spectrum its own cluster id in the CSV, as the development "
                            "head of falcon does (cluster.py:144-155), instead of -1 (published releases).")
        # The development head of falcon replaced the nearest-neighbour + DBSCAN pipeline by exact
        # hierarchical clustering (config.py:96-124).  This build implements the published pipeline: those
        # flags are recognised and refused, not silently ignored.
        p.add_argument("--linkage", default=None, help=argparse.SUPPRESS)
        p.add_argument("--min_matched_peaks", default=None, help=argparse.SUPPRESS)
        # preprocessing (config.py:127-183)
        p.add_argument("--min_peaks", default=5, type=int)
        p.add_argument("--min_mz_range", default=250.0, type=float)
        p.add_argument("--min_mz", default=101.0, type=float)
        p.add_argument("--max_mz", default=1500.0, type=float)
        p.add_argument("--remove_precursor_tol", default=1.5, type=float)
        p.add_argument("--min_intensity", default=0.01, type=float)
        p.add_argument("--max_peaks_used", default=50, type=int)
        p.add_argument("--scaling", default="off", type=str, choices=["off", "root", "log", "rank"])
        self._parser = p
        self._namespace: Optional[dict] = None

    def parse(self, args: Optional[List[str]] = None) -> None:
        import sys

        argv = list(sys.argv[1:] if args is None else args)
        ini = None
        for flag in ("-c", "--config"):
            if flag in argv:
                ini = argv[argv.index(flag) + 1]
        if ini is None and os.path.exists("config.ini"):
            ini = "config.ini"
        if ini is not None:
            argv = _ini_args(ini) + argv  # the command line comes last and wins
        ns = vars(self._parser.parse_args(argv))
        ns["precursor_tol"] = [float(ns["precursor_tol"][0]), str(ns["precursor_tol"][1])]
        if ns["precursor_tol"][1] not in ("ppm", "Da"):
            raise ValueError("Unknown precursor tolerance mode")
        for flag in ("linkage", "min_matched_peaks"):
            if ns[flag] is not None:
                raise ValueError(f"--{flag} belongs to the hierarchical-clustering pipeline of falcon's development "
                                 "head; this build clusters with the published nearest-neighbour + DBSCAN pipeline "
                                 "(--eps, --n_probe, ...)")
        if ns["distance_threshold"] is not None:
            ns["eps"] = ns["distance_threshold"]
        if ns["n_neighbors_ann"] < ns["n_neighbors"]:
            raise ValueError("n_neighbors_ann should be equal or greater than n_neighbors")
        self._namespace = ns

    def __getattr__(self, option):
        ns = self.__dict__.get("_namespace")
        if ns is None:
            raise RuntimeError("The configuration has not been initialized")
        return ns[option]

    def __getitem__(self, item):
        return self.__getattr__(item)


config = Config()
