"""The ``falcon`` command -- mirror of ``falcon.falcon.main``
(/root/reference/falcon/falcon.py:70-245): read the peak files, preprocess, cluster per precursor
charge, make the labels of different charges disjoint, export the assignments (CSV with the ``#``
settings header, falcon.py:483-524) and optionally the cluster representatives (MGF).

Everything numeric runs on the GPU through ``falcon_b200.pipeline.HotPath`` (preprocessing,
vectorisation, nearest-neighbour search, DBSCAN + split, medoids); nothing is spilled to a work
directory.  Only MGF input is read in this build.
"""
from __future__ import annotations

import logging
import os
import sys
from typing import List, Optional

import numpy as np
import torch

import glob
import re

from . import __version__, pipeline, synth
from .config import config
from .ms_io import mgf_io

logger = logging.getLogger("falcon")


def _read_inputs(filenames: List[str]):
    sets, idents, files = [], [], []
    for fn in filenames:
        ext = os.path.splitext(fn.lower())[1]
        if ext != ".mgf":
            raise ValueError(f"Unsupported peak file format {ext!r} (this build reads .mgf)")
        sp, ids, fns = mgf_io.read_mgf(fn)
        sets.append(sp)
        idents.extend(ids)
        files.extend(os.path.basename(f) for f in fns)
    return synth.concat(sets), idents, files


def _write_cluster_info(path: str, rows) -> None:
    import pandas as pd

    with open(path, "w") as out:
        out.write(f"# falcon version {__version__}\n")
        for key in ("work_dir", "overwrite", "export_representatives"):
            out.write(f"# {key} = {config[key]}\n")
        out.write(f"# precursor_tol = {config.precursor_tol[0]:.2f} {config.precursor_tol[1]}\n")
        out.write(f"# rt_tol = {config.rt_tol}\n")
        out.write(f"# fragment_tol = {config.fragment_tol:.2f}\n")
        out.write(f"# eps = {config.eps:.3f}\n")
        for key in ("mz_interval", "low_dim", "n_neighbors", "n_neighbors_ann", "batch_size", "n_probe", "min_peaks"):
            out.write(f"# {key} = {config[key]}\n")
        out.write(f"# min_mz_range = {config.min_mz_range:.2f}\n")
        out.write(f"# min_mz = {config.min_mz:.2f}\n")
        out.write(f"# max_mz = {config.max_mz:.2f}\n")
        out.write(f"# remove_precursor_tol = {config.remove_precursor_tol:.2f}\n")
        out.write(f"# min_intensity = {config.min_intensity:.2f}\n")
        out.write(f"# max_peaks_used = {config.max_peaks_used}\n")
        out.write(f"# scaling = {config.scaling}\n#\n")
        pd.DataFrame(rows).to_csv(out, index=False)


def main(args: Optional[List[str]] = None) -> int:
    logging.basicConfig(format="{asctime} {levelname} [{name}/{processName}] {module}.{funcName} : {message}",
                        style="{", level=logging.INFO)
    config.parse(args)
    csv_path = f"{config.output_filename}.csv"
    # existing results are an error unless --overwrite (/root/reference/falcon/falcon.py:90-122), the representatives file included
    outputs = [csv_path] + ([f"{config.output_filename}.mgf"] if config.export_representatives else [])
    if not config.overwrite:
        for path in outputs:
            if os.path.isfile(path):
                logger.error("Output file %s already exists (use --overwrite)", path)
                return 1
    logger.info("falcon version %s", __version__)
    # input patterns are expanded like the reference does (/root/reference/falcon/falcon.py:262-264)
    input_filenames = [fn for pattern in config.input_filenames for fn in (sorted(glob.glob(pattern)) or [pattern])]
    raw, idents, files = _read_inputs(input_filenames)
    logger.info("Read %d spectra from %d peak file(s)", len(raw), len(input_filenames))
    settings = pipeline.Settings(
        precursor_tol_mass=config.precursor_tol[0], precursor_tol_mode=config.precursor_tol[1], rt_tol=config.rt_tol,
        fragment_tol=config.fragment_tol, eps=config.eps, mz_interval=config.mz_interval, low_dim=config.low_dim,
        n_neighbors=config.n_neighbors, n_neighbors_ann=config.n_neighbors_ann, batch_size=config.batch_size,
        n_probe=config.n_probe, min_mz=config.min_mz, max_mz=config.max_mz, exhaustive=config.exhaustive,
        representatives=config.export_representatives)
    # Under torchrun (one process per GPU) every charge is clustered by all GPUs together
    # (distributed.cluster_sharded: buckets dealt to the ranks, labels and representatives gathered); every rank
    # reads and preprocesses the input, rank 0 writes the results.
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    device = None
    if world > 1:
        import torch.distributed as dist

        device = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(device)
        if not dist.is_initialized():
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=torch.device("cuda", device))
    hp = pipeline.HotPath(settings, device)
    dev = hp.device
    up = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a, dtype=dt)).to(dev)  # noqa: E731
    if len(raw) == 0:
        logger.error("No spectra to cluster")
        return 1
    # ---- preprocessing on the device (spectrum.py:73-169)
    pmz_d, z_d = up(raw.precursor_mz, np.float64), up(raw.precursor_charge, np.int32)
    mz_d, in_d, indptr_d, valid = hp.preprocess(
        up(raw.mz, np.float32), up(raw.intensity, np.float32), up(raw.indptr, np.int64), pmz_d, z_d,
        # the window get_dim adjusted to the bin grid, as falcon.py:120-133 passes it on
        config.min_peaks, config.min_mz_range, hp.min_mz, hp.max_mz, config.remove_precursor_tol,
        config.min_intensity, config.max_peaks_used, config.scaling)
    valid_h = valid.cpu().numpy().astype(bool)
    has_charge = raw.precursor_charge != 0  # falcon clusters per charge; spectra without one are skipped
    logger.info("%d spectra pass the quality filters, %d lack a precursor charge",
                int(valid_h.sum()), int((valid_h & ~has_charge).sum()))
    counts = (indptr_d[1:] - indptr_d[:-1])
    rows = {"filename": [], "spectrum_id": [], "precursor_charge": [], "precursor_mz": [], "retention_time": [],
            "cluster": []}
    representatives, current_label = [], 0
    # ---- cluster per charge (falcon.py:151-203)
    for charge in sorted(set(raw.precursor_charge[valid_h & has_charge].tolist())):
        sel = np.flatnonzero(valid_h & (raw.precursor_charge == charge))
        sel_d = up(sel, np.int64)
        cnt = counts[sel_d]
        sub_indptr = torch.zeros(sel.shape[0] + 1, dtype=torch.int64, device=dev)
        torch.cumsum(cnt, 0, out=sub_indptr[1:])
        # gather the peak ranges of this charge (plumbing: index arithmetic in torch)
        owner = torch.repeat_interleave(torch.arange(sel.shape[0], device=dev), cnt)
        src = indptr_d[:-1][sel_d][owner] + (torch.arange(int(sub_indptr[-1]), device=dev) - sub_indptr[:-1][owner])
        rt = up(raw.retention_time[sel], np.float32) if config.rt_tol is not None else None
        if world > 1:
            from . import distributed as fdist

            part = synth.SpectrumSet(mz_d[src].cpu().numpy(), in_d[src].cpu().numpy(), sub_indptr.cpu().numpy(),
                                     raw.precursor_mz[sel], raw.precursor_charge[sel], raw.retention_time[sel])
            labels_np, n_clusters, reps_np = fdist.cluster_sharded(part, settings, device=dev)
            labels_h = labels_np.astype(np.int64)
        else:
            labels, n_clusters = hp.run(mz_d[src], in_d[src], sub_indptr, pmz_d[sel_d], z_d[sel_d], rt,
                                        max_peaks=int(cnt.max().item()))
            labels_h = labels.cpu().numpy().astype(np.int64)
            reps_np = hp.representatives.cpu().numpy() if config.export_representatives and n_clusters > 0 else None
        labels_h[labels_h >= 0] += current_label  # disjoint labels across charges, noise stays -1
        if config.export_representatives and n_clusters > 0:
            for r in reps_np:
                i = int(sel[r])
                a, b = int(raw.indptr[i]), int(raw.indptr[i + 1])
                representatives.append({
                    "identifier": f"mzspec:{config.usi_pxd}:{os.path.splitext(files[i])[0]}:scan:{idents[i]}",
                    "precursor_mz": float(raw.precursor_mz[i]), "precursor_charge": int(charge),
                    "retention_time": float(raw.retention_time[i]), "mz": raw.mz[a:b], "intensity": raw.intensity[a:b],
                    "cluster": int(labels_h[r])})
        current_label += n_clusters
        rows["filename"].extend(files[i] for i in sel)
        rows["spectrum_id"].extend(idents[i] for i in sel)
        rows["precursor_charge"].extend([int(charge)] * sel.shape[0])
        rows["precursor_mz"].extend(raw.precursor_mz[sel].tolist())
        rows["retention_time"].extend(raw.retention_time[sel].tolist())
        rows["cluster"].extend(labels_h.tolist())
        logger.info("charge %d: %d spectra grouped in %d clusters, %d spectra remain as singletons", charge,
                    int((labels_h >= 0).sum()), n_clusters, int((labels_h < 0).sum()))
    if not rows["cluster"]:
        logger.error("No valid spectra found for clustering")
        return 1
    if rank != 0:  # rank 0 exports
        return 0
    if config.singletons_as_clusters:  # development-head convention: noise -> new singleton cluster ids
        lab = np.asarray(rows["cluster"], np.int64)
        noise = np.flatnonzero(lab < 0)
        lab[noise] = current_label + np.arange(noise.shape[0])
        rows["cluster"] = lab.tolist()

    def natural(text):  # natsort's order (/root/reference/falcon/falcon.py:206-208): digit runs compare as numbers
        return [(0, int(t), "") if t.isdigit() else (1, 0, t) for t in re.split(r"(\d+)", str(text)) if t != ""]

    order = sorted(range(len(rows["cluster"])),
                   key=lambda i: (natural(rows["filename"][i]), natural(rows["spectrum_id"][i])))
    rows = {k: [v[i] for i in order] for k, v in rows.items()}
    logger.info("Export cluster assignments of %d spectra to %d unique clusters to output file %s",
                len(rows["cluster"]), current_label, csv_path)
    _write_cluster_info(csv_path, rows)
    if config.export_representatives:
        logger.info("Export %d cluster representative spectra to output file %s.mgf", len(representatives),
                    config.output_filename)
        mgf_io.write_spectra(f"{config.output_filename}.mgf", representatives)
    return 0


if __name__ == "__main__":
    sys.exit(main())
