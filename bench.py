#!/usr/bin/env python
"""Benchmark of the falcon clustering hot path (BASELINE.json metric: spectra/sec).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W    # CPU reference arm (oracle)

One "step" = one pass of the whole hot path (buckets -> vectors -> IVF -> scan ->
k-NN CSR -> DBSCAN -> precursor split [-> NCCL label gather]) over the workload
of BASELINE.json configs[1]: 1 M synthetic spectra with falcon's defaults per
GPU.  With N > 1 every rank clusters its own 1 M spectra (precursor buckets are
independent units, so there is no data-path collective; "weak" scaling) and the
labels are gathered over NCCL inside the step.

Prints ONE JSON line (rank 0).  `value` = spectra/s with the inputs resident in
HBM; `e2e` = the same through the host-buffer API (pinned H2D + D2H inside the
timed region); `roofline` = the dominant hand-written kernel against the
measured peaks of MEASURED_PEAKS.json; `cpu_baseline` = the CPU oracle timed on
a bounded sample of the same workload on this box's host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "spectra_per_sec_end_to_end"
UNIT = "spectra/s"
NCU_TRAFFIC_FILE = "r02_ncu_traffic.json"  # DRAM bytes per launch from this round's ncu --set full captures
FALLBACK_PEAKS = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        d["_source"] = "measured"
        return d
    d = dict(FALLBACK_PEAKS)
    d["_source"] = "fallback"
    return d


# --------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index: int):
        self.idx = device_index
        self.proc = None
        self.path = None
        self.skip = 0

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.idx)], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def wait_first_sample(self, timeout_s: float):
        t0 = time.time()
        while self.proc is not None and time.time() - t0 < timeout_s:
            try:
                if os.path.getsize(self.path) > 0:
                    return
            except OSError:
                return
            time.sleep(0.02)

    def mark(self):
        try:
            self.skip = sum(1 for _ in open(self.path))
        except OSError:
            self.skip = 0

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln, line in enumerate(open(self.path)):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9 or ln < self.skip:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        os.unlink(self.path)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


# --------------------------------------------------------------------------- CPU reference
def bucket_sample(spectra, target: int):
    """A contiguous range of whole precursor buckets of `spectra` holding about
    `target` spectra (same bucket-size distribution as the full workload)."""
    from oracle import ivf as oivf

    order, bptr, _ = oivf.bucket_sort(spectra.precursor_mz, spectra.precursor_charge)
    nb = bptr.shape[0] - 1
    b0 = nb // 3
    b1 = int(np.searchsorted(bptr, bptr[b0] + target, "left"))
    b1 = min(max(b1, b0 + 1), nb)
    idx = order[bptr[b0]: bptr[b1]]
    return spectra.take(np.sort(idx)), b1 - b0


def cpu_reference_time(sample, cores: int, exhaustive: bool):
    from oracle import pipeline as opipe

    t0 = time.perf_counter()
    labels, stages, _ = opipe.run(sample, n_jobs=cores, exhaustive=exhaustive)
    return time.perf_counter() - t0, stages, int(labels.max()) + 1


def run_reference_arm(args):
    """CPU arm: the oracle (restatement of the reference's faiss/numba/sklearn
    path; the reference itself cannot be installed offline) on a bounded sample
    of the same workload, all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from falcon_b200 import synth

    cores = len(os.sched_getaffinity(0))
    kw = {"mass_range": tuple(args.mass_range)} if args.mass_range else {}
    full = synth.generate(args.n, 42, **kw)
    sample, n_buckets = bucket_sample(full, args.cpu_sample)
    times = []
    for i in range(args.warmup + args.steps):
        dt, stages, _ = cpu_reference_time(sample, cores, args.exhaustive)
        log(f"[reference] step {i}: {dt:.2f}s {stages}")
        if i >= args.warmup:
            times.append(dt)
    t = float(np.mean(times))
    value = len(sample) / t
    desc = (f"{len(sample)} spectra = {n_buckets} whole precursor buckets of the {args.n}-spectrum workload")
    out = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc,
                         "note": "CPU restatement (faiss-cpu/lance/fastcluster unavailable offline): "
                                 "numpy/OpenBLAS per-bucket IVF + sklearn dbscan_inner, joblib over buckets"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(out), flush=True)


def workload_config(args):
    return {
        "workload": (f"{args.total} synthetic spectra in total, precursor-mass range split across the GPUs, "
                     if args.total else "") +
                    f"{args.n} synthetic spectra per GPU (seed 42+rank, charge 2/3, 101-1500 m/z, <=50 peaks), "
                    "falcon defaults low_dim=400 eps=0.10 n_probe=32 n_neighbors=64/128 precursor_tol=20ppm"
                    + (" exhaustive (n_probe=nlist)" if args.exhaustive else "")
                    + (f" neutral mass {args.mass_range[0]:g}-{args.mass_range[1]:g} Da" if args.mass_range else ""),
        "baseline_config": "configs[1]: 1M synthetic spectra, defaults, single B200",
        "spectra_per_gpu": args.n, "low_dim": 400, "eps": 0.1, "n_probe": 32, "exhaustive": bool(args.exhaustive),
        "l2": "inputs larger than L2 (peaks 270 MB, vectors 2.4 GB per step vs 126 MB L2); no flush",
    }


# --------------------------------------------------------------------------- roofline
def kernel_rooflines(stats: dict, peaks: dict) -> dict:
    """Algorithmic work per launch of every hand-written kernel."""
    n, p, d, ldb = stats["n"], stats["n_peaks"], stats["low_dim"], stats["ld_bf16"]
    pairs, nnz = stats["n_pairs"], stats["nnz"]
    hbm = peaks["hbm_gbs"]
    work = {
        # bytes: peaks (m/z + intensity) + indptr + order + f32 row + bf16 row
        "vectorize": ("hbm", p * 8 + (n + 1) * 8 + n * 4 + n * d * 4 * stats.get("dense_f32", 1) + n * ldb * 2
                      + n * (stats.get("ell_width", 0) * 6 + 2)),
        # HBM floor of the scan: every bf16 row read once, pairs written once
        "scan_tc": ("hbm", n * ldb * 2 + pairs * 8),
        # re-score + selection (refine_block and, for the query groups it defers, refine: one logical launch):
        # pairs in / out, every sparse row ONCE (a candidate row is re-read once per pair it takes part in, but
        # those re-reads are served by L2 -- ncu: 0.38 GB of DRAM reads at 1 M), m/z + row counts, CSR entries out
        "refine_block": ("hbm", pairs * 16 + n * (stats.get("ell_width", 0) * 6 + 8 + 8 + 4) + nnz * 8),
        "pair_hist": ("hbm", pairs * 8 + pairs * 4),
        "pair_scatter": ("hbm", pairs * 16 + pairs * 12),
        "csr_compact": ("hbm", nnz * 16 + (n + 1) * 24),
        "dbscan_core": ("hbm", nnz * 4 + (n + 1) * 8 + n * 5),
        "dbscan_propagate": ("hbm", nnz * 8 + (n + 1) * 8 + n * 9),
        "ivf_assign": ("hbm", n * d * 4 + n * 4 * (1 + stats.get("max_nprobe", 1))),
        # fused trainer (both size classes = one logical launch): every populated sparse slot
        # (6 bytes) + the row population read once, centroids + list_id + probes written once
        "kmeans_fused": ("hbm", stats.get("ivf_nnz", 0) * 6 + stats.get("ivf_rows", 0) * 2
                         + stats.get("total_centroids", 0) * d * 4
                         + stats.get("ivf_rows", 0) * 4 * (1 + stats.get("max_nprobe", 1))),
        # tiled trainer, per iteration: sparse rows + previous assignment in, 8-byte atomics out
        "kmeans_tiled_assign": ("hbm", stats.get("ivf_nnz", 0) * 6 + stats.get("ivf_rows", 0) * 10
                                + stats.get("ivf_nnz", 0) * 8),
        # tensor-core assignment from the sparse rows, per iteration: whole ELL rows (padding included: the
        # kernel loads a tile's block contiguously) + populations in, one int32 per row out.  Credited with
        # every IVF row, i.e. exact when all IVF buckets are tiled (the dense-bucket workloads)
        "kmeans_tc_sparse": ("hbm", stats.get("ivf_rows", 0) * (stats.get("ell_width", 0) * 6 + 2 + 4)),
        # float32 pass of the tiled final assignment: sparse rows in, list + probes out
        "ivf_assign_tiled_f32": ("hbm", stats.get("ivf_nnz", 0) * 6 + stats.get("ivf_rows", 0) * 2
                                 + stats.get("ivf_rows", 0) * 4 * (1 + stats.get("max_nprobe", 1))),
    }
    out = {}
    if "kmeans_fused" in stats["kernels"]:  # some IVF buckets are not tiled: the tiled kernels' row counts are unknown
        work.pop("kmeans_tc_sparse")
        work.pop("ivf_assign_tiled_f32")
    for name, (bound, bytes_) in work.items():
        if name in stats["kernels"]:
            ms, launches = stats["kernels"][name]
            if launches and ms > 0:
                ach = bytes_ / (ms / launches * 1e-3) / 1e9
                out[name] = {"bound": bound, "achieved": ach, "peak": hbm, "unit": "GB/s", "frac": ach / hbm,
                             "ms_per_launch": ms / launches, "launches_per_step": launches / stats["steps"],
                             "algorithmic_bytes_per_launch": bytes_}
    if "vectorize" in out:
        # SURVEY 8(d)'s per-unit figure counts the dense float32 row the reference writes (N * d * 4); this kernel
        # emits the sparse row instead, so `frac` above is over the bytes it really moves and this is the same time
        # against the survey's formula: peaks + indptr + f32 row + bf16 row
        b8d = p * 8 + (n + 1) * 8 + n * d * 4 + n * ldb * 2
        a8d = b8d / (out["vectorize"]["ms_per_launch"] * 1e-3) / 1e9
        out["vectorize"]["survey_8d"] = {"algorithmic_bytes_per_launch": b8d, "achieved": a8d, "frac": a8d / hbm}
    # tensor view of the scan: FLOPs the IVF semantics require
    if "scan_tc" in stats["kernels"]:
        ms, launches = stats["kernels"]["scan_tc"]
        tf = 2.0 * d * stats["required_pairs"] / (ms / launches * 1e-3) / 1e12
        pk = peaks["bf16_tflops"]
        # The tensor roofline is over the FLOPs the tensor pipe EXECUTES (`computed_pairs`: whole buckets, but only
        # the tiles at or above the diagonal of the symmetric product) -- a fraction of peak by construction <= 1.
        # `required` is SURVEY 8(d)'s count, the (query, member) pairs the IVF semantics ask for: in exhaustive
        # mode the symmetric kernel executes about half of them (ratio > 1), with the IVF index it executes a
        # superset (ratio < 1); it is reported as a rate, not as a fraction of peak.
        ex = 2.0 * d * stats["computed_pairs"] / (ms / launches * 1e-3) / 1e12
        out["scan_tc"]["tensor"] = {"achieved": ex, "peak": pk, "unit": "TFLOP/s", "frac": ex / pk,
                                    "executed_frac": ex / pk,
                                    "required_pairs": stats["required_pairs"],
                                    "computed_pairs": stats["computed_pairs"],
                                    "required_tflops": tf,
                                    "computed_over_required": stats["computed_pairs"] / max(stats["required_pairs"], 1.0)}
    return out


def bind_to_gpu_numa_node(torch, local: int) -> None:
    """Multi-GPU runs: keep this rank's threads (and so its pinned host buffers) on the NUMA node its GPU hangs
    off, so that N concurrent host->device copies do not cross the socket interconnect.  Best effort: any
    missing piece of information leaves the affinity untouched."""
    try:
        p = torch.cuda.get_device_properties(local)
        bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read().strip())
        if node < 0:
            return
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            print(f"[rank {os.environ.get('RANK', '0')}] GPU {bdf} on NUMA node {node}: bound to {len(cpus)} CPUs",
                  file=sys.stderr, flush=True)
    except Exception as exc:  # noqa: BLE001 -- placement is an optimisation, never a failure
        print(f"[rank {os.environ.get('RANK', '0')}] NUMA binding skipped: {exc}", file=sys.stderr, flush=True)


# --------------------------------------------------------------------------- main arm
class Workload:
    """One rank's batch of spectra, resident in HBM and in pinned host memory, and the step functions
    over it (resident / host API / host API software-pipelined)."""

    def __init__(self, torch, fdist, hp, sp, dev, world):
        self.torch, self.fdist, self.hp, self.sp, self.dev, self.world = torch, fdist, hp, sp, dev, world
        self.n = len(sp)
        self.host = {k: torch.from_numpy(np.ascontiguousarray(v)).pin_memory() for k, v in dict(
            mz=sp.mz, intensity=sp.intensity, indptr=sp.indptr, precursor_mz=sp.precursor_mz,
            charge=sp.precursor_charge).items()}
        self.h2d_bytes = sum(t.numel() * t.element_size() for t in self.host.values())
        self.d = {k: v.to(dev) for k, v in self.host.items()}
        self.labels_host = torch.empty(self.n, dtype=torch.int32).pin_memory()
        self.max_peaks = int(np.diff(sp.indptr).max())  # falcon's max_peaks_used setting (known up front)
        self.gather_stream = torch.cuda.Stream(device=dev) if world > 1 else None
        self.peer_gather = None
        if world > 1 and os.environ.get("FLC_GATHER", "peer") == "peer":
            try:  # labels stored into the peers' buffers over NVLink by the step's last kernel
                self.peer_gather = fdist.PeerLabelGather(self.n, dev)
                hp.label_sink = self.peer_gather
            except Exception as exc:  # noqa: BLE001 -- no symmetric memory on this platform: NCCL all-gather
                log(f"[bench] peer-memory label gather unavailable ({exc!r}); using NCCL all_gather")
        self.gather_mode = "nvlink peer memory (flc_scatter_labels_peers + barrier + flc_relabel_gathered pull)" if self.peer_gather else \
            ("nccl all_gather" if world > 1 else "none")

    def gather(self, labels, n_clusters):
        if self.peer_gather is not None:
            return self.peer_gather.last[0]
        # Every rank's batch has the same number of spectra: one collective, label offsets computed on the
        # device, no host sync.  The collective runs on a side stream behind this batch's labels, so a rank that
        # finishes its batch early starts the next one instead of idling in the all-gather until its peers
        # arrive; the timed region ends only after every gather has completed (wait_gathers).
        if self.world > 1:
            torch = self.torch
            done = torch.cuda.Event()
            done.record()
            with torch.cuda.stream(self.gather_stream):
                self.gather_stream.wait_event(done)
                out = self.fdist.gather_labels_padded(labels, n_clusters, max_len=self.n)[0]
            labels.record_stream(self.gather_stream)
            return out
        return labels

    def wait_gathers(self):
        if self.peer_gather is not None:
            self.peer_gather.wait()
        elif self.gather_stream is not None:
            self.torch.cuda.current_stream().wait_stream(self.gather_stream)

    def step_resident(self):
        d = self.d
        labels, nc = self.hp.run(d["mz"], d["intensity"], d["indptr"], d["precursor_mz"], d["charge"],
                                 max_peaks=self.max_peaks)
        return self.gather(labels, nc), nc

    def step_e2e(self):
        # host (pinned) buffers in, labels back on the host: chunked H2D overlapped with vectorisation
        h = self.host
        labels, nc = self.hp.run_host(h["mz"], h["intensity"], h["indptr"], h["precursor_mz"], h["charge"],
                                      labels_out=self.labels_host, max_peaks=self.max_peaks)
        out = self.gather(labels, nc)
        self.torch.cuda.current_stream().synchronize()
        return out, nc

    def run_e2e_pipelined(self, steps):
        """K batches through the host API, software-pipelined two deep: batch i + 1 is staged
        (its H2D copies enqueued on the copy stream) before batch i's kernels run.  Every
        batch's H2D, kernels and D2H are inside the caller's timed region."""
        h, hp = self.host, self.hp
        stage = lambda: hp.stage_host(h["mz"], h["intensity"], h["indptr"], h["precursor_mz"], h["charge"],  # noqa: E731
                                      max_peaks=self.max_peaks)
        out = None
        staged = stage()
        for i in range(steps):
            nxt = stage() if i + 1 < steps else None
            labels, nc = hp.run_staged(staged, labels_out=self.labels_host)
            out = (self.gather(labels, nc), nc)
            staged = nxt
        self.torch.cuda.current_stream().synchronize()
        return out

    def plain_copy(self):
        """The same host buffers copied to the device and the labels copied back, nothing else: the PCIe floor
        of the host API at this number of concurrently copying ranks."""
        for k, v in self.host.items():
            self.d[k].copy_(v, non_blocking=True)
        self.labels_host.copy_(self.d["charge"][: self.n].view(self.torch.int32), non_blocking=True)
        return None, 0


def scan_pair_stats(torch, keep, n, dev):
    """(computed_pairs, required_pairs, sizes): what the scan multiplies vs what the IVF semantics require."""
    sizes = (keep["buckets"].bucket_ptr[1:] - keep["buckets"].bucket_ptr[:-1]).double()
    # pairs the scan actually multiplies: query tile i (128 rows) of a bucket meets the candidate rows from its
    # own first row to the bucket's end (S = X X^T is symmetric: scan_tc.cu)
    nb_np = sizes.cpu().numpy()
    computed_pairs = 0.0
    for i in range(int(np.ceil(nb_np.max() / 128.0)) if nb_np.size else 0):
        rows_i = np.clip(nb_np - 128.0 * i, 0.0, 128.0)
        computed_pairs += float((rows_i * np.maximum(nb_np - 128.0 * i, 0.0)).sum())
    required_pairs = float((sizes * sizes).sum().item())
    ivf = keep["ivf"]
    extra = {}
    if ivf is not None:
        bp = keep["buckets"].bucket_ptr
        b_of_row = torch.searchsorted(bp, torch.arange(n, device=dev), right=True) - 1
        max_l = int(ivf.nlist.max().item()) + 1
        lsize = torch.bincount(b_of_row * max_l + ivf.list_id.long(), minlength=int(bp.shape[0]) * max_l)
        pr = ivf.probes.long()
        valid = pr >= 0
        idx = (b_of_row[:, None] * max_l + pr.clamp(min=0))
        required_pairs = float((lsize[idx] * valid).sum().item())
        nl = ivf.nlist[:-1].long()
        row_in_ivf = (nl > 0)[b_of_row]
        ell_nnz = keep["vectors"].ell_nnz.long() & 0xFFFF
        extra = {"ell_width": keep["vectors"].ell_width, "max_nprobe": ivf.max_nprobe,
                 "ivf_rows": int(sizes[nl > 0].sum().item()), "ivf_nnz": int(ell_nnz[row_in_ivf].sum().item()),
                 "total_centroids": ivf.total_centroids}
    return computed_pairs, required_pairs, sizes, extra


def merge_kernel_classes(kernels: dict) -> dict:
    if "kmeans_fused_large" in kernels:  # the two size classes of the fused trainer: one logical launch
        big = kernels.pop("kmeans_fused_large")
        small = kernels.get("kmeans_fused", (0.0, big[1]))
        kernels["kmeans_fused"] = (small[0] + big[0], small[1])
    if "refine_block" in kernels and "refine" in kernels:  # fast path + the query groups it defers
        rest = kernels.pop("refine")
        kernels["refine_block"] = (kernels["refine_block"][0] + rest[0], kernels["refine_block"][1])
    return kernels


def run_ours(args):
    import torch

    from falcon_b200 import _lib, distributed as fdist, pipeline, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: falcon_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        bind_to_gpu_numa_node(torch, local)  # before any pinned allocation (first touch decides the node)
    dist = None
    if world > 1:
        import torch.distributed as dist  # noqa: F811

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    peaks = load_peaks()
    settings = pipeline.Settings(exhaustive=args.exhaustive)
    hp = pipeline.HotPath(settings, dev)

    t0 = time.perf_counter()
    kw = {"mass_range": tuple(args.mass_range)} if args.mass_range else {}
    if args.total:
        # strong scaling (BASELINE configs[3]): args.total spectra over the whole precursor-mass range; rank r
        # owns the r-th slice of the range, i.e. whole precursor buckets, with the full data set's bucket sizes
        lo, hi = args.mass_range if args.mass_range else (700.0, 3500.0)
        args.n = args.total // world
        kw = {"mass_range": (lo + (hi - lo) * rank / world, lo + (hi - lo) * (rank + 1) / world)}
    sp = synth.generate(args.n, 42 + rank, **kw)
    log(f"[rank {rank}] generated {len(sp)} spectra / {sp.n_peaks} peaks in {time.perf_counter() - t0:.1f}s")
    wl = Workload(torch, fdist, hp, sp, dev, world)
    d, h2d_bytes, gather_mode = wl.d, wl.h2d_bytes, wl.gather_mode

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(wl_, fn, steps, warmup, sample_clocks=False, profile=False):
        # nvidia-smi takes ~0.1-0.5 s to initialise NVML (longer on a multi-GPU box) and holds driver locks while
        # it does: start it BEFORE the warm-up, wait for its first sample, and only then run warm-up + timed steps,
        # so that it samples during the timed region without its start-up landing inside it.  Rank 0 only.
        sampler = ClockSampler(local) if (sample_clocks and rank == 0) else None
        if sampler:
            sampler.start()
            sampler.wait_first_sample(5.0)
            sampler.mark()  # samples before this point are idle clocks: not reported
        for _ in range(warmup):
            fn()
        barrier()
        if profile:
            _lib.profile_reset()
            _lib.profile_enable(True)
        _lib.reset_launch_count()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(steps):
            out = fn()
        wl_.wait_gathers()
        b.record()
        torch.cuda.synchronize()
        launches = _lib.launch_count()
        if profile:
            _lib.profile_enable(False)
        clocks = sampler.stop() if sampler else None
        barrier()
        ms = a.elapsed_time(b) / steps
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, out, launches, clocks

    ms, (labels, n_clusters), launches, clocks = timed(wl, wl.step_resident, args.steps, args.warmup,
                                                       sample_clocks=True, profile=True)
    kernels = merge_kernel_classes(_lib.profile_summary())
    ms_e2e, _, _, _ = timed(wl, wl.step_e2e, args.steps, max(1, args.warmup - 1))
    # the same K batches, software-pipelined two deep (one call = K steps)
    ms_pipe, _, _, _ = timed(wl, lambda: wl.run_e2e_pipelined(args.steps), 1, 1)
    ms_pipe /= args.steps
    # PCIe floor of the host API: the same pinned buffers copied in (and the labels out) by all ranks at once
    ms_copy, _, _, _ = timed(wl, wl.plain_copy, args.steps, 1)
    total = args.n * world

    # one instrumented pass for the roofline statistics (outside the timed regions)
    hp2 = pipeline.HotPath(settings, dev, profile=False)
    hp2.run(d["mz"], d["intensity"], d["indptr"], d["precursor_mz"], d["charge"], max_peaks=wl.max_peaks)  # warm scratch
    hp2.timer = pipeline._Timer(True)
    _, _, keep = hp2.run(d["mz"], d["intensity"], d["indptr"], d["precursor_mz"], d["charge"], keep=True)
    stage_ms = hp2.timer.result()
    computed_pairs, required_pairs, sizes, stats_extra = scan_pair_stats(torch, keep, args.n, dev)
    stats = {"n": args.n, "n_peaks": sp.n_peaks, "low_dim": settings.low_dim, "ld_bf16": hp.ld_bf16,
             "dense_f32": 1 if settings.dense_f32 else 0, "ell_width": keep["vectors"].ell_width,
             "n_pairs": keep["graph"].n_pairs, "nnz": keep["graph"].nnz, "kernels": kernels,
             "steps": args.steps, "required_pairs": required_pairs, "computed_pairs": computed_pairs,
             **stats_extra}
    roof_all = kernel_rooflines(stats, peaks)
    hand = {k: v for k, v in kernels.items() if k in roof_all}
    dominant = max(hand, key=lambda k: hand[k][0]) if hand else None
    roofline = None
    if dominant:
        r = roof_all[dominant]
        roofline = {"kernel": dominant, "bound": r["bound"], "achieved": r["achieved"], "peak": r["peak"],
                    "unit": r["unit"], "frac": r["frac"], "traffic": None,
                    "peak_source": f"{peaks['_source']} (MEASURED_PEAKS.json hbm_gbs; kernel timed alone -> burst figure)",
                    "ms_per_launch": r["ms_per_launch"],
                    "share_of_step": hand[dominant][0] / args.steps / ms}
        if "tensor" in r:
            roofline["tensor"] = r["tensor"]
            if r["tensor"]["frac"] > r["frac"]:  # the binding roofline is the one the kernel sits closer to
                roofline.update(bound="tensor", achieved=r["tensor"]["achieved"], peak=r["tensor"]["peak"],
                                unit="TFLOP/s", frac=r["tensor"]["frac"], hbm_view={
                                    "achieved": r["achieved"], "peak": r["peak"], "unit": "GB/s", "frac": r["frac"]},
                                peak_source=f"{peaks['_source']} (MEASURED_PEAKS.json bf16_tflops, burst: kernel timed alone)")
        if dominant == "kmeans_fused":
            roofline["note"] = ("fused k-means keeps a bucket's sparse rows in shared memory for all iterations: "
                                "HBM sees every row once (the algorithmic bytes used here), the kernel itself is "
                                "bound by shared-memory gathers and barriers -- see profiles/ for the smem-pipe figures")
    ncu_traffic = os.path.join(ROOT, "profiles", NCU_TRAFFIC_FILE)
    if os.path.exists(ncu_traffic) and not args.mass_range and not args.total and args.n == 1_000_000:
        # DRAM bytes per launch from the committed ncu --set full capture of this workload's steady-state step
        traffic = json.load(open(ncu_traffic))
        for name, r in roof_all.items():
            if name in traffic and traffic[name].get("workload") == "default":
                r["traffic"] = traffic[name]["dram_bytes_per_launch"]
        t = traffic.get(roofline["kernel"]) if roofline else None
        if t and t.get("workload") == "default":
            roofline["traffic"] = t["dram_bytes_per_launch"]
            roofline["traffic_over_algorithmic"] = t["dram_bytes_per_launch"] / roof_all[roofline["kernel"]]["algorithmic_bytes_per_launch"]
            roofline["traffic_source"] = t.get("source")

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = len(os.sched_getaffinity(0))
        sample, nb = bucket_sample(sp, args.cpu_sample)
        cpu_reference_time(bucket_sample(sp, 4000)[0], cores, args.exhaustive)  # warm the worker pool / imports
        dt, stages, _ = cpu_reference_time(sample, cores, args.exhaustive)
        cpu_baseline = {"value": len(sample) / dt, "unit": UNIT, "cores": cores, "kind": "port",
                        "sample": f"{len(sample)} spectra = {nb} whole precursor buckets of the workload, "
                                  f"{dt:.1f}s of CPU work",
                        "stages_s": stages}

    out = None
    if rank == 0:
        bsz = sizes.cpu().numpy()
        out = {
            "metric": METRIC, "value": total / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "strong" if args.total else "weak",
            "vs_baseline": None, "dtype": "bf16 scan / f32 vectors / f64 re-score", "data": "synthetic",
            "config": workload_config(args),
            "e2e": {"value": total / (ms_pipe * 1e-3), "unit": UNIT, "ms_per_step": ms_pipe,
                    "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": args.n * 4,
                    "h2d_gbs_per_gpu": h2d_bytes / (ms_pipe * 1e-3) / 1e9,
                    "mode": "host API, batches software-pipelined two deep (stage_host of batch i+1 before "
                            "run_staged of batch i); all K batches' H2D + kernels + D2H inside the timed region",
                    "unpipelined": {"value": total / (ms_e2e * 1e-3), "ms_per_step": ms_e2e},
                    "copy_floor": {"ms_per_step": ms_copy, "gbs_per_gpu": (h2d_bytes + args.n * 4) / (ms_copy * 1e-3) / 1e9,
                                   "value": total / (ms_copy * 1e-3),
                                   "what": "the same pinned buffers copied host->device (labels device->host) and "
                                           "nothing else, by all ranks at once, max over ranks: the platform's "
                                           "ceiling for the host API at this GPU count"}},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu_baseline,
            "n_clusters": int(n_clusters), "label_gather": gather_mode,
            "stage_ms": stage_ms,
            "kernels_ms_per_step": {k: v[0] / args.steps for k, v in kernels.items()},
            "kernel_rooflines": {k: {"frac": v["frac"], "achieved": v["achieved"], "unit": v["unit"],
                                     "ms_per_launch": v["ms_per_launch"],
                                     "algorithmic_bytes_per_launch": v["algorithmic_bytes_per_launch"],
                                     **({"traffic": v["traffic"]} if "traffic" in v else {}),
                                     **({"survey_8d": v["survey_8d"]} if "survey_8d" in v else {}),
                                     **({"tensor": v["tensor"]} if "tensor" in v else {})}
                                 for k, v in roof_all.items()},
            "bucket_sizes": {"n_buckets": int(bsz.shape[0]), "mean": float(bsz.mean()), "p50": float(np.median(bsz)),
                             "p99": float(np.percentile(bsz, 99)), "max": float(bsz.max())},
            "n_pairs": keep["graph"].n_pairs, "nnz": keep["graph"].nnz,
        }
    # ---- the same step with cluster representatives (medoids by the published rule: uncut matrix), gathered
    # over NCCL with the labels when there are several ranks -- north_star's "final gather of labels and cluster
    # representatives"; kept beside the headline because building the uncut matrix multiplies the step time
    reps_rec = None
    if not args.no_representatives:
        hp_r = pipeline.HotPath(pipeline.Settings(exhaustive=args.exhaustive, representatives=True), dev)

        def step_reps():
            lab, nc = hp_r.run(d["mz"], d["intensity"], d["indptr"], d["precursor_mz"], d["charge"],
                               max_peaks=wl.max_peaks)
            reps = hp_r.representatives
            if world > 1:
                lab = fdist.gather_labels_padded(lab, nc, max_len=args.n)[0]
                reps = fdist.gather_representatives(reps, args.n)
            return (lab, reps), nc

        ms_reps, ((_, reps), _), _, _ = timed(wl, step_reps, 2, 1)
        if rank == 0:
            reps_rec = {"ms_per_step": ms_reps, "value": total / (ms_reps * 1e-3), "unit": UNIT,
                        "n_representatives": int(reps.shape[0]),
                        "what": "whole step + medoid of every cluster from the uncut n_neighbors matrix "
                                "(HotPath.representatives_exact)" + (", labels and representatives gathered over NCCL"
                                                                     if world > 1 else "")}
        del hp_r, reps
    # ---- free the headline workload, then the witnesses of the other configurations
    del keep, hp2, d, wl, labels
    torch.cuda.empty_cache()
    mg = multi_gpu_check(args, torch, fdist, pipeline, synth, dev, world, rank) if world > 1 else None
    ns = north_star_record(args, torch, fdist, pipeline, synth, _lib, dev, world, rank, dist, timed, peaks) \
        if args.north_star_total else None
    if rank == 0:
        out["representatives"] = reps_rec
        out["multi_gpu_check"] = mg
        out["north_star_10m"] = ns
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


def multi_gpu_check(args, torch, fdist, pipeline, synth, dev, world, rank):
    """Outside every timed region: ONE data set clustered by all ranks through the product's sharded path
    (`distributed.cluster_sharded`: buckets dealt to the ranks, oversized buckets cut with a tolerance halo,
    labels + representatives gathered) must give the partition of the single-GPU run (rank 0, same data)."""
    n = args.check_n
    a = synth.generate(int(n * 0.6), 50, mass_range=(1000.0, 1012.0))  # buckets of several thousand rows: forced cuts
    sp = synth.concat([a, synth.generate(n - len(a), 51)])
    res = {"n": n, "world": world, "bucket_cap": 2000}
    for exhaustive in (True, False):
        s = pipeline.Settings(exhaustive=exhaustive, representatives=True)
        # with the IVF index a cut bucket would train one index per piece: only whole buckets are dealt out
        # (cluster_sharded's default there)
        cap = 2000 if exhaustive else None
        labels, nc, reps = fdist.cluster_sharded(sp, s, device=dev, bucket_cap=cap)
        key = "exhaustive" if exhaustive else "default_nprobe"
        if rank == 0:
            ref, nc_ref, _ = pipeline.cluster_host(sp, s, device=dev)
            units = fdist.plan_units(*_bucket_plan(torch, pipeline, sp, s, dev), world, s.precursor_tol_mass,
                                     s.precursor_tol_mode, exhaustive, cap)
            res[key] = {"equal_to_single_gpu": bool(nc == nc_ref and fdist.same_partition(labels, ref)),
                        "n_clusters": int(nc), "n_clusters_single_gpu": int(nc_ref),
                        "representatives_ok": bool(reps.shape[0] == nc and (labels[reps] == np.arange(nc)).all()),
                        "units": int(units["owner"].shape[0]), "halo_pieces": int(units["piece"].sum())}
        torch.distributed.barrier()
    if rank == 0:
        res["equal_to_single_gpu"] = bool(res["exhaustive"]["equal_to_single_gpu"]
                                          and res["default_nprobe"]["equal_to_single_gpu"])
    return res


def _bucket_plan(torch, pipeline, sp, s, dev):
    hp = pipeline.HotPath(s, dev)
    b = hp.bucket_sort(torch.from_numpy(sp.precursor_mz).to(dev), torch.from_numpy(sp.precursor_charge).to(dev))
    return b.bucket_ptr.cpu().numpy(), b.mz.cpu().numpy()


def north_star_record(args, torch, fdist, pipeline, synth, _lib, dev, world, rank, dist, timed, peaks):
    """BASELINE configs[3] beside the headline: ONE data set of `--north-star-total` spectra (80 bucket-aligned
    chunks over the whole precursor-mass range, the same data set at every GPU count), chunks dealt to the
    ranks in order -- whole precursor buckets, so no halo is needed -- each rank clusters its share, labels
    gathered over NCCL inside the step.  Strong scaling: the total is fixed."""
    n_chunks = 80
    total = args.north_star_total // n_chunks * n_chunks
    mine = range(n_chunks * rank // world, n_chunks * (rank + 1) // world)
    workers = max(1, min(16, len(os.sched_getaffinity(0)) // world))
    t0 = time.perf_counter()
    sp = synth.generate_chunks(total, n_chunks, mine, workers=workers)
    log(f"[rank {rank}] north star: generated {len(sp)} of {total} spectra in {time.perf_counter() - t0:.1f}s "
        f"({workers} workers)")
    settings = pipeline.Settings(exhaustive=args.exhaustive)
    hp = pipeline.HotPath(settings, dev)
    wl = Workload(torch, fdist, hp, sp, dev, world)
    steps, warmup = max(2, min(args.steps, 3)), 3
    ms, (_, n_clusters), launches, _ = timed(wl, wl.step_resident, steps, warmup, profile=True)
    kernels = merge_kernel_classes(_lib.profile_summary())
    ms_pipe, _, _, _ = timed(wl, lambda: wl.run_e2e_pipelined(steps), 1, 1)
    ms_pipe /= steps
    d = wl.d
    hp2 = pipeline.HotPath(settings, dev)
    labels, nc, keep = hp2.run(d["mz"], d["intensity"], d["indptr"], d["precursor_mz"], d["charge"], keep=True)
    computed, required, sizes, _ = scan_pair_stats(torch, keep, len(sp), dev)
    # order-independent fingerprint of the partition: cluster-size histogram (sizes 2..64, 65+) and noise count
    lab = labels.long()
    csize = torch.bincount(lab[lab >= 0], minlength=max(int(nc), 1))
    hist = torch.bincount(csize.clamp(max=65), minlength=66).double()
    agg = torch.tensor([float(nc), float((lab < 0).sum().item()), computed, required, float(len(sp)),
                        float(sizes.max().item()), float(sizes.shape[0])], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(agg, op=dist.ReduceOp.SUM)
        dist.all_reduce(hist, op=dist.ReduceOp.SUM)
    out = None
    if rank == 0:
        import hashlib

        scan_ms = kernels.get("scan_tc", (0.0, 0))[0] / steps
        ex_tf = 2.0 * settings.low_dim * computed / (scan_ms * 1e-3) / 1e12 if scan_ms > 0 else None
        out = {
            "config": f"configs[3]: {total} synthetic spectra in 80 bucket-aligned chunks, chunks (whole precursor "
                      f"buckets) dealt to {world} GPU(s), falcon defaults" + (" exhaustive" if args.exhaustive else ""),
            "scaling": "strong", "total_spectra": int(agg[4].item()), "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": ms, "value": agg[4].item() / (ms * 1e-3), "unit": UNIT,
            "e2e": {"value": agg[4].item() / (ms_pipe * 1e-3), "ms_per_step": ms_pipe, "h2d_bytes_per_step": wl.h2d_bytes,
                    "d2h_bytes_per_step": len(sp) * 4},
            "gpu_launches": int(launches),
            "n_clusters": int(agg[0].item()), "noise": int(agg[1].item()),
            "partition_fingerprint": hashlib.sha1(hist.cpu().numpy().astype(np.int64).tobytes()).hexdigest()[:16],
            "buckets": {"n": int(agg[6].item()), "max_rows_rank0": float(sizes.max().item()),
                        "mean_rows": agg[4].item() / max(agg[6].item(), 1.0)},
            "scan": {"ms_rank0": scan_ms, "required_pairs": agg[3].item(), "computed_pairs": agg[2].item(),
                     "computed_over_required": agg[2].item() / max(agg[3].item(), 1.0),
                     "executed_tflops_rank0": ex_tf,
                     "executed_frac": (ex_tf / peaks["bf16_tflops"]) if ex_tf else None,
                     "note": "rank 0's scan_tc time and the pairs of rank 0's share; the fingerprint and the cluster "
                             "count must not depend on the GPU count (same data set, whole buckets per rank)"},
            "kernels_ms_per_step_rank0": {k: v[0] / steps for k, v in kernels.items()},
        }
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=1_000_000, help="spectra per GPU")
    ap.add_argument("--exhaustive", action="store_true", help="n_probe = nlist (BASELINE configs[2])")
    ap.add_argument("--cpu-sample", type=int, default=1_000_000,
                    help="spectra in the CPU-baseline sample (whole buckets; ~10 s of CPU work at 1M)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-representatives", action="store_true", help="skip the step-with-representatives record")
    ap.add_argument("--total", type=int, default=0,
                    help="strong scaling: this many spectra in total, the precursor-mass range (whole buckets) "
                         "split across the ranks -- BASELINE configs[3] is --total 10000000")
    ap.add_argument("--north-star-total", type=int, default=10_000_000,
                    help="spectra of the configs[3] witness record (`north_star_10m`) measured after the headline "
                         "workload; 0 skips it")
    ap.add_argument("--check-n", type=int, default=200_000,
                    help="spectra of the multi-GPU parity self-check (N > 1 only, outside the timed regions)")
    ap.add_argument("--mass-range", type=float, nargs=2, default=None, metavar=("LO", "HI"),
                    help="neutral mass range of the synthetic peptides (default 700-3500 Da); a narrow range "
                         "makes large precursor buckets (tensor-core regime of the scan)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
