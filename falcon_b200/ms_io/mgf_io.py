"""MGF reading and writing -- mirror of ``falcon.ms_io.mgf_io``
(/root/reference/falcon/ms_io/mgf_io.py:10-116).  The reference parses with
pyteomics (not installed here); this is a small parser of the same fields:
TITLE -> identifier, PEPMASS (first number) -> precursor m/z, CHARGE (``2+``)
-> precursor charge (absent: None), RTINSECONDS -> retention time (absent: -1),
peak lines ``m/z intensity``.  Spectra that cannot be parsed are skipped, like
the reference's ``except (ValueError, KeyError): pass`` (mgf_io.py:26-30).
Peaks are sorted by m/z (what ``MsmsSpectrum`` does on construction).
"""
from __future__ import annotations

import math
import re
from typing import IO, Dict, Iterable, Iterator, List, Tuple, Union

import numpy as np

from .. import synth

_CHARGE = re.compile(r"^\s*(\d+)\s*([+-]?)")


def _parse_charge(text: str) -> int:
    m = _CHARGE.match(text)
    if not m:
        raise ValueError(f"bad CHARGE {text!r}")
    z = int(m.group(1))
    return -z if m.group(2) == "-" else z


def _blocks(lines: Iterable[str]) -> Iterator[Tuple[Dict[str, str], List[str]]]:
    params, peaks, inside = {}, [], False
    for line in lines:
        line = line.strip()
        if not line or line[0] in "#;!/":
            continue
        if line == "BEGIN IONS":
            params, peaks, inside = {}, [], True
        elif line == "END IONS":
            if inside:
                yield params, peaks
            inside = False
        elif inside:
            if "=" in line and not (line[0].isdigit() or line[0] in "+-."):
                k, v = line.split("=", 1)
                params[k.strip().lower()] = v.strip()
            else:
                peaks.append(line)


def _parse(params: Dict[str, str], peaks: List[str], filename: str) -> dict:
    identifier = params["title"]
    precursor_mz = float(params["pepmass"].split()[0])
    charge = _parse_charge(params["charge"]) if "charge" in params else None
    rt = float(params.get("rtinseconds", -1))
    if peaks:
        arr = np.array([p.split()[:2] for p in peaks], dtype=np.float64)
        order = np.argsort(arr[:, 0], kind="stable")
        mz, inten = arr[order, 0].astype(np.float32), arr[order, 1].astype(np.float32)
    else:
        mz = inten = np.zeros(0, np.float32)
    return {"identifier": identifier, "precursor_mz": precursor_mz, "precursor_charge": charge, "mz": mz,
            "intensity": inten, "retention_time": rt, "filename": filename}


def get_spectra(source: Union[IO, str]) -> Iterator[dict]:
    """Iterate over the MS/MS spectra of an MGF file (name or open text file) as spectrum dicts
    (the keys ``process_spectrum`` returns, /root/reference/falcon/cluster/spectrum.py:161-169)."""
    if isinstance(source, str):
        with open(source) as fh:
            yield from get_spectra_named(fh, source)
    else:
        yield from get_spectra_named(source, getattr(source, "name", "<stream>"))


def get_spectra_named(fh: IO, filename: str) -> Iterator[dict]:
    for params, peaks in _blocks(fh):
        try:
            yield _parse(params, peaks, filename)
        except (ValueError, KeyError, IndexError):
            pass


def read_mgf(source: Union[IO, str]) -> Tuple[synth.SpectrumSet, List[str], List[str]]:
    """All spectra of an MGF file as one ``SpectrumSet`` (CSR peak arrays) + identifiers + file names.
    An absent charge is stored as 0."""
    dicts = list(get_spectra(source))
    for d in dicts:
        if d["precursor_charge"] is None:
            d["precursor_charge"] = 0
    return synth.SpectrumSet.from_dicts(dicts), [d["identifier"] for d in dicts], [d["filename"] for d in dicts]


def write_spectra(filename: str, spectra: Iterable[dict]) -> None:
    """Write spectra (dicts, or objects with the same attributes) to an MGF file
    (/root/reference/falcon/ms_io/mgf_io.py:70-116: TITLE, PEPMASS, CHARGE, RTINSECONDS, SCAN, CLUSTER)."""
    def get(s, k, default=None):
        return s.get(k, default) if isinstance(s, dict) else getattr(s, k, default)

    with open(filename, "w") as out:
        for s in spectra:
            out.write("BEGIN IONS\n")
            out.write(f"TITLE={get(s, 'identifier')}\n")
            out.write(f"PEPMASS={get(s, 'precursor_mz')}\n")
            z = get(s, "precursor_charge")
            if z is not None and not (isinstance(z, float) and math.isnan(z)) and int(z) != 0:
                out.write(f"CHARGE={abs(int(z))}{'-' if int(z) < 0 else '+'}\n")
            rt = get(s, "retention_time")
            if rt is not None:
                out.write(f"RTINSECONDS={rt}\n")
            for key in ("scan", "cluster"):
                v = get(s, key)
                if v is not None:
                    out.write(f"{key.upper()}={v}\n")
            mz, inten = np.asarray(get(s, "mz")), np.asarray(get(s, "intensity"))
            out.write("".join(f"{float(a)!r} {float(b)!r}\n" for a, b in zip(mz, inten)))
            out.write("END IONS\n\n")
