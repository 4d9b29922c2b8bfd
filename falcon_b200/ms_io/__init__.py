"""Peak-file IO: the minimal MGF reader / writer the CLI needs."""
from .mgf_io import get_spectra, read_mgf, write_spectra  # noqa: F401
