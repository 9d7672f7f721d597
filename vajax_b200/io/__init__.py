"""Output side of the hot path: result container and the raw / CSV wire formats.

Mirrors the reference's ``vajax.io`` names (vajax/io/__init__.py:3-15) for the writers that sit directly
behind ``run_transient``; the ngspice / Xyce *readers* of that package serve its regression harness
and are out of scope."""
from .csv_format import read_csv, write_csv
from .rawfile import read_rawfile, read_rawfile_header, write_rawfile
from .result import TransientResult, extract_results

__all__ = ["TransientResult", "extract_results", "write_rawfile", "read_rawfile_header", "read_rawfile",
           "write_csv", "read_csv"]
