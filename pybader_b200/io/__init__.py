"""CHGCAR / cube readers with the reference's `read()` signatures and return
values (pybader/io/vasp.py:15-164, pybader/io/cube.py:18-156); the numeric
blocks are converted on the GPU (`bdr_parse_text`)."""
from . import cube, vasp  # noqa: F401
