"""fp64 grid -> text through the C ABI (bdr_format_grid), plus the reference's
'Fortran' number format restated with numpy."""
import numpy as np

from .. import _lib
from .._lib import check


def append_block(path, data, x_fastest, row_len, per_line, prec, sign_space):
    """append the formatted block of `data` (C-ordered float64 [nx][ny][nz]) to `path`"""
    lib = _lib.load()
    a = np.ascontiguousarray(data, dtype=np.float64)
    nx, ny, nz = (int(s) for s in a.shape)
    check(lib.bdr_format_grid(path.encode(), a.ctypes.data, nx, ny, nz, int(bool(x_fastest)),
                              int(row_len), int(per_line), int(prec), int(bool(sign_space))))


def fortran_lines(a, prec):
    """utils.fortran_format (utils.py:40-82): ' 0.ddddE+xx' / ' -.ddddE+xx' with the digits
    from int(0.5 + |a| / 10**(exp - prec)), exp = floor(log10|a|) + 1; rows of `a` are lines"""
    a = np.asarray(a, dtype=np.float64)
    flat = a.reshape(-1)
    absa = np.abs(flat)
    nz = flat != 0
    exp = np.zeros(flat.shape, dtype=np.int64)
    exp[nz] = np.floor(np.log10(absa[nz])) + 1
    value = np.zeros(flat.shape, dtype=np.int64)
    value[nz] = 0.5 + absa[nz] / np.power(10.0, exp[nz] - prec)
    out = []
    width = a.shape[1]
    for i in range(flat.shape[0]):
        sign = ' -.' if flat[i] < 0 else ' 0.'
        digits = ('0' * prec) if not nz[i] else str(int(value[i]))[:prec]
        ae = abs(int(exp[i]))
        e = ('E-' if exp[i] < 0 else 'E+') + ('0' if ae < 10 else '') + (str(ae)[:2] if nz[i] else '0')
        out.append(sign + digits + e)
        if (i + 1) % width == 0:
            out.append('\n')
    return ''.join(out)
