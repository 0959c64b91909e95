"""text block -> fp64 grid through the C ABI (bdr_parse_text)."""
import ctypes

import numpy as np

from .. import _lib
from .._lib import check

OP_NONE, OP_DIVIDE, OP_MULTIPLY = 0, 1, 2


def pinned_empty(shape, dtype):
    """numpy array on page-locked memory (bdr_host_alloc), freed with the array;
    falls back to ordinary memory if the driver refuses.  Worth it only for buffers
    that are reused: page-locking costs ~0.3 ms per MB, a pageable copy 0.1-0.2."""
    import weakref
    lib = _lib.load()
    dtype = np.dtype(dtype)
    n = int(np.prod(shape)) * dtype.itemsize
    ptr = lib.bdr_host_alloc(max(n, 1))
    if not ptr:
        return np.empty(shape, dtype=dtype)
    raw = (ctypes.c_char * max(n, 1)).from_address(ptr)
    arr = np.frombuffer(raw, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
    weakref.finalize(raw, lib.bdr_host_free, ptr)
    return arr


def parse_block(text, shape, x_fastest, op=OP_NONE, operand=1.0, device=0):
    """Convert the first prod(shape) whitespace-separated tokens of `text` (bytes /
    bytearray / uint8 array) into a C-ordered float64 array of `shape`.

    x_fastest: the tokens run with the first axis fastest (CHGCAR); otherwise in
    C order (cube).  op/operand: the reader's arithmetic on every value.
    Returns (array, bytes_consumed).  Raises ValueError like the reference when a
    token is not a number or the text holds too few tokens."""
    lib = _lib.load()
    buf = np.frombuffer(text, dtype=np.uint8) if not isinstance(text, np.ndarray) else text
    n = int(np.prod(shape))
    # (page-locking a fresh result array costs more than the pageable copy it would save)
    out = np.empty(shape, dtype=np.float64)
    found, used, nfb = ctypes.c_int64(0), ctypes.c_int64(0), ctypes.c_int64(0)
    cap = 1 << 16
    while True:
        fb = np.zeros((cap, 3), dtype=np.int64)
        check(lib.bdr_parse_text(int(device), buf.ctypes.data, buf.size, n, *map(int, shape),
                                 int(bool(x_fastest)), int(op), float(operand), out.ctypes.data,
                                 ctypes.byref(found), ctypes.byref(used), ctypes.byref(nfb),
                                 fb.ctypes.data, cap))
        if nfb.value <= cap:
            break
        cap = int(nfb.value)          # rare: a file full of tokens the device hands back
    if found.value < n:
        raise ValueError(f"could not broadcast input array from shape ({found.value},) into shape ({n},)")
    if nfb.value:
        # tokens the device does not convert exactly: Python's float() decides (and raises
        # the reference's ValueError on '****' and the like)
        nx, ny, nz = (int(s) for s in shape)
        flat = out.reshape(-1)
        for t, off, ln in fb[:nfb.value]:
            # the device reports at most 64 bytes of a token: the whole whitespace-delimited
            # token decides, as in the reference (a long valid number converts in full, junk
            # behind the 64th byte still raises)
            end = int(off + ln)
            while end < buf.size and buf[end] not in (32, 10, 9, 13, 12, 11):
                end += 1
            v = float(bytes(buf[off:end]))
            if op == OP_DIVIDE:
                v = v / operand
            elif op == OP_MULTIPLY:
                v = v * operand
            if x_fastest:
                x, y, z = t % nx, (t // nx) % ny, t // (nx * ny)
                flat[(x * ny + y) * nz + z] = v
            else:
                flat[t] = v
    return out, int(used.value)
