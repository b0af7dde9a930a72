"""ctypes mirror of the liblzma-named entry points of libxzb200.so (include/xzb200_lzma.h), driven
the way src/xz/coder.c:836, 956, 1226 drives liblzma: lzma_stream_encoder_mt / lzma_stream_decoder,
lzma_code in a loop, lzma_end.  Struct layouts follow src/liblzma/api/lzma/base.h:498-569 and
container.h:64-256 (checked field by field in tests/test_api_cpu.py)."""
import ctypes as C

from . import lib

LZMA_RUN, LZMA_FULL_FLUSH, LZMA_FINISH = 0, 2, 3
LZMA_OK, LZMA_STREAM_END = 0, 1


class LzmaStream(C.Structure):
    _fields_ = [("next_in", C.c_void_p), ("avail_in", C.c_size_t), ("total_in", C.c_uint64), ("next_out", C.c_void_p),
                ("avail_out", C.c_size_t), ("total_out", C.c_uint64), ("allocator", C.c_void_p), ("internal", C.c_void_p),
                ("reserved_ptr1", C.c_void_p), ("reserved_ptr2", C.c_void_p), ("reserved_ptr3", C.c_void_p), ("reserved_ptr4", C.c_void_p),
                ("seek_pos", C.c_uint64), ("reserved_int2", C.c_uint64), ("reserved_int3", C.c_size_t), ("reserved_int4", C.c_size_t),
                ("reserved_enum1", C.c_int), ("reserved_enum2", C.c_int)]


class LzmaMt(C.Structure):
    _fields_ = [("flags", C.c_uint32), ("threads", C.c_uint32), ("block_size", C.c_uint64), ("timeout", C.c_uint32), ("preset", C.c_uint32),
                ("filters", C.c_void_p), ("check", C.c_int), ("reserved_enum1", C.c_int), ("reserved_enum2", C.c_int), ("reserved_enum3", C.c_int),
                ("reserved_int1", C.c_uint32), ("reserved_int2", C.c_uint32), ("reserved_int3", C.c_uint32), ("reserved_int4", C.c_uint32),
                ("memlimit_threading", C.c_uint64), ("memlimit_stop", C.c_uint64), ("reserved_int7", C.c_uint64), ("reserved_int8", C.c_uint64),
                ("reserved_ptr1", C.c_void_p), ("reserved_ptr2", C.c_void_p), ("reserved_ptr3", C.c_void_p), ("reserved_ptr4", C.c_void_p)]


def encode_mt(src_ptr, n, preset, block_size, dst_ptr, cap, threads=4, check=4, timeout=0):
    """lzma_stream_encoder_mt + lzma_code(LZMA_FINISH) over host buffers; returns the Stream's size."""
    L = lib()
    s = LzmaStream()
    m = LzmaMt()
    m.threads, m.preset, m.check, m.block_size, m.timeout = threads, preset, check, block_size, timeout
    r = L.lzma_stream_encoder_mt(C.byref(s), C.byref(m))
    if r != LZMA_OK:
        raise RuntimeError(f"lzma_stream_encoder_mt: {r}")
    s.next_in, s.avail_in = src_ptr, n
    s.next_out, s.avail_out = dst_ptr, cap
    try:
        while True:
            r = L.lzma_code(C.byref(s), LZMA_FINISH)
            if r == LZMA_STREAM_END:
                return int(s.total_out)
            if r != LZMA_OK:
                raise RuntimeError(f"lzma_code: {r}")
    finally:
        L.lzma_end(C.byref(s))


def decode(src_ptr, n, dst_ptr, cap, mt=False, flags=0):
    """lzma_stream_decoder[_mt] + lzma_code(LZMA_FINISH) over host buffers; returns (ret, size)."""
    L = lib()
    s = LzmaStream()
    if mt:
        m = LzmaMt()
        m.threads, m.flags, m.memlimit_threading, m.memlimit_stop = 4, flags, (1 << 64) - 1, (1 << 64) - 1
        r = L.lzma_stream_decoder_mt(C.byref(s), C.byref(m))
    else:
        r = L.lzma_stream_decoder(C.byref(s), C.c_uint64((1 << 64) - 1), C.c_uint32(flags))
    if r != LZMA_OK:
        raise RuntimeError(f"lzma_stream_decoder: {r}")
    s.next_in, s.avail_in = src_ptr, n
    s.next_out, s.avail_out = dst_ptr, cap
    try:
        while True:
            r = L.lzma_code(C.byref(s), LZMA_FINISH)
            if r != LZMA_OK:
                return r, int(s.total_out)
    finally:
        L.lzma_end(C.byref(s))
