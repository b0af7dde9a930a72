"""CPU tests of the boundary: the C-ABI library loads and exports everything include/*.h declares,
liblzma struct layouts match the reference headers, option validation returns liblzma's codes
(no compute calls -- there is no GPU here, and no CPU fallback to call)."""
import ctypes as C
import os
import re
import subprocess
import tempfile

import pytest

import xzlibs as X

ROOT = X.ROOT


def _declared(header):
    txt = open(os.path.join(ROOT, "include", header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b((?:xzb|lzma)_[a-z0-9_]+)\s*\(", txt)) - {"lzma_internal_s"})


def test_library_exports_every_declared_symbol():
    import xz_b200
    lib = xz_b200.lib()
    names = _declared("xzb200.h") + _declared("xzb200_lzma.h")
    assert len(names) > 25
    for n in names:
        assert hasattr(lib, n), n


def test_no_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import xz_b200
    with pytest.raises(xz_b200.XzError):
        xz_b200.Context(0)


LAYOUT_PROG = r'''
#include <stdio.h>
#include <stddef.h>
#include HEADER
#define P(T, f) printf(#T "." #f " %zu %zu\n", offsetof(T, f), sizeof(((T *)0)->f))
int main(void) {
	printf("sizeof %zu %zu %zu %zu %zu\n", sizeof(lzma_stream), sizeof(lzma_mt), sizeof(lzma_options_lzma), sizeof(lzma_filter), sizeof(lzma_allocator));
	P(lzma_stream, next_in); P(lzma_stream, avail_in); P(lzma_stream, total_in); P(lzma_stream, next_out); P(lzma_stream, avail_out);
	P(lzma_stream, total_out); P(lzma_stream, allocator); P(lzma_stream, internal); P(lzma_stream, seek_pos); P(lzma_stream, reserved_enum2);
	P(lzma_mt, flags); P(lzma_mt, threads); P(lzma_mt, block_size); P(lzma_mt, timeout); P(lzma_mt, preset); P(lzma_mt, filters); P(lzma_mt, check);
	P(lzma_mt, memlimit_threading); P(lzma_mt, memlimit_stop); P(lzma_mt, reserved_ptr4);
	P(lzma_options_lzma, dict_size); P(lzma_options_lzma, preset_dict); P(lzma_options_lzma, preset_dict_size); P(lzma_options_lzma, lc);
	P(lzma_options_lzma, lp); P(lzma_options_lzma, pb); P(lzma_options_lzma, mode); P(lzma_options_lzma, nice_len); P(lzma_options_lzma, mf);
	P(lzma_options_lzma, depth); P(lzma_options_lzma, ext_flags); P(lzma_options_lzma, reserved_ptr2);
	P(lzma_filter, id); P(lzma_filter, options);
	printf("enums %d %d %d %d %d %d %d %d\n", LZMA_FINISH, LZMA_FULL_BARRIER, LZMA_FULL_FLUSH, LZMA_BUF_ERROR, LZMA_PROG_ERROR, LZMA_CHECK_CRC64, LZMA_MF_BT4, LZMA_MODE_NORMAL);
	return 0;
}
'''


@pytest.mark.skipif(not os.path.isdir("/root/reference/src/liblzma/api"), reason="reference headers not present")
def test_struct_layouts_match_reference_headers():
    outs = []
    with tempfile.TemporaryDirectory() as d:
        for i, (hdr, inc) in enumerate((("<lzma.h>", "/root/reference/src/liblzma/api"), ('"xzb200_lzma.h"', os.path.join(ROOT, "include")))):
            src = os.path.join(d, f"l{i}.c")
            open(src, "w").write(LAYOUT_PROG.replace("HEADER", hdr))
            exe = os.path.join(d, f"l{i}")
            subprocess.check_call(["gcc", "-I", inc, src, "-o", exe])
            outs.append(subprocess.check_output([exe], text=True))
    assert outs[0] == outs[1]


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


def _mt(**kw):
    m = LzmaMt()
    m.threads, m.preset, m.check, m.block_size = 1, 6, 4, 1 << 20
    for k, v in kw.items():
        setattr(m, k, v)
    return m


BAD_OPTIONS = [({"flags": 1}, 8), ({"threads": 0}, 8), ({"threads": 20000}, 8), ({"preset": 10}, 8), ({"preset": 6 | 0x40000000}, 8),
               ({"check": 10}, None), ({"check": 16}, 11), ({"block_size": (1 << 64) - 1}, 8)]


@pytest.mark.parametrize("kw,want", BAD_OPTIONS)
def test_encoder_mt_option_validation(kw, want):
    """lzma_stream_encoder_mt rejects bad lzma_mt fields with the reference's codes
    (stream_encoder_mt.c:955-1000, 1052-1060) before any device work."""
    import xz_b200
    lib = xz_b200.lib()
    s = LzmaStream()
    m = _mt(**kw)
    got = lib.lzma_stream_encoder_mt(C.byref(s), C.byref(m))
    if X.have_ref():
        ref = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "liblzma_ref.so"))
        rs = LzmaStream()
        ref_ret = ref.lzma_stream_encoder_mt(C.byref(rs), C.byref(_mt(**kw)))
        ref.lzma_end(C.byref(rs))
        if want is not None:
            assert ref_ret == want
        else:
            want = 3  # SHA-256 is outside the GPU path's scope: LZMA_UNSUPPORTED_CHECK (the reference supports it)
    assert got == (want if want is not None else 3)
    assert not s.internal


def test_lzma_code_argument_checks_without_coder():
    import xz_b200
    lib = xz_b200.lib()
    s = LzmaStream()
    assert lib.lzma_code(C.byref(s), 0) == 11  # internal == NULL -> LZMA_PROG_ERROR (common.c:206-213)
    lib.lzma_end(C.byref(s))  # no-op on a fresh stream


def test_preset_table_matches_oracle():
    import xz_b200
    for p in list(range(10)) + [i | 0x80000000 for i in range(10)]:
        a, b = xz_b200.lzma_lzma_preset(p), X.preset_options(p)
        assert [getattr(a, f) for f, _ in a._fields_] == [getattr(b, f) for f, _ in b._fields_]
    assert xz_b200.lib().xzb_block_bound(1 << 24) == X.oracle().xzo_block_bound(1 << 24) == 16778080


def test_framing_helpers_match_oracle():
    """Stream Header / Index / Footer built by the product's host code == oracle's."""
    import xz_b200
    recs = [(16778072, 1 << 24), (100, 5), (70000, 1 << 20), (3000005, 1 << 24)]
    idx = xz_b200.index_encode(recs)
    U = (C.c_uint64 * len(recs))(*[r[0] for r in recs])
    V = (C.c_uint64 * len(recs))(*[r[1] for r in recs])
    o = X.oracle()
    o.xzo_index_encode.restype = C.c_size_t
    n = o.xzo_index_encode(U, V, C.c_size_t(len(recs)), None)
    buf = (C.c_uint8 * n)()
    o.xzo_index_encode(U, V, C.c_size_t(len(recs)), buf)
    assert idx == bytes(buf)
    hdr = (C.c_uint8 * 12)(); o.xzo_stream_header(hdr, C.c_uint32(4))
    ftr = (C.c_uint8 * 12)(); o.xzo_stream_footer(ftr, C.c_uint32(4), C.c_uint64(len(idx)))
    assert xz_b200.stream_header(4) == bytes(hdr) and xz_b200.stream_footer(4, len(idx)) == bytes(ftr)
