"""The drop-in boundary under the UNMODIFIED reference `xz` (src/xz, built by oracle/Makefile.ref): CPU-side checks.

oracle/_ref/xz      = reference xz + reference liblzma                       (the binary to compare with)
oracle/_ref/xz_gpu  = the same objects with libxzb200.so ahead of liblzma    (the hybrid of INTEGRATION.md)

No GPU work here: layout of lzma_stream.internal against the reference header, symbol resolution of the hybrid,
the reference's own coders driven through this library's generic lzma_code, and the loud failure without CUDA."""
import os
import re
import subprocess
import tempfile

import pytest

import xzlibs as X

ROOT = X.ROOT
REF = "/root/reference"
XZ = os.path.join(ROOT, "oracle", "_ref", "xz")
XZ_GPU = os.path.join(ROOT, "oracle", "_ref", "xz_gpu")
needs_bins = pytest.mark.skipif(not (os.path.exists(XZ) and os.path.exists(XZ_GPU)), reason="oracle/_ref/xz[_gpu] not built")


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present")
def test_internal_layout_matches_reference_header(tmp_path):
    """xzb_lzma_api.cpp restates lzma_next_coder_s / lzma_internal_s (common/common.h:222-324); a C probe compiled
    against the reference's own header must report the offsets the shim's static_assert and code assume."""
    src = tmp_path / "probe.c"
    src.write_text(r'''
#include "common.h"
#include <stdio.h>
#include <stddef.h>
int main(void) {
	printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu ", offsetof(lzma_next_coder, coder), offsetof(lzma_next_coder, id),
		offsetof(lzma_next_coder, init), offsetof(lzma_next_coder, code), offsetof(lzma_next_coder, end),
		offsetof(lzma_next_coder, get_progress), offsetof(lzma_next_coder, get_check), offsetof(lzma_next_coder, memconfig),
		offsetof(lzma_next_coder, update), offsetof(lzma_next_coder, set_out_limit), sizeof(lzma_next_coder));
	printf("%zu %zu %zu %zu %zu %zu %d %d\n", offsetof(lzma_internal, next), offsetof(lzma_internal, sequence),
		offsetof(lzma_internal, avail_in), offsetof(lzma_internal, supported_actions), offsetof(lzma_internal, allow_buf_error),
		sizeof(lzma_internal), (int)LZMA_ACTION_MAX, (int)LZMA_TIMED_OUT);
	return 0;
}''')
    L = REF + "/src/liblzma"
    exe = tmp_path / "probe"
    subprocess.run(["gcc", "-std=gnu99", "-DHAVE_STDBOOL_H", "-DHAVE_STDINT_H", "-DHAVE_INTTYPES_H", "-DHAVE_STRING_H", "-DHAVE_LIMITS_H",
                    "-DMYTHREAD_POSIX", "-DSIZEOF_SIZE_T=8", "-DHAVE_VISIBILITY=1", f"-I{L}/api", f"-I{L}/common", f"-I{REF}/src/common",
                    str(src), "-o", str(exe)], check=True)
    got = [int(v) for v in subprocess.run([str(exe)], stdout=subprocess.PIPE, text=True, check=True).stdout.split()]
    assert got == [0, 8, 16, 24, 32, 40, 48, 56, 64, 72, 80, 0, 80, 88, 96, 101, 104, 4, 101]


@needs_bins
def test_hybrid_resolves_every_symbol_xz_imports():
    """Every lzma_* symbol the unmodified xz imports is defined by libxzb200.so or by the reference liblzma behind it."""
    def syms(path, flag):
        out = subprocess.run(["nm", "-D", flag, path], stdout=subprocess.PIPE, text=True, check=True).stdout
        return {ln.split()[-1] for ln in out.splitlines() if re.search(r"\blzma_", ln)}
    imports = syms(XZ_GPU, "--undefined-only")
    ours = syms(os.path.join(ROOT, "xz_b200", "libxzb200.so"), "--defined-only")
    theirs = syms(os.path.join(ROOT, "oracle", "_ref", "liblzma_ref.so"), "--defined-only")
    assert len(imports) >= 37
    assert imports <= (ours | theirs), sorted(imports - ours - theirs)
    # the stream coders of the hot path and the generic drivers come from this library
    for name in ("lzma_stream_encoder_mt", "lzma_stream_decoder_mt", "lzma_code", "lzma_end", "lzma_memusage", "lzma_get_progress",
                 "lzma_filters_update", "lzma_stream_encoder_mt_memusage", "lzma_mt_block_size"):
        assert name in ours and name in imports | ours
    out = subprocess.run(["ldd", XZ_GPU], stdout=subprocess.PIPE, text=True).stdout
    assert out.index("libxzb200.so") < out.index("liblzma_ref.so")   # lookup order = link order


@needs_bins
def test_reference_coders_run_through_this_librarys_lzma_code(tmp_path):
    """In the hybrid, lzma_code/lzma_end bind to libxzb200.so.  They are generic over the reference's coder vtable, so the
    reference's single-threaded .xz encoder, its .lzma coders and its file-info decoder (xz -l) behave exactly as under
    the reference's own lzma_code -- byte-identical output, no GPU involved."""
    buf = X.gendata("T", 600000)
    f = tmp_path / "in.bin"
    f.write_bytes(bytes(buf[:600000]))
    a = subprocess.run([XZ, "-6", "-T1", "-c", str(f)], stdout=subprocess.PIPE, check=True).stdout
    b = subprocess.run([XZ_GPU, "-6", "-T1", "-c", str(f)], stdout=subprocess.PIPE, check=True).stdout
    assert a == b
    a = subprocess.run([XZ, "--format=lzma", "-c", str(f)], stdout=subprocess.PIPE, check=True).stdout
    b = subprocess.run([XZ_GPU, "--format=lzma", "-c", str(f)], stdout=subprocess.PIPE, check=True).stdout
    assert a == b
    back = subprocess.run([XZ_GPU, "--format=lzma", "-dc"], input=b, stdout=subprocess.PIPE, check=True).stdout
    assert back == f.read_bytes()
    g = tmp_path / "x.xz"
    g.write_bytes(subprocess.run([XZ, "-6", "-T1", "-c", str(f)], stdout=subprocess.PIPE, check=True).stdout)
    la = subprocess.run([XZ, "-l", str(g)], stdout=subprocess.PIPE, text=True, check=True).stdout
    lb = subprocess.run([XZ_GPU, "-l", str(g)], stdout=subprocess.PIPE, text=True, check=True).stdout
    assert la == lb


@needs_bins
def test_hybrid_fails_loudly_without_cuda(tmp_path):
    """No CPU fallback: the threaded encoder of the hybrid needs the GPU; without one xz reports an error and exits 1."""
    import xz_b200
    try:
        ctx = xz_b200.Context(0)
        ctx.close()
        pytest.skip("a CUDA device is present")
    except Exception:
        pass
    f = tmp_path / "in.bin"
    f.write_bytes(b"hello " * 1000)
    r = subprocess.run([XZ_GPU, "-6", "-T2", "-c", str(f)], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert r.returncode == 1 and b"no usable CUDA device" in r.stderr and r.stdout == b""
