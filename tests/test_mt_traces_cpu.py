"""The reference's THREADED decoder (lzma_stream_decoder_mt, stream_decoder_mt.c) against the recorded single-threaded
traces: for every (corpus file, flags) case of tests/golden/stream_trace_golden.json its lzma_code() return-code
sequence, lzma_get_check() values and output bytes are the same with threads = 4.  That is what lets one golden file
pin both liblzma-named decoder entry points of libxzb200.so (tests/test_gpu_lzma_api.py).  Runs the unmodified
reference (oracle/_ref) on the CPU; no GPU involved."""
import ctypes as C
import hashlib
import json
import os
import sys

import pytest

import xzlibs as X

GOLD = os.path.join(X.ROOT, "tests", "golden")
pytestmark = pytest.mark.skipif(not X.have_ref(), reason="oracle/_ref not built")


def test_reference_mt_decoder_sequences_equal_the_recorded_ones():
    sys.path.insert(0, GOLD)
    import make_golden as MG
    gold = json.load(open(os.path.join(GOLD, "stream_trace_golden.json")))
    inputs = {n: open(os.path.join(GOLD, "ref_files", n), "rb").read() for n in os.listdir(os.path.join(GOLD, "ref_files"))}
    inputs.update(MG.trace_inputs())
    lib = X.ref()
    norm = lambda cs: [c if c[0] <= 4 else [c[0], None] for c in cs]   # the check ID after an error code is whatever was there
    cap = 1 << 22
    out = (C.c_uint8 * cap)()
    bad = []
    for key, want in sorted(gold.items()):
        name, fl = key.split("|")
        data = inputs[name]
        sz = C.c_size_t(); codes = (C.c_uint32 * 32)(); n = C.c_uint32()
        lib.ref_decode_trace_mt(data, C.c_size_t(len(data)), C.c_uint32(int(fl, 16)), C.c_uint32(4), out, C.c_size_t(cap), C.byref(sz), codes, 32, C.byref(n))
        got = [[codes[i] & 0xFF, codes[i] >> 8] for i in range(n.value)]
        o = bytes(out[:sz.value])
        if not (norm(got) == norm([list(c) for c in want["codes"]]) and len(o) == want["out_size"]
                and hashlib.sha256(o).hexdigest() == want["out_sha256"]):
            bad.append(key)
    assert len(gold) >= 700 and not bad, bad[:10]
