"""Development probe: one 4 MiB Block at -6 through the library given in XZB200_LIB (default: the XZB_DP_PROF build),
checked against the oracle; the kernel prints its DPPROF cycle counters."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import xz_b200
xz_b200.LIB_PATH = os.path.abspath(os.environ.get("XZB200_LIB", "xz_b200/libxzb200_prof.so"))
import xzlibs as X
MiB = 1 << 20
for kind, preset, n in (("T", 6, 4 * MiB), ("E", 6, 4 * MiB)):
    buf = X.gendata(kind, n)
    ctx = xz_b200.Context(0)
    got = ctx.stream_encode(buf, preset=preset, block_size=n, n=n)
    s = ctx.stats().as_dict()
    ctx.close()
    ok = got == X.oracle_encode(buf, n, preset, n)
    print(kind, preset, "OK" if ok else "MISMATCH", "ms_parse %.1f ms_total %.1f" % (s["ms_parse"], s["ms_total"]), flush=True)
