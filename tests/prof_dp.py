import os, sys, time
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import xz_b200
xz_b200.LIB_PATH = os.path.abspath(os.environ.get("XZB200_LIB", "xz_b200/libxzb200_prof.so"))
import xzlibs as X
MiB = 1 << 20
for kind, preset, n in (("T", 6, 4 * MiB), ("E", 6, 4 * MiB)):
    buf = X.gendata(kind, n)
    ctx = xz_b200.Context(0)
    ctx.stream_encode(buf, preset=preset, block_size=n, n=n)
    s = ctx.stats().as_dict()
    print(kind, preset, s, flush=True)
    ctx.close()
