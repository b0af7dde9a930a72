"""Ad-hoc A/B probe (not a test): time encode/decode kernels with the library given in XZB200_LIB."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import xz_b200, xzlibs as X
if os.environ.get("XZB200_LIB"):
    xz_b200.LIB_PATH = os.path.abspath(os.environ["XZB200_LIB"])
ctx = xz_b200.Context(0)
for a in sys.argv[1:]:
    kind, preset, n, bs = [int(v, 0) if v[0].isdigit() else v for v in a.split(",")]
    buf = X.gendata(kind, n)
    mine = ctx.stream_encode(buf, preset=preset, block_size=bs, n=n)
    s = ctx.stats().as_dict()
    line = f"{os.environ.get('AB_TAG', '')} {kind} -{preset} n={n} bs={bs}: total {s['ms_total']:.1f} prep {s['ms_mf_prep']:.1f} mf {s['ms_mf']:.1f} parse {s['ms_parse']:.1f}"
    if os.environ.get("AB_NO_DECODE"):
        print(line, flush=True)
        continue
    best = 1e9
    for _ in range(3):
        r, back = ctx.stream_decode(mine, n)
        assert r == 0 and back == bytes(buf[:n])
        best = min(best, ctx.stats().as_dict()["ms_decode"])
    print(line + f" decode {best:.1f} ms", flush=True)
