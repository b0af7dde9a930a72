"""Ad-hoc GPU probe (not a test): timing of the encode/decode paths at a few sizes."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as ge
ge.build()
import xz_b200, xzlibs as X

ctx = xz_b200.Context(0)
MiB = 1 << 20
cases = [("T", 6, 1 * MiB, 1 * MiB), ("T", 1, 4 * MiB, 1 * MiB), ("R", 3, 4 * MiB, 1 * MiB), ("T", 6, 8 * MiB, 1 * MiB), ("E", 6, 8 * MiB, 1 * MiB)]
if len(sys.argv) > 1:
    cases = [tuple(int(v, 0) if v[0].isdigit() else v for v in a.split(",")) for a in sys.argv[1:]]
for kind, preset, n, bs in cases:
    buf = X.gendata(kind, n)
    t = time.time(); mine = ctx.stream_encode(buf, preset=preset, block_size=bs, n=n); dt = time.time() - t
    s = ctx.stats().as_dict()
    t = time.time(); want = X.oracle_encode(buf, n, preset, bs); dto = time.time() - t
    print(f"{kind} -{preset} n={n} bs={bs}: match={mine == want} gpu {dt:.2f}s oracle {dto:.2f}s  mf_prep {s['ms_mf_prep']:.1f} mf {s['ms_mf']:.1f} parse {s['ms_parse']:.1f} other {s['ms_other']:.1f} ms; fallback {s['n_fallback_blocks']}", flush=True)
    t = time.time(); r, back = ctx.stream_decode(mine, n); dt = time.time() - t
    s = ctx.stats().as_dict()
    print(f"   decode ret={r} ok={back == bytes(buf[:n])} {dt:.2f}s decode kernel {s['ms_decode']:.1f} ms", flush=True)
