"""Development probe: one filter chain through the GPU encoder (argument: chain as id.arg+id.arg, hex), checked against the reference."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import xz_b200, xzlibs as X
from test_gpu_filters import mixed_input, ref_chain_encode
chain = [tuple(int(v, 16) for v in part.split(".")) for part in sys.argv[1].split("+")]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 700 * 1024 + 123
data = mixed_input(n, 17)
ctx = xz_b200.Context(0)
ctx.set_filters(chain)
got = ctx.stream_encode(data, preset=6, block_size=256 * 1024, n=n)
want = ref_chain_encode(data, chain, 6, 256 * 1024)
print("chain", chain, "encode", "OK" if got == want else "MISMATCH", len(got), len(want), flush=True)
r, back = ctx.stream_decode(want, n)
print("decode", r, back == data, flush=True)
