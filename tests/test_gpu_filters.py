"""Delta / BCJ filter chains on the GPU path (-m gpu): Streams made with chains are byte-identical to the unmodified
reference's lzma_stream_encoder_mt with the same chain (oracle/_ref, run live on the same inputs), and Streams of the
reference decode to the input through xzb_k_decode + xzb_k_filter."""
import ctypes as C
import os
import random

import pytest

import xzlibs as X

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not X.have_ref(), reason="oracle/_ref not built")]
DELTA, X86, POWERPC, IA64, ARM, ARMTHUMB, SPARC, ARM64, RISCV = 3, 4, 5, 6, 7, 8, 9, 10, 11
KiB = 1 << 10


@pytest.fixture(scope="module")
def ctx():
    import xz_b200
    c = xz_b200.Context(0)
    yield c
    c.close()


def ref_chain_encode(data, chain, preset, bs, check=4):
    n = len(data)
    cap = n + n // 2 + 65536
    out = (C.c_uint8 * cap)()
    sz = C.c_size_t()
    ids = (C.c_uint32 * len(chain))(*[c[0] for c in chain])
    args = (C.c_uint32 * len(chain))(*[c[1] for c in chain])
    r = X.ref().ref_encode_mt_chain(data, C.c_size_t(n), ids, args, C.c_uint32(len(chain)), C.c_uint32(preset), C.c_uint64(bs), C.c_uint32(check),
                                    C.c_uint32(4), out, C.c_size_t(cap), C.byref(sz))
    assert r == 0, r
    return bytes(out[:sz.value])


def mixed_input(n, seed):
    """Text, a ramp (delta-friendly), and random bytes salted with branch opcodes of every architecture."""
    from test_filters_cpu import codeish
    rnd = random.Random(seed)
    parts = []
    t = bytes(X.gendata("T", n // 3)[: n // 3])
    parts.append(t)
    parts.append(bytes((i * 3 + (i >> 8)) & 0xFF for i in range(n // 3)))
    rest = n - 2 * (n // 3)
    per = rest // 9
    for fid in (X86, ARM, ARMTHUMB, POWERPC, SPARC, ARM64, IA64, RISCV):
        parts.append(codeish(fid, per, seed + fid))
    parts.append(bytes(rnd.getrandbits(8) for _ in range(rest - 8 * per)))
    return b"".join(parts)


CHAINS = [[(DELTA, 1)], [(DELTA, 4)], [(DELTA, 256)], [(X86, 0)], [(X86, 0x1000)], [(ARM, 0)], [(ARMTHUMB, 0)], [(POWERPC, 0)], [(SPARC, 0)],
          [(ARM64, 0)], [(ARM64, 0x40000)], [(IA64, 0)], [(RISCV, 0)], [(RISCV, 0x2000), (DELTA, 2)], [(DELTA, 2), (X86, 0)], [(ARM64, 0), (DELTA, 4)], [(X86, 0), (DELTA, 1), (ARM, 0)]]


@pytest.mark.parametrize("chain", CHAINS, ids=lambda c: "+".join(f"{i:x}.{a:x}" for i, a in c))
def test_chain_encode_identical_and_decode(ctx, chain):
    n = 700 * KiB + 123
    data = mixed_input(n, 17)
    want = ref_chain_encode(data, chain, 6, 256 * KiB)
    ctx.set_filters(chain)
    try:
        got = ctx.stream_encode(data, preset=6, block_size=256 * KiB, n=n)
    finally:
        ctx.set_filters(())
    assert got == want
    r, back = ctx.stream_decode(want, n)
    assert r == 0 and back == data


@pytest.mark.parametrize("preset", [0, 3])
def test_chain_fast_presets_and_incompressible_fallback(ctx, preset):
    """Random bytes: every Block falls back to uncompressed LZMA2 chunks of the UNFILTERED input under an LZMA2-only
    header (block_buffer_encoder.c:87-162), whatever the chain says."""
    n = 300 * KiB
    rnd = random.Random(5)
    data = bytes(rnd.getrandbits(8) for _ in range(n))
    for chain in ([(X86, 0)], [(DELTA, 3), (ARM64, 0)]):
        want = ref_chain_encode(data, chain, preset, 128 * KiB)
        ctx.set_filters(chain)
        try:
            got = ctx.stream_encode(data, preset=preset, block_size=128 * KiB, n=n)
        finally:
            ctx.set_filters(())
        assert got == want
        r, back = ctx.stream_decode(want, n)
        assert r == 0 and back == data


def test_bad_chains_are_refused(ctx):
    import xz_b200
    for chain in ([(DELTA, 0)], [(DELTA, 257)], [(ARM, 2)], [(IA64, 8)], [(RISCV, 1)], [(0x0C, 0)], [(0x21, 0)], [(X86, 0)] * 4):
        with pytest.raises(xz_b200.XzError):
            ctx.set_filters(chain)
    ctx.set_filters(())
