#!/usr/bin/env python
"""Per-Block golden vectors of the bench workloads, from the UNMODIFIED reference (oracle/_ref).

Runs lzma_stream_encoder_mt of the reference (build container only) over the synthetic inputs of
BASELINE.json's GPU configs and records, for every .xz Block, its total size (header + data +
padding + check) and SHA-256, plus the SHA-256 of the whole Stream.  bench.py compares ALL Blocks of
a run against these without any CPU work inside the run; tests/test_gpu_parity.py uses them too.

    python tests/golden/make_bench_golden.py [T6 E9e R3 ...]

Entries: name -> {kind, preset, block_size, nblocks, blocks: [[size, sha256], ...], stream_sha256}.
The first `nblocks` Blocks of the infinite synthetic stream `kind` (xz_b200/csrc/xzgen.c).
"""
import hashlib, json, os, sys, time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import xzlibs as X

MiB = 1 << 20
OUT = os.path.join(HERE, "bench_golden.json")
CASES = {
    # name: (kind, preset, block_size, nblocks)
    "T6": ("T", 6, 16 * MiB, 64),                        # configs[1] / [2]: the whole 1 GiB
    "E9e": ("E", 9 | X.XZ_PRESET_EXTREME, 16 * MiB, 32),  # configs[3]: one GPU's share (32 of 256 Blocks)
    "R3": ("R", 3, 16 * MiB, 64),                        # configs[4]: one GPU's share (64 of 512 Blocks)
}


def split_blocks(xz, n, bs):
    """Cut a Stream produced with block sizes known a priori into its Blocks using the Block Headers."""
    out = []
    pos = 12
    nb = (n + bs - 1) // bs
    for _ in range(nb):
        hs = (xz[pos] + 1) * 4
        # compressed size VLI follows the flags byte (MT encoder always stores both sizes)
        p = pos + 2
        comp = 0; shift = 0
        while True:
            b = xz[p]; p += 1
            comp |= (b & 0x7F) << shift; shift += 7
            if not b & 0x80:
                break
        total = hs + ((comp + 3) & ~3) + 8
        out.append(xz[pos:pos + total])
        pos += total
    return out


def main():
    assert X.have_ref(), "build oracle/_ref first"
    want = sys.argv[1:] or list(CASES)
    db = json.load(open(OUT)) if os.path.exists(OUT) else {}
    for name in want:
        kind, preset, bs, nb = CASES[name]
        n = nb * bs
        t = time.time()
        buf = X.gendata(kind, n)
        xz = X.ref_encode(buf, n, preset, bs, threads=0)
        blocks = split_blocks(xz, n, bs)
        db[name] = {"kind": kind, "preset": preset, "block_size": bs, "nblocks": nb,
                    "blocks": [[len(b), hashlib.sha256(b).hexdigest()] for b in blocks],
                    "stream_sha256": hashlib.sha256(xz).hexdigest(), "stream_size": len(xz)}
        print(name, "done in %.0f s, %d bytes" % (time.time() - t, len(xz)), flush=True)
        json.dump(db, open(OUT, "w"), indent=0)


if __name__ == "__main__":
    main()
