"""world_size-2 gloo test of the N>1 path's host logic: contiguous block sharding, the all-gather
of 16-byte Index records, and Stream reassembly.  The per-rank "encoder" here is the oracle (no
GPU in this container); on the GPU the same functions run over NCCL (bench.py)."""
import hashlib
import os
import socket
import sys

import pytest
import torch.multiprocessing as mp

import xzlibs as X


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, n, bs, preset, q):
    import ctypes as C
    import torch.distributed as dist
    sys.path.insert(0, X.ROOT); sys.path.insert(0, os.path.join(X.ROOT, "tests"))
    from xz_b200 import sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    nblocks = (n + bs - 1) // bs
    lo, hi = sharding.shard_blocks(nblocks, world, rank)
    off, end = lo * bs, min(hi * bs, n)
    buf = X.gendata("T", n)
    shard = (C.c_uint8 * max(end - off, 1)).from_buffer_copy(bytes(buf[off:end]) or b"\0")
    xz = X.oracle_encode(shard, end - off, preset, bs) if end > off else b""
    # strip Stream Header / Index / Footer: recover this shard's blocks + records from the oracle's stream
    recs, blocks = [], b""
    if xz:
        r, _ = X.oracle_decode(xz, end - off)
        assert r == 0
        pos = 12
        for b in range(hi - lo):
            hs = (xz[pos] + 1) * 4
            # compressed size VLI at header offset 2
            v, sh, p = 0, 0, pos + 2
            while True:
                c = xz[p]; p += 1; v |= (c & 0x7F) << sh; sh += 7
                if not c & 0x80: break
            unp = hs + v + 8
            recs.append((unp, min(bs, end - off - b * bs)))
            pos += (unp + 3) // 4 * 4
        blocks = xz[12:pos]
    full = sharding.gather_records(recs)
    offs, total = sharding.block_offsets(full)
    gathered = [None] * world
    dist.all_gather_object(gathered, blocks)
    if rank == 0:
        stream = sharding.assemble_stream(4, gathered, full)
        q.put((hashlib.sha256(stream).hexdigest(), len(full), offs[lo:hi][:1], total))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n,bs", [(5 * 65536 + 100, 65536), (65536, 65536), (3 * 65536, 65536)])
def test_two_rank_index_gather_and_reassembly(n, bs):
    world, preset = 2, 1
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, bs, preset, q)) for r in range(world)]
    for p in procs: p.start()
    sha, nrec, first_off, total = q.get(timeout=120)
    for p in procs: p.join(timeout=60)
    assert all(p.exitcode == 0 for p in procs)
    want = X.oracle_encode(X.gendata("T", n), n, preset, bs)
    assert sha == hashlib.sha256(want).hexdigest()
    assert nrec == (n + bs - 1) // bs and first_off == [12]


def test_shard_blocks_partition():
    from xz_b200 import sharding
    for nb in (0, 1, 7, 64, 65, 512):
        for w in (1, 2, 4, 8):
            r = [sharding.shard_blocks(nb, w, k) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == nb and all(r[i][1] == r[i + 1][0] for i in range(w - 1))
            assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1
