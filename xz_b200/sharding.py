"""Multi-GPU plumbing of the path: independent .xz Blocks are sharded over ranks (contiguous
ranges, no halo); the only exchange is one all-gather of the 16-byte Index records
(unpadded_size, uncompressed_size) -- what lzma_index_append() collects from the worker threads in
the reference (common/stream_encoder_mt.c:756-770).  Works with any torch.distributed backend
(NCCL on GPUs, gloo in the CPU tests)."""
import torch
import torch.distributed as dist


def shard_blocks(nblocks, world, rank):
    """Contiguous block range [lo, hi) of `rank`; earlier ranks take the remainder."""
    base, rem = divmod(nblocks, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_records(records, device="cpu"):
    """All-gather of per-rank Index records (list of (unpadded, uncompressed)) -> full list in block order."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return [tuple(r) for r in records]
    world = dist.get_world_size()
    counts = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(counts, torch.tensor([len(records)], dtype=torch.int64, device=device))
    mx = max(int(c.item()) for c in counts)
    t = torch.zeros((mx, 2), dtype=torch.int64, device=device)
    if records:
        t[: len(records)] = torch.tensor(records, dtype=torch.int64, device=device)
    outs = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(outs, t)
    full = []
    for c, o in zip(counts, outs):
        full += [tuple(int(v) for v in row) for row in o[: int(c.item())].cpu().tolist()]
    return full


def block_offsets(records, start=12):
    """Byte offset of every Block in the Stream: 12 + sum of padded sizes before it (block_util.c:80-89)."""
    offs, pos = [], start
    for u, _ in records:
        offs.append(pos)
        pos += (u + 3) // 4 * 4
    return offs, pos


def assemble_stream(check, shard_bytes_in_rank_order, records):
    """Stream Header + Blocks (rank order == block order) + Index + Stream Footer."""
    from . import index_encode, stream_footer, stream_header
    idx = index_encode(records)
    return stream_header(check) + b"".join(shard_bytes_in_rank_order) + idx + stream_footer(check, len(idx))
