#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric on BASELINE.json's config.

    python bench.py --gpus N --steps K --warmup W [--impl reference]

Workload (config.workload): configs[1] = "xz -6 bt4, 8 MiB dict, 1 GiB synthetic text, 64 .xz
blocks" (16 MiB blocks, CRC64), strong-scaled over N GPUs: rank r owns blocks
[r*64/N, (r+1)*64/N); the only exchange is one all-gather of the 16-byte Index records.
A step = one encode pass of the hot path over the whole batch:
  * `value`     : MB/s (1e6 B of uncompressed input / s) with the input resident in HBM and the
                  Blocks left in HBM (xzb_encode_blocks_device), CUDA-event timed on the library's
                  stream, max over ranks;
  * `e2e.value` : same metric through the reference-facing host-buffer call (pinned host input ->
                  finished Blocks in host memory + Index), host<->device copies inside the timing;
  * `decode`    : decode MB/s of the stream just produced (configs[2]), device-timed and e2e;
  * `roofline`  : the match-finder kernel (xzb_k_bt runs as one launch per 2^20-position segment on its own
                  stream beside the parser kernel; time = CUDA events on that stream over the step's launches);
  * `cpu_baseline`: the unmodified reference on all host threads over the same blocks, encode (`value`) and
                  threaded decode of the produced Stream (`decode_value`).
`--impl reference` times the UNMODIFIED reference (oracle/_ref, lzma_stream_encoder_mt with all
host threads) on a bounded sample of the same workload.
Input larger than L2 (1 GiB vs 126 MB) => no explicit L2 flush between timed iterations.
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

MiB = 1 << 20
METRIC = "encode MB/s at -6, 1 GiB synthetic text, 64 x 16 MiB .xz blocks (bit-exact); decode MB/s beside it"


def measured_hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(opts, kind, algorithmic_bytes):
    """dram__bytes_read.sum + dram__bytes_write.sum of the match-finder kernel per launch, from the
    round's `ncu --set full` captures (profiles/r01_{bt,hc}_ncu.txt), scaled by inserted positions
    (the captures ran 8 x 16 MiB 'T' at -6 for xzb_k_bt and 256 MiB 'R' at -3 for xzb_k_hc).  None
    for workloads that were not captured."""
    is_bt = bool(opts.mf & 0x10)
    if is_bt and kind == "T":
        per_pos = (261.271145e9 + 51.672717e9) / (8 * (16 * MiB - 3))
        return per_pos * algorithmic_bytes / 33.0
    if not is_bt and kind == "R":
        per_pos = (3.926867e9 + 1.241704e9) / (16 * (16 * MiB - 3))
        return per_pos * algorithmic_bytes / 29.0
    return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for i, nme in enumerate(names):
                if f[2 + i].lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def run_reference(args):
    """Reference arm: lzma_stream_encoder_mt of the unmodified reference on the host cores."""
    import xzlibs as X
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if not X.have_ref():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libref_shim.so is not built"}))
        return
    cores = X.ref().ref_cputhreads() or os.cpu_count() or 1
    nblocks_total = args.size // args.block_size
    # One step = one pass of lzma_stream_encoder_mt over `sample_blocks` Blocks of the same input with
    # all host threads (threads in use = min(cores, blocks)).  The whole workload (64 Blocks) takes
    # about a minute per pass on a 128-thread host, so with many steps the sample shrinks to keep
    # the run within a few minutes; the sample actually used is reported.
    sample_blocks = max(1, min(nblocks_total, cores))
    passes = args.warmup + args.steps
    if passes > 4:
        sample_blocks = max(min(16, sample_blocks), sample_blocks * 4 // passes)
    n = sample_blocks * args.block_size
    buf = X.gendata(args.kind, n)
    times = []
    out_len = 0
    for it in range(args.warmup + args.steps):
        t = time.perf_counter()
        out = X.ref_encode(buf, n, args.preset, args.block_size, threads=0)
        dt = time.perf_counter() - t
        out_len = len(out)
        if it >= args.warmup:
            times.append(dt)
    t_step = sum(times) / len(times)
    val = n / 1e6 / t_step
    # decode side of the reference on the same sample (output buffer allocated outside the timing)
    dbuf = (C.c_uint8 * n)(); dsz = C.c_size_t()
    t = time.perf_counter()
    r = X.ref().ref_decode_mt(out, C.c_size_t(len(out)), C.c_uint32(0), dbuf, C.c_size_t(n), C.byref(dsz))
    dt_dec = time.perf_counter() - t
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "MB/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": t_step * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": workload_config(args),
        "cpu_baseline": {"value": val, "unit": "MB/s", "cores": int(cores), "kind": "reference",
                         "sample": f"{sample_blocks} x {args.block_size // MiB} MiB blocks of the same input, lzma_stream_encoder_mt threads={cores}, in RAM"},
        "e2e": {"value": val, "unit": "MB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "decode": {"value": n / 1e6 / dt_dec if r == 0 else None, "unit": "MB/s", "note": "lzma_stream_decoder_mt, same sample"},
        "xz_bytes": out_len,
    }
    print(json.dumps(line))


def workload_config(args):
    return {"workload": f"xz -{args.preset & 0x1F}{'e' if args.preset & 0x80000000 else ''} LZMA2, {args.size // MiB} MiB synthetic '{args.kind}' "
                        f"({ {'T': 'Lorem-word text, tests/create_compress_files.c generator scaled', 'E': 'enwik-style', 'R': 'random'}[args.kind] }), "
                        f"{args.size // args.block_size} x {args.block_size // MiB} MiB .xz blocks, CRC64",
            "preset": args.preset & 0x1F, "block_size": args.block_size, "total_bytes": args.size, "check": "crc64",
            "l2_policy": "inputs (>= 128 MiB per GPU) exceed the 126 MB L2; no explicit flush"}


def run_ours(args):
    import torch
    import torch.distributed as dist
    import __graft_entry__ as ge
    import xz_b200
    from xz_b200 import sharding
    import xzlibs as X

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if rank == 0:
        ge.build()
    if world > 1:
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        dist.barrier()
    if rank != 0:
        ge.build()
    torch.cuda.set_device(local_rank)
    ctx = xz_b200.Context(local_rank)
    opts = xz_b200.lzma_lzma_preset(args.preset)
    bs = args.block_size
    nblocks = args.size // bs
    lo, hi = sharding.shard_blocks(nblocks, world, rank)
    my_blocks = hi - lo
    my_off = lo * bs
    my_n = my_blocks * bs

    # synthetic shard, generated on the host (excluded from all timings), pinned for the e2e leg
    h_in = torch.empty(my_n, dtype=torch.uint8).pin_memory()
    assert X.gen().xzgen_fill(C.c_char(args.kind.encode()), C.c_void_p(h_in.data_ptr()), C.c_size_t(my_n), C.c_uint64(my_off)) == 0
    cap = my_blocks * xz_b200.lzma_block_buffer_bound(bs)
    h_out = torch.empty(cap, dtype=torch.uint8).pin_memory()
    d_in = torch.empty(my_n, dtype=torch.uint8, device="cuda")
    d_in.copy_(h_in)
    d_out = torch.empty(cap, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    dev = torch.device("cuda", local_rank)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def gather_records(recs):
        """The one exchange step of the path: all-gather of the 16-byte Index records (NCCL)."""
        return sharding.gather_records(recs, device=dev)

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    dev_times, e2e_times, wall_dev = [], [], []
    stat_acc = {}
    launches = 0
    sampler = ClockSampler(local_rank)
    all_recs = None
    total_steps = args.warmup + args.steps
    for it in range(total_steps):
        timed = it >= args.warmup
        if it == args.warmup and rank == 0:
            sampler.start()
        # ---- device-resident leg ----
        barrier()
        t0 = time.perf_counter()
        size_d, recs = ctx.encode_blocks_device(d_in.data_ptr(), my_n, opts, 4, bs, d_out.data_ptr(), cap)
        all_recs = gather_records(recs)
        barrier()
        t1 = time.perf_counter()
        s = ctx.stats().as_dict()
        if timed:
            dev_times.append(s["ms_total"] / 1e3)
            wall_dev.append(t1 - t0)
            launches += s["gpu_launches"]
            for k, v in s.items():
                stat_acc[k] = stat_acc.get(k, 0) + v
        if not timed and it + 1 < total_steps:
            continue  # warm-up steps exercise the same kernels through the device leg only
        # ---- e2e leg: pinned host in -> Blocks + records in host memory ----
        barrier()
        t0 = time.perf_counter()
        size_h, recs_h = ctx.encode_blocks_host(h_in.data_ptr(), my_n, opts, 4, bs, h_out.data_ptr(), cap)
        gather_records(recs_h)
        barrier()
        t1 = time.perf_counter()
        if timed:
            e2e_times.append(t1 - t0)
            launches += ctx.stats().gpu_launches
        assert size_h == size_d and recs_h == recs
    clocks = sampler.stop() if rank == 0 else None

    # ---- correctness inside the run: SHA-256 of this rank's blocks == oracle's on a sample ----
    import hashlib
    mine = bytes(h_out[:size_h].numpy().tobytes())
    parity = "unchecked"
    if args.verify_blocks > 0:
        vb = min(args.verify_blocks, my_blocks)
        want = X.oracle_encode(h_in.numpy().ctypes.data_as(C.POINTER(C.c_uint8)), vb * bs, args.preset, bs)
        want_blocks = want[12:12 + sum((u + 3) // 4 * 4 for u, _ in recs[:vb])]
        parity = "bit-exact" if mine[:len(want_blocks)] == want_blocks else "MISMATCH"

    # ---- decode legs on this rank's shard (configs[2]) ----
    idx = xz_b200.index_encode(recs)
    stream = xz_b200.stream_header(4) + mine + idx + xz_b200.stream_footer(4, len(idx))
    h_xz = torch.frombuffer(bytearray(stream), dtype=torch.uint8).pin_memory()
    h_back = torch.empty(my_n, dtype=torch.uint8).pin_memory()
    dec_e2e, dec_dev = [], []
    for it in range(max(1, min(args.steps, 2)) + 1):
        barrier()
        t0 = time.perf_counter()
        r, sz = ctx.stream_decode_into(h_xz.data_ptr(), len(stream), h_back.data_ptr(), my_n)
        barrier()
        t1 = time.perf_counter()
        assert r == 0 and sz == my_n
        if it > 0:
            dec_e2e.append(t1 - t0)
            sd = ctx.stats().as_dict()
            dec_dev.append((sd["ms_total"] - sd["ms_h2d"] - sd["ms_d2h"]) / 1e3)
            launches_dec = sd["gpu_launches"]
    assert bytes(h_back.numpy().tobytes()) == bytes(h_in.numpy().tobytes()), "decode round trip failed"

    t_dev = max_over_ranks(sum(dev_times) / len(dev_times))
    t_e2e = max_over_ranks(sum(e2e_times) / len(e2e_times))
    t_dec_dev = max_over_ranks(sum(dec_dev) / len(dec_dev))
    t_dec_e2e = max_over_ranks(sum(dec_e2e) / len(dec_e2e))
    total_launches = sum_over_ranks(launches)
    xz_total = sum_over_ranks(size_h)
    n_steps = len(dev_times)
    ms_mf = stat_acc.get("ms_mf", 0) / n_steps
    ms_parse = stat_acc.get("ms_parse", 0) / n_steps
    ms_prep = stat_acc.get("ms_mf_prep", 0) / n_steps
    mf_bytes = stat_acc.get("mf_bytes_algorithmic", 0) / n_steps
    peak, peak_src = measured_hbm_peak()
    mf_gbs = mf_bytes / 1e9 / (ms_mf / 1e3) if ms_mf > 0 else 0.0
    parse_bytes = my_n + size_h + stat_acc.get("n_positions", 0) / n_steps * (4 + 8 * 8)  # input + output + match store stream
    parse_gbs = parse_bytes / 1e9 / (ms_parse / 1e3) if ms_parse > 0 else 0.0

    if rank == 0:
        cpu_baseline = None
        if world == 1 and not args.no_cpu_baseline:
            cores = X.ref().ref_cputhreads() if X.have_ref() else 1
            sample_blocks = max(1, min(nblocks, cores or 1))
            n_s = sample_blocks * bs
            sbuf = h_in.numpy().ctypes.data_as(C.POINTER(C.c_uint8))
            t0 = time.perf_counter()
            if X.have_ref():
                X.ref_encode(sbuf, n_s, args.preset, bs, threads=0)
                kind = "reference"
            else:
                X.oracle_encode(sbuf, bs, args.preset, bs)
                n_s, cores, kind = bs, 1, "port"
            dt = time.perf_counter() - t0
            dec_ref = None
            if X.have_ref():  # the reference's threaded decoder on the Stream this run produced (world == 1: whole Stream)
                dbuf = (C.c_uint8 * my_n)(); dsz = C.c_size_t()
                t0 = time.perf_counter()
                rr = X.ref().ref_decode_mt(stream, C.c_size_t(len(stream)), C.c_uint32(0), dbuf, C.c_size_t(my_n), C.byref(dsz))
                dt_dec = time.perf_counter() - t0
                dec_ref = my_n / 1e6 / dt_dec if rr == 0 and dsz.value == my_n else None
                del dbuf
            cpu_baseline = {"value": n_s / 1e6 / dt, "decode_value": dec_ref, "unit": "MB/s", "cores": int(cores), "kind": kind,
                            "sample": f"{n_s // bs} x {bs // MiB} MiB blocks of the same input, one pass, "
                                      + ("lzma_stream_encoder_mt (oracle/_ref) all threads" if kind == "reference" else "oracle port, 1 thread")}
        line = {
            "metric": METRIC, "value": args.size / 1e6 / t_dev, "unit": "MB/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": t_dev * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": workload_config(args),
            "e2e": {"value": args.size / 1e6 / t_e2e, "unit": "MB/s", "h2d_bytes_per_step": args.size, "d2h_bytes_per_step": int(xz_total),
                    "ms_per_step": t_e2e * 1e3},
            "decode": {"value": args.size / 1e6 / t_dec_dev, "e2e_value": args.size / 1e6 / t_dec_e2e, "unit": "MB/s",
                       "h2d_bytes_per_step": int(xz_total), "d2h_bytes_per_step": args.size},
            "gpu_launches": int(total_launches),
            "roofline": {"kernel": "xzb_k_bt (match finder)" if opts.mf & 0x10 else "xzb_k_hc (match finder)", "bound": "hbm",
                         "achieved": mf_gbs, "peak": peak, "unit": "GB/s", "frac": mf_gbs / peak if peak else None,
                         "traffic": ncu_traffic(opts, args.kind, mf_bytes),
                         "peak_source": peak_src, "bytes_per_launch": mf_bytes, "ms_per_launch": ms_mf,
                         "note": "algorithmic bytes = inserted positions x (29 hc | 33 bt) B (SURVEY 8d lower bound); rank 0's shard. "
                                 "xzb_k_bt runs as one launch per 2^20-position segment of the blocks on its own stream beside the "
                                 "parser kernel; bytes/ms are the step's totals over those launches (CUDA events on that stream)"},
            "kernels_ms": {"mf_prep(sort+heads)": ms_prep, "match_finder": ms_mf, "parse+rangecode": ms_parse,
                           "other": stat_acc.get("ms_other", 0) / n_steps,
                           "parse_streamed_GBps": parse_gbs},
            "cpu_baseline": cpu_baseline,
            "clocks": clocks,
            "parity": parity,
            "xz_bytes": int(xz_total) + 12 + 12,
            "index_records": len(all_recs),
        }
        print(json.dumps(line))
    ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size", type=int, default=1024 * MiB)
    ap.add_argument("--block-size", type=int, default=16 * MiB)
    ap.add_argument("--preset", type=lambda s: int(s, 0), default=6)
    ap.add_argument("--kind", default="T")
    ap.add_argument("--verify-blocks", type=int, default=1, help="blocks per rank checked against the oracle inside the run")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
