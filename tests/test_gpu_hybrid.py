"""The UNMODIFIED reference `xz` (src/xz) running on the GPU path: oracle/_ref/xz_gpu is the reference's xz objects linked
with libxzb200.so ahead of the reference liblzma (oracle/Makefile.ref), so coder.c:836 lzma_stream_encoder_mt,
coder.c:956 lzma_stream_decoder_mt and coder.c:1226 lzma_code are this library's.  Compared with oracle/_ref/xz (reference
only) on the same command lines: identical .xz bytes, identical decoded bytes, identical verdicts on the corpus."""
import os
import subprocess

import pytest

import xzlibs as X

ROOT = X.ROOT
XZ = os.path.join(ROOT, "oracle", "_ref", "xz")
XZ_GPU = os.path.join(ROOT, "oracle", "_ref", "xz_gpu")
FILES = os.path.join(ROOT, "tests", "golden", "ref_files")
pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not (os.path.exists(XZ) and os.path.exists(XZ_GPU)), reason="oracle/_ref/xz[_gpu] not built")]
MiB = 1 << 20
# filter chains outside the GPU path: none (Delta and all BCJ filters run as xzb_k_filter in front of / behind LZMA2)
OTHER_FILTERS = set()


def run(args, **kw):
    return subprocess.run(args, stdout=subprocess.PIPE, stderr=subprocess.PIPE, **kw)


@pytest.mark.parametrize("kind,preset,n,bs", [("T", "-6", 24 * MiB + 12345, "4MiB"), ("E", "-4", 6 * MiB, "1MiB"), ("R", "-1", 3 * MiB + 1, "1MiB")])
def test_xz_threaded_encode_is_byte_identical_and_decodes(tmp_path, kind, preset, n, bs):
    f = tmp_path / "in.bin"
    buf = X.gendata(kind, n)
    f.write_bytes(bytes(buf[:n]))
    args = [preset, "-T4", f"--block-size={bs}", "-c", str(f)]
    want = run([XZ] + args)
    got = run([XZ_GPU] + args)
    assert want.returncode == 0 and got.returncode == 0, got.stderr
    assert got.stdout == want.stdout
    g = tmp_path / "gpu.xz"
    g.write_bytes(got.stdout)
    back = run([XZ_GPU, "-dc", "-T4", str(g)])
    assert back.returncode == 0 and back.stdout == f.read_bytes(), back.stderr
    # a single-threaded reference Stream (Blocks without sizes) through the hybrid's decoder
    st = tmp_path / "st.xz"
    st.write_bytes(run([XZ, preset, "-T1", "-c", str(f)]).stdout)
    back = run([XZ_GPU, "-dc", str(st)])
    assert back.returncode == 0 and back.stdout == f.read_bytes(), back.stderr
    # xz -l reads the Index through the reference's lzma_file_info_decoder driven by this library's lzma_code
    assert run([XZ_GPU, "-l", str(g)]).stdout == run([XZ, "-l", str(g)]).stdout


@pytest.mark.parametrize("fopts", [["--x86", "--lzma2=preset=6"], ["--delta=dist=4", "--lzma2=preset=4"], ["--arm64", "--delta=dist=2", "--lzma2=preset=1"]])
def test_xz_filter_chains(tmp_path, fopts):
    """xz --x86 / --delta / --arm64 in front of --lzma2: coder.c hands the chain to lzma_stream_encoder_mt (filter chain table)."""
    from test_gpu_filters import mixed_input
    n = 3 * MiB + 77
    data = mixed_input(n, 3)
    f = tmp_path / "in.bin"
    f.write_bytes(data)
    args = fopts + ["-T4", "--block-size=1MiB", "-c", str(f)]
    want = run([XZ] + args)
    got = run([XZ_GPU] + args)
    assert want.returncode == 0 and got.returncode == 0, got.stderr
    assert got.stdout == want.stdout
    back = run([XZ_GPU, "-dc", "-T4"], input=got.stdout)
    assert back.returncode == 0 and back.stdout == data, back.stderr


def test_xz_encoder_fans_out_over_several_contexts(tmp_path):
    """lzma_stream_encoder_mt deals the Blocks of a wave over the GPUs of the process, one host thread and context each
    (stream_encoder_mt.c:362-595 / :716-888 behind the one call).  XZB_DEVICES names them; the same device twice
    exercises the split / threads / packing on a one-GPU box, every visible GPU is the default."""
    n = 11 * MiB + 321
    data = bytes(X.gendata("T", n)[:n])
    args = ["-6", "-T4", "--block-size=1MiB"]
    want = run([XZ] + args, input=data)
    for devs in ("0,0", "0,0,0", None):
        env = dict(os.environ)
        env.pop("XZB_DEVICE", None)
        if devs:
            env["XZB_DEVICES"] = devs
        got = run([XZ_GPU] + args, input=data, env=env)
        assert got.returncode == 0 and got.stdout == want.stdout, (devs, got.stderr)


def test_xz_default_threads_and_block_size(tmp_path):
    """xz -T0 with the preset's own block size (3 x dict), stdin to stdout, as in `xz -6 -T0 < in > out`."""
    n = 5 * MiB
    data = bytes(X.gendata("T", n)[:n])
    want = run([XZ, "-6", "-T0"], input=data)
    got = run([XZ_GPU, "-6", "-T0"], input=data)
    assert got.returncode == 0 and got.stdout == want.stdout, got.stderr
    back = run([XZ_GPU, "-dc", "-T0"], input=got.stdout)
    assert back.returncode == 0 and back.stdout == data


def test_high_ratio_unsized_stream(tmp_path):
    """100 MB of zeros from the single-threaded encoder: ratio ~ 7000 : 1, no sizes in the Block Header."""
    n = 100 * 1000 * 1000
    z = tmp_path / "z.xz"
    z.write_bytes(run([XZ, "-6", "-T1"], input=bytes(n)).stdout)
    assert z.stat().st_size < 20000
    back = run([XZ_GPU, "-dc", str(z)])
    assert back.returncode == 0 and len(back.stdout) == n and back.stdout.count(0) == n, back.stderr


def test_corpus_verdicts_match_reference_xz():
    """tests/test_files.sh of the reference, restated: `xz -t` on every corpus file gives the reference's exit code."""
    diffs = []
    for name in sorted(os.listdir(FILES)):
        if not name.endswith(".xz"):
            continue
        p = os.path.join(FILES, name)
        a = run([XZ, "-t", "-qq", p]).returncode
        r = run([XZ_GPU, "-t", p])
        if name in OTHER_FILTERS:
            assert a == 0 and r.returncode == 1 and b"Unsupported" in r.stderr, (name, r.stderr)
            continue
        if a != r.returncode:
            diffs.append((name, a, r.returncode, r.stderr[-200:]))
    assert not diffs, diffs
