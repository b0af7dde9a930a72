#!/usr/bin/env python
"""Regenerates tests/golden/ from the UNMODIFIED reference (run in the build container only).

1. Copies the reference's own decoder fixtures for the LZMA2 path
   (/root/reference/tests/files/{good,bad,unsupported}-*.xz; semantics in tests/files/README:44-176)
   into tests/golden/ref_files/ and records the reference's verdict (lzma_ret of
   lzma_stream_decoder + lzma_code(FINISH), SHA-256 of the output) in decode_verdicts.json.
2. Runs the reference encoder (lzma_stream_encoder_mt via oracle/_ref) on seeded synthetic
   inputs and records SHA-256 + size of the .xz it produces in encode_golden.json.
   The inputs are regenerated from xz_b200/csrc/xzgen.c, so only hashes are stored.
3. Stores the reference's known-answer values for CRC32/CRC64 (tests/test_check.c:74,112) and the
   MicroLZMA encoder KAT (tests/test_microlzma.c:20-32) in kat.json.
"""
import glob, hashlib, json, os, shutil, sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import xzlibs as X

REF_FILES = "/root/reference/tests/files"


def main():
    assert X.have_ref(), "build oracle/_ref first (make -f oracle/Makefile.ref all)"
    dst = os.path.join(HERE, "ref_files")
    os.makedirs(dst, exist_ok=True)
    verdicts = {}
    for f in sorted(glob.glob(os.path.join(REF_FILES, "*.xz"))):
        name = os.path.basename(f)
        data = open(f, "rb").read()
        if len(data) > 64 * 1024:
            continue  # good-1-delta-lzma2.tiff.xz: big and a Delta chain (out of scope)
        shutil.copyfile(f, os.path.join(dst, name))
        r, out = X.ref_decode(data, 1 << 22)
        verdicts[name] = {"ret": r, "out_size": len(out), "out_sha256": hashlib.sha256(out).hexdigest() if r == 0 else None}
        # same with LZMA_CONCATENATED (what `xz -d` uses): Stream Padding / multi-Stream verdicts
        import ctypes as C
        o2 = (C.c_uint8 * (1 << 22))(); s2 = C.c_size_t()
        r2 = X.ref().ref_decode_flags(data, C.c_size_t(len(data)), C.c_uint32(0x08), o2, C.c_size_t(1 << 22), C.byref(s2))
        verdicts[name]["ret_concat"] = r2
        verdicts[name]["out_concat_sha256"] = hashlib.sha256(bytes(o2[: s2.value])).hexdigest() if r2 == 0 else None
    json.dump(verdicts, open(os.path.join(HERE, "decode_verdicts.json"), "w"), indent=1, sort_keys=True)

    enc = []
    MiB = 1 << 20
    cases = []
    for kind in "TER":
        for preset in (0, 1, 3, 4, 6, 9 | X.XZ_PRESET_EXTREME):
            for n in (0, 1, 2, 3, 4, 5, 273, 4096, 65535, 65536, 65537, 300000):
                cases.append((kind, preset, n, 256 * 1024))
        for preset in (1, 3, 6):
            cases.append((kind, preset, 2 * MiB - 273, 4 * MiB))
            cases.append((kind, preset, 2 * MiB + 273, 1 * MiB))
            cases.append((kind, preset, 3 * MiB + 1, MiB))
    # BASELINE.json configs[0] exactly, plus one full 16 MiB block per GPU config's preset
    cases += [("T", 1, 16 * MiB, 16 * MiB), ("T", 6, 16 * MiB, 16 * MiB), ("T", 6, 16 * MiB + 1, 16 * MiB),
              ("R", 3, 16 * MiB, 16 * MiB), ("E", 6, 16 * MiB, 16 * MiB), ("E", 9 | X.XZ_PRESET_EXTREME, 4 * MiB, 16 * MiB)]
    for kind, preset, n, bs in cases:
        buf = X.gendata(kind, n)
        out = X.ref_encode(buf, n, preset, bs)
        enc.append({"kind": kind, "preset": preset, "size": n, "block_size": bs, "check": 4,
                    "xz_size": len(out), "xz_sha256": hashlib.sha256(out).hexdigest()})
        print(kind, hex(preset), n, bs, len(out), flush=True)
    json.dump(enc, open(os.path.join(HERE, "encode_golden.json"), "w"), indent=1)

    kat = {"crc32_123456789": 0xCBF43926, "crc64_123456789": 0x995DC9BBDF1939FA,
           "ref_crc32_123456789": X.ref().ref_crc32(b"123456789", 9, 0),
           "ref_crc64_123456789": X.ref().ref_crc64(b"123456789", 9, 0),
           "microlzma_hello_world_crc32": 0x3CDE40A8}
    json.dump(kat, open(os.path.join(HERE, "kat.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
