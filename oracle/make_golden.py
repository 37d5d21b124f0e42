"""Generates tests/golden/ from the UNMODIFIED reference (oracle/_ref, built in place from
/root/reference).  Run in the build container only:  python oracle/make_golden.py
TEST INFRASTRUCTURE ONLY.

Outputs (all small, committed):
  test_flac_s16le.bin.gz   tests/test.flac decoded by the reference's vendored dr_flac (192000 x 2 s16le)
  ref_test3_s32le.bin.gz   the reference's own golden tests/test3 (== tests/test4), s32le, legacy normaliser
  ref_vectors.npz          outputs of the reference on seeded synthetic cases (low-level incl. start state
                           and early stop, high-level incl. odd chunk sizes), with the inputs
  tripwires.json           sha256 of the table, the decoded fixture and the reference outputs for the
                           four CTest parameter sets (tests/CMakeLists.txt:25-47)
"""
import gzip
import hashlib
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.cro import Reference, build_ref  # noqa: E402

REFERENCE = "/root/reference"
GOLD = os.path.join(ROOT, "tests", "golden")


def sha(a: bytes) -> str:
    return hashlib.sha256(a).hexdigest()


def gz_write(path, data: bytes):
    with open(path, "wb") as f:
        with gzip.GzipFile(fileobj=f, mode="wb", mtime=0, compresslevel=9) as g:
            g.write(data)


def main():
    os.makedirs(GOLD, exist_ok=True)
    build_ref(REFERENCE)
    ref = Reference()
    trip = {"abi": ref.abi(), "table_sha256_int32le": sha(ref.table.astype("<i4").tobytes())}

    # 1. fixture
    tmp = "/tmp/_flac.s16"
    subprocess.check_call([os.path.join(ROOT, "oracle/_ref/flac-dump"), os.path.join(REFERENCE, "tests/test.flac"), tmp])
    pcm_bytes = open(tmp, "rb").read()
    gz_write(os.path.join(GOLD, "test_flac_s16le.bin.gz"), pcm_bytes)
    trip["fixture_sha256_s16le"] = sha(pcm_bytes)
    pcm = np.frombuffer(pcm_bytes, dtype="<i2").reshape(-1, 2)

    # 2. the reference's shipped golden
    t3 = open(os.path.join(REFERENCE, "tests/test3"), "rb").read()
    assert t3 == open(os.path.join(REFERENCE, "tests/test4"), "rb").read()
    gz_write(os.path.join(GOLD, "ref_test3_s32le.bin.gz"), t3)
    trip["ref_test3_sha256"] = sha(t3)

    # 3. the reference's own test programs on the four CTest parameter sets
    ctest = {}
    for name, prog in (("low", "ref-test-low-level"), ("high", "ref-test-high-level")):
        for (i, o, l) in [(8000, 44100, 44100), (8000, 44100, 8000), (44100, 8000, 44100), (44100, 8000, 8000)]:
            out = f"/tmp/_ref_{name}_{i}_{o}_{l}"
            subprocess.check_call([os.path.join(ROOT, "oracle/_ref", prog), os.path.join(REFERENCE, "tests/test.flac"), out, str(i), str(o), str(l)],
                                  stderr=subprocess.DEVNULL)
            b = open(out, "rb").read()
            ctest[f"{name}:{i}:{o}:{l}"] = {"bytes": len(b), "sha256": sha(b)}
    trip["ctest_outputs"] = ctest

    # 4. seeded synthetic vectors through the reference
    rng = np.random.default_rng(20261017)
    rates = [8000, 11025, 16000, 22050, 32000, 44100, 48000, 88200, 96000, 176400, 192000, 384000, 1, 2, 3, 7, 1000, 65537]
    vec = {}
    meta = []
    k = 0
    for case in range(96):
        ch = int(rng.choice([1, 2, 2, 3, 4, 6, 8, 16]))
        i, o = int(rng.choice(rates)), int(rng.choice(rates))
        l = int(rng.choice([i, o, max(1, min(i, o) // 2), 44100]))
        cfg = ref.configure(i, o, l)
        if cfg is None or cfg[3] == 0 or cfg[1] > 300:
            continue
        R = cfg[1]
        T = int(rng.integers(1, 600))
        T = max(1, min(T, 2500 * ref.ratio(i, o) // 65536))   # keep each output under ~2500 frames
        kind = case % 4
        data = rng.integers(-32768, 32768, size=(T, ch), dtype=np.int16)
        if kind == 1:   # full-scale square-ish, worst case for the accumulators
            data = np.where(rng.random((T, ch)) < 0.5, -32768, 32767).astype(np.int16)
        if kind == 2:   # low-amplitude signal: exercises truncation toward zero around 0
            data = rng.integers(-3, 4, size=(T, ch), dtype=np.int16)
        padded = np.zeros((T + 2 * R, ch), dtype=np.int16)
        padded[R:R + T] = data
        pos_int = int(rng.integers(0, 4)) if case % 3 == 0 else 0
        pos_frac = int(rng.integers(0, 65536)) if case % 3 == 0 else 0
        limit = int(rng.integers(1, 400)) if case % 5 == 0 else 0
        out, ret, remaining, pi, pf = ref.lowlevel(ch, i, o, l, padded, T, pos_int, pos_frac, limit)
        vec[f"ll{k}_in"] = padded
        vec[f"ll{k}_out"] = out
        meta.append({"kind": "lowlevel", "id": k, "channels": ch, "in": i, "out": o, "lpf": l, "T": T, "pos_int": pos_int, "pos_frac": pos_frac,
                     "limit": limit, "ret": ret, "remaining": remaining, "end_pos_int": pi, "end_pos_frac": pf, "frames": int(out.shape[0])})
        k += 1
    for case in range(24):
        ch = int(rng.choice([1, 2, 3, 8]))
        i, o = int(rng.choice(rates[:12])), int(rng.choice(rates[:12]))
        l = int(rng.choice([i, o, 44100]))
        cfg = ref.configure(i, o, l)
        if cfg is None or cfg[3] == 0 or cfg[1] * 2 * ch >= 4096:
            continue
        T = int(rng.integers(0, 6000)) if case else 0
        T = min(T, 6000 * ref.ratio(i, o) // 65536)
        data = rng.integers(-32768, 32768, size=(T, ch), dtype=np.int16)
        chunk = int(rng.choice([0, 0, 1, 7, 333]))
        out = ref.highlevel(ch, i, o, l, data, chunk)
        vec[f"hl{k}_in"] = data
        vec[f"hl{k}_out"] = out
        meta.append({"kind": "highlevel", "id": k, "channels": ch, "in": i, "out": o, "lpf": l, "T": T, "chunk": chunk, "frames": int(out.shape[0])})
        k += 1
    np.savez_compressed(os.path.join(GOLD, "ref_vectors.npz"), **vec)
    json.dump(meta, open(os.path.join(GOLD, "ref_vectors.json"), "w"), indent=0)

    # 5. ratio / configuration table over a rate grid (host-side integer code must match exactly)
    grid = [1, 2, 3, 5, 7, 1000, 8000, 11025, 22050, 44100, 48000, 96000, 192000, 384000, 65535, 65536, 65537, 3000000, 0xFFFFFFFF, 0]
    rat = [[a, b, ref.ratio(a, b)] for a in grid for b in grid]
    cfgs = [[a, b, l, list(ref.configure(a, b, l) or [])] for a in grid[:-2] for b in grid[:-2] for l in (a, b, 44100)]
    json.dump({"ratio": rat, "configure": cfgs}, open(os.path.join(GOLD, "ref_ratio_config.json"), "w"))

    json.dump(trip, open(os.path.join(GOLD, "tripwires.json"), "w"), indent=1)
    print(json.dumps({k: v for k, v in trip.items() if k != "ctest_outputs"}, indent=1))
    print(len(meta), "vectors")


if __name__ == "__main__":
    main()
