"""Pins oracle/cr_oracle.c (the CPU restatement used as the parity checker) against
(1) vectors generated from the unmodified reference, (2) the reference's shipped golden
tests/test3 through the legacy normaliser, (3) sha256 tripwires of the reference's own test
programs, and (4) the live reference in oracle/_ref when present.  CPU only."""
import gzip
import hashlib
import json
import os

import numpy as np
import pytest

from conftest import CTEST_CASES, GOLD, pad


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).astype("<i4").tobytes()).hexdigest()


def test_table_matches_reference_hash(oracle, tripwires):
    assert hashlib.sha256(oracle.table.astype("<i4").tobytes()).hexdigest() == tripwires["table_sha256_int32le"]
    t = oracle.table
    assert t[0] == 0 and t[3072] == 65536 and t.min() == -9651 and t.max() == 65536
    assert np.array_equal(t[1:], t[1:][::-1])                                           # symmetric about 3072


def test_fixture_hash(flac_pcm, tripwires):
    assert flac_pcm.shape == (192000, 2)
    assert hashlib.sha256(flac_pcm.astype("<i2").tobytes()).hexdigest() == tripwires["fixture_sha256_s16le"]


@pytest.mark.parametrize("rates", CTEST_CASES)
def test_ctest_workload_lowlevel_and_highlevel(oracle, flac_pcm, tripwires, rates):
    """The reference's own test workload (tests/test-low-level.c / test-high-level.c on tests/test.flac)."""
    i, o, l = rates
    R = oracle.configure(i, o, l)[1]
    out, ret, remaining, pi, pf = oracle.lowlevel(2, i, o, l, pad(flac_pcm, R), flac_pcm.shape[0])
    want = tripwires["ctest_outputs"][f"low:{i}:{o}:{l}"]
    assert out.size * 4 == want["bytes"] and sha(out) == want["sha256"]
    assert ret == 1 and remaining == 0
    hl = oracle.highlevel(2, i, o, l, flac_pcm)
    assert sha(hl) == tripwires["ctest_outputs"][f"high:{i}:{o}:{l}"]["sha256"]


def test_legacy_normaliser_reproduces_shipped_golden(oracle, flac_pcm, tripwires):
    """tests/test3 (== tests/test4) predates the tap-sum normaliser; with the legacy one the oracle
    reproduces it bit-exactly, which pins table, tap geometry, per-tap truncation, stepping and count."""
    raw = gzip.open(os.path.join(GOLD, "ref_test3_s32le.bin.gz"), "rb").read()
    assert hashlib.sha256(raw).hexdigest() == tripwires["ref_test3_sha256"]
    gold = np.frombuffer(raw, dtype="<i4").reshape(-1, 2)
    R = oracle.configure(44100, 8000, 44100)[1]
    out = oracle.lowlevel(2, 44100, 8000, 44100, pad(flac_pcm, R), flac_pcm.shape[0], norm=1, legacy_scale=oracle.ratio(8000, 44100))[0]
    assert np.array_equal(out, gold)


def test_reference_vectors(oracle, ref_vectors):
    meta, data = ref_vectors
    assert len(meta) > 80
    for m in meta:
        k = m["id"]
        if m["kind"] == "lowlevel":
            out, ret, remaining, pi, pf = oracle.lowlevel(m["channels"], m["in"], m["out"], m["lpf"], data[f"ll{k}_in"], m["T"],
                                                          m["pos_int"], m["pos_frac"], m["limit"])
            assert np.array_equal(out, data[f"ll{k}_out"]), m
            assert (ret, remaining, pi, pf) == (m["ret"], m["remaining"], m["end_pos_int"], m["end_pos_frac"]), m
        else:
            out = oracle.highlevel(m["channels"], m["in"], m["out"], m["lpf"], data[f"hl{k}_in"], m["chunk"])
            assert np.array_equal(out, data[f"hl{k}_out"]), m


def test_ratio_and_configure_grid(oracle):
    g = json.load(open(os.path.join(GOLD, "ref_ratio_config.json")))
    for a, b, want in g["ratio"]:
        assert oracle.ratio(a, b) == want, (a, b)
    for a, b, l, want in g["configure"]:
        got = oracle.configure(a, b, l)
        assert (list(got) if got else []) == want, (a, b, l)


def test_closed_form_count_and_state(oracle):
    """N = ceil((T*65536 - P0)/inc); end state from P(N); early stop from P(k)."""
    rng = np.random.default_rng(7)
    rates = [8000, 11025, 22050, 44100, 48000, 96000, 192000, 384000, 1, 2, 3, 7, 1000]
    checked = 0
    for _ in range(600):
        i, o = int(rng.choice(rates)), int(rng.choice(rates))
        cfg = oracle.configure(i, o, o)
        if cfg is None or cfg[3] == 0 or cfg[1] > 200:
            continue
        R, T = cfg[1], int(rng.integers(1, 400))
        p_int, p_frac = int(rng.integers(0, 5)), int(rng.integers(0, 65536))
        limit = int(rng.integers(1, 200)) if rng.random() < 0.5 else 0
        inc = oracle.ratio(i, o)
        padded = np.zeros((T + 2 * R + 8, 1), dtype=np.int16)
        out, ret, remaining, pi, pf = oracle.lowlevel(1, i, o, o, padded, T, p_int, p_frac, limit)
        P0 = p_int * 65536 + p_frac
        N = 0 if P0 >= T * 65536 else -((P0 - T * 65536) // inc)
        assert oracle.count(p_int, p_frac, inc, T) == N
        if limit and limit <= N:
            Pk = P0 + limit * inc
            delta = min(Pk >> 16, T)
            assert (out.shape[0], ret, remaining, pi, pf) == (limit, 0, T - delta, (Pk >> 16) - delta, Pk & 0xFFFF)
        else:
            PN = P0 + N * inc
            assert (out.shape[0], ret, remaining, pi, pf) == (N, 1, 0, (PN >> 16) - T, PN & 0xFFFF)
        checked += 1
    assert checked > 300


def test_live_reference_random(oracle, reference):
    """Bit-exact agreement with the unmodified reference compiled from /root/reference (oracle/_ref)."""
    assert np.array_equal(oracle.table, reference.table)
    rng = np.random.default_rng(99)
    rates = [8000, 16000, 22050, 44100, 48000, 96000, 192000, 384000, 5, 13]
    n = 0
    for _ in range(120):
        ch = int(rng.integers(1, 17))
        i, o = int(rng.choice(rates)), int(rng.choice(rates))
        l = int(rng.choice([i, o, 44100, 1000]))
        cfg = reference.configure(i, o, l)
        assert (oracle.configure(i, o, l) or None) == (cfg or None)
        if cfg is None or cfg[3] == 0 or cfg[1] > 400:
            continue
        T = int(rng.integers(1, 800))
        T = max(1, min(T, 3000 * reference.ratio(i, o) // 65536))
        data = rng.integers(-32768, 32768, size=(T, ch), dtype=np.int16)
        p = pad(data, cfg[1])
        pi, pf = int(rng.integers(0, 3)), int(rng.integers(0, 65536))
        lim = int(rng.integers(0, 50))
        a = oracle.lowlevel(ch, i, o, l, p, T, pi, pf, lim)
        b = reference.lowlevel(ch, i, o, l, p, T, pi, pf, lim)
        assert np.array_equal(a[0], b[0]) and a[1:] == b[1:], (ch, i, o, l, T, pi, pf, lim)
        n += 1
    assert n > 60


def test_noise_generator_is_counter_based(oracle):
    a = oracle.noise(5, 3, 1000, 64, 2)
    b = oracle.noise(5, 3, 1010, 54, 2)
    assert np.array_equal(a[10:], b)
    assert not np.array_equal(oracle.noise(5, 4, 1000, 64, 2), a)
    big = oracle.noise(1, 0, 0, 1 << 16, 1).astype(np.int64)
    assert abs(big.mean()) < 400 and 17000 < big.std() < 20500   # roughly uniform s16
