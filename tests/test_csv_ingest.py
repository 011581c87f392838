"""Scan ingest: the reference reads its dataset with fscanf("%f,") (readDatasetLineByLine,
Subsystem_1/main.c:22-30).  The oracle restates that loop and is pinned to the reference's own function; the
GPU parser (b200slam_csv_ingest) must return the same floats bit for bit -- on the replay-sized dataset and on
tokens chosen to hit every branch: values a double division cannot settle, more digits than a double holds,
exponents, signs, odd separators."""
import os

import numpy as np
import pytest

from conftest import bits


def _write(path, text):
    with open(path, "wb") as f:
        f.write(text)
    return text


def _adversarial_text(rng):
    toks = ["0", "-0.0", "+1.5", "000012.3400", ".5", "5.", "0.1", "0.10000000149011612", "16777216", "16777217",
            "16777219.0", "33554433", "0.30000001192092896", "1e-3", "1E+2", "-2.5e-1", "123456789012345678",
            "0.000000000000000000001", "1.17549435e-38", "3.4028234e38", "1e-45", "7.006492321624085e-46",
            "2.00000011920928955078125", "1.00000005960464477539062", "1.000000059604644775390625",
            "1.0000000596046447753906250000001", "8388608.5", "8388609.5", "4194304.25", "0.3333333432674408",
            "24.0000", "0.0230", "0.0229999", "inf", "-inf", "1.0000001", "9999999", "99999999", "0.0000001"]
    # floats printed with 9 significant digits round-trip; 17 digits (a double's worth) exceed the device's 15
    vals = rng.uniform(0.0, 30.0, 4000).astype(np.float32)
    toks += ["%.9g" % v for v in vals[:2000]] + ["%.17g" % float(v) for v in vals[2000:3000]] + ["%.4f" % v for v in vals[3000:]]
    # midpoints between adjacent floats, exactly and +- a hair: the double-rounding guard's territory
    for v in vals[:300]:
        a = np.float32(v)
        b = np.nextafter(a, np.float32(np.inf), dtype=np.float32)
        mid = (float(a) + float(b)) / 2.0
        toks += ["%.25g" % mid, "%.17g" % np.nextafter(mid, np.inf), "%.17g" % np.nextafter(mid, -np.inf)]
    # separators fscanf("%f,") itself gets through: an optional ',' right behind the value, then any white space
    seps = [",", ",\n", ", ", "\t", ",\r\n", "  ", ",\t ", "\n\n"]
    out = []
    for i, t in enumerate(toks):
        out.append(t)
        out.append(seps[i % len(seps)])
    return "".join(out).encode(), len(toks)


def test_oracle_csv_reader_is_the_references(oracle, synth, tmp_path):
    """Pin: orc_read_csv == the reference's readDatasetLineByLine, compiled unmodified, on the same file."""
    from oracle import pyoracle
    if not pyoracle.reference_available():
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    ranges = synth.lidar_dataset(7)
    path = str(tmp_path / "d.csv")
    synth.write_lidar_csv(path, ranges)
    got = oracle.read_csv(path, 7 * 1079)
    want = pyoracle.Reference("main").read_dataset_rows(path, 7)
    assert np.array_equal(bits(got), bits(want.ravel()))
    # and on awkward but valid numeric text (exponents, signs, odd separators), 1079 values = one reference row
    text, n = _adversarial_text(np.random.default_rng(5))
    text = text.replace(b"inf", b"1e3")                      # keep it finite for the == below
    path2 = str(tmp_path / "a.csv")
    _write(path2, text)
    got2 = oracle.read_csv(path2, n)
    assert len(got2) == n
    k = (n // 1079) * 1079
    want2 = pyoracle.Reference("main").read_dataset_rows(path2, n // 1079)
    assert np.array_equal(bits(got2[:k]), bits(want2.ravel()))


def _correctly_rounded_f32(tok: str) -> np.float32:
    """Nearest binary32 of a decimal string, ties to even, by exact rational arithmetic (numpy's float32(str)
    rounds twice, through a double)."""
    from fractions import Fraction
    t = tok.lower()
    if "inf" in t or "nan" in t:
        return np.float32(t)
    v = Fraction(t)
    g = np.float32(float(v))                              # within one ulp of the answer
    best = None
    with np.errstate(over="ignore"):
        cands = [np.nextafter(g, np.float32(-np.inf), dtype=np.float32), g, np.nextafter(g, np.float32(np.inf), dtype=np.float32)]
    for c in cands:
        if not np.isfinite(c):
            continue
        d = abs(Fraction(float(c)) - v)
        even = (int(np.float32(c).view(np.uint32)) & 1) == 0
        if best is None or d < best[0] or (d == best[0] and even):
            best = (d, c)
    r = best[1]
    if r == 0 and tok.strip().startswith("-"):
        r = np.float32(-0.0)
    return np.float32(r)


def test_oracle_csv_reader_is_correctly_rounded(oracle, tmp_path):
    """fscanf's %f is a correctly rounded decimal -> float conversion (what the GPU parser has to reproduce)."""
    text, n = _adversarial_text(np.random.default_rng(6))
    text = text.replace(b"3.4028234e38", b"3.4028234e37")          # stay clear of the overflow threshold
    path = str(tmp_path / "a.csv")
    _write(path, text)
    got = oracle.read_csv(path, n)
    toks = [t for t in text.replace(b",", b" ").split()]
    want = np.array([_correctly_rounded_f32(t.decode()) for t in toks], np.float32)
    assert len(got) == n == len(want)
    assert np.array_equal(bits(got), bits(want)), [toks[i] for i in np.flatnonzero(bits(got) != bits(want))[:5]]


@pytest.mark.gpu
def test_gpu_csv_ingest_replay_sized_dataset(ctx, oracle, synth, tmp_path):
    """The 3480 x 1079 dataset of the replay (configs[0]): every value bit-identical to fscanf's, and the
    resident values feed readAScan exactly like an upload of the same row."""
    ranges = synth.lidar_dataset(3480)
    path = str(tmp_path / "lidar.csv")
    synth.write_lidar_csv(path, ranges)
    text = open(path, "rb").read()
    want = oracle.read_csv(path, 3480 * 1079)
    got = ctx.csv_ingest(text)
    assert len(got) == len(want) == 3480 * 1079
    assert np.array_equal(bits(got), bits(want))
    # max_values cuts the result like the reference's fixed row count does
    assert np.array_equal(bits(ctx.csv_ingest(text, max_values=5 * 1079)), bits(want[:5 * 1079]))
    got = ctx.csv_ingest(text)
    ctx.lidar_set(oracle.lidar_angles(), 0.023)
    for r in (0, 17, 3479):
        ctx.scan_read_resident_async(r * 1079)
        xa, ya = ctx.scan_download(transformed=False)
        n = ctx.scan_read(want[r * 1079:(r + 1) * 1079])
        xb, yb = ctx.scan_download(transformed=False)
        assert len(xa) == n and np.array_equal(bits(xa), bits(xb)) and np.array_equal(bits(ya), bits(yb))


@pytest.mark.gpu
def test_gpu_csv_ingest_adversarial_tokens(ctx, oracle, tmp_path):
    text, n = _adversarial_text(np.random.default_rng(7))
    path = str(tmp_path / "a.csv")
    _write(path, text)
    want = oracle.read_csv(path, n)
    got = ctx.csv_ingest(text)
    assert len(got) == len(want) == n
    assert np.array_equal(bits(got), bits(want)), np.flatnonzero(bits(got) != bits(want))[:10]
    # empty text, separators only, one value without a trailing separator
    assert len(ctx.csv_ingest(b"")) == 0
    assert len(ctx.csv_ingest(b" ,\n, ")) == 0
    assert np.array_equal(bits(ctx.csv_ingest(b"3.25")), bits(np.array([3.25], np.float32)))
    # a token that is not a number is an error, not a silent zero
    with pytest.raises(Exception):
        ctx.csv_ingest(b"1.0,abc,2.0")
