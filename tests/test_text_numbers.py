"""CPU: the number conversion of the device-side text parser (gpu-pruner_b200/csrc/gpr_text.cuh: Clinger's fast
path + Eisel-Lemire with the generated 128-bit table) against Python's correctly rounded float().

The parser's contract is "equal to strtod + (float), or decline" (a declined number marks the span hard and the
CPU's strtod decides), so every ACCEPTED conversion must match bit for bit; the rate of declines on realistic input
(17-digit DCGM_FI_PROF_GR_ENGINE_ACTIVE ratios: shortest-round-trip doubles) must be negligible."""
import os
import random
import struct
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def driver(tmp_path_factory):
    out = tmp_path_factory.mktemp("num") / "number_check"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fsanitize=undefined", "-fno-sanitize-recover=all",
                           os.path.join(ROOT, "tests", "cpp", "number_check.cpp"), "-o", str(out)])
    return str(out)


def _run(driver, lines):
    r = subprocess.run([driver], input="\n".join(lines) + "\n", capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    out = r.stdout.splitlines()
    assert len(out) == len(lines)
    return out


def _bits64(x):
    return struct.unpack("<Q", struct.pack("<d", x))[0]


def test_pow10_table_is_what_the_generator_writes():
    """the committed header is the generator's output (exact big-integer arithmetic)"""
    import importlib.util
    spec = importlib.util.spec_from_file_location("gen", os.path.join(ROOT, "tools", "gen_pow10_table.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    text = open(os.path.join(ROOT, "gpu-pruner_b200", "csrc", "gpr_pow10_table.h")).read()
    for q in (-348, -27, -1, 0, 1, 22, 27, 55, 347):
        m = gen.mantissa128(q)
        assert f"{{0x{m >> 64:016X}ull, 0x{m & (2**64 - 1):016X}ull}}, /* 1e{q} */" in text
    assert gen.mantissa128(0) == 1 << 127 and gen.mantissa128(-1) >> 64 == 0xCCCCCCCCCCCCCCCC


def test_eisel_lemire_matches_correct_rounding(driver):
    rng = random.Random(20260921)
    cases = []
    realistic = set()
    for _ in range(120_000):
        nd = rng.randrange(1, 20)
        man = rng.randrange(10 ** (nd - 1), 10 ** nd)
        if man >= 1 << 64:
            continue
        cases.append((man, rng.randrange(-340, 300)))
    # shortest-round-trip representations of random doubles (what Prometheus prints), as mantissa / exponent
    for _ in range(80_000):
        x = rng.random() if rng.random() < 0.7 else rng.uniform(0, 1000)
        s = repr(x)
        if "e" in s or "." not in s:
            continue
        ip, fp = s.split(".")
        cases.append((int(ip + fp), -len(fp)))
        realistic.add(cases[-1])
    # numbers at and around rounding boundaries
    for k in (53, 54, 60, 63):
        for d in (-1, 0, 1):
            cases.append(((1 << k) + d, 0))
            cases.append(((1 << k) + d, -5))
    cases += [(9007199254740993, 0), (9007199254740993, -3), (1, -324), (1, 308), (17976931348623157, 292),
              (22250738585072014, -324), (4, -324), (12345678901234567890, 0), (1, 0), (5, -1)]
    out = _run(driver, [f"E {m} {e}" for m, e in cases])
    declined = 0
    for (m, e), line in zip(cases, out):
        ok, bits = line.split()
        if ok == "0":
            declined += 1
            assert (m, e) not in realistic, (m, e)     # what Prometheus prints is always decided on the device
            continue   # subnormal / overflowing results (random exponents reach both) or a genuine half-way case
        want = float(f"{m}e{e}")
        assert int(bits, 16) == _bits64(want), (m, e, bits, hex(_bits64(want)))
    assert declined < len(cases) // 20


def test_parse_value_matches_strtod_then_float(driver):
    rng = random.Random(7)
    texts = ["0", "-0", "100", "37", "0.5", "0.25", "12.25", "0.30000000000000004", "123456789012345678",
             "0.1", "0.07", "99.99999999999999", "1234567.1234567", "16777216", "16777217", "4294967296.5",
             "0.000000000000000000000000000000000000000000001", "0.0000000000000000000000000000000000000000000001",
             "340282350000000000000000000000000000000", "0.1234567890123456789", "1.7976931348623157",
             "000123", "0.000", "5.0000000000000000000", "5e-07", "1.2345e+21", "1e2", "1E1", "1e-50", "1e23",
             "9.999999e-07", "1e+21", "0e0", "3.4028235e+38", "1e39", "4.9e-324", "1.5e-46"]
    for _ in range(60_000):
        x = rng.random() if rng.random() < 0.6 else rng.uniform(0, 700)
        texts.append(repr(x))          # includes Go-style exponent forms for the tiny ones ("5e-07")
        if rng.random() < 0.05:
            texts.append(repr(x * 10.0 ** rng.randrange(-30, 30)))
        if rng.random() < 0.1:
            texts.append(str(rng.randrange(0, 101)))
        if rng.random() < 0.05:
            texts.append("-" + texts[-1])
    out = _run(driver, ["V " + t for t in texts])
    declined = []
    for t, line in zip(texts, out):
        q, bits, tiny = line.split()
        if q == "0":
            declined.append(t)
            continue
        want = np.float32(float(t))
        if float(t) != 0.0 and want == 0.0:      # below the f32 denormal range: kept non-zero (to_f32 in ingest.cpp)
            assert int(bits, 16) in (0x00000001, 0x80000001) and tiny == "1", t
            continue
        assert int(bits, 16) == struct.unpack("<I", struct.pack("<f", want))[0], (t, bits)
    # declined: more than 19 significant digits, results in the subnormal range, and the handful of exactly-half-way
    # products Eisel-Lemire leaves to a slower method — never a shortest-round-trip decimal below 1e21
    realistic = {t for t in texts if "e" not in t and 0 <= float(t) <= 1000}
    for t in declined:
        digits = len(t.lstrip("-+").split("e")[0].replace(".", "").lstrip("0"))
        assert digits > 19 or t not in realistic, t
    assert len(declined) < len(texts) // 100, declined[:10]
    assert "0.30000000000000004" not in declined and "123456789012345678" not in declined


def test_parse_value_declines_what_it_cannot_decide(driver):
    out = _run(driver, ["V .5", "V 5.", "V 12345678901234567890", "V 1.2.3", "V abc", "V +", "V 0x10",
                        "V 1234567890123456789012", "V 1e", "V 1e+", "V 1e1234", "V e5", "V 1.e5"])
    assert all(line.split()[0] == "0" for line in out), out


def test_parse_timestamp_is_exact_in_milliseconds(driver):
    texts = ["1700000000", "1700000000.4", "1700000000.5", "1700000000.499", "1700000000.500", "1700000000.123",
             "0", "0.001", "9999999999999", "4000000000000", "3999999999999.5", "1700000000.05"]
    out = _run(driver, ["T " + t for t in texts])
    for t, line in zip(texts, out):
        q, ts = line.split()
        assert int(q) == len(t) + 1, t
        x = float(t)
        if not (x < 4e12):
            assert int(ts) < -(1 << 60)          # "no sane epoch time": far outside every window
        else:
            assert int(ts) == round(x * 1000), t  # what the CPU path computes: llround(strtod(t) * 1000)
    # finer than a millisecond, signs, exponents: declined (the span goes to the CPU parser)
    out = _run(driver, ["T 1700000000.1234", "T 17000000001234", "T -5", "T +5", "T 1e9", "T ", "T 5."])
    assert all(line.split()[0] == "0" for line in out), out
