"""The collecting side of the fused exchange (exchange_bitmaps_ll in gpu-pruner_b200/csrc/gpr_kernels.cuh), run on the
CPU: the function's source text is cut out of the kernel header verbatim and compiled under a host shim
(tests/cpp/exchange_emul.cpp) that maps a CTA onto real threads and the peers onto writers that fill the tagged
slots late and in random order.  Checks indexing (rank-major outputs, stride != span, odd sizes, fewer threads than
slots), rejection of stale tags, re-polling until everything arrived, and that with the late ordering nothing
reaches the caller's buffers before the previous fold is done."""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HDR = os.path.join(ROOT, "gpu-pruner_b200", "csrc", "gpr_kernels.cuh")


def _extract():
    src = open(HDR).read()
    peers = re.search(r"^constexpr int kMaxPeers = \d+;.*$", src, re.M).group(0)
    params = re.search(r"^struct FoldParams \{.*?^\};\n", src, re.M | re.S).group(0)
    timeout = re.search(r"^constexpr unsigned long long kPeerTimeoutNs = [^;]+;", src, re.M).group(0)
    begin = src.index("constexpr int kCollectBatch")
    end = src.index("// The fold kernel.")
    body = src[begin:end]
    assert "exchange_bitmaps_ll" in body and "ld_relaxed_sys_u64" in body
    return "\n".join([peers, params, timeout, body])


def test_collector_source_under_host_shim(tmp_path):
    (tmp_path / "exchange_extract.inc").write_text(_extract())
    exe = tmp_path / "exchange_emul"
    subprocess.run(["g++", "-std=c++20", "-O1", "-pthread", "-Wall", "-I", str(tmp_path),
                    os.path.join(ROOT, "tests", "cpp", "exchange_emul.cpp"), "-o", str(exe)],
                   check=True, capture_output=True, text=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "ALL OK" in out.stdout, out.stdout + out.stderr
