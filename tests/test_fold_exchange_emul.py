"""The fold kernel and its fused multi-GPU exchange (fold_words, exchange_bitmaps, exchange_bitmaps_ll, k_fold in
gpu-pruner_b200/csrc/gpr_kernels.cuh) run on the CPU: the functions' source text is cut out of the kernel header and
compiled under a host shim (tests/cpp/fold_emul.cpp: CTAs and warps on real threads, several emulated ranks driven
through back-to-back decisions the way decide_impl launches them).  Every protocol — tagged slots in order,
pipelined, flags — must deliver every decision's counters and gathered bitmaps, keep writes to a shared output buffer
in launch order, leave the scratch state zeroed, and never deadlock, with ranks drifting steps apart."""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HDR = os.path.join(ROOT, "gpu-pruner_b200", "csrc", "gpr_kernels.cuh")

REWRITES = [
    # the one inline-PTX statement of fold_words: the tagged 64-bit peer store
    ('asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(dst), "l"(v) : "memory");', "st_relaxed_sys_u64(dst, v);"),
    # shared memory of a CTA lives in the shim's per-CTA context
    ("__shared__ unsigned long long s_cnt[3];", "unsigned long long* s_cnt = tl_cta->s_cnt;"),
    ("__shared__ unsigned int s_last;", "unsigned int& s_last = tl_cta->s_last;"),
]


def _extract():
    src = open(HDR).read()
    peers = re.search(r"^constexpr int kMaxPeers = \d+;.*$", src, re.M).group(0)
    params = re.search(r"^struct FoldParams \{.*?^\};\n", src, re.M | re.S).group(0)
    timeout = re.search(r"^constexpr unsigned long long kPeerTimeoutNs = [^;]+;", src, re.M).group(0)
    begin = src.index("template <int BATCH>\n__device__ __forceinline__ void fold_words")
    end = src.index("// Device-side rendezvous + time mark")
    body = src[begin:end]
    for old, new in REWRITES:
        assert body.count(old) == 1, old
        body = body.replace(old, new)
    assert "asm" not in body and "__shared__" not in body
    return "\n".join([peers, params, timeout, body])


def test_fold_and_exchange_source_under_host_shim(tmp_path):
    (tmp_path / "fold_extract.inc").write_text(_extract())
    exe = tmp_path / "fold_emul"
    subprocess.run(["g++", "-std=c++20", "-O1", "-pthread", "-Wall", "-Wno-unknown-pragmas", "-I", str(tmp_path),
                    os.path.join(ROOT, "tests", "cpp", "fold_emul.cpp"), "-o", str(exe)],
                   check=True, capture_output=True, text=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=1200)
    assert out.returncode == 0 and "ALL OK" in out.stdout, out.stdout + out.stderr


def test_no_data_race_under_thread_sanitizer(tmp_path):
    """The same source under ThreadSanitizer: with CTAs, warps and ranks on real threads and the CUDA barriers / scoped
    atomics mapped to C++ ones, a missing __syncthreads or fence in the kernel text is a reported race (removing the two
    barriers between collecting and assembling, for instance, is).  All protocols, 2 to 8 ranks, the lagging rank."""
    (tmp_path / "fold_extract.inc").write_text(_extract())
    exe = tmp_path / "fold_emul_tsan"
    subprocess.run(["g++", "-std=c++20", "-O1", "-g", "-pthread", "-fsanitize=thread", "-Wno-unknown-pragmas",
                    "-I", str(tmp_path), os.path.join(ROOT, "tests", "cpp", "fold_emul.cpp"), "-o", str(exe)],
                   check=True, capture_output=True, text=True)
    env = dict(os.environ, TSAN_OPTIONS="halt_on_error=1")
    for scenario in (0, 1, 2, 4, 7, 10, 12):
        out = subprocess.run([str(exe), str(scenario)], capture_output=True, text=True, timeout=900, env=env)
        assert out.returncode == 0 and "ALL OK" in out.stdout and "ThreadSanitizer" not in out.stderr, \
            (scenario, out.stdout, out.stderr[-3000:])
