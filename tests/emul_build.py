"""Builds tests/cpp/text_emul.cpp (TEST INFRASTRUCTURE): the emulated device behind ingest_device.cpp.
flavour "tiles":  the parse pass calls the host/device parser core (gpr_text.cuh) tile by tile, candidate by candidate;
flavour "kernel": the parse pass runs the SOURCE of k_text_parse (gpr_text_kernels.cuh), cut out verbatim, as real
                  threads under tests/cpp/cuda_shim.hpp — warp compaction of the '[' offsets, the bulk-copy ring,
                  one candidate per lane, atomics into the plane."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "gpu-pruner_b200", "host")


def extract_parse_kernel():
    src = open(os.path.join(ROOT, "gpu-pruner_b200", "csrc", "gpr_text_kernels.cuh")).read()
    body = src[src.index("// NaN-aware max into a cell that starts as kFillBits"):src.index("}  // namespace text")]
    for old, new in (('asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");', ";"),
                     ("extern __shared__ __align__(128) unsigned char smem[];", "unsigned char* smem = tl_cta->smem;")):
        assert body.count(old) == 1, old
        body = body.replace(old, new)
    assert "asm" not in body and "__shared__" not in body and "k_text_parse" in body and "k_fill_columns" in body
    return body


def build(out_dir, flavour, sanitize="address,undefined"):
    out = os.path.join(str(out_dir), "text_emul_" + flavour)
    cmd = ["g++", "-O1", "-g", "-fsanitize=" + sanitize, "-fno-omit-frame-pointer", "-I", HOST]
    if sanitize != "thread":
        cmd.append("-fno-sanitize-recover=all")
    if flavour == "kernel":
        with open(os.path.join(str(out_dir), "text_kernel_extract.inc"), "w") as f:
            f.write(extract_parse_kernel())
        cmd += ["-std=c++20", "-DEMUL_PARSE_KERNEL", "-Wno-unknown-pragmas", "-I", os.path.join(ROOT, "tests", "cpp"),
                "-I", str(out_dir)]
    else:
        cmd.append("-std=c++17")
    cmd += [os.path.join(ROOT, "tests", "cpp", "text_emul.cpp"), os.path.join(HOST, "ingest.cpp"),
            os.path.join(HOST, "ingest_device.cpp"), os.path.join(HOST, "json.cpp"), "-o", out, "-lpthread"]
    subprocess.check_call(cmd)
    return out
