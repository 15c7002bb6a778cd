"""Builds oracle/_ref/libref_sogp.so: the reference's own SOGP sources (compiled from /root/reference/src,
where they lie) over oracle/eigen_shim.  Runs only where /root/reference exists (this container); the
built .so is git-ignored but travels to the GPU box with the snapshot.  Test infrastructure."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/src"
OUT = os.path.join(HERE, "_ref", "libref_sogp.so")


def build():
    if not os.path.isdir(REF):
        return None
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    srcs = [os.path.join(HERE, "ref_sogp.cpp"), os.path.join(REF, "rbf_kernel.cpp"), os.path.join(REF, "gaussian_noise.cpp"),
            os.path.join(REF, "gaussian_noise_3d.cpp")]
    deps = srcs + [os.path.join(HERE, "eigen_shim", "Eigen", "Dense"), os.path.join(REF, "sparse_gp.hpp"), os.path.join(REF, "sparse_gp.h"),
                   os.path.join(REF, "sparse_gp_field.hpp"), os.path.join(REF, "sparse_gp_field.h")]
    if os.path.exists(OUT) and all(os.path.getmtime(d) <= os.path.getmtime(OUT) for d in deps):
        return OUT
    cmd = ["g++", "-O2", "-std=c++11", "-ffp-contract=off", "-fPIC", "-shared", "-w", "-I" + os.path.join(HERE, "eigen_shim"), "-I" + REF] + srcs + ["-o", OUT]
    subprocess.check_call(cmd)
    return OUT


OUT_FRAMES = os.path.join(HERE, "_ref", "libref_frames.so")


def frames_slice():
    """(first, last) 1-based line numbers of compute_rotation .. project_points in the reference's gp_compressor.cpp."""
    lines = open(os.path.join(REF, "gp_compressor.cpp")).read().split("\n")
    first = next(i for i, l in enumerate(lines) if l.startswith("void gp_compressor::compute_rotation")) + 1
    last = next(i for i, l in enumerate(lines) if l.startswith("// do this for to_be_added instead")) + 1 - 1
    return first, last


def build_frames():
    """gp_compressor::compute_rotation and ::project_points (gp_compressor.cpp:29-118) compiled from the reference's own text:
    prelude + that line range (streamed from /root/reference, never written to disk) + C wrappers, piped to g++ on stdin."""
    if not os.path.isdir(REF):
        return None
    os.makedirs(os.path.dirname(OUT_FRAMES), exist_ok=True)
    pre, post = os.path.join(HERE, "ref_frames_prelude.h"), os.path.join(HERE, "ref_frames_post.cpp")
    src = os.path.join(REF, "gp_compressor.cpp")
    deps = [pre, post, src, os.path.join(HERE, "eigen_shim", "Eigen", "Dense"), os.path.abspath(__file__)]
    if os.path.exists(OUT_FRAMES) and all(os.path.getmtime(d) <= os.path.getmtime(OUT_FRAMES) for d in deps):
        return OUT_FRAMES
    first, last = frames_slice()
    lines = open(src).read().split("\n")
    unit = open(pre).read() + "\n".join(lines[first - 1:last]) + "\n" + open(post).read()
    cmd = ["g++", "-O2", "-std=c++11", "-ffp-contract=off", "-fPIC", "-shared", "-w", "-I" + os.path.join(HERE, "eigen_shim"),
           "-x", "c++", "-", "-o", OUT_FRAMES]
    subprocess.run(cmd, input=unit.encode(), check=True)
    return OUT_FRAMES


if __name__ == "__main__":
    print(build())
    print(build_frames())
