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


if __name__ == "__main__":
    print(build())
