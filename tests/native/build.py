"""Builds tests/native/libiai_cpu.so: the product's C++ IAI engine header driven by an oracle-backed CPU backend
(test double; see iai_engine_cpu.cpp)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))


def build():
    import orc
    orc.build()
    so = os.path.join(HERE, "libiai_cpu.so")
    srcs = [os.path.join(HERE, "iai_engine_cpu.cpp"), os.path.join(ROOT, "autobzcore.jl_b200", "csrc", "abz_iai_engine.hpp")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        odir = os.path.join(ROOT, "oracle")
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-o", so, srcs[0],
                               "-I", os.path.dirname(srcs[1]), "-L", odir, "-lorc", f"-Wl,-rpath,{odir}"])
    return so
