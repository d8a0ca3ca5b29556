#!/usr/bin/env python
"""Test infrastructure: compile the REFERENCE's own MISE (reg_slices/src_convonet/utils/libmise/mise.pyx) from where
it lies under /root/reference into oracle/_ref/ (git-ignored), so that slice3d_b200/mise.py can be pinned against
the real thing (tests/test_mise.py) and golden vectors can be generated (oracle/make_golden_mise.py).

    python oracle/build_ref_mise.py        # -> oracle/_ref/mise.<abi>.so

Only a translation (Cython -> C++) and a compile of the reference's file; no reference source is copied into
the repository.
"""
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
SRC = "/root/reference/reg_slices/src_convonet/utils/libmise/mise.pyx"


def build():
    if not os.path.exists(SRC):
        return None
    os.makedirs(OUT, exist_ok=True)
    so = os.path.join(OUT, "mise" + sysconfig.get_config_var("EXT_SUFFIX"))
    if os.path.exists(so) and os.path.getmtime(so) >= os.path.getmtime(SRC):
        return so
    cpp = os.path.join(OUT, "mise.cpp")
    subprocess.check_call([sys.executable, "-m", "cython", "--cplus", "-3", SRC, "-o", cpp])
    inc = sysconfig.get_paths()["include"]
    subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-w", "-I", inc, cpp, "-o", so])
    os.remove(cpp)
    return so


if __name__ == "__main__":
    print(build())
