"""include/sgufp_b200.hpp compiled against the reference's own headers and objects: the C++
drop-in adapters must build with the reference's types and agree with the reference's DD on
structure (CPU, SGUFP_DEVICE_NONE).  Only where /root/reference exists."""
import os
import subprocess
import tempfile

import pytest

from sgufp_solver_b200 import instances as I

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"


@pytest.mark.skipif(not os.path.isdir(REF), reason="/root/reference absent")
def test_cpp_adapter_builds_and_agrees(built_lib, tmp_path):
    exe = str(tmp_path / "adapter_check")
    cmd = ["g++", "-std=gnu++20", "-O1", "-w", f"-I{REF}", f"-I{ROOT}/include", os.path.join(ROOT, "tests", "cpp", "adapter_check.cpp"),
           f"{REF}/Network.cpp", f"{REF}/DD.cpp", f"{REF}/optimized.cpp", built_lib, "-o", exe, f"-Wl,-rpath,{os.path.dirname(built_lib)}"]
    subprocess.check_call(cmd)
    for inst in (I.config1(S=2), I.config2(S=1)):
        f = str(tmp_path / f"{inst.name}.txt")
        inst.write_text(f)
        out = subprocess.run([exe, f], capture_output=True, text=True)
        assert out.returncode == 0, out.stdout + out.stderr
        assert "adapter ok" in out.stdout
