"""Without libnccl the multi-GPU entry points must fail with SLAMRS_E_NCCL and a message, not crash
(comm.cu reads dlerror() once: a second call returns NULL)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_unique_id_without_nccl_is_an_error_not_a_crash():
    code = ("import sys; sys.path.insert(0, %r)\n"
            "from slamrs_b200 import _lib, nccl_unique_id\n"
            "try:\n"
            "    nccl_unique_id()\n"
            "except _lib.SlamrsGpuError as e:\n"
            "    assert e.code == _lib.E_NCCL, e.code\n"
            "    assert 'dlopen(' in str(e) and 'no-such-nccl' in str(e), str(e)\n"
            "    print('OK')\n" % ROOT)
    env = dict(os.environ, SLAMRS_NCCL_LIB="/no-such-nccl/libnccl.so.2")
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0 and "OK" in out.stdout, out.stdout + out.stderr
