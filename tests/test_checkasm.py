"""The reference's own parity harness, tests/checkasm, run against the CUDA DSP tables.

oracle/build_checkasm.py (run by __graft_entry__.build() where /root/reference exists) builds
checkasm from the reference's sources with the three hook patches of SURVEY.md section 7 step 1
(a "cuda" cpu flag; the flag reported by dav1d_init_cpu; dav1d_cuda_*_dsp_init_* called at the
tail of the three C init functions) and links it with libdav1d_cuda.so.  checkasm then compares
every function of the CUDA tables with the C templates on its own inputs, guard bands included."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "oracle", "_ref", "checkasm", "checkasm")


def test_checkasm_is_built_and_knows_the_cuda_flag():
    """CPU: the harness exists (prebuilt binary travels to the GPU box) and lists the new flag."""
    assert os.path.exists(EXE), "run `python oracle/build_checkasm.py`"
    r = subprocess.run([EXE, "--list-cpuflags"], capture_output=True, text=True, timeout=60)
    assert "cuda" in r.stdout + r.stderr
    r = subprocess.run([EXE, "--list-tests"], capture_output=True, text=True, timeout=60)
    assert set(r.stdout.split()) >= {"mc_8bpc", "itx_8bpc", "ipred_8bpc", "mc_16bpc", "itx_16bpc", "ipred_16bpc"}


@pytest.mark.gpu
@pytest.mark.parametrize("test", ["mc_8bpc", "mc_16bpc", "itx_8bpc", "itx_16bpc", "ipred_8bpc", "ipred_16bpc"])
def test_checkasm_cuda_tables(test):
    for seed in (1, 20261019):
        r = subprocess.run([EXE, f"--test={test}", str(seed)], capture_output=True, text=True, timeout=1500)
        out = r.stdout + r.stderr
        m = re.search(r"all (\d+) tests passed", out)
        assert r.returncode == 0 and m, out[-3000:]
        assert "CUDA" in out, out[-2000:]          # the pass for the new flag ran
        print(f"checkasm --test={test} {seed}: all {m.group(1)} tests passed")
