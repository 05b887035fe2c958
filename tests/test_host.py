"""CPU: the product's transform code compiled as plain C++ (every lane of a group run in turn)
against the reference's own functions from oracle/_ref/libdav1d_ref.so."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libdav1d_ref.so")


def _build_and_run(tmp_path, name, *args):
    exe = tmp_path / name
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-o", str(exe),
                           os.path.join(ROOT, "tests", "host", name + ".cpp"), "-ldl"])
    return subprocess.run([str(exe), REF_SO, *args], capture_output=True, text=True)


def test_itx_1d_transforms_match_the_reference(tmp_path):
    """dav1d-mirror_b200/csrc/itx_1d.cuh vs src/itx_1d.c, 20000 vectors per 1-D function."""
    r = _build_and_run(tmp_path, "itx1d_check")
    assert r.returncode == 0, r.stdout[-2000:]


def test_itx2_all_slots_match_the_reference(tmp_path):
    """dav1d-mirror_b200/csrc/itx2.cuh (the 2-D transforms every CUDA path uses) vs the reference's
    itxfm_add table: 156 slots x 8/10/12 bit, dense blocks and packed boxes, both prediction sources."""
    r = _build_and_run(tmp_path, "itx2_check", "120")
    assert r.returncode == 0, r.stdout[-2000:]
    assert "468 slots" in r.stdout and "ok" in r.stdout
