"""GPU tests of the C++ host facade (icp-symm_b200/host): the MyICP / regist.h shaped interface
reproduces the reference's only known answer on its own fixture (main.cpp:44-52)."""
import os
import subprocess

import numpy as np
import pytest

import workloads as wl

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "icp-symm_b200", "host")


def _write_pcd(path, xyz, extra_fields=False):
    with open(path, "w") as f:
        n = len(xyz)
        if extra_fields:   # the cat_out.pcd layout (cat_out.pcd:1-11)
            f.write("# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS x y z label normal_x normal_y normal_z curvature\n"
                    "SIZE 4 4 4 4 4 4 4 4\nTYPE F F F U F F F F\nCOUNT 1 1 1 1 1 1 1 1\n")
        else:
            f.write("# .PCD v.7 - Point Cloud Data file format\nVERSION .7\nFIELDS x y z\nSIZE 4 4 4\nTYPE F F F\nCOUNT 1 1 1\n")
        f.write(f"WIDTH {n}\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS {n}\nDATA ascii\n")
        for p in xyz:
            f.write("%.9g %.9g %.9g" % tuple(p) + (" 0 0 0 0 0\n" if extra_fields else "\n"))


def _parse_T(stdout):
    line = [ln for ln in stdout.splitlines() if ln.startswith("T ")][-1]
    return np.array([float(v) for v in line.split()[1:]]).reshape(4, 4)


def test_cli_main_cpp_scenario(tmp_path, cat, cat_out, gold):
    _write_pcd(tmp_path / "cat.pcd", cat)
    _write_pcd(tmp_path / "cat_out.pcd", cat_out, extra_fields=True)
    exe = os.path.join(HOST, "icp_symm_cli")
    # no arguments: the reference's hard-coded file names and defaults (main.cpp:8, myicp.cpp:6)
    r = subprocess.run([exe], cwd=tmp_path, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    assert "Result transform:" in r.stdout and "iters#" in r.stdout        # myicp.cpp:125,146
    T = _parse_T(r.stdout)
    assert wl.rotation_angle(T, gold["truth_R"]) < 1e-5
    assert np.abs(T[:3, 3] - gold["truth_t"]).max() < 1e-5 * wl.extent(cat)
    assert "iterations 5 converged 1" in r.stdout
    # --out writes the aligned source: it must coincide with the target
    out = tmp_path / "aligned.pcd"
    r = subprocess.run([exe, "cat.pcd", "cat_out.pcd", "--iters", "50", "--quiet", "--out", str(out)], cwd=tmp_path,
                       capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    from oracle import oracle as orc
    aligned = orc.pcd_read_xyz(str(out))
    assert np.abs(aligned - cat_out).max() < 5e-5
    # point-to-point mode (RegisterP2P): also recovers the fixture's transform
    r = subprocess.run([exe, "cat.pcd", "cat_out.pcd", "--iters", "60", "--thr", "0.5", "--p2p", "--quiet"], cwd=tmp_path,
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    assert wl.rotation_angle(_parse_T(r.stdout), gold["truth_R"]) < 1e-3


def test_register_test_cpp_scenario(tmp_path, cat, cat_out, gold):
    _write_pcd(tmp_path / "cat.pcd", cat)
    _write_pcd(tmp_path / "cat_out.pcd", cat_out, extra_fields=True)
    r = subprocess.run([os.path.join(HOST, "register_test")], cwd=tmp_path, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    R = np.array([float(v) for v in [ln for ln in r.stdout.splitlines() if ln.startswith("R ")][0].split()[1:]]).reshape(3, 3)
    T = np.array([float(v) for v in [ln for ln in r.stdout.splitlines() if ln.startswith("T ")][0].split()[1:]])
    assert np.allclose(R, gold["kabsch_R"], atol=1e-9) and np.allclose(T, gold["kabsch_T"], atol=1e-7)
