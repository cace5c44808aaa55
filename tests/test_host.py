"""CPU tests of the boundary: the C-ABI library loads without a GPU and exports every symbol
include/icpsymm.h declares, the host logic (PCD I/O) works, and nothing falls back to the CPU."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "icp-symm_b200")


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as g
    g.build()
    return g


def test_library_exports_every_declared_symbol(built):
    import icpsymm
    lib = icpsymm.load()
    hdr = open(os.path.join(ROOT, "include", "icpsymm.h")).read()
    declared = sorted(set(re.findall(r"\b(icpsymm_[a-z_0-9]+)\s*\(", hdr)))
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/icpsymm.h but not exported"
    # the binding covers the same set
    assert set(declared) == set(built.declared_symbols())


def test_release_library_exports_nothing_but_the_header(built):
    """developer instrumentation (ICP_BLOCK_TIMES, ICP_TILE_TIMES, ICP_EXPERIMENT) must not leak into the shipped library:
    its dynamic symbol table holds exactly the icpsymm_* entry points of include/icpsymm.h"""
    out = subprocess.run(["nm", "-D", "--defined-only", os.path.join(PKG, "libicpsymm_b200.so")], capture_output=True, text=True)
    assert out.returncode == 0
    exported = {ln.split()[-1] for ln in out.stdout.splitlines() if " T " in ln}
    ours = {s for s in exported if "icp" in s.lower() and not s.startswith("_Z")}
    assert ours == set(built.declared_symbols()), sorted(ours ^ set(built.declared_symbols()))
    assert not any("debug" in s for s in exported)


def test_header_cites_the_reference_interfaces():
    hdr = open(os.path.join(ROOT, "include", "icpsymm.h")).read()
    for cite in ("myicp.h:7-36", "regist.h:8", "myicp.cpp:100-150", "myicp.cpp:152-172", "func.cpp:43-60"):
        assert cite in hdr


def test_python_binding_constants_match_the_header():
    """option keys and status codes of the ctypes binding are the ones include/icpsymm.h declares"""
    import re
    import icpsymm as icp
    hdr = open(os.path.join(ROOT, "include", "icpsymm.h")).read()
    enums = dict((m.group(1), int(m.group(2))) for m in re.finditer(r"\b(ICPSYMM_[A-Z0-9_]+)\s*=\s*(-?\d+)", hdr))
    opts = {k: v for k, v in enums.items() if k.startswith("ICPSYMM_OPT_")}
    assert len(opts) >= 10
    for name, value in opts.items():
        assert getattr(icp, name[len("ICPSYMM_"):]) == value, name


def test_status_strings_and_struct_layout(built):
    import icpsymm
    lib = icpsymm.load()
    assert lib.icpsymm_status_string(0) == b"ok"
    assert b"no CPU fallback" in lib.icpsymm_status_string(icpsymm.ERR_CUDA)
    assert ctypes.sizeof(icpsymm.Stats) == 72      # must match struct icpsymm_stats


def test_no_cpu_fallback_without_a_device(built):
    import icpsymm
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("a GPU is present")
    with pytest.raises(icpsymm.IcpError) as e:
        icpsymm.Context(0)
    assert e.value.status == icpsymm.ERR_CUDA
    with pytest.raises(icpsymm.IcpError):
        icpsymm.device_info(0)


def test_product_never_links_or_imports_the_oracle(built):
    # the shared library has no dependency on oracle/, and no product source mentions it
    out = subprocess.run(["ldd", os.path.join(PKG, "libicpsymm_b200.so")], capture_output=True, text=True).stdout
    assert "oracle" not in out
    for dirpath, _, files in os.walk(PKG):
        for f in files:
            if f.endswith((".cu", ".cuh", ".h", ".cpp", ".py")) and f != "workloads.py":
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "liboracle" not in txt and "icp_oracle" not in txt and "from oracle" not in txt, f


def test_host_pcd_io_selftest(built, tmp_path):
    exe = os.path.join(PKG, "host", "host_selftest")
    r = subprocess.run([exe, str(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "host selftest ok" in r.stdout


def test_txt2pcd_matches_fixture(built, tmp_path, bunny):
    # test.cpp:33-58: za.txt -> txt2pcd_bunny1.pcd
    txt = tmp_path / "za.txt"
    np.savetxt(txt, bunny, fmt="%.9g", delimiter="\t")
    pcd = tmp_path / "bunny.pcd"
    r = subprocess.run([os.path.join(PKG, "host", "txt2pcd"), str(txt), str(pcd)], capture_output=True, text=True)
    assert r.returncode == 0 and "Saved 93 data points" in r.stdout
    from oracle import oracle as orc
    assert np.array_equal(orc.pcd_read_xyz(str(pcd)), bunny)


def test_cli_reports_missing_device_or_file(built, tmp_path):
    exe = os.path.join(PKG, "host", "icp_symm_cli")
    r = subprocess.run([exe, str(tmp_path / "nope.pcd"), str(tmp_path / "nope2.pcd")], capture_output=True, text=True)
    assert r.returncode != 0      # 2 = no CUDA device, 1 = unreadable file: never a silent success
