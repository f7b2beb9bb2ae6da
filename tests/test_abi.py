"""The C-ABI libraries load on a CPU-only box and export every symbol their headers declare (no compute calls)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header, prefix):
    txt = open(os.path.join(ROOT, "include", header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(" + prefix + r"[a-z0-9_]+)\s*\(", txt)))


def test_device_library_exports_every_declared_symbol():
    from lt_mapper_b200 import binding
    names = _declared("ltr_b200.h", "ltr_")
    assert len(names) >= 35
    L = ctypes.CDLL(binding.LIB_PATH)
    for n in names:
        assert hasattr(L, n), n
    assert set(binding.EXPORTS) == set(names)


def test_host_library_exports_every_declared_symbol():
    from lt_mapper_b200 import removert
    names = _declared("ltr_removert.h", "ltrh_")
    L = removert.host_lib()
    for n in names:
        assert hasattr(L, n), n
    assert set(removert.HOST_EXPORTS) == set(names)


def test_no_cpu_fallback():
    """Without a GPU the product fails loudly instead of computing on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import lt_mapper_b200 as ltr
    with pytest.raises(ltr.LtrError):
        ltr.Context()
    from lt_mapper_b200 import removert
    with pytest.raises(ltr.LtrError):
        removert.Removerter()


def test_product_never_imports_oracle():
    for dp, _, files in os.walk(os.path.join(ROOT, "lt_mapper_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert "import oracle" not in txt and "oracle/" not in txt and "liboracle" not in txt, f


def test_reset_rimg_size_host_side():
    import lt_mapper_b200 as ltr
    for a, rc in {2.5: (125, 900), 2.375: (119, 855), 2.0: (100, 720), 1.9: (95, 684), 1.5: (75, 540), 1.425: (71, 513), 3.0: (150, 1080)}.items():
        assert ltr.reset_rimg_size(a) == rc


def test_params_defaults_match_reference_yaml():
    from lt_mapper_b200 import removert
    p = removert.Params()
    removert.host_lib().ltrh_params_default(ctypes.byref(p))
    assert (p.sequence_vfov, p.sequence_hfov) == (50.0, 360.0)                     # RosParamServer.cpp:15-16
    assert p.num_nn_points_within == 2 and abs(p.dist_nn_points_within - 0.01) < 1e-9  # params_ltmapper.yaml:65-66
    assert abs(p.downsample_voxel_size - 0.05) < 1e-9 and p.n_schedule == 1 and p.schedule_res[0] == 2.5
    assert list(p.ExtrinsicLiDARtoPoseBase) == [1.0 if i % 5 == 0 else 0.0 for i in range(16)]
    s = removert.selfremovert_schedule([2.5, 2.0])
    assert [o for o, _ in s] == [0, 1, 0, 0, 1, 0] and abs(s[1][1] - 2.375) < 1e-6


def test_pcl_adapter_builds_against_stand_in_headers(tmp_path):
    """include/ltr_pcl_adapter.hpp (the glue a ROS build of the node would include) compiles against the stand-in PCL /
    Eigen headers, links against both libraries, and its host-side conversions round-trip."""
    import subprocess
    exe = str(tmp_path / "adapter_check")
    libdir = os.path.join(ROOT, "lt_mapper_b200")
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                           "-I", os.path.join(ROOT, "oracle", "ref_shim", "include"), os.path.join(ROOT, "tests", "adapter_check.cpp"),
                           "-o", exe, "-L", libdir, "-lltr_removert", "-lltr_b200", "-Wl,-rpath," + libdir])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and "adapter ok" in out.stdout, out.stdout + out.stderr
