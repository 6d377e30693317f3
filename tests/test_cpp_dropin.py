"""The C++ drop-in header (include/intp_b200/Interpolation.hpp): compiles with the host
compiler alone (CPU check) and reproduces the reference's interpolation-test scenarios on
the GPU (gpu check)."""
import json
import os
import subprocess

import pytest

from conftest import ROOT

CXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"


def _build(tmp_path, lib_built, source="drop_in_test.cpp"):
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_vectors.json")))["interpolation"]
    # outputs of the reference instantiated with float values on double coordinates (make_ref_outputs_mixed.py)
    g.update(json.load(open(os.path.join(ROOT, "tests", "golden", "ref_outputs_mixed.json"))))
    inc = tmp_path / "golden_vectors.inc"
    with open(inc, "w") as fh:
        fh.write("#include <vector>\nnamespace golden {\n")
        for k, v in g.items():
            fh.write("const std::vector<double> %s = {%s};\n" % (k, ", ".join(repr(float(x)) for x in v)))
        fh.write("}\n")
    exe = tmp_path / os.path.splitext(source)[0]
    pkg = os.path.join(ROOT, "bsplineinterpolation_b200")
    cmd = [CXX, "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), "-I", str(tmp_path),
           os.path.join(ROOT, "tests", "cpp", source), "-o", str(exe),
           "-L", pkg, "-lbspline_b200", "-Wl,-rpath," + pkg]
    subprocess.check_call(cmd)
    return exe


def _build_plain(tmp_path, source, lib_built):
    exe = tmp_path / os.path.splitext(source)[0]
    pkg = os.path.join(ROOT, "bsplineinterpolation_b200")
    subprocess.check_call([CXX, "-std=c++17", "-O1", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", source), "-o", str(exe),
                           "-L", pkg, "-lbspline_b200", "-Wl,-rpath," + pkg])
    return exe


def test_host_helpers_of_the_header(tmp_path, lib_built):
    """Dummy-sample stripping and value-type conversion: pure host code, runs here."""
    exe = _build_plain(tmp_path, "host_helpers_test.cpp", lib_built)
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0 and "all host helper checks passed" in r.stdout, r.stdout + r.stderr


def test_dummy_point_convention_compiles(tmp_path, lib_built):
    """The reference's default periodic convention (INTP_PERIODIC_NO_DUMMY_POINT undefined)."""
    assert os.path.exists(_build_plain(tmp_path, "dummy_point_test.cpp", lib_built))


@pytest.mark.gpu
def test_dummy_point_convention_on_the_gpu(tmp_path, lib_built):
    exe = _build_plain(tmp_path, "dummy_point_test.cpp", lib_built)
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    print(r.stdout[-4000:])
    assert r.returncode == 0, r.stdout[-4000:] + r.stderr[-2000:]


def test_header_compiles_and_links(tmp_path, lib_built):
    assert os.path.exists(_build(tmp_path, lib_built))
    assert os.path.exists(_build(tmp_path, lib_built, "drop_in_test2.cpp"))


@pytest.mark.gpu
def test_reference_scenarios_through_cpp_header(tmp_path, lib_built):
    exe = _build(tmp_path, lib_built)
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=600)
    print(r.stdout[-4000:])
    assert r.returncode == 0, r.stdout[-4000:] + r.stderr[-2000:]


def _cmake_consumer(tmp_path, lib_built):
    """The reference README's CMake snippet (find_package + target_link_libraries) and include line
    (<BSplineInterpolation/Interpolation.hpp>), unchanged, against cmake/BSplineInterpolationConfig.cmake."""
    import shutil
    cmake = shutil.which("cmake")
    if cmake is None:
        pytest.skip("cmake not available")
    build = tmp_path / "consumer_build"
    subprocess.check_call([cmake, "-S", os.path.join(ROOT, "tests", "cmake_consumer"), "-B", str(build),
                           "-DBSplineInterpolation_DIR=" + os.path.join(ROOT, "cmake"), "-DCMAKE_CXX_COMPILER=" + CXX],
                          stdout=subprocess.DEVNULL)
    subprocess.check_call([cmake, "--build", str(build)], stdout=subprocess.DEVNULL)
    return build / "main"


def test_cmake_package_builds_reference_style_project(tmp_path, lib_built):
    assert os.path.exists(_cmake_consumer(tmp_path, lib_built))
