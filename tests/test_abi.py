"""CPU-side checks of the drop-in boundary: liborbx.so builds/loads, exports every symbol include/orbx.h declares,
record layouts equal cv::KeyPoint / cv::DMatch, and the library refuses to run without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "orbx.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(orbx_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported():
    from rgbd_visualodometry_b200 import _lib
    lib = _lib.load()
    decl = _declared()
    assert len(decl) >= 18
    for name in decl:
        assert hasattr(lib, name), f"{name} declared in include/orbx.h but not exported"
    assert sorted(_lib.SYMBOLS) == decl, "python binding table and header disagree"
    assert b"sm_100a" in lib.orbx_version()


def test_record_layouts():
    from rgbd_visualodometry_b200 import _lib, orb
    assert C.sizeof(_lib.Keypoint) == 28 == orb.KP_DTYPE.itemsize          # cv::KeyPoint
    assert C.sizeof(_lib.Match) == 16 == orb.DMATCH_DTYPE.itemsize         # cv::DMatch
    assert [orb.KP_DTYPE.fields[n][1] for n in ("x", "y", "size", "angle", "response", "octave", "class_id")] == [0, 4, 8, 12, 16, 20, 24]
    assert [orb.DMATCH_DTYPE.fields[n][1] for n in ("queryIdx", "trainIdx", "imgIdx", "distance")] == [0, 4, 8, 12]


def test_no_cpu_fallback():
    """Without a CUDA device the product path must fail loudly, never compute on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from rgbd_visualodometry_b200 import orb
    with pytest.raises(orb.OrbxError) as e:
        orb.Context()
    assert e.value.code == orb.E_CUDA
    with pytest.raises(orb.OrbxError):
        orb.ORB_create(500).detectAndCompute(np.zeros((480, 640, 3), np.uint8))
    with pytest.raises(orb.OrbxError):
        orb.BFMatcher(orb.NORM_HAMMING)


def test_create_rejects_bad_arguments():
    from rgbd_visualodometry_b200 import _lib
    lib = _lib.load()
    h = C.c_void_p()
    assert lib.orbx_create(C.byref(h), 0, 500, 1.2, 0, 640, 480, 1) == -1      # nlevels < 1
    assert lib.orbx_create(C.byref(h), 0, 500, 1.0, 8, 640, 480, 1) == -1      # scale factor must exceed 1
    assert lib.orbx_create(None, 0, 500, 1.2, 8, 640, 480, 1) == -1
    assert lib.orbx_last_error(None) == b"null context"


def test_filter_matches_host_glue(oracle):
    """orbx_filter_matches == the reference's threshold loop (src/frontend.cpp:190-211) == oracle restatement."""
    from rgbd_visualodometry_b200 import orb
    from rgbd_visualodometry_b200.synth import synth_descriptors, synth_map_queries
    t = synth_descriptors(400, 3)
    m = oracle.match_hamming(synth_map_queries(t, 250, 4), t)
    assert orb.filter_matches(m.astype(orb.DMATCH_DTYPE), 2.0).tobytes() == oracle.filter_matches(m, 2.0).tobytes()
    assert len(orb.filter_matches(np.zeros(0, orb.DMATCH_DTYPE))) == 0


def test_filter_matches_skips_empty_frame_placeholders(oracle):
    """The batched matchers emit trainIdx = -1, distance = 0 for a frame without keypoints; such a record is neither a match
    nor a candidate for the minimum (a shim would otherwise index keypointsCurr_[-1], src/frontend.cpp:208-209)."""
    from rgbd_visualodometry_b200 import orb
    from rgbd_visualodometry_b200.synth import synth_descriptors, synth_map_queries
    t = synth_descriptors(300, 5)
    m = oracle.match_hamming(synth_map_queries(t, 200, 6), t).astype(orb.DMATCH_DTYPE)
    want = oracle.filter_matches(m, 2.0)
    holes = m.copy()
    holes = np.insert(holes, [0, 50, 200], np.array([(7, -1, 0, 0.0)], orb.DMATCH_DTYPE))
    got = orb.filter_matches(holes, 2.0)
    assert (got["trainIdx"] >= 0).all()
    assert got[["trainIdx", "distance"]].tobytes() == want.astype(orb.DMATCH_DTYPE)[["trainIdx", "distance"]].tobytes()
    only = np.array([(0, -1, 0, 0.0), (1, -1, 0, 0.0)], orb.DMATCH_DTYPE)
    assert len(orb.filter_matches(only, 2.0)) == 0


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under the package may reference it (or cv2)."""
    pkg = os.path.join(ROOT, "rgbd_visualodometry_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                s = open(os.path.join(dp, f)).read()
                assert "import cv2" not in s and "from oracle" not in s and "import oracle" not in s and "orb_oracle" not in s, f


def test_header_is_plain_c_and_a_c_caller_links(tmp_path):
    """The boundary is a C ABI: the header must compile as C99 and as C++11 (the reference is C++11, CMakeLists.txt:8), and a
    plain C translation unit using it must link against liborbx.so (no C++ / torch types in the signatures)."""
    import shutil
    import subprocess
    from rgbd_visualodometry_b200 import _lib
    _lib.load()
    hdr = os.path.join(ROOT, "include", "orbx.h")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-x", "c", hdr])
    subprocess.check_call(["g++", "-std=c++11", "-Wall", "-Werror", "-fsyntax-only", "-x", "c++", hdr])
    src = tmp_path / "caller.c"
    src.write_text('''#include "orbx.h"
#include <stdio.h>
int main(void) {
    orbx_ctx* c = 0;
    int rc = orbx_create(&c, 0, 500, 1.2f, 8, 640, 480, 1);      /* no GPU in the build container: must fail, not fall back */
    printf("%s rc=%d sizeof(kp)=%d sizeof(m)=%d\\n", orbx_version(), rc, (int)sizeof(orbx_keypoint), (int)sizeof(orbx_match));
    if (c) orbx_destroy(c);
    return (sizeof(orbx_keypoint) == 28 && sizeof(orbx_match) == 16) ? 0 : 1;
}
''')
    exe = tmp_path / "caller"
    libdir = os.path.dirname(_lib.SO)
    subprocess.check_call(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe), "-L", libdir, "-lorbx",
                           "-Wl,-rpath," + libdir])
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0 and "sm_100a" in out.stdout, out.stdout + out.stderr


def _build_shim(tmp_path):
    import subprocess
    from rgbd_visualodometry_b200 import _lib
    _lib.load()
    exe = str(tmp_path / "frontend_shim")
    libdir = os.path.dirname(_lib.SO)
    subprocess.check_call(["g++", "-std=c++11", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "frontend_shim.cpp"),
                           "-o", exe, "-L", libdir, "-lorbx", "-Wl,-rpath," + libdir])
    return exe


def test_cpp_frontend_shim_builds_and_has_no_fallback(tmp_path):
    """INTEGRATION.md's FrontEnd shim as a C++11 translation unit (examples/frontend_shim.cpp) compiles warning-free against
    include/orbx.h and links liborbx.so; without a GPU its constructor throws (rc 3) instead of falling back to anything."""
    import subprocess
    import torch
    exe = _build_shim(tmp_path)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    if torch.cuda.is_available():
        assert out.returncode == 0 and "shim ok" in out.stdout, out.stdout + out.stderr
    else:
        assert out.returncode == 3 and "no sm_100 CUDA device" in out.stdout, out.stdout + out.stderr
