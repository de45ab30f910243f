"""CPU tests of the drop-in boundary: liblgrast.so loads and exports every symbol include/lgrast.h declares,
the Python surface has the reference's names / field order / argument validation, and nothing in the product
package reaches into oracle/."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "lgrast.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(lgr_[a-z_]+)\s*\(", text)) - {"lgr_alloc_fn"})


def test_library_exports_every_declared_symbol():
    from lightgaussian_b200 import build
    path = build.build_library()
    lib = C.CDLL(path)
    names = _declared_symbols()
    assert {"lgr_forward", "lgr_forward_count", "lgr_backward", "lgr_mark_visible"} <= set(names)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/lgrast.h but not exported"
    lib.lgr_abi_version.restype = C.c_int
    assert lib.lgr_abi_version() == 2


def test_layout_queries_need_no_gpu():
    from lightgaussian_b200 import capi
    gl, total = capi.geometry_layout(1000)
    assert total > 0 and all(v % 256 == 0 for v in gl.values())
    assert gl["means2D"] - gl["depth"] >= 4 * 1000
    il, itotal = capi.image_layout(1920, 1080)
    assert il["n_contrib"] - il["final_T"] >= 4 * 1920 * 1080 and itotal > il["ranges"]
    bl, btotal = capi.binning_layout(12345, 1920, 1080)
    assert btotal >= 12345 * (4 + 4 + 2 + 2)


def test_python_surface_matches_reference_names():
    from lightgaussian_b200.rasterizer import GaussianRasterizationSettings, GaussianRasterizer, rasterize_gaussians, _C
    assert GaussianRasterizationSettings._fields == ("image_height", "image_width", "tanfovx", "tanfovy", "bg", "scale_modifier",
                                                     "viewmatrix", "projmatrix", "sh_degree", "campos", "prefiltered", "debug", "f_count")
    for n in ("rasterize_gaussians", "count_gaussians", "rasterize_gaussians_backward", "mark_visible"):
        assert callable(getattr(_C, n))
    for n in ("forward", "forward_count", "markVisible"):
        assert callable(getattr(GaussianRasterizer, n))
    assert callable(rasterize_gaussians)


def test_argument_validation_matches_reference():
    import torch
    from lightgaussian_b200.rasterizer import GaussianRasterizationSettings, GaussianRasterizer
    rs = GaussianRasterizationSettings(8, 8, 1.0, 1.0, torch.zeros(3), 1.0, torch.eye(4), torch.eye(4), 0, torch.zeros(3), False, False, False)
    r = GaussianRasterizer(rs)
    x = torch.zeros(4, 3)
    with pytest.raises(Exception, match="SHs or precomputed colors"):
        r(means3D=x, means2D=x, opacities=torch.zeros(4, 1), scales=x, rotations=torch.zeros(4, 4))
    with pytest.raises(Exception, match="SHs or precomputed colors"):
        r(means3D=x, means2D=x, opacities=torch.zeros(4, 1), shs=torch.zeros(4, 1, 3), colors_precomp=x, scales=x, rotations=torch.zeros(4, 4))
    with pytest.raises(Exception, match="scale/rotation pair or precomputed 3D covariance"):
        r(means3D=x, means2D=x, opacities=torch.zeros(4, 1), shs=torch.zeros(4, 1, 3))
    with pytest.raises(Exception, match="scale/rotation pair or precomputed 3D covariance"):
        r(means3D=x, means2D=x, opacities=torch.zeros(4, 1), shs=torch.zeros(4, 1, 3), scales=x, rotations=torch.zeros(4, 4),
          cov3D_precomp=torch.zeros(4, 6))
    # no CPU path: CPU tensors are rejected loudly instead of silently falling back
    with pytest.raises(RuntimeError, match="CUDA"):
        r(means3D=x, means2D=x, opacities=torch.zeros(4, 1), shs=torch.zeros(4, 1, 3), scales=x, rotations=torch.zeros(4, 4))
    with pytest.raises(RuntimeError, match="num_points, 3"):
        r(means3D=torch.zeros(4, 2), means2D=x, opacities=torch.zeros(4, 1), shs=torch.zeros(4, 1, 3), scales=x, rotations=torch.zeros(4, 4))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "lightgaussian_b200")
    bad = []
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                if re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M) or "oracle/" in text and f.endswith(".py") and "import" in text and re.search(r"oracle[/.]lgo", text):
                    bad.append(os.path.join(dirpath, f))
    assert not bad, bad
    for d in ("dropin",):
        for dirpath, _, files in os.walk(os.path.join(ROOT, d)):
            for f in files:
                if f.endswith(".py"):
                    assert "oracle" not in open(os.path.join(dirpath, f)).read()


def test_build_hash_covers_every_kernel_header():
    """a header missing from the staleness hash once let an edited kernel run as its previous binary"""
    from lightgaussian_b200 import build
    on_disk = {f for f in os.listdir(build.CSRC) if f.endswith((".cuh", ".h"))}
    assert on_disk and on_disk <= set(build.HEADERS), on_disk - set(build.HEADERS)
    for inc in re.findall(r'#include "([^"]+)"', open(os.path.join(build.CSRC, "lgrast.cu")).read()):
        if not inc.startswith(".."):
            assert inc in build.HEADERS, inc


def test_bench_cli_parses():
    """bench.py is the driver's contract: a syntax error or a broken option must show up on the CPU box"""
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--help"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    for opt in ("--gpus", "--steps", "--warmup", "--impl", "--mode", "--bin-mode", "--kback-mode"):
        assert opt in r.stdout
