"""Build libznll.so (the C-ABI library with the sm_100a kernels) in-tree with nvcc.

``python -m zfit_b200.build`` or ``zfit_b200.build.build()``.  nvcc cross-compiles without a GPU.
The built library lands next to this file (git-ignored, but it travels to the GPU box).
"""

from __future__ import annotations

import concurrent.futures as cf
import hashlib
import os
import shutil
import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
# ZNLL_LIB_OUT / ZNLL_OBJ_DIR / ZNLL_EXTRA_CFLAGS: tuning experiments only (a variant library next to the product one,
# selected at run time with ZNLL_LIB, e.g. ZNLL_EXTRA_CFLAGS=-DZNLL_EXP_TABLE=1)
OBJ = Path(os.environ.get("ZNLL_OBJ_DIR", HERE / "_build"))
LIB = Path(os.environ.get("ZNLL_LIB_OUT", HERE / "libznll.so"))

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
BPS = os.environ.get("ZNLL_BPS", "2")  # resident CTAs (256 threads) per SM the kernels are tuned for
CFLAGS = ["-std=c++17", "-O3", "-lineinfo", "-Xcompiler", "-fPIC", f"-DZNLL_MIN_BLOCKS={BPS}", f"-DZNLL_BLOCKS_PER_SM={BPS}",
          *os.environ.get("ZNLL_EXTRA_CFLAGS", "").split()]

SOURCES = ["znll_host.cu", "znll_prep.cu", "znll_catalog_a.cu", "znll_catalog_b.cu", "znll_catalog_c.cu", "znll_migrad.cpp"]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    msg = "nvcc not found"
    raise RuntimeError(msg)


def _embed_device_source() -> Path:
    """The NVRTC path compiles the very same header at run time: embed it as a string."""
    text = (CSRC / "znll_device.cuh").read_text()
    out = OBJ / "znll_device_src.cpp"
    body = 'extern "C" __attribute__((visibility("default"))) const char znll_device_src[] = R"ZNLLSRC(\n' + text + '\n)ZNLLSRC";\n'
    if not out.exists() or out.read_text() != body:
        out.write_text(body)
    return out


def _digest(paths) -> str:
    h = hashlib.sha256()
    for p in sorted(paths):
        h.update(Path(p).read_bytes())
    h.update(" ".join(ARCH + CFLAGS).encode())
    return h.hexdigest()


def build(verbose: bool = False, force: bool = False) -> Path:
    OBJ.mkdir(parents=True, exist_ok=True)
    deps = [CSRC / f for f in os.listdir(CSRC) if (CSRC / f).is_file()] + [HERE.parent / "include" / "znll.h", Path(__file__)]
    stamp = OBJ / "stamp"
    digest = _digest(deps)
    if not force and LIB.exists() and stamp.exists() and stamp.read_text() == digest:
        return LIB
    nvcc = _nvcc()
    embed = _embed_device_source()

    def compile_one(src: Path) -> Path:
        obj = OBJ / (src.stem + ".o")
        cmd = [nvcc, *ARCH, *CFLAGS, "-c", str(src), "-o", str(obj)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or r.returncode:
            sys.stderr.write(r.stdout + r.stderr)
        if r.returncode:
            msg = f"nvcc failed on {src.name}"
            raise RuntimeError(msg)
        return obj

    srcs = [CSRC / s for s in SOURCES] + [embed]
    with cf.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(compile_one, srcs))
    cmd = [nvcc, *ARCH, "-shared", "-cudart", "static", "-o", str(LIB), *map(str, objs), "-ldl", "-lpthread"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode:
        sys.stderr.write(r.stdout + r.stderr)
        msg = "link of libznll.so failed"
        raise RuntimeError(msg)
    stamp.write_text(digest)
    return LIB


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv, force="-f" in sys.argv))
