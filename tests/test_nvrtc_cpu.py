"""The run-time instantiation path compiles: the device source embedded in ``libznll.so`` goes through NVRTC for sm_100a
with the exact name expressions and options ``znll_host.cu::rtc::build`` uses.  NVRTC needs no GPU to compile; loading the
cubin and launching it is covered by the ``-m gpu`` tests (``tests/test_random_trees_gpu.py``)."""

import ctypes as C

import pytest

from zfit_b200 import _znll as Z


def _nvrtc():
    for name in ("libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12"):
        try:
            return C.CDLL(name)
        except OSError:
            continue
    pytest.skip("libnvrtc not found")


def _compile(tree, n_slots, extra=()):
    nv = _nvrtc()
    lib = C.CDLL(str(Z.LIB_PATH))
    src = C.string_at(C.addressof(C.c_char.in_dll(lib, "znll_device_src")))
    prog = C.c_void_p()
    assert nv.nvrtcCreateProgram(C.byref(prog), src + b"\n", b"znll_rtc.cu", 0, None, None) == 0
    import re

    q = re.sub(r"[A-Za-z_][A-Za-z_0-9]*", lambda m: "znll::" + m.group(0), tree)
    names = [f"znll::nll_kernel<{q},{w},{g}>" for w in ("false", "true") for g in ("false", "true")]
    names.append(f"znll::pdf_kernel<{q}>")
    if 2 + n_slots + (n_slots + 1) * (n_slots + 2) // 2 <= 256:
        names += [f"znll::cov_kernel<{q},false>", f"znll::cov_kernel<{q},true>"]
    for n in names:
        assert nv.nvrtcAddNameExpression(prog, n.encode()) == 0
    flags = [b"--gpu-architecture=sm_100a", b"--std=c++17", b"-lineinfo", *extra]
    opts = (C.c_char_p * len(flags))(*flags)
    rc = nv.nvrtcCompileProgram(prog, len(flags), opts)
    size = C.c_size_t()
    nv.nvrtcGetProgramLogSize(prog, C.byref(size))
    log = C.create_string_buffer(size.value + 1)
    nv.nvrtcGetProgramLog(prog, log)
    assert rc == 0, log.value.decode(errors="replace")[-3000:]
    lowered = []
    for n in names:
        low = C.c_char_p()
        assert nv.nvrtcGetLoweredName(prog, n.encode(), C.byref(low)) == 0 and low.value
        lowered.append(low.value)
    nv.nvrtcGetCUBINSize(prog, C.byref(size))
    nv.nvrtcDestroyProgram(C.byref(prog))
    return size.value, lowered


@pytest.mark.parametrize(
    "tree,n_slots",
    [
        ("Sum<Gauss<0>,Gauss<0>,Cheb<0,2>>", 10),  # the `rtc` parity case of tests/_cases.py
        ("Prod<Sum<CBall<0>,Expo<0>>,Sum<Gauss<1>,Legd<1,3>>>", 7 + 8),  # sums under a product: log-native and not
        ("Sum<GCBall<0>,GGETail<0>,TGauss<0>,Bern<0,2>>", 4 + 7 + 5 + 4 + 3),
        ("Sum<Gauss<0>,Herm<0,3>,Lagr<0,2>>", 3 + 2 + 4 + 3),
    ],
)
def test_embedded_source_compiles_with_nvrtc(tree, n_slots):
    cubin_bytes, lowered = _compile(tree, n_slots)
    assert cubin_bytes > 10_000
    assert all(name.startswith(b"_ZN4znll") for name in lowered)


def test_ring_experiment_compiles_with_nvrtc():
    """ZNLL_RING=1 (off by default): the cp.async staging loop for three- and four-stream trees."""
    cubin_bytes, _ = _compile("Prod<Sum<Gauss<0>,Expo<0>>,Gauss<1>,Cheb<2,3>>", 4 + 1 + 2 + 4, extra=(b"-DZNLL_RING=1",))
    assert cubin_bytes > 10_000


def test_compiled_images_are_cached_on_disk(tmp_path, monkeypatch):
    """``rtc::obtain_image`` (what ``znll_plan_set_model`` uses for trees outside the catalogue): first call compiles and
    writes ``$ZNLL_CACHE_DIR/rtc-<key>.cubin``, the second reads it back; a damaged, truncated or foreign file is ignored
    (recompiled and replaced); ``ZNLL_CACHE_DIR=off`` never touches the disk."""
    import time

    _nvrtc()
    tree, n_slots = "Sum<Gauss<0>,Expo<0>,Cheb<0,1>>", 3 + 2 + 1 + 2
    cache = tmp_path / "nested" / "cache"
    monkeypatch.setenv("ZNLL_CACHE_DIR", str(cache))
    t0 = time.perf_counter()
    size, names, cached = Z.test_rtc_image(tree, n_slots)
    t_compile = time.perf_counter() - t0
    assert size > 10_000 and names == 7 and not cached
    files = list(cache.glob("rtc-*.cubin"))
    assert len(files) == 1 and files[0].stat().st_size > size and not list(cache.glob("*.tmp.*"))
    t0 = time.perf_counter()
    assert Z.test_rtc_image(tree, n_slots) == (size, names, True)
    t_hit = time.perf_counter() - t0
    assert t_hit < 0.25 * t_compile, (t_hit, t_compile)
    # a different tree is a different key
    assert Z.test_rtc_image("Sum<Gauss<0>,Expo<0>,Cheb<0,2>>", n_slots + 1)[2] is False
    assert len(list(cache.glob("rtc-*.cubin"))) == 2
    good = files[0].read_bytes()
    for bad in (good[: len(good) // 2], good[:-1] + bytes([good[-1] ^ 1]), b"ZNLLRTC1" + b"\0" * 64, b""):
        files[0].write_bytes(bad)
        assert Z.test_rtc_image(tree, n_slots)[1:] == (names, False)  # not trusted: compiled again ...
        assert Z.test_rtc_image(tree, n_slots)[1:] == (names, True)  # ... and replaced by a valid file
    monkeypatch.setenv("ZNLL_CACHE_DIR", "off")
    assert Z.test_rtc_image(tree, n_slots)[1:] == (names, False)
    # an unwritable cache location is not an error
    monkeypatch.setenv("ZNLL_CACHE_DIR", "/proc/znll-cannot-exist")
    assert Z.test_rtc_image(tree, n_slots)[1:] == (names, False)


def test_flat_trees_outside_the_catalogue_get_the_hessian_kernels(tmp_path, monkeypatch):
    """A flat tree built at run time carries the two ``hess_kernel`` instantiations too (``znll_plan_set_model`` passes the
    tree's count of second-order accumulators); the count is part of the cache key, and a tree that needs more than the
    reduction's 256 accumulators is built without them."""
    _nvrtc()
    monkeypatch.setenv("ZNLL_CACHE_DIR", str(tmp_path))
    tree, n_slots = "Sum<Gauss<0>,Gauss<0>,Cheb<0,2>>", 3 + 2 + 2 + 3
    size, names, cached = Z.test_rtc_image(tree, n_slots, hess_nh=6)
    assert names == 9 and not cached
    plain = Z.test_rtc_image(tree, n_slots)
    assert plain[1:] == (7, False) and plain[0] < size
    assert Z.test_rtc_image(tree, n_slots, hess_nh=6) == (size, 9, True)
    # 2 + (21 + 1)(21 + 2)/2 + 3 = 258 accumulators: neither Hessian nor covariance kernels, the step kernels only
    wide = "Sum<Gauss<0>,Cheb<0,16>>"
    assert Z.test_rtc_image(wide, 2 + 2 + 17, hess_nh=3)[1] == 5


def test_built_step_kernels_fit_two_blocks_per_sm_without_spills():
    """Resource usage of every kernel in the built ``libznll.so`` (``cuobjdump --dump-resource-usage``, no GPU needed): the
    step kernels are tuned for two resident 256-thread blocks per SM, i.e. at most 128 registers,
    and none of them -- nor the pdf / covariance kernels -- may touch local memory (spills) or a stack frame; static shared
    memory has to leave room for two blocks within the 227 KB an SM offers."""
    import re
    import shutil
    import subprocess

    tool = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    out = subprocess.run([tool, "--dump-resource-usage", str(Z.LIB_PATH)], capture_output=True, text=True)
    if out.returncode:
        pytest.skip("cuobjdump not usable: " + out.stderr[-200:])
    rows = re.findall(r"Function (\S+):\s*\n\s*REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", out.stdout)
    kinds = {"nll": [], "pdf": [], "cov": []}
    for name, reg, stack, shared, local in rows:
        for kind in kinds:
            if f"znll{len(kind) + 7}{kind}_kernel" in name:  # _ZN4znll10nll_kernelI...
                kinds[kind].append((name, int(reg), int(stack), int(shared), int(local)))
    assert len(kinds["nll"]) >= 4 * 25 and len(kinds["pdf"]) >= 25 and len(kinds["cov"]) >= 2 * 25  # the whole catalogue was seen
    for kind, found in kinds.items():
        for name, reg, stack, shared, local in found:
            # one 8-byte slot is tolerated (round 2: the weighted Sum<GGETail, Expo> step kernel keeps one double of its
            # rarely taken careful re-evaluation on the stack at the 128-register cap); the four BASELINE kernels have none
            assert stack <= 8 and local == 0, (name, stack, local)
            if any(t in name for t in ("5CBallILi0EEENS_4ExpoILi0", "5GaussILi0EEENS_4PolyILi0ELi4ELi0", "4ProdIJNS_5GaussILi0")):
                assert stack == 0, (name, stack)
            assert shared <= 48 * 1024, (name, shared)
            if kind == "nll":
                assert reg <= 128, (name, reg)
