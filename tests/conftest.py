import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # build artefacts are git-ignored: make sure they exist (no-op when up to date or when nvcc is absent
    # and the prebuilt library travelled with the snapshot)
    import shutil

    from dorylus_b200 import build as product_build

    if shutil.which("nvcc") or os.path.exists(product_build.NVCC):
        product_build.build()
    elif not os.path.exists(product_build.LIB):
        raise pytest.UsageError("libdorylus_b200.so is missing and nvcc is not available to build it")


def _cuda_device_visible() -> bool:
    """True when dory_create can find an sm_100 device (asked of the library itself: no torch needed)."""
    try:
        from dorylus_b200 import _lib
        from dorylus_b200.engine import DoryError, Engine

        try:
            Engine([4, 4, 2]).close()
            return True
        except DoryError as ex:
            return ex.code != _lib.ENODEV
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    """`pytest tests` on a box without a GPU: the gpu-marked tests are skipped instead of erroring in their
    fixtures with DORY_ENODEV (the product has no CPU fallback to run them on)."""
    gpu_items = [it for it in items if it.get_closest_marker("gpu")]
    if gpu_items and not _cuda_device_visible():
        skip = pytest.mark.skip(reason="no CUDA device visible (dory_create: DORY_ENODEV)")
        for it in gpu_items:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    from oracle.pyoracle import Oracle

    return Oracle()


@pytest.fixture(scope="session")
def ref():
    from oracle.pyoracle import Ref

    if not Ref.available():
        pytest.skip("oracle/_ref/libdoryref.so not available (no /root/reference and no prebuilt library)")
    return Ref()


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    d = os.path.join(ROOT, "tests", "golden")
    return {n: np.load(os.path.join(d, n + ".npz")) for n in ("xavier", "masks", "reference_runs", "weight_dump", "numpy_gnn", "funcs_ops", "input_tools", "ref_engine")}


@pytest.fixture()
def hostcheck():
    """The product's engine object linked against a host-memory fake of the CUDA runtime and scalar CPU
    statements of the kernel launchers (tests/hostcheck/): `Engine` then runs its HOST logic here, without
    a GPU.  Test infrastructure: the product binding never loads this library by itself."""
    import ctypes as C
    import importlib.util

    from dorylus_b200 import _lib

    spec = importlib.util.spec_from_file_location("hostcheck_build", os.path.join(ROOT, "tests", "hostcheck", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    # DORY_HOSTCHECK_LIB: an alternative build of the same library (e.g. with -fsanitize=address)
    lib = C.CDLL(os.environ.get("DORY_HOSTCHECK_LIB") or mod.build())
    for name, (res, args) in _lib.SYMBOLS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    saved = _lib._lib
    _lib._lib = lib
    try:
        yield lib
    finally:
        _lib._lib = saved


@pytest.fixture()
def launchcheck():
    """Like `hostcheck`, but with the product's own kernel launchers (spmm.cu / dense.cu / gat.cu host
    side) linked in: every launch is checked against the hardware's launch limits, nothing is executed.
    Yields a function returning (launches, violations, first offender) since the last call."""
    import ctypes as C
    import importlib.util

    from dorylus_b200 import _lib

    spec = importlib.util.spec_from_file_location("hostcheck_build", os.path.join(ROOT, "tests", "hostcheck", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    lib = C.CDLL(mod.build_launchcheck())
    for name, (res, args) in _lib.SYMBOLS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args

    def log():
        n, v = C.c_ulonglong(), C.c_ulonglong()
        buf = C.create_string_buffer(512)
        lib.hostcheck_launch_log(C.byref(n), C.byref(v), buf, 512)
        return int(n.value), int(v.value), buf.value.decode()

    saved = _lib._lib
    _lib._lib = lib
    try:
        log()
        yield log
    finally:
        _lib._lib = saved


@pytest.fixture()
def commcheck():
    """Like `hostcheck`, with the product's REAL communicator object (comm.cu) linked in over an NCCL
    whose ranks are threads of this process, pointer-carrying CUDA IPC handles and scalar statements of
    comm.cu's kernels (tests/hostcheck/fake_nccl.cpp)."""
    import ctypes as C
    import importlib.util

    from dorylus_b200 import _lib

    spec = importlib.util.spec_from_file_location("hostcheck_build", os.path.join(ROOT, "tests", "hostcheck", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    lib = C.CDLL(os.environ.get("DORY_COMMCHECK_LIB") or mod.build_commcheck())
    for name, (res, args) in _lib.SYMBOLS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    saved = _lib._lib
    _lib._lib = lib
    try:
        yield lib
    finally:
        _lib._lib = saved
