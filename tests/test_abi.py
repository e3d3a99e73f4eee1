"""CPU: the C-ABI shared library builds, loads, and exports every symbol include/kllms_b200.h declares.
No compute call is made (there is no GPU here)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "kllms_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(kc_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as entry
    from k_llms_b200 import _native
    if not os.path.exists(_native.LIB_PATH):
        entry.build()
    lib = ctypes.CDLL(_native.LIB_PATH)
    names = _declared()
    assert len(names) >= 10
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/kllms_b200.h but not exported"
    assert sorted(_native.EXPORTS) == names
    _native.load()
    assert _native.load().kc_version() == 200


def test_sass_is_blackwell_native():
    """The built library carries sm_100a code with TMA bulk-tensor copies and mbarrier transactions."""
    import shutil
    import subprocess
    from k_llms_b200 import _native
    if shutil.which("cuobjdump") is None or not os.path.exists(_native.LIB_PATH):
        import pytest
        pytest.skip("cuobjdump or library not available")
    out = subprocess.run(["cuobjdump", "-sass", _native.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    assert "UTMALDG" in out and "SYNCS" in out


def test_product_refuses_to_run_without_cuda():
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from k_llms_b200.utils.consensus_utils import ConsensusSettings, consensus_values
    with pytest.raises(RuntimeError, match="no CUDA device"):
        consensus_values(["a", "a", "b"], ConsensusSettings(), lambda t: [[0.0]], None)
