"""C-ABI checks that need no GPU: the library loads and exports every function include/mmdk.h declares."""
import os
import re

from mmd_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "mmdk.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mmdk_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = _declared()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"libmmdk.so does not export {n}"
    assert sorted(_lib.EXPORTS) == names, "mmd_b200/_lib.py EXPORTS out of sync with include/mmdk.h"


def test_struct_layouts_match_header_field_order():
    src = open(os.path.join(ROOT, "include", "mmdk.h")).read()
    for cname, struct in [("mmdk_step_scalars", _lib.StepScalars), ("mmdk_groups", _lib.Groups),
                          ("mmdk_guide_env", _lib.GuideEnv), ("mmdk_unet_config", _lib.UnetConfig),
                          ("mmdk_chain_desc", _lib.ChainDesc), ("mmdk_peer_exchange", _lib.PeerExchange)]:
        end = src.index("} " + cname + ";")
        body = src[src.rindex("typedef struct {", 0, end) + len("typedef struct {"):end]
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        fields = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            decl = re.sub(r"^(const\s+)?(float|int|void|uint16_t|uint8_t|uint32_t|int64_t|int32_t|mmdk_step_scalars|mmdk_peer_exchange)\s*\*?\s*", "", decl)
            for f in decl.split(","):
                fields.append(re.sub(r"\[.*\]", "", f).strip().lstrip("*").strip())
        assert fields == [f[0] for f in struct._fields_], cname


def test_no_gpu_means_loud_failure():
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_lib.MMDKError):
        _lib.lib()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "mmd_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f


def test_sass_contains_blackwell_native_instructions():
    """The built library must carry tcgen05 / TMA / mbarrier SASS (B200_PROFILING.md: UTCHMMA, LDTM, UBLKCP), i.e. the
    tensor-core executor is really in the binary the tests load."""
    import shutil
    import subprocess
    import pytest
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(exe):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([exe, "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    for mnemonic in ("UTCHMMA", "LDTM", "UBLKCP", "UTCBAR", "SYNCS"):
        assert mnemonic in sass, f"{mnemonic} missing from libmmdk.so"
    assert "HGMMA" not in sass  # no Hopper wgmma
