"""The C-ABI shared library: loads without a GPU, exports every symbol include/iamrx.h
declares, refuses to compute without a device (no CPU fallback), and the product package
never touches the oracle."""
import ctypes as C
import os
import re
import subprocess

import pytest

import iamr_b200 as ix

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    hdr = open(os.path.join(ROOT, "include", "iamrx.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b(iamrx_[a-z0-9_]+)\s*\(", hdr))
    return {n for n in names if not n.endswith("_fn")}


def test_library_exports_every_declared_symbol():
    if not os.path.exists(ix.lib_path()):
        subprocess.check_call(["make", "-s", "-j8", "-C", ROOT])
    lib = ix.load()
    names = _declared()
    assert len(names) >= 40
    for n in names:
        assert hasattr(lib.dll, n), f"{n} declared in include/iamrx.h but not exported"
    # and the binding knows all of them
    assert names <= set(ix.binding.SIGNATURES)
    assert lib.iamrx_version() >= 100


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = ix.load()
    assert lib.iamrx_device_ok() == 0
    g = ix.Geom.make((8, 8, 8))
    lev = ix.Level(lib, g, [((0, 0, 0), (7, 7, 7))])  # pure host object: allowed
    p = ix.NSParams()
    lib.iamrx_ns_params_default(C.byref(p))
    h = C.c_void_p()
    rc = lib.iamrx_ns_create(lev.h, C.byref(p), C.byref(h))
    assert rc == -3 and b"no CPU fallback" in lib.iamrx_last_error()
    rc = lib.iamrx_fill_boundary(lev.h, None, 0, 1, 1, None)
    assert rc == -3
    lev.close()


def test_missing_library_fails_loudly(tmp_path):
    with pytest.raises(FileNotFoundError):
        ix.load(str(tmp_path / "libiamrx.so"))


def test_product_never_references_oracle():
    pkg = os.path.join(ROOT, "iamr_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".cpp")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert "liboracle" not in txt and "import orc" not in txt and "oracle/_build" not in txt, f
    out = subprocess.run(["ldd", ix.lib_path()], capture_output=True, text=True).stdout
    assert "oracle" not in out and "emul" not in out


def test_level_argument_checks():
    lib = ix.load()
    g = ix.Geom.make((8, 8, 8))
    with pytest.raises(ix.IamrxError):
        ix.Level(lib, g, [((0, 0, 0), (7, 7, 7)), ((4, 4, 4), (7, 7, 7))])  # overlap
    with pytest.raises(ix.IamrxError):
        ix.Level(lib, g, [((0, 0, 0), (8, 7, 7))])  # outside the domain
    with pytest.raises(ix.IamrxError):
        ix.Level(lib, g, [((0, 0, 0), (7, 7, 7))], owners=[3])  # rank out of range


def _build_abi_check():
    exe = os.path.join(ROOT, "tests", "abi", "_build", "abi_check")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.check_call([cxx, "-std=c++17", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "abi", "abi_check.cpp"), "-o", exe,
                           "-L", os.path.join(ROOT, "iamr_b200"), "-liamrx",
                           "-Wl,-rpath," + os.path.join(ROOT, "iamr_b200"), "-Wl,-rpath,/usr/local/cuda/lib64",
                           "-Wl,--allow-shlib-undefined"])
    return exe


def test_cxx_program_links_against_the_header_and_library():
    """include/iamrx.h as a C++ consumer sees it (not ctypes): compiled with -Wall -Werror, linked with libiamrx.so;
    host-side entry points work, compute entries fail loudly without a device."""
    exe = _build_abi_check()
    out = subprocess.run([exe, "host"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "abi_check host ok" in out.stdout


@pytest.mark.gpu
def test_cxx_program_runs_a_step_on_the_device():
    exe = _build_abi_check()
    out = subprocess.run([exe, "device"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "abi_check device ok" in out.stdout


def test_header_compiles_as_c():
    """The ABI header is plain C (extern "C" guards only under __cplusplus)."""
    src = os.path.join(ROOT, "tests", "abi", "_build", "hdr.c")
    os.makedirs(os.path.dirname(src), exist_ok=True)
    open(src, "w").write('#include "iamrx.h"\nint main(void) { iamrx_bcrec b; b.lo[0] = IAMRX_BC_EXT_DIR; return b.lo[0] == 3 ? 0 : 1; }\n')
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"), "-c", src,
                           "-o", src + ".o"])


def test_integration_snippets_compile():
    """INTEGRATION.md's call-site replacements are excerpts of tests/abi/adapter_sites.cpp, which compiles (-Wall -Werror)
    against include/iamrx.h + include/IamrxAdapter.H and a stand-in with the shape of the AMReX classes, and links with the
    library: the documented binding cannot drift from the ABI."""
    import re
    abi = os.path.join(ROOT, "tests", "abi")
    exe = os.path.join(abi, "_build", "adapter_sites")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    subprocess.check_call(["g++", "-std=c++17", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
                           "-I", os.path.join(abi, "amrex_stub"), os.path.join(abi, "adapter_sites.cpp"), "-o", exe,
                           "-L", os.path.join(ROOT, "iamr_b200"), "-liamrx", "-Wl,-rpath," + os.path.join(ROOT, "iamr_b200")])
    md = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    src = open(os.path.join(abi, "adapter_sites.cpp")).read()
    blocks = re.findall(r"// \[site (\d) begin\]\n(.*?)// \[site \1 end\]", src, re.S)
    assert len(blocks) == 6
    for n, body in blocks:
        assert body in md, f"INTEGRATION.md section {n} differs from tests/abi/adapter_sites.cpp"
    adapter = open(os.path.join(ROOT, "include", "IamrxAdapter.H")).read()
    ns = re.search(r"(namespace iamrx_adapt \{.*?\n\}  // namespace iamrx_adapt)", adapter, re.S).group(1)
    assert ns in md
