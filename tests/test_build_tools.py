"""CPU checks of the build-time PTX rewrite (tools/ptx_brx.py) and of what the built library contains (SASS mnemonics)."""
import os
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import ptx_brx  # noqa: E402

PTX = """
.visible .entry kern(
	.param .u64 p
)
{
	.reg .pred %p<9>;
	.reg .b32 %r<9>;
	and.b32  	%r7, %r8, 255;
	// begin inline asm
	// BT_DISPATCH %r7;
	// end inline asm
	setp.gt.s32 	%p1, %r7, 1;
	@%p1 bra 	$L__BB0_3;
	setp.eq.s32 	%p2, %r7, 0;
	@%p2 bra 	$L__BB0_1;
	bra.uni 	$L__BB0_2;
$L__BB0_1:
	// begin inline asm
	// BT_CASE 0;
	// end inline asm
	add.s32 %r1, %r1, 1;
	bra.uni 	$L__BB0_9;
$L__BB0_2:
	// begin inline asm
	// BT_CASE 1;
	// end inline asm
	add.s32 %r1, %r1, 2;
	bra.uni 	$L__BB0_9;
$L__BB0_3:
	setp.eq.s32 	%p3, %r7, 3;
	@%p3 bra 	$L__BB0_4;
	bra.uni 	$L__BB0_9;
$L__BB0_4:
	// begin inline asm
	// BT_CASE 3;
	// end inline asm
	add.s32 %r1, %r1, 3;
$L__BB0_9:
	// begin inline asm
	// BT_CASE 255;
	// end inline asm
	ret;
}
"""


def test_switch_becomes_one_indirect_branch():
    out = ptx_brx.patch(PTX)
    assert "brx.idx %bt_idx, $BT_TBL_0;" in out
    line = [l for l in out.splitlines() if ".branchtargets" in l][0]
    # sites 0, 1, hole (2 -> default), 3, and the clamp entry for indices past the table
    assert line.split(".branchtargets")[1].strip() == "$L__BB0_1, $L__BB0_2, $L__BB0_9, $L__BB0_4, $L__BB0_9;"
    assert "min.u32 %bt_idx, %r7, 4;" in out
    # the compare tree is left in place (unreachable after the indirect branch; ptxas drops it)
    assert out.count("setp") == PTX.count("setp")


def test_rewrite_refuses_unsafe_trees():
    # an instruction with side effects inside the compare tree: the function must be left untouched
    bad = PTX.replace("	setp.eq.s32 	%p3, %r7, 3;", "	st.global.u32 [%r2], %r3;\n	setp.eq.s32 	%p3, %r7, 3;")
    assert ptx_brx.patch(bad) == bad
    # duplicated dispatch marker (the compiler cloned the loop): untouched
    dup = PTX.replace("	ret;", "	// BT_DISPATCH %r7;\n	ret;")
    assert ptx_brx.patch(dup) == dup
    # no default marker: untouched
    nodef = PTX.replace("// BT_CASE 255;", "// nothing")
    assert ptx_brx.patch(nodef) == nodef


def test_library_holds_tensor_copies_and_indirect_branch():
    """the fused tile kernel must contain the TMA tensor load/store (UTMALDG / UTMASTG), the mbarrier wait and the BRX dispatch"""
    lib = os.path.join(ROOT, "bluetangle.jl_b200", "lib", "libbluetangle_cuda.so")
    if not os.path.exists(lib) or shutil.which("cuobjdump") is None:
        pytest.skip("library or cuobjdump not present")
    sass = subprocess.run(["cuobjdump", "-sass", "-fun", "_Z10k_tile_tmaILb0EEv14CUtensorMap_st10TileParams", lib], capture_output=True, text=True).stdout
    assert "UTMALDG.5D" in sass and "UTMASTG.5D" in sass and "SYNCS.PHASECHK" in sass
    if os.environ.get("BT_NO_BRX"):
        return
    assert "BRX" in sass, "micro-op dispatch was not rewritten to brx.idx (tools/ptx_brx.py)"


def test_pass_specialiser_generates_and_compiles_on_the_host():
    """bt_jit_selftest: a synthetic pass with every micro-op site kind is turned into CUDA text and compiled by NVRTC for
    sm_100a -- no device needed.  Skipped when libnvrtc is not installed (the library then simply keeps interpreting)."""
    import ctypes as C
    sys.path.insert(0, ROOT)
    import __graft_entry__ as ge
    bt = ge.load_package()
    lib = bt._lib.load()
    buf = C.create_string_buffer(1 << 20)
    rc = lib.bt_jit_selftest(buf, 1 << 20)
    if rc == -2:
        pytest.skip("libnvrtc not available")
    assert rc == 0, lib.bt_last_error().decode()
    src = buf.value.decode()
    for needle in ("bt_jit_pass", "cp.async.bulk.tensor.5d", "mbarrier.try_wait", "fma(C.c[", "(base & 0x100000ull) == 0x100000ull", "make_double2("):  # the condition is a branch or a select (BT_JIT_VARIANT)
        assert needle in src
    # the CX of the synthetic pass is a renaming: amplitudes are stored from permuted variables, no swap code is emitted
    stores = [l for l in src.splitlines() if "make_double2(" in l]
    assert len(stores) == 16 and any("xr[5]" in l for l in stores) and "xr[1] = xr[5]" not in src


def test_pass_specialiser_disk_cache_round_trip(tmp_path, monkeypatch):
    """On-disk cubin cache (BT_JIT_CACHE_DIR; default under ~/.cache, "" = off): the cubin of a compiled pass is written under a
    content-hashed name (temporary file + rename) and read back byte for byte; with the cache off nothing is written; with the
    variable unset the default directory under $XDG_CACHE_HOME is used."""
    import ctypes as C
    sys.path.insert(0, ROOT)
    import __graft_entry__ as ge
    bt = ge.load_package()
    lib = bt._lib.load()
    monkeypatch.setenv("BT_JIT_CACHE_DIR", "")
    monkeypatch.setenv("XDG_CACHE_HOME", str(tmp_path / "xdg"))
    rc = lib.bt_jit_selftest(None, 0)
    if rc == -2:
        pytest.skip("libnvrtc not available")
    assert rc == 0
    assert list(tmp_path.iterdir()) == []
    monkeypatch.delenv("BT_JIT_CACHE_DIR", raising=False)
    assert lib.bt_jit_selftest(None, 0) == 0
    d = C.create_string_buffer(512)
    hits = C.c_uint64()
    assert lib.bt_jit_cache_info(C.byref(hits), d, 512) == 0
    assert d.value.decode() == str(tmp_path / "xdg" / "bluetangle_cuda")
    assert len(list((tmp_path / "xdg" / "bluetangle_cuda").iterdir())) == 1
    import shutil
    shutil.rmtree(tmp_path / "xdg")
    monkeypatch.setenv("BT_JIT_CACHE_DIR", str(tmp_path))
    assert lib.bt_jit_selftest(None, 0) == 0
    files = list(tmp_path.iterdir())
    assert len(files) == 1 and files[0].name.startswith("btjit_") and files[0].suffix == ".cubin"
    assert files[0].read_bytes()[:4] == b"\x7fELF" and files[0].stat().st_size > 10000
    # a truncated / foreign file is not accepted (the library then recompiles and rewrites it)
    files[0].write_bytes(b"garbage")
    assert lib.bt_jit_selftest(None, 0) == 0
    assert files[0].read_bytes()[:4] == b"\x7fELF"
    # an unwritable directory: the self test reports that nothing came back (-5); the product path just compiles as usual
    monkeypatch.setenv("BT_JIT_CACHE_DIR", "/proc/nonexistent-dir")
    assert lib.bt_jit_selftest(None, 0) == -5


def test_compile_workers_finish_their_jobs_and_the_process_still_exits(tmp_path):
    """The pass specialiser's worker threads (csrc/bt_jit.cu) are detached and wait on a condition variable for the life of the
    process.  A process that has used them must exit normally: a static std::condition_variable with waiters blocks forever in its
    destructor at exit (seen on a 4-GPU box: every rank hung after its last line of output)."""
    import subprocess
    code = ("import sys; sys.path.insert(0, %r)\n"
            "import __graft_entry__ as ge\n"
            "lib = ge.load_package()._lib.load()\n"
            "rc = lib.bt_jit_selftest_workers(6)\n"
            "print('workers rc', rc, flush=True)\n"
            "sys.exit(0 if rc in (0, -2) else 1)\n") % ROOT
    env = dict(os.environ, BT_JIT_CACHE_DIR=str(tmp_path), BT_JIT_THREADS="3")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=240, env=env)
    assert r.returncode == 0, (r.stdout, r.stderr)
    if "workers rc 0" in r.stdout:
        assert len([f for f in os.listdir(tmp_path) if f.endswith(".cubin")]) == 6
