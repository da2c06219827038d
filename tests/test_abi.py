"""CPU tests of the C-ABI boundary: the library builds, loads without a GPU, exports exactly what
include/c2m_warp.h declares, and validates arguments before touching the device."""
import ctypes
import os
import re
import subprocess

import pytest

from c2m_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "c2m_warp.h")


def _declared_symbols():
    src = open(HEADER).read()
    return sorted(set(re.findall(r"C2M_API\s+[\w\s\*]+?\b(c2m_\w+)\s*\(", src)))


def test_header_and_binding_agree():
    assert _declared_symbols() == sorted(_lib.SYMBOLS)


def test_library_loads_and_exports_every_symbol():
    lib = _lib.load()
    for name in _declared_symbols():
        assert getattr(lib, name) is not None
    assert lib.c2m_warp_version() == 210


def test_dynamic_symbol_table_has_only_the_abi():
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = sorted(l.split()[-1] for l in out.splitlines() if " T " in l)
    assert exported == _declared_symbols()


def test_no_link_time_dependency_on_libcuda():
    out = subprocess.run(["ldd", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "libcuda.so" not in out


def test_sass_is_sm100a_and_uses_tma():
    sass = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in sass or "SM100a" in sass or "sm_100" in sass
    assert "UTMALDG" in sass  # cp.async.bulk.tensor staging of the flow/mask tiles
    assert "REDG" in sass or "RED." in sass or "ATOMG" in sass


def test_flag_constants_match_header():
    src = open(HEADER).read()
    for name in ("DETERMINISTIC", "COORD_GRID", "TRUE_DIV", "NO_FMA", "FORCE_GENERIC", "NO_TMA", "BWD_ATOMIC"):
        m = re.search(r"#define\s+C2M_FLAG_%s\s+(0x[0-9a-fA-F]+)" % name, src)
        assert m and int(m.group(1), 16) == getattr(_lib, "FLAG_" + name)
    assert int(re.search(r"#define\s+C2M_PAD_ZEROS\s+(\d+)", src).group(1)) == _lib.PAD_ZEROS


def test_argument_validation_happens_before_any_device_work():
    s = _lib.strides4((1, 1, 1, 1))
    lib = _lib.load()
    # negative size
    rc = lib.c2m_warp_blend_fwd(None, None, None, None, None, -1, 1, 1, 1, 0, s, s, 0, 0, None)
    assert rc == 1 and b"invalid sizes" in lib.c2m_warp_last_error()
    # bad padding
    rc = lib.c2m_warp_blend_fwd(None, None, None, None, None, 1, 1, 1, 1, 0, s, s, 7, 0, None)
    assert rc == 1 and b"padding" in lib.c2m_warp_last_error()
    # x_batch must divide N
    rc = lib.c2m_warp_blend_fwd(None, None, None, None, None, 5, 1, 1, 1, 2, s, s, 0, 0, None)
    assert rc == 1 and b"x_batch" in lib.c2m_warp_last_error()
    # null pointers with a non-empty problem
    rc = lib.c2m_warp_blend_fwd(None, None, None, None, None, 1, 1, 2, 2, 0, s, s, 0, 0, None)
    assert rc == 1 and b"null" in lib.c2m_warp_last_error()
    rc = lib.c2m_warp_blend_bwd(None, None, None, None, None, None, None, None, None, 1, 1, 2, 2, 0, s, s, 0, 0,
                                None, 0, None)
    assert rc == 1
    # empty problems are a no-op, not an error (reference: empty tensors pass through grid_sample)
    assert lib.c2m_warp_blend_fwd(None, None, None, None, None, 0, 3, 4, 4, 0, s, s, 0, 0, None) == 0
    assert lib.c2m_warp_blend_bwd(None, None, None, None, None, None, None, None, None, 0, 3, 4, 4, 0, s, s, 0, 0,
                                  None, 0, None) == 0
    with pytest.raises(_lib.C2MWarpError):
        _lib.warp_blend_fwd(None, None, None, None, None, 1, 1, 2, 2, 0, (1, 1, 1, 1), (1, 1, 1, 1), 0, 0, None)


def test_workspace_size_contract():
    # default (gather-form) backward: the larger of the two schemes (the call is not told the layout).
    #  NCHW global lists: per destination pixel a 4 B counter (+ the overflow-list length and one per-frame
    #  "tiles binned" counter behind the counters) and 8 in-line (src, w) entries; per output pixel a 1 B
    #  overflow flag and a 4 B overflow-list slot; each block rounded up to 256 B.
    #  channels-last local binning: tile counters (candidates, registrations, fill cursors), overflow flags, candidate segments (8 B each, 96 per tile and
    #  source frame), overflow list, 16 B pixel records; levels with fewer than 1184 tiles (8 per SM) and at least
    #  128 channels are channel-sliced (2, 4 or 8 slices of >= 64 channels): 1 + slices flag arrays / list segments
    #  and one grad-flow / grad-mask partial-sum buffer per slice.
    up = lambda v: (v + 255) // 256 * 256  # noqa: E731

    def global_lists(N, H, W, B):
        npd, npo = B * H * W, N * H * W
        return up(4 * (npd + 1 + 4)) + up(64 * npd) + up(npo) + up(4 * npo)

    def local(N, C, H, W, B, det=False):
        tiles_per = ((H + 7) // 8) * ((W + 31) // 32)
        ntile, npo, npd = B * tiles_per, N * H * W, B * H * W
        cap = min(256, 96 * (N // B))
        slices, c4 = 1, C // 4
        while (not det and C % 4 == 0 and N * tiles_per * slices < 1184 and slices < 8 and c4 % 2 == 0
               and c4 // 2 >= 16):
            slices, c4 = slices * 2, c4 // 2
        nov = 1 + slices if slices > 1 else 1
        v = up(4 * (3 * ntile + 4)) + nov * up(npo) + up(8 * ntile * cap) + up(4 * nov * npo) + up(16 * npo)
        if det:
            # incoherent flows (counting sort of their pixels by destination tile, deterministic mode): tile offsets,
            # heavy-tile list, the list of incoherent segments, the pool of registered pixels (<= four tiles per pixel)
            v += 2 * up(4 * ntile) + up(4 * N * H * ((W + 31) // 32)) + up(16 * npo)
        if slices > 1:
            v += up(slices * 3 * npo * 4)
        if det:  # + scale bits, touched flags, per-destination corner counts, per-image incoherent-segment counters and the int64 overflow rows
            v += 256 + up(npd) + up(4 * npd) + up(4 * B) + up(8 * npd * C)
        return v

    for (N, C, H, W, B) in [(4, 8, 16, 32, 4), (40, 64, 256, 512, 40), (10, 8, 16, 32, 2), (2, 256, 8, 16, 2),
                            (40, 512, 8, 16, 40), (40, 128, 64, 128, 40)]:
        assert _lib.bwd_workspace_bytes(N, C, H, W, B, True, 0) == 256 + max(global_lists(N, H, W, B),
                                                                            local(N, C, H, W, B))
    assert _lib.bwd_workspace_bytes(4, 8, 16, 32, 4, True, _lib.FLAG_BWD_ATOMIC) == 256
    # deterministic: the larger of (a) one int64 per grad-input element (generic fixed-point scatter) and
    # (b) the channels-last gather's bookkeeping
    det = _lib.bwd_workspace_bytes(4, 8, 16, 32, 4, True, _lib.FLAG_DETERMINISTIC)
    assert det == 256 + max(4 * 8 * 16 * 32 * 8, local(4, 8, 16, 32, 4, True))
    # repeat: gx only has x_batch images
    assert _lib.bwd_workspace_bytes(10, 8, 16, 32, 2, True, _lib.FLAG_DETERMINISTIC) == (
        256 + max(2 * 8 * 16 * 32 * 8, local(10, 8, 16, 32, 2, True)))
    assert _lib.bwd_workspace_bytes(40, 64, 256, 512, 40, True, _lib.FLAG_DETERMINISTIC) == (
        256 + max(40 * 64 * 256 * 512 * 8, local(40, 64, 256, 512, 40, True)))
    assert _lib.bwd_workspace_bytes(2, 256, 8, 16, 2, True, _lib.FLAG_DETERMINISTIC) == (
        256 + max(2 * 256 * 8 * 16 * 8, local(2, 256, 8, 16, 2, True)))
    assert _lib.bwd_workspace_bytes(4, 8, 16, 32, 4, False, _lib.FLAG_DETERMINISTIC) == 256


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.C2MWarpError):
        _lib.load()


def test_header_is_plain_c_and_a_c_program_links(tmp_path):
    """include/c2m_warp.h compiles as C99 (no C++ / torch types cross the boundary) and a plain C program
    linked against libc2m_warp.so can call into it (no GPU needed for the entry points used here)."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "t.c"
    src.write_text(r'''
#include <stdio.h>
#include <string.h>
#include "c2m_warp.h"
int main(void) {
  const int64_t s[4] = {1, 1, 1, 1};
  /* invalid sizes are rejected before anything touches a device */
  int rc = c2m_warp_blend_fwd(NULL, NULL, NULL, NULL, NULL, -1, 1, 1, 1, 0, s, s, C2M_PAD_BORDER, 0, NULL);
  printf("%d %d %zu %d\n", c2m_warp_version(), rc, c2m_occlusion_map_workspace_bytes(2, 3, 5),
         (int)(strlen(c2m_warp_last_error()) > 0));
  return 0;
}
''')
    exe = tmp_path / "t"
    libdir = os.path.join(root, "c2m_b200")
    subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", "-I", os.path.join(root, "include"), str(src), "-o", str(exe),
                    "-L", libdir, "-lc2m_warp", "-Wl,-rpath," + libdir], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()
    assert out == ["210", "1", str(2 * 3 * 5 * 8), "1"]
