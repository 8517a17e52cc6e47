"""CPU: include/tetgs_rast.h is a valid plain-C header (C99, no C++), a C program links against libtetgs_rast.so and
calls it, and the struct layouts the C compiler sees are the ones the ctypes mirrors in _lib.py use — the drop-in
boundary is a C ABI, not a C++ or torch one."""
import ctypes as C
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

PROGRAM = r'''
#include <stddef.h>
#include <stdio.h>
#include <string.h>
#include "tetgs_rast.h"

#define OFF(s, f) printf(#s "." #f " %zu\n", offsetof(s, f))
int main(void) {
  printf("abi %d\n", tgr_abi_version());
  printf("sizeof tgr_params %zu\n", sizeof(tgr_params));
  printf("sizeof tgr_binding %zu\n", sizeof(tgr_binding));
  printf("sizeof tgr_adam_group %zu\n", sizeof(tgr_adam_group));
  OFF(tgr_params, background); OFF(tgr_params, geom_bytes); OFF(tgr_params, out_color); OFF(tgr_params, host_num_rendered);
  OFF(tgr_binding, verts); OFF(tgr_binding, dL_dverts);
  OFF(tgr_adam_group, count); OFF(tgr_adam_group, lr); OFF(tgr_adam_group, period);
  printf("geom_bytes %llu\n", (unsigned long long)tgr_geom_bytes(1000));
  printf("loss_bytes %llu\n", (unsigned long long)tgr_image_loss_bytes(2, 64, 48));
  /* argument validation happens on the host, before any CUDA call: usable without a GPU */
  tgr_adam_group g;
  memset(&g, 0, sizeof g);
  printf("adam_bad_step %d\n", tgr_adam_step(&g, 1, 0, 0.9, 0.999, 1e-15, 1.0f, NULL));
  printf("err %s\n", tgr_last_error());
  printf("loss_bad %d\n", tgr_image_loss(0, 8, 8, NULL, NULL, 0, NULL, 0.8f, 0.0f, 0.2f, NULL, NULL, NULL, 0, NULL));
  printf("cams_none %d\n", tgr_build_cameras(0, NULL, NULL, 1e-4f, 100.0f, NULL, NULL));
  return 0;
}
'''


@pytest.mark.skipif(shutil.which("gcc") is None, reason="no C compiler")
def test_header_is_plain_c_and_layouts_match_the_ctypes_mirrors(tmp_path):
    from youreditableavatar_b200 import _lib
    _lib.lib()
    src = tmp_path / "abi.c"
    src.write_text(PROGRAM)
    exe = str(tmp_path / "abi")
    libdir = os.path.dirname(_lib.LIB_PATH)
    cmd = ["gcc", "-std=c99", "-pedantic", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"), str(src),
           "-o", exe, "-L", libdir, "-ltetgs_rast", "-Wl,-rpath," + libdir, "-Wl,-rpath,/usr/local/cuda/lib64"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    out = dict(line.rsplit(" ", 1) for line in r.stdout.strip().splitlines() if not line.startswith("err "))
    assert int(out["abi"]) == _lib.ABI_VERSION
    for name, cls in (("tgr_params", _lib.TgrParams), ("tgr_binding", _lib.TgrBinding), ("tgr_adam_group", _lib.TgrAdamGroup)):
        assert int(out["sizeof " + name]) == C.sizeof(cls), name
    for key, (cls, field) in {"tgr_params.background": (_lib.TgrParams, "background"),
                              "tgr_params.geom_bytes": (_lib.TgrParams, "geom_bytes"),
                              "tgr_params.out_color": (_lib.TgrParams, "out_color"),
                              "tgr_params.host_num_rendered": (_lib.TgrParams, "host_num_rendered"),
                              "tgr_binding.verts": (_lib.TgrBinding, "verts"),
                              "tgr_binding.dL_dverts": (_lib.TgrBinding, "dL_dverts"),
                              "tgr_adam_group.count": (_lib.TgrAdamGroup, "count"),
                              "tgr_adam_group.lr": (_lib.TgrAdamGroup, "lr"),
                              "tgr_adam_group.period": (_lib.TgrAdamGroup, "period")}.items():
        assert int(out[key]) == getattr(cls, field).offset, key
    L = _lib.lib()
    assert int(out["geom_bytes"]) == L.tgr_geom_bytes(1000) and int(out["loss_bytes"]) == L.tgr_image_loss_bytes(2, 64, 48)
    assert int(out["adam_bad_step"]) == 1 and int(out["loss_bad"]) == 1 and int(out["cams_none"]) == 0
    assert "step must be >= 1" in r.stdout
