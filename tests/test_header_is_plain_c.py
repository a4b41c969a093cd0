"""include/amsweep.h must be bindable from cgo: plain C (C99, pedantic), no C++isms,
and a C translation unit must link against libamsweep.so using only that header."""
import os
import subprocess
import tempfile

from conftest import ROOT

PROG = r'''
#include <stdio.h>
#include <string.h>
#include "amsweep.h"
int main(void) {
  am_cron_t c; char err[128];
  if (am_abi_version() != AMSWEEP_ABI_VERSION) return 1;
  if (am_cron_parse("*/15 9-17 * * 1-5", 17, &c, err, sizeof err) != AM_OK) return 2;
  if (c.kind != AM_CRON_SPEC || c.minute != 0x200040008001ull || c.dow != 0x3eull) return 3;
  if (!am_cron_matches(&c, 1789982100ll)) return 4;           /* 2026-09-21 09:15:00 Mon */
  if (am_cron_parse("NOT_A_VALID_CRON", 16, &c, err, sizeof err) != AM_E_PARSE) return 5;
  if (strlen(err) == 0) return 6;
  {
    am_healthcheck_t hc; am_record_t r;
    memset(&hc, 0, sizeof hc);
    hc.has_resource = 1; hc.cron = "@every 5s"; hc.cron_len = 9;
    if (am_healthcheck_classify(&hc, &r) != AM_OK) return 7;
    if ((r.flags & AM_KIND_MASK) != AM_KIND_CRON_EVERY || r.ras != 5) return 8;
  }
  printf("ok\n");
  return 0;
}
'''


def test_header_compiles_as_c99_and_links():
    inc = os.path.join(ROOT, "include")
    libdir = os.path.join(ROOT, "active-monitor_b200", "lib")
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "t.c")
        exe = os.path.join(d, "t")
        open(src, "w").write(PROG)
        subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-I", inc, src, "-L", libdir,
                        "-lamsweep", f"-Wl,-rpath,{libdir}", "-o", exe], check=True)
        out = subprocess.run([exe], capture_output=True, text=True)
        assert out.returncode == 0 and out.stdout.strip() == "ok", (out.returncode, out.stdout, out.stderr)
