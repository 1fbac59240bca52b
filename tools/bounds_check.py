"""Run the parity sweep on the -DFFVD_BOUNDS_CHECK build (make check): every hand-computed shared / global index of the
fused kernels is asserted on the device; a violation traps and the call fails.  In-tree replacement for compute-sanitizer,
which is closed on the development pool.
usage (GPU box): python tools/bounds_check.py"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "ffvd_b200", "lib", "libffvd_b200_check.so")
if not os.path.exists(lib):
    sys.exit("build it first: make check")
env = dict(os.environ, FFVD_B200_LIB=lib)
rc = 0
for dl in ("0", "1"):
    env["FFVD_DL"] = dl
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "dev_check.py"), "dev"], env=env, capture_output=True, text=True)
    tail = "\n".join(r.stdout.strip().splitlines()[-10:])
    print("FFVD_DL=%s rc=%d\n%s" % (dl, r.returncode, tail))
    if "FFVD_ASSERT" in r.stdout or "FFVD_ASSERT" in r.stderr or r.returncode != 0:
        print(r.stderr[-2000:])
        rc = 1
sys.exit(rc)
