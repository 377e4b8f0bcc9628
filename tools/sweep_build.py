"""Dev tool: build tiling variants of libwsocean.so for one tile size (in parallel) into build/variants/.
    python tools/sweep_build.py LOGN [extra -D flags ...]
Variants: K1 (CP,NF), K2 RI, K2h RH around the defaults in wso_kernels.cu."""
import itertools, os, subprocess, sys
from concurrent.futures import ThreadPoolExecutor
logn = int(sys.argv[1]); N = 1 << logn
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEF = {9: (4, 4, 4, 8), 10: (4, 2, 4, 8), 11: (4, 2, 2, 4)}[logn]
def flags(cp, nf, ri, rh):
    return [f"-DWSO_ONLY_LOGN={logn}", f"-DWSO_TUNE_CP{logn}={cp}", f"-DWSO_TUNE_NF{logn}={nf}",
            f"-DWSO_TUNE_RI{logn}={ri}", f"-DWSO_TUNE_RH{logn}={rh}"]
jobs = {}
cp0, nf0, ri0, rh0 = DEF
jobs[f"n{logn}_base"] = flags(*DEF)
for cp, nf in itertools.product((1, 2, 4, 8), (1, 2, 4)):
    T = cp * nf * N // 16
    if 64 <= T <= 1024 and (cp, nf) != (cp0, nf0):
        jobs[f"n{logn}_k1_cp{cp}nf{nf}"] = flags(cp, nf, ri0, rh0)
for ri in (1, 2, 4, 8):
    T = ri * 2 * N // 16
    if 64 <= T <= 1024 and ri != ri0:
        jobs[f"n{logn}_k2_ri{ri}"] = flags(cp0, nf0, ri, rh0)
for rh in (1, 2, 4, 8, 16):
    T = rh * N // 16
    if 64 <= T <= 1024 and rh != rh0:
        jobs[f"n{logn}_kh_rh{rh}"] = flags(cp0, nf0, ri0, rh)
for sk in ("EVOLVE", "FFT1", "STORE1"):
    jobs[f"n{logn}_skip_{sk.lower()}"] = flags(*DEF) + [f"-DWSO_EXP_SKIP_{sk}"]
def run(item):
    name, fl = item
    r = subprocess.run(["bash", os.path.join(ROOT, "tools/tune_build.sh"), name, *fl, *sys.argv[2:]], capture_output=True, text=True)
    return name, r.returncode, r.stderr[-300:]
with ThreadPoolExecutor(8) as ex:
    for name, rc, err in ex.map(run, jobs.items()):
        print(name, "ok" if rc == 0 else "FAIL " + err)
