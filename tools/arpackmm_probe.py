"""Quick sweep of arpackmm_b200's solver options on small generated problems (debug aid; the real tests are in
tests/test_gpu_arpackmm.py).  Prints one line per command: exit code, mode / found / iterations, last error line."""
import os, re, subprocess, sys, tempfile, time
import numpy as np, scipy.sparse as sp
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
from test_mmio import write_mtx
from test_gpu_arpackmm import write_mtx_complex, EXE

d = tempfile.mkdtemp()
n = 60
As = sp.diags([-np.ones(n - 1), 2.0 + 0.05 * np.arange(n), -np.ones(n - 1)], [-1, 0, 1]).tocsr()
An = sp.diags([-1.3 * np.ones(n - 1), 2.0 + 0.03 * np.arange(n), -0.4 * np.ones(n - 1), 0.2 * np.ones(n - 5)], [-1, 0, 1, 5]).tocsr()
Bm = sp.diags([np.ones(n - 1) / 6, 4 * np.ones(n) / 6, np.ones(n - 1) / 6], [-1, 0, 1]).tocsr()
k = np.arange(48)
Az = sp.diags([(-1.0 - 0.1j) * np.ones(47), (2.0 + 0.05 * k) + 0.3j * np.cos(k), (-1.0 + 0.3j) * np.ones(47)], [-1, 0, 1]).tocsr()
Bz = sp.diags([np.ones(47) / 6, 4 * np.ones(48) / 6, np.ones(47) / 6], [-1, 0, 1]).astype(complex).tocsr()
write_mtx(os.path.join(d, "As.mtx"), As, base=1); write_mtx(os.path.join(d, "An.mtx"), An, base=0, banner=False, with_nnz=False)
write_mtx(os.path.join(d, "B.mtx"), Bm, base=1)
write_mtx_complex(os.path.join(d, "Az.mtx"), Az, base=0); write_mtx_complex(os.path.join(d, "Bz.mtx"), Bz, base=1)
S = "--A As.mtx --genPb --nbEV 3 --nbCV 20 --maxIt 500 --verbose 1"
Z = "--nonSymPb --cpxPb --A Az.mtx --nbEV 3 --nbCV 20 --maxIt 1000 --verbose 1"
cmds = [S + " --slv BiCG --slvItrTol 1.e-12 --slvItrMaxIt 2000", S + " --slv CG --slvItrTol 1.e-12 --slvItrMaxIt 2000",
        S + " --slv BiCG --slvItrPC ILU#1.e-06#2 --slvItrTol 1.e-12", S + " --slv CG --slvItrPC ILU", S + " --slv LU", S + " --slv QR",
        S + " --slv LLT", S + " --slv LDLT", S + " --slv LDLT --shiftReal 1.0", S + " --slv LLT --shiftReal 1.0", S + " --slv LU --shiftReal 1.0",
        S + " --slv LU --simplePrec", S + " --slv LLT --simplePrec",
        S + " --slv LU --dense true", S + " --slv QR --dense false", S + " --slv LLT --dense true", S + " --slv LDLT --dense true --shiftReal 1.0",
        S + " --slv QR --dense true --simplePrec",
        "--nonSymPb --A An.mtx --genPb --nbEV 4 --nbCV 24 --shiftReal 2.0 --maxIt 1000 --slv LU --verbose 1",
        Z, Z + " --simplePrec", Z + " --shiftReal 3.0", Z + " --B Bz.mtx --genPb --slv BiCG --slvItrTol 1.e-12 --slvItrMaxIt 2000",
        Z + " --B Bz.mtx --genPb --slv LU --shiftReal 6.0 --shiftImag 1.0", Z + " --B Bz.mtx --genPb --slv QR --dense true",
        Z + " --B Bz.mtx --genPb --slv LLT", Z + " --B Bz.mtx --genPb --slv LLT --restart",
        Z + " --B Bz.mtx --genPb --slv BiCG --slvItrPC ILU#1.e-08#4 --slvItrTol 1.e-12 --simplePrec"]
only = os.environ.get('PROBE_ONLY')
for c in cmds:
    if only and only not in c:
        continue
    t0 = time.time()
    p = subprocess.run([EXE] + c.split(), cwd=d, capture_output=True, text=True, timeout=120)
    m = re.search(r"OUT: mode (\d+), nb EV found (\d+), nb iterations (\d+)", p.stdout)
    inner = re.search(r"inner solver ([^\n]*)", p.stdout)
    vals = re.findall(r"Ritz value\s+\d+: \(([-+.\de]+),([-+.\de]+)\)", p.stdout)
    errs = [l for l in p.stderr.splitlines() if l.strip()]
    print(f"rc={p.returncode} {time.time() - t0:4.1f}s | {c[:110]:110s} | {m.groups() if m else None} | {inner.group(1) if inner else ''} | "
          f"{[f'{float(a):.5g}{float(b):+.3g}j' for a, b in vals]} | {errs[-3:] if p.returncode else errs[-1:]}", flush=True)
