import sys, os
sys.path.insert(0, '/root/repo')
import arpack_ng_b200 as ab
L = ab.lib()
n = int(sys.argv[1]); j = int(sys.argv[2]); what = int(sys.argv[3]) if len(sys.argv) > 3 else 0
kout = int(sys.argv[4]) if len(sys.argv) > 4 else 0
rc = L.ab200_kernel_probe_f64(n, j, 40, 2, what, kout)
print("probe", n, j, what, kout, "rc", rc)
