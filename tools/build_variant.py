import sys, os, subprocess
sys.path.insert(0,'/root/repo')
from lc_b200 import _native as nat
tag, defs = sys.argv[1], sys.argv[2:]
out = os.path.join(nat.BUILD_DIR, f"liblc_b200_{tag}.so")
objs, procs = [], []
for src in nat.SOURCES:
    obj = os.path.join(nat.BUILD_DIR, os.path.basename(src)[:-3] + f".{tag}.o"); objs.append(obj)
    procs.append(subprocess.Popen(["nvcc"] + nat.NVCC_FLAGS + defs + ["-c", "-o", obj, src], cwd='/root/repo'))
assert all(p.wait() == 0 for p in procs)
subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", out] + objs, check=True)
print(out)
