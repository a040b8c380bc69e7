import time, numpy as np, torch, sys
sys.path.insert(0, '/root/repo')
from thermonucleotideblast_b200 import Engine, FragmentList
n = 1_000_000_000
buf = torch.empty(n, dtype=torch.uint8, pin_memory=True).numpy()
buf[:] = 1
eng = Engine()
for nf in (2000, 200, 31):
    L = n // nf
    fl = FragmentList([buf[i*L:(i+1)*L] for i in range(nf)])
    for rep in range(3):
        eng.clear_targets()
        t0 = time.perf_counter(); eng.add_targets(fl); t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
    print(nf, "frags: call %.2f ms, until device idle %.2f ms" % ((t1-t0)*1e3, (t2-t0)*1e3))
