import time, sys
t0=time.perf_counter()
import torch
torch.cuda.init(); x=torch.zeros(1,device="cuda"); torch.cuda.synchronize()
t1=time.perf_counter()
sys.path.insert(0,"/root/repo")
import nitorch_fastmath_b200 as nfm
from nitorch_fastmath_b200 import _lib
_lib.load()
t2=time.perf_counter()
m=torch.rand(4096,6,device="cuda")+3; v=torch.rand(4096,3,device="cuda")
nfm.sym_solve(m,v); torch.cuda.synchronize()
t3=time.perf_counter()
nfm.sym_invert(m); torch.cuda.synchronize()
t4=time.perf_counter()
print(f"torch+cuda init {t1-t0:.2f}s | import + dlopen {t2-t1:.2f}s | first sym_solve {t3-t2:.3f}s | first sym_invert {t4-t3:.3f}s")
