import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, time
from crfconv_b200 import _lib
L=_lib.lib()
n=6*40960*17
idx=torch.randint(0,40960,(n,),dtype=torch.int64).pin_memory()
out=torch.empty(n,dtype=torch.int16).pin_memory()
for th in (1,2,4,8,16):
    ts=[]
    for _ in range(8):
        t=time.perf_counter(); rc=L.crfconv_pack_index_host(idx.data_ptr(),n,16,out.data_ptr(),th); ts.append(time.perf_counter()-t)
    print('pack threads',th, rc, round(min(ts)*1e3,3),'ms', flush=True)
