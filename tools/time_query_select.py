"""Times destr_select_queries against the same tail of MiniDetector.forward written with torch ops on the GPU."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from object_detection_destr_b200 import query_select as QS
for B,N,C,k in ((8,1050,91,100),(2,4200,91,300),(64,1050,91,300)):
    g=torch.Generator().manual_seed(0)
    sc=torch.rand(B,N,C,generator=g).cuda(); mask=torch.zeros(B,N,dtype=torch.bool).cuda()
    cf=torch.randn(B,N,256,generator=g).cuda(); rf=torch.randn(B,N,256,generator=g).cuda(); co=torch.rand(B,N,4,generator=g).cuda()
    for _ in range(3): QS.select_queries(sc,mask,cf,rf,co,top_k=k,valid0=N,want_bf16=True)
    torch.cuda.synchronize()
    s,e=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(20): QS.select_queries(sc,mask,cf,rf,co,top_k=k,valid0=N,want_bf16=True)
    e.record(); e.synchronize()
    t_ours=s.elapsed_time(e)/20
    # torch restatement of the reference tail (GPU): sigmoid, max, topk, gathers
    def ref():
        key=sc.sigmoid().max(-1).values
        _,idx=torch.topk(key,k=k,dim=1)
        bi=torch.arange(B,device='cuda').repeat_interleave(k)
        feats=torch.concat([cf,rf],-1)
        return feats[(bi,idx.flatten())].reshape(B,k,-1), co[...,:2][(bi,idx.flatten())]
    for _ in range(3): ref()
    torch.cuda.synchronize(); s.record()
    for _ in range(20): ref()
    e.record(); e.synchronize()
    print(f"B={B} N={N} k={k}: kernel {t_ours*1000:.1f} us   torch ops on the GPU {s.elapsed_time(e)/20*1000:.1f} us")
