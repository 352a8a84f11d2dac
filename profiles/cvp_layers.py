import sys, os, torch
sys.path.insert(0, os.getcwd())
from wild_deep_mvs_b200 import ops
dev="cuda:0"
torch.manual_seed(0)
def t(fn, reps=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b)/reps
for cin,cout,eng in ((32,64,"zm"),(64,64,"zm"),(64,64,"tc"),(64,32,"zm"),(64,32,"tc")):
    dims=(1,4,592,800) if cin>16 else (1,8,1184,1600)
    x=torch.randn(*dims,cin,device=dev)
    w=torch.randn(cout,cin,3,3,3,device=dev)/(cin*27)**0.5
    layer=ops.PackedConv(w,None,relu=True)
    y=ops.conv3d(x,layer,engine=eng)
    ms=t(lambda: ops.conv3d(x,layer,engine=eng,out=y))
    vox=dims[1]*dims[2]*dims[3]
    print("%d->%d %s on %s: %.3f ms  (%.1f TFLOP/s fp32-equivalent)"%(cin,cout,eng,dims,ms, vox*cin*cout*54/ms/1e9))
