import sys, numpy as np, torch, time
sys.path.insert(0,''+__import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__)))+'')
import occuseg_b200.sparseconvnet as scn
from occuseg_b200.sparseconvnet import SCN
from occuseg_b200 import _lib, scenes
def lt(v): return torch.LongTensor([v,v,v])
nsc=int(sys.argv[1]) if len(sys.argv)>1 else 8
prec=sys.argv[2] if len(sys.argv)>2 else 'bf16'
coords,feats=scenes.make_batch("S250k",tuple(range(nsc)))
m=SCN.Metadata_3(); out=torch.empty(0,device='cuda')
SCN.InputLayer_updateOutput(m, lt(4096), torch.from_numpy(coords), torch.from_numpy(feats).cuda(), out, nsc, 4, None)
scn.set_precision(prec); print('precision', prec)
size=4096
def ev(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(True),torch.cuda.Event(True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/reps
for lvl,C in enumerate([64,128,192,256,320,384]):
    N=m.getNActive(lt(size))
    nbr,R=m.submanifoldNeighbourTable(lt(size))
    x=torch.randn(N,C,device='cuda'); g=torch.randn(N,C,device='cuda'); w=torch.randn(27,C,C,device='cuda')*0.05
    y=torch.empty(0,device='cuda'); dx=torch.empty(0,device='cuda'); dw=torch.zeros_like(w)
    tf=ev(lambda: SCN.SubmanifoldConvolution_updateOutput(lt(size),lt(3),m,x,y,w,torch.empty(0),1))
    _lib.profile(True)
    for _ in range(3): SCN.SubmanifoldConvolution_backward(lt(size),lt(3),m,x,dx,g,w,dw,torch.empty(0),1)
    torch.cuda.synchronize(); pr=_lib.profile_read(); _lib.profile(False)
    td=pr['conv_tc']['ms']/3; tw=pr['wgrad_tc']['ms']/3
    fl=2.0*R*C*C
    byt=4.0*(R*C+N*C+R+27*C*C)
    print(f"L{lvl} N={N} R={R} C={C}: fwd {tf:.3f} ms ({fl/tf/1e9:.0f} TF/s, {byt/tf/1e6:.0f} GB/s)  dgrad {td:.3f} ms  wgrad {tw:.3f} ms ({fl/tw/1e9:.0f} TF/s)")
    if lvl<5:
        C2=C+64
        w8=torch.randn(8,C,C2,device='cuda')*0.05; yc=torch.empty(0,device='cuda')
        tc_=ev(lambda: SCN.Convolution_updateOutput(lt(size),lt(size//2),lt(2),lt(2),m,x,yc,w8,torch.empty(0)))
        gc=torch.randn_like(yc); dxc=torch.empty(0,device='cuda'); dw8=torch.zeros_like(w8)
        _lib.profile(True)
        for _ in range(3): SCN.Convolution_backward(lt(size),lt(size//2),lt(2),lt(2),m,x,dxc,gc,w8,dw8,torch.empty(0))
        torch.cuda.synchronize(); pr=_lib.profile_read(); _lib.profile(False)
        print(f"   conv {C}->{C2}: fwd {tc_:.3f} ms  dgrad {pr['conv_tc']['ms']/3:.3f} ms  wgrad {pr['wgrad_tc']['ms']/3:.3f} ms")
        wd=torch.randn(8,C2,C,device='cuda')*0.05; yd=torch.empty(0,device='cuda')
        tdc=ev(lambda: SCN.Deconvolution_updateOutput(lt(size//2),lt(size),lt(2),lt(2),m,yc,yd,wd,torch.empty(0)))
        dxd=torch.empty(0,device='cuda'); dwd=torch.zeros_like(wd)
        _lib.profile(True)
        for _ in range(3): SCN.Deconvolution_backward(lt(size//2),lt(size),lt(2),lt(2),m,yc,dxd,g,wd,dwd,torch.empty(0))
        torch.cuda.synchronize(); pr=_lib.profile_read(); _lib.profile(False)
        print(f"   deconv {C2}->{C}: fwd {tdc:.3f} ms  dgrad {pr['conv_tc']['ms']/3:.3f} ms  wgrad {pr['wgrad_tc']['ms']/3:.3f} ms")
        # BN
        gam=torch.ones(C,device='cuda'); bet=torch.zeros(C,device='cuda'); rm=torch.zeros(C,device='cuda'); rv=torch.ones(C,device='cuda')
        sm=torch.empty(C,device='cuda'); si=torch.empty(C,device='cuda'); yb=torch.empty(0,device='cuda')
        tb=ev(lambda: SCN.BatchNormalization_updateOutput(x,yb,sm,si,rm,rv,gam,bet,1e-4,0.9,True,0.0))
        dxb=torch.empty(0,device='cuda'); dg=torch.zeros(C,device='cuda'); db=torch.zeros(C,device='cuda')
        tbb=ev(lambda: SCN.BatchNormalization_backward(x,dxb,yb,g,sm,si,rm,rv,gam,bet,dg,db,0.0))
        print(f"   bn C={C}: fwd {tb:.3f} ms ({3*4*N*C/tb/1e6:.0f} GB/s)  bwd {tbb:.3f} ms ({5*4*N*C/tbb/1e6:.0f} GB/s)")
    size//=2
