import sys, os, time, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
import mkfbodytracker_pdaf_b200 as mk
T,N,K,W=4096,500,50,5
dev=torch.device("cuda",0); torch.cuda.set_device(0)
model=mk.Model.load(mk.LEFT_ARM_MODEL, mk.RIGHT_ARM_MODEL)
st=torch.cuda.current_stream()
b=mk.TrackBatch(model,T,N,0,st.cuda_stream)
F=K+W
meas=torch.empty((F,T,6),dtype=torch.float64,device=dev); ui=torch.empty((F,T),dtype=torch.float64,device=dev); up=torch.empty((F,T),dtype=torch.float64,device=dev)
for f in range(F): b.synth_fill(0x5EED0002,0,f,1,mk.MEAS_SHARED,meas[f],ui[f],up[f])
u0=torch.rand(T,dtype=torch.float64,device=dev)
pose=torch.empty((T,22),dtype=torch.float64,device=dev)
def run(tag, est=True, prof=False, smi=None):
    b.reset(u0)
    for f in range(W):
        b.update(meas[f],ui[f],up[f]); b.estimate_into(None,pose)
    torch.cuda.synchronize()
    p=None
    if smi=="early":
        p=subprocess.Popen(["nvidia-smi","--query-gpu=clocks.sm","--format=csv,noheader","-lms","100"],stdout=subprocess.DEVNULL); time.sleep(1.0)
    if prof: b.profile(K)
    if smi=="late":
        p=subprocess.Popen(["nvidia-smi","--query-gpu=clocks.sm","--format=csv,noheader","-lms","100"],stdout=subprocess.DEVNULL)
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    t0=time.perf_counter(); e0.record()
    for f in range(W,F):
        b.update(meas[f],ui[f],up[f])
        if est: b.estimate_into(None,pose)
    t1=time.perf_counter(); e1.record(); torch.cuda.synchronize(); t2=time.perf_counter()
    if prof: print("   prof",b.profile_read()); b.profile(0)
    if p: p.terminate()
    print(f"{tag:30s} gpu {e0.elapsed_time(e1)/K*1e3:8.1f} us/step  host-issue {(t1-t0)/K*1e6:8.1f} us/step  wall {(t2-t0)/K*1e6:8.1f}")
run("plain")
run("plain again")
run("no estimate", est=False)
run("profile events", prof=True)
run("smi late", smi="late")
run("smi early", smi="early")
run("plain end")
