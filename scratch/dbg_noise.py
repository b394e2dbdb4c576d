import numpy as np, torch, sys
sys.path.insert(0, '.')
from simpleds_b200.engine import RedundantPipeline
N=256
out={}
for name,(S,P,gs,T) in dict(rows=(1,1,[4,7,1],1), times=(1,1,[1],3), pols=(1,2,[1],1), wins=(2,1,[1],1)).items():
    F0=N*S
    freq=100e6+97656.25*np.arange(F0)
    go=np.concatenate([[0],np.cumsum(gs)])
    B=int(go[-1])
    pipe=RedundantPipeline(freq,[(w*N,(w+1)*N-1) for w in range(S)],[-5,-6][:P],go,trcvr_k=144.0,beam_area_sr=0.76,precision="complex64",seed=0x1234567890)
    dev=torch.device("cuda:0")
    vis=torch.zeros((P,B,T,F0),dtype=torch.complex64,device=dev)
    fl=torch.zeros((P,B,T,F0),dtype=torch.uint8,device=dev)
    ns=torch.ones((P,B,T,F0),dtype=torch.float32,device=dev)
    it=torch.ones((B,T),dtype=torch.float32,device=dev)
    pipe.process(vis,fl,ns,it,noise="philox",thermal=False)
    torch.cuda.synchronize()
    out[name+"_auto"]=pipe.noise_auto.cpu().numpy(); out[name+"_sc"]=pipe.sigma_coef32.cpu().numpy()
    out["taper"]=pipe.taper_values; out["df"]=pipe.delta_f
np.savez("gpurun_out/dbg_noise.npz", **out)
print("ok")
