import sys; sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np, torch
from test_gpu_engine import make_array, run_pipeline
N=64
a = make_array(100+N, 2, [1], 1, N)
a["nsample"][:]=40; a["int_time"][:]=10.0
pipe,res = run_pipeline(a,[(0,N-1)],"complex64",None)
ws=pipe._ws.cpu().numpy()
off=256+256
wl=ws[off:off+2*N*8].view(np.float64).reshape(2,N); off+=((2*N*8+255)//256)*256
wr=ws[off:off+2*N*8].view(np.float64).reshape(2,N)
tc=pipe.thermal_coef.cpu().numpy()[0]; fr=pipe.freq_windows[0]; hd=0.5*np.diff(fr)
print("wl err", np.abs(wl[:,:-1]-hd*tc[:,:-1]).max()/np.abs(wl).max(), "wr err", np.abs(wr[:,:-1]-hd*tc[:,1:]).max()/np.abs(wr).max(), wl[:,-1], wr[:,-1])
print("sum", (wl+wr).sum(-1)*1e-6/np.sqrt(2)/400, res["thermal_all"][0].ravel())
np.set_printoptions(linewidth=200, precision=4)
print(wl[0,28:36]); print((hd*tc[0,:-1])[28:36]); print(wr[0,28:36]); print((hd*tc[0,1:])[28:36])
print(wl[1,28:36]); 
