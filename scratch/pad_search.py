import itertools
def ilog2(n): return n.bit_length()-1
def radices(N):
    lg=ilog2(N); out=[]
    while lg>=3: out.append(8); lg-=3
    if lg: out.append(1<<lg)
    return out
def patterns(N):
    """list of (name, [list of logical indices per lane for one warp instruction])"""
    TPR=N//8; W=max(32,TPR); rows=W//TPR
    rad=radices(N); pats=[]
    NS=1
    for P,R in enumerate(rad[:-1]):
        NB=8//R
        for m in range(NB):
            for p in range(R):
                for w0 in range(0,TPR,32) if TPR>=32 else [0]:
                    lanes=[]
                    for lane in range(32):
                        tau=(w0+lane)%TPR if TPR>=32 else lane%TPR
                        sub=0 if TPR>=32 else lane//TPR
                        j=tau+TPR*m
                        o=(j//NS)*NS*R+(j%NS)+p*NS
                        lanes.append((sub,o))
                    pats.append(("W%d"%P,lanes))
        NS*=R
    for r in range(8):
        for w0 in range(0,TPR,32) if TPR>=32 else [0]:
            lanes=[]
            for lane in range(32):
                tau=(w0+lane)%TPR if TPR>=32 else lane%TPR
                sub=0 if TPR>=32 else lane//TPR
                lanes.append((sub,tau+TPR*r))
            pats.append(("R",lanes))
    return pats
def conflicts(N, phi, elem):
    """extra wavefronts summed over patterns; elem = 8 or 16 bytes"""
    group=16 if elem==8 else 8   # lanes per wavefront
    nb=128//elem
    buf=phi(N-1)+1
    buf=(buf+1)//2*2
    tot=0;cnt=0
    for name,lanes in patterns(N):
        for g0 in range(0,32,group):
            banks={}
            for sub,o in lanes[g0:g0+group]:
                a=sub*2*buf+phi(o)
                banks.setdefault(a%nb,set()).add(a)
            tot+=max(len(v) for v in banks.values())-1; cnt+=1
    return tot,cnt
cands={
 'o+(o>>3)':lambda o:o+(o>>3),
 'o+(o>>4)':lambda o:o+(o>>4),
 'o+(o>>3)+(o>>6)':lambda o:o+(o>>3)+(o>>6),
 'o+(o>>4)+(o>>6)':lambda o:o+(o>>4)+(o>>6),
 'o+(o>>4)+(o>>7)':lambda o:o+(o>>4)+(o>>7),
 'o^((o>>4)&7)':lambda o:o^((o>>4)&7),
 'o^((o>>3)&7)':lambda o:o^((o>>3)&7),
 'o^((o>>3)&15)':lambda o:o^((o>>3)&15),
 'o^((o>>4)&15)':lambda o:o^((o>>4)&15),
 'o^((o>>3)&7)^((o>>6)&7)':lambda o:o^((o>>3)&7)^((o>>6)&7),
 'o^(((o>>3)^(o>>6))&7)':lambda o:o^(((o>>3)^(o>>6))&7),
 'o^((o>>4)&7)^(((o>>7)&1)<<3)':lambda o:o^((o>>4)&7)^(((o>>7)&1)<<3),
 'o+(o>>4)+(o>>5)':lambda o:o+(o>>4)+(o>>5),
}
for N in (64,128,256,512,1024,2048):
    print("N",N,radices(N))
    for name,f in cands.items():
        print("   %-32s  f32 extra %4d/%d   f64 extra %4d/%d"%((name,)+conflicts(N,f,8)+conflicts(N,f,16)))
