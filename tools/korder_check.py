"""CPU emulation of the generic tcgen05 GEMM's K order (shape_for / korder_true_k / gen_values in dctn_b200/csrc/eps_tc_gemm.cu):
for a grid of (Q, #factors, with/without gout, Q_out) checks that every true k appears exactly once among the packed positions, that
padding positions are marked, that sections are 32-aligned and that the producer's decode (section, table index, value) agrees with the
pack kernel's.  Design check only: the GPU parity tests are the test of the kernels."""
import numpy as np
def ipow(b,e):
    r=1
    for _ in range(e): r*=b
    return r
def shape_for(Q,nf,withG,O):
    best=-1; s={}
    for cr in range(nf+1):
        G=ipow(Q,cr)*(O if withG else 1)
        if G>64: break
        H=ipow(Q,nf-cr)
        for RB in (8,4):
            GP=(G+RB-1)//RB*RB; hq=32//RB; Hpad=(H+hq-1)//hq*hq
            cost=GP*Hpad
            if best<0 or cost<best or (cost==best and (RB>s['RB'] or (RB==s['RB'] and cr>s['cr']))):
                best=cost; s=dict(cr=cr,G=G,GP=GP,RB=RB,H=H,Hpad=Hpad)
    if best<0:
        s=dict(cr=0,G=O,RB=8,GP=(O+7)//8*8,H=ipow(Q,nf)); s['Hpad']=(s['H']+3)//4*4
    s['Kp']=s['GP']*s['Hpad']
    return s
def true_k_vec(s):
    kp=np.arange(s['Kp']); sec=s['Hpad']*s['RB']; ob=kp//sec; rem=kp-ob*sec; h=rem//s['RB']; r=ob*s['RB']+(rem-h*s['RB'])
    return np.where((h<s['H'])&(r<s['G']), h*s['G']+r, -1), h, r
bad=0; n=0; worst=0
for Q in range(2,26):
    for nf in range(1,6):
        if ipow(Q,nf)>300000: continue
        for withG,O in [(0,1)]+[(1,o) for o in (2,3,4,5,6,8,10,12,23,24,70)]:
            s=shape_for(Q,nf,withG,O); n+=1
            Kdim=ipow(Q,nf)*(O if withG else 1)
            assert s['Kp']%32==0 and (s['Hpad']*s['RB'])%32==0, (Q,nf,withG,O,s)
            tk,h,r=true_k_vec(s)
            real=np.sort(tk[tk>=0])
            if real.size!=Kdim or not np.array_equal(real,np.arange(Kdim)): bad+=1; print("BAD",Q,nf,withG,O,s)
            # producer decode: within every run of 16 positions the section index ob is constant and h advances every RB positions
            kp=np.arange(s['Kp']); sec=s['Hpad']*s['RB']
            ob=kp//sec
            assert np.all(ob.reshape(-1,16)==ob.reshape(-1,16)[:,:1])
            h0=((kp-ob*sec)//s['RB']).reshape(-1,16)[:,0]; 
            hh=h.reshape(-1,16); rr=r.reshape(-1,16)
            i=np.arange(16)//s['RB']; rl=np.arange(16)%s['RB']
            assert np.array_equal(hh, h0[:,None]+i[None,:]) and np.array_equal(rr, ob.reshape(-1,16)[:,:1]*s['RB']+rl[None,:])
            worst=max(worst, s['Kp']/Kdim if Kdim>=64 else 0)
print("shapes",n,"bad",bad,"worst padding factor (K>=64)",round(worst,2))
