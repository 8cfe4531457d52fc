"""TEST INFRASTRUCTURE, not collected by pytest: time-boxed fuzz of the oracle port against the compiled reference, stage by
stage (pairs with every filter, quads, rigid fits incl. max_angle, Verify, TryCongruentSet) over random clouds / deltas / seeds.
  python tests/fuzz_port_vs_reference.py [seed] [seconds]      round 1: 1767 configurations in 240 s, 0 differences"""
import sys, time
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import port as oport, ref as oref
from super4pcs_b200 import synth
bits=lambda a: np.ascontiguousarray(a,np.float32).view(np.uint32)
def same(a,b):
    a=np.asarray(a,np.float32); b=np.asarray(b,np.float32); na,nb=np.isnan(a),np.isnan(b)
    return np.array_equal(na,nb) and np.array_equal(bits(a[~na]),bits(b[~nb]))
rng=np.random.RandomState(int(sys.argv[1]) if len(sys.argv)>1 else 0)
t0=time.time(); n_cfg=0; bad=0
while time.time()-t0 < float(sys.argv[2]) if len(sys.argv)>2 else 120:
    n=int(rng.randint(40,700)); ov=float(rng.choice([0.3,0.5,0.8])); normals=bool(rng.randint(0,2))
    delta=float(rng.choice([0.005,0.02,0.05,0.1])); seed=int(rng.randint(1,1e6))
    d=synth.make_pair(n,ov,seed=seed,with_normals=normals,noise_sigma=float(rng.choice([0,0.003])),outlier_frac=float(rng.choice([0,0.15])))
    filt={}
    if normals and rng.randint(0,2): filt['max_normal_difference']=float(rng.choice([15.0,40.0,90.0]))
    if rng.randint(0,3)==0: filt['max_translation_distance']=float(rng.choice([0.5,2.0]))
    if rng.randint(0,3)==0: filt['max_angle']=float(rng.choice([30.0,80.0]))
    opt=oref.make_options(delta=delta,sample_size=10**8,overlap=ov,random_seed=seed,**filt)
    m=oref.RefMatcher(d['P'],d['Q'],opt,Pn=d['Pn'],Qn=d['Qn'])
    P,Pn,_=m.sampled_p(); Q,Qn,Qrgb=m.sampled_q()
    pt=oport.Port(P,Q,delta,Qn=Qn if normals else None)
    f4=(filt.get('max_normal_difference',-1),filt.get('max_translation_distance',-1),filt.get('max_angle',-1),-1)
    for _ in range(2):
        ok,inv1,inv2,ids=m.select_quadrilateral()
        if not ok: continue
        bx,bn,brgb=m.base3d()
        b9=lambda i: np.concatenate([bx[i],bn[i],brgb[i]]).astype(np.float32)
        en=lambda v: np.sqrt(np.float32(v[0]*v[0])+(np.float32(v[1]*v[1])+np.float32(v[2]*v[2])))
        d1,d2=en(bx[0]-bx[1]),en(bx[2]-bx[3]); a1,a2=en(bn[0]-bn[1]),en(bn[2]-bn[3])
        eps=2*delta
        p1r,p2r=m.extract_pairs(d1,a1,eps,0,1),m.extract_pairs(d2,a2,eps,2,3)
        p1,p2=pt.extract_pairs(d1,a1,eps,b9(0),b9(1),f4),pt.extract_pairs(d2,a2,eps,b9(2),b9(3),f4)
        if not (np.array_equal(p1,p1r) and np.array_equal(p2,p2r)): bad+=1; print('PAIRS DIFF',n,delta,seed,filt); continue
        if len(p1r)*len(p2r)>4e7: continue
        qr=m.find_quads(inv1,inv2,eps,eps,p1r,p2r); qp=pt.find_quads(inv1,inv2,eps,bx,p1,p2)
        if not np.array_equal(qr,qp): bad+=1; print('QUADS DIFF',n,delta,seed,len(qr),len(qp)); continue
        if len(qr):
            q=qr[:2000]
            Tr,rr,okr=m.rigid_batch(ids,q); Tp,rp,okp=pt.rigid_batch(ids,q,max_angle_deg=filt.get('max_angle',-1.0))
            sel=okr&(rr<1e8)
            if not (np.array_equal(okr,okp) and same(rr,rp) and same(Tr[sel],Tp[sel])): bad+=1; print('RIGID DIFF',n,delta,seed,filt); continue
            if sel.any():
                T=Tr[sel][:24]
                lr,_=m.verify_batch(T,0.0); lp,good,_=pt.verify_batch(T,0.0)
                if not np.array_equal(lr,lp): bad+=1; print('VERIFY DIFF',n,delta,seed); continue
                m.set_best_lcp(0.0)
                tr=m.try_congruent_set(ids,q); tp=pt.try_congruent_set(ids,q,best_lcp_in=0.0,max_angle_deg=filt.get('max_angle',-1.0))
                if not (tr['n_gate']==tp['n_gate'] and np.float32(tr['best_lcp'])==np.float32(tp['best_lcp'])): bad+=1; print('TCS DIFF',n,delta,seed,tr['n_gate'],tp['n_gate'],tr['best_lcp'],tp['best_lcp'])
    m.close(); n_cfg+=1
print('configs',n_cfg,'bad',bad,'secs',round(time.time()-t0,1))
