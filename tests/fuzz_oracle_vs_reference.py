"""Randomised differential of the oracle (oracle/libtntoracle.so) against the compiled reference
(oracle/_ref/libtntref.so) with the generator of tests/fuzz_parity.py -- the other link of the chain engine ==
oracle == reference.  CPU only; needs /root/reference to have been built (oracle/Makefile).
    python tests/fuzz_oracle_vs_reference.py <seconds> <seed>
Searches in which the reference throws (`deflate_dna_seq: Unknown symbol`: hits with dangling-end columns or
degenerate oligo letters cannot be packed into a hybrid_sig) are counted as exceptions, not as
differences."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np
import gen, harness as H
O=H.oracle(); R=H.ref()
budget=float(sys.argv[1]); seed0=int(sys.argv[2])
t_end=time.time()+budget
it=ncase=nhits=nbad=nexc=0
while time.time()<t_end:
    it+=1
    rng=np.random.default_rng(seed0*100003+it)
    kind=["pcr","taqman","probe","padlock"][int(rng.integers(0,4))]
    W=int(rng.choice([4,5,6,7,7,7,8]))
    pick=int(rng.integers(0,8))
    T=float(rng.choice([310.15,310.15,300.15,325.15])); na=float(rng.choice([0.05,0.05,0.2,1.0]))
    dink=bool(rng.integers(0,8)==0)
    d5,d3=(int(rng.integers(0,2)),int(rng.integers(0,2))) if rng.integers(0,5)==0 else (0,0)
    lens=(int(rng.integers(14,31)),int(rng.integers(14,31)),int(rng.integers(16,41)))
    if rng.integers(0,6)==0: lens=(int(rng.integers(14,55)),int(rng.integers(14,55)),int(rng.integers(14,55)))
    top=150000 if rng.integers(0,10)==0 else 40000
    db=[gen.random_codes(int(rng.integers(4000,top)),rng) for _ in range(int(rng.integers(1,4)))]
    if rng.integers(0,3)==0:
        gen.sprinkle_degenerate(db[0],rng,frac=float(rng.choice([5e-4,5e-3])),n_runs_per_50kb=int(rng.integers(0,8)))
    assays=gen.make_assays(rng,db,int(rng.integers(1,5)),kind,lens=lens,variants=int(rng.integers(1,5)))
    if rng.integers(0,5)==0:
        def degen(ol):
            if ol is None: return None
            ol=list(ol)
            for _ in range(int(rng.integers(1,3))):
                i=int(rng.integers(0,len(ol)))
                ol[i]="I" if rng.integers(0,2) else {"A":"R","G":"R","C":"Y","T":"Y"}[ol[i]] if ol[i] in "ACGT" else ol[i]
            return "".join(ol)
        assays=[tuple(degen(x) for x in a) for a in assays]
    kw=dict(word_size=W,target_T=T,salt=na,dangle5=d5,dangle3=d3,
        min_primer_tm=float(rng.choice([0.0,30.0,38.0,45.0])),min_probe_tm=float(rng.choice([0.0,30.0,40.0])),
        max_len=int(rng.choice([300,2000])),single_primer_pcr=int(rng.integers(0,2)),
        primer_clamp=int(rng.integers(0,4)),min_max_primer_clamp=int(rng.choice([-1,-1,3])),
        probe_clamp_5=int(rng.integers(0,3)),probe_clamp_3=int(rng.integers(0,3)),
        max_gap=int(rng.choice([999,999,0,1])),max_mismatch=int(rng.choice([999,999,2])),
        target_strand=int(rng.choice([3,3,1,2])))
    if kw["min_primer_tm"]==0.0: kw["max_primer_dg"]=float(rng.choice([-6.0,-9.0]))
    if kw["min_probe_tm"]==0.0: kw["max_probe_dg"]=float(rng.choice([-6.0,-9.0]))
    if kind=="probe": kw["assay_format"]=H.ASSAY_PROBE
    elif kind=="padlock":
        kw["assay_format"]=int(rng.choice([H.ASSAY_PADLOCK,H.ASSAY_MIPS])); kw["max_len"]=int(rng.choice([0,3,50]))
    o=H.default_options(**kw)
    O.set_dinkelbach(dink); R.set_dinkelbach(dink)
    for t,codes in enumerate(db):
        for i,a in enumerate(assays):
            try:
                x=R.search(codes,a[0],a[1],a[2],o); y=O.search(codes,a[0],a[1],a[2],o)
            except Exception as ex:
                nexc+=1; continue
            ncase+=1; nhits+=len(x)
            if [(h.exact_key(),h.floats()) for h in x]!=[(h.exact_key(),h.floats()) for h in y]:
                nbad+=1; print('DIFF it=%d kind=%s W=%d T=%g na=%g dink=%d d5=%d d3=%d assay=%s nref=%d norc=%d kw=%s'%(it,kind,W,T,na,dink,d5,d3,a,len(x),len(y),kw),flush=True)
O.set_dinkelbach(False); R.set_dinkelbach(False)
print('oracle vs reference: %d iterations, %d searches, %d hits, %d differences, %d searches in which the reference threw'%(it,ncase,nhits,nbad,nexc))
