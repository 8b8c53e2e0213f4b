#!/usr/bin/env python
"""Differential fuzzing of the CPU oracle against the live unmodified reference (oracle/_ref) over random option
combinations -- TEST INFRASTRUCTURE.  Part 1: DD / DDrppi / DDsmu (precision, periodicity incl. one non-periodic axis,
auto/cross, refine factors, max_cells_per_dim, rmin = 0).  Part 2: DDrppi_mocks / DDsmu_mocks / DDtheta (link modes).
  python tools/fuzz_oracle_vs_reference.py box SEED NTRIALS     |     ... sky SEED NTRIALS
End of round 1: box 1 150 and sky 2 120 -> 0 mismatches."""
import sys
which = sys.argv.pop(1)
if which == "box":
    import sys, os, numpy as np
    sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
    import harness as H
    from corrfunc_b200 import _capi
    ref=H.load_ref()
    rng=np.random.default_rng(int(sys.argv[1]))
    bad=0; ran=0
    for trial in range(int(sys.argv[2])):
        dtype=[np.float64,np.float32][int(rng.integers(2))]
        n=int(rng.integers(200,4000)); L=np.array([rng.uniform(40,120) for _ in range(3)])
        x,y,z=[(rng.random(n)*L[a]).astype(dtype) for a in range(3)]
        periodic=bool(rng.integers(2)); autocorr=int(rng.integers(2))
        n2=int(rng.integers(100,3000)); x2,y2,z2=[(rng.random(n2)*L[a]).astype(dtype) for a in range(3)]
        refine=tuple(int(v) for v in rng.integers(1,4,3)); custom=bool(rng.integers(2)); maxc=int(rng.choice([5,11,100]))
        rmax=float(rng.uniform(3,0.45*L.min())); edges=np.sort(np.concatenate([[rng.choice([0.0,0.3])],rng.uniform(0.5,rmax,4),[rmax]]))
        box=tuple(float(v) for v in L)
        if periodic and rng.integers(3)==0:
            a=int(rng.integers(3)); box=tuple(-1.0 if i==a else box[i] for i in range(3))
        stat=str(rng.choice(["DD","DDrppi","DDsmu"]))
        pimax=float(int(rng.uniform(2,0.45*L[2]))); mu_max=float(rng.uniform(0.3,1.0)); nmu=int(rng.integers(1,8))
        if not custom: refine=(2,2,1)
        o=_capi.default_options(dtype,periodic=periodic,boxsize=box if periodic else None,isa=H.ref_isa(),bin_refine_factors=refine,custom_refine=custom,max_cells_per_dim=maxc,need_avg_sep=True)
        kw=dict(options=o)
        if not autocorr: kw.update(X2=x2,Y2=y2,Z2=z2)
        try:
            if stat=="DD": r=_capi.call_DD(ref,autocorr,4,edges,x,y,z,**kw)
            elif stat=="DDrppi": r=_capi.call_DDrppi(ref,autocorr,4,pimax,edges,x,y,z,**kw)
            else: r=_capi.call_DDsmu(ref,autocorr,4,edges,mu_max,nmu,x,y,z,**kw)
            rn=r["npairs"]
        except RuntimeError: rn=None
        okw=dict(autocorr=bool(autocorr),periodic=periodic,boxsize=box if periodic else None,refine=refine,custom_refine=custom,max_cells=maxc,need_avg=True)
        if not autocorr: okw.update(X2=x2,Y2=y2,Z2=z2)
        try:
            a=H.oracle_theory(stat,x,y,z,edges,pimax=pimax,mu_max=mu_max,nmu_bins=nmu,**okw); an=a["npairs"]
        except RuntimeError: an=None
        ran+=1
        same=(rn is None and an is None) or (rn is not None and an is not None and np.array_equal(np.asarray(rn).ravel(),np.asarray(an).ravel()))
        if not same:
            bad+=1; print("MISMATCH",trial,stat,dtype.__name__,"per",periodic,box,"auto",autocorr,refine,custom,maxc,"ref",None if rn is None else int(np.asarray(rn).sum()),"orc",None if an is None else int(np.asarray(an).sum()),flush=True)
    print("ran",ran,"mismatches",bad)

else:
    import sys, os, numpy as np
    sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
    import harness as H
    from corrfunc_b200 import _capi
    ref=H.load_ref()
    rng=np.random.default_rng(int(sys.argv[1]))
    bad=0; ran=0
    for trial in range(int(sys.argv[2])):
        dtype=[np.float64,np.float32][int(rng.integers(2))]
        n=int(rng.integers(300,5000)); n2=int(rng.integers(100,3000)); autocorr=int(rng.integers(2))
        def cat(m):
            ra=(rng.uniform(0,300)+rng.uniform(5,60)*rng.random(m)).astype(dtype); dec=(rng.uniform(-60,30)+rng.uniform(5,30)*rng.random(m)).astype(dtype)
            d=(rng.uniform(100,500)+rng.uniform(20,200)*rng.random(m)).astype(dtype); return ra,dec,d
        ra,dec,d=cat(n); ra2,dec2,d2=cat(n2)
        kind=str(rng.choice(["rppi_mocks","smu_mocks","theta"]))
        refine=tuple(int(v) for v in rng.integers(1,4,3)); custom=bool(rng.integers(2)); maxc=int(rng.choice([7,20,100]))
        if not custom: refine=(2,2,1)
        rmax=float(rng.uniform(5,40)); edges=np.sort(np.concatenate([[0.4],rng.uniform(0.5,rmax,4),[rmax]]))
        pimax=float(int(rng.uniform(2,40))); mu_max=float(rng.uniform(0.3,1.0)); nmu=int(rng.integers(1,8))
        try:
            if kind=="theta":
                li=[(1,1),(1,0),(0,0)][int(rng.integers(3))]; tb=np.sort(rng.uniform(0.05,8.0,6))
                rr=refine[:2] if custom else (2,2)
                o=_capi.default_options(dtype,isa=H.ref_isa(),need_avg_sep=True,link_in_dec=li[0],link_in_ra=li[1],bin_refine_factors=(rr[0],rr[1],1),custom_refine=custom,max_cells_per_dim=maxc)
                kw=dict(options=o)
                if not autocorr: kw.update(RA2=ra2,DEC2=dec2)
                r=_capi.call_DDtheta(ref,autocorr,4,tb,ra,dec,**kw)["npairs"]
                okw=dict(autocorr=bool(autocorr),link_in_dec=bool(li[0]),link_in_ra=bool(li[1]),ra_refine=rr[0],dec_refine=rr[1],max_cells=maxc,need_avg=True)
                if not autocorr: okw.update(RA2=ra2,DEC2=dec2)
                a=H.oracle_theta(ra,dec,tb,**okw)["npairs"]
            else:
                o=_capi.default_options(dtype,isa=H.ref_isa(),need_avg_sep=True,is_comoving_dist=True,bin_refine_factors=refine,custom_refine=custom,max_cells_per_dim=maxc)
                kw=dict(options=o)
                if not autocorr: kw.update(RA2=ra2,DEC2=dec2,CZ2=d2)
                okw=dict(autocorr=bool(autocorr),periodic=False,refine=refine,custom_refine=custom,max_cells=maxc,need_avg=True)
                if not autocorr: okw.update(X2=ra2,Y2=dec2,Z2=d2)
                if kind=="rppi_mocks":
                    r=_capi.call_DDrppi_mocks(ref,autocorr,1,4,pimax,edges,ra,dec,d,**kw)["npairs"]; a=H.oracle_theory("DDrppi_mocks",ra,dec,d,edges,pimax=pimax,**okw)["npairs"]
                else:
                    r=_capi.call_DDsmu_mocks(ref,autocorr,1,4,mu_max,nmu,edges,ra,dec,d,**kw)["npairs"]; a=H.oracle_theory("DDsmu_mocks",ra,dec,d,edges,mu_max=mu_max,nmu_bins=nmu,**okw)["npairs"]
        except RuntimeError as e:
            print("ERR",trial,kind,e); continue
        ran+=1
        if not np.array_equal(np.asarray(r).ravel(),np.asarray(a).ravel()):
            bad+=1; print("MISMATCH",trial,kind,dtype.__name__,"auto",autocorr,refine,custom,maxc,int(np.asarray(r).sum()),int(np.asarray(a).sum()),flush=True)
    print("ran",ran,"mismatches",bad)
