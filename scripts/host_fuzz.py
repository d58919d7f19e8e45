"""Random sweep of the kernels' code on the CPU (no GPU needed): builds the host harness of tests/test_generic_solver_host.py —
the generic solver's sweeps and Controller, and optionally the single-pass kernels' per-unit code, pasted verbatim from the .cu
sources — and holds random launches (families, d from 1 to 69, L-BFGS memory 1/3/10, every start kind, atol from 1 to 1e-9,
θ_sim ≠ θ_eval) against the oracle, unit by unit.

    python scripts/host_fuzz.py SEED TRIALS        # generic solver only
    python scripts/host_fuzz.py SEED TRIALS sp     # single pass first, generic solver on the hand-backs (launch_solver)
"""
import os, re, subprocess, sys, tempfile
import ctypes as C

import numpy as np
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0,ROOT); sys.path.insert(0,os.path.join(ROOT,'tests'))
import oracle as O
import test_generic_solver_host as T
csrc=ROOT+'/museinference.jl_b200/csrc'
ctl=open(csrc+'/muse_iso_ctl.cuh').read(); sol=open(csrc+'/muse_iso_solver.cu').read()
def block(src,tag): return re.search(r"// \[host-test:begin %s\][^\n]*\n(.*?)// \[host-test:end %s\]"%(tag,tag),src,flags=re.S).group(1)
stream=open(csrc+'/muse_iso_stream.cu').read()
src=open(ROOT+'/tests/csrc/generic_solver_host.cpp.in').read().replace('@CTL_A@',block(ctl,'ctl-a')).replace('@CTL_B@',block(ctl,'ctl-b')).replace('@SWEEPS@',block(sol,'sweeps'))
for tag,key in (('red-slots','@RED_SLOTS@'),('item-desc','@ITEM_DESC@'),('elem3','@ELEM3@'),('fast-replay','@FAST_REPLAY@'),('publish','@PUBLISH@'),('warp-item','@WARP_ITEM@')): src=src.replace(key,block(stream,tag))
open(os.path.join(tempfile.gettempdir(),'muse_gs.cpp'),'w').write(src)
subprocess.run(["g++","-O2","-std=c++17","-fPIC","-shared","-ffp-contract=off","-I",csrc,"-I","/usr/local/cuda/include","-o",os.path.join(tempfile.gettempdir(),'libmuse_gs.so'),os.path.join(tempfile.gettempdir(),'muse_gs.cpp')],check=True)
lib=C.CDLL(os.path.join(tempfile.gettempdir(),'libmuse_gs.so'))
rng=np.random.default_rng(int(sys.argv[1]) if len(sys.argv)>1 else 0)
nbad=0; nunits=0; near=0
for trial in range(int(sys.argv[2]) if len(sys.argv)>2 else 200):
    family='funnel' if rng.random()<0.5 else 'hiergauss'
    d=int(rng.integers(1,70)); nsims=int(rng.integers(2,8))
    fam=O.make_family(family,d)
    draws=O.Draws.from_numpy(int(rng.integers(1<<30)),nsims,d)
    xd,_=fam.sample(np.zeros(fam.ntheta),rng.standard_normal(d),rng.standard_normal(d))
    prob=O.OracleProblem(fam,xd,draws)
    m=int(rng.choice([1,3,10]))
    hs=T.HostSolver(lib,family,d,draws,xd,lbfgs_m=m,single_pass=(len(sys.argv)>3))
    zcur=[np.zeros(d) for _ in range(nsims+1)]
    for p in range(3):
        th=np.array([rng.normal(0,1.2)]) if family=='funnel' else np.array([rng.normal(0,1.5),rng.normal(0,0.7)])
        ths=th+ (rng.normal(0,0.1,th.size) if rng.random()<0.3 else 0)
        atol=float(rng.choice([1.0,1e-2,1e-5,1e-9]))
        kind=int(rng.choice([T.ZERO,T.OWN,T.TRUTH,T.KEEP])) if p>0 else int(rng.choice([T.ZERO,T.TRUTH,T.KEEP]))
        incl=bool(rng.random()<0.7)
        z0u=rng.normal(0,1,d) if kind==T.KEEP else None
        out=hs.map_score(ths,th,atol,incl,kind,zshared=z0u)
        units=([0] if incl else [])+list(range(1,nsims+1))
        for i,u in enumerate(units):
            x=xd if u==0 else prob.sample_x_z(u-1,ths)[0]
            if kind==T.ZERO: z0=np.zeros(d)
            elif kind==T.OWN: z0=zcur[u]
            elif kind==T.TRUTH: z0=prob.sample_x_z(u-1,ths)[1] if u>0 else np.zeros(d)
            else: z0=z0u
            _,g,soln=O.map_score_unit(prob,x,z0,th,atol)
            nunits+=1
            ok=(out['iters'][i],out['fg'][i],out['status'][i])==(soln.iterations,soln.f_calls,T._status(soln))
            zh=hs.z(u)
            if not ok:
                # near a threshold?
                if abs(out['gnorm'][i]-atol)<1e-6*atol or abs(soln.g_residual-atol)<1e-6*atol: near+=1
                else:
                    nbad+=1; print('MISMATCH',family,d,m,p,kind,atol,u,(out['iters'][i],out['fg'][i],out['status'][i]),(soln.iterations,soln.f_calls,T._status(soln)),out['gnorm'][i],soln.g_residual)
            else:
                if not np.allclose(zh,soln.minimizer,rtol=1e-10,atol=1e-11): nbad+=1; print('ZDIFF',family,d,kind,atol,u,np.abs(zh-soln.minimizer).max())
                if not np.allclose(out['g'][i],g,rtol=1e-9,atol=1e-8*max(1,np.abs(x).max())**2): nbad+=1; print('GDIFF',family,d,kind,atol,u,out['g'][i],g)
            zcur[u]=zh
print('units',nunits,'bad',nbad,'near-threshold',near,'handed back in total',int(hs.redo_total[0]) if hs.single_pass else '-')
