import numpy as np, sys, time
sys.path.insert(0,'/root/repo')
from gleipnir_b200 import networks, jarm
from oracle import ode

def ros4(f, jac, y, p, t0, t1, h, rtol, atol, log=None, norm="max"):
    GAM=0.5; A21=2.0; A31=48/25; A32=6/25
    C21=-8.0; C31=372/25; C32=12/5; C41=-112/125; C42=-54/125; C43=-2/5
    B1=19/9; B2=0.5; B3=25/108; B4=125/108
    E1=17/54; E2=7/36; E4=125/108
    t=t0; n=len(y); steps=0; rej=0
    I=np.eye(n)
    while t < t1:
        rem=t1-t; last = h>=rem; hh = rem if last else h
        steps+=1
        J=jac(y,t,p); Minv=np.linalg.inv(I/(GAM*hh)-J); ih=1/hh
        f0=f(y,t,p); g1=Minv@f0
        g2=Minv@(f(y+A21*g1,t,p)+C21*ih*g1)
        f3=f(y+A31*g1+A32*g2,t,p)
        g3=Minv@(f3+C31*ih*g1+C32*ih*g2)
        g4=Minv@(f3+C41*ih*g1+C42*ih*g2+C43*ih*g3)
        ynew=y+B1*g1+B2*g2+B3*g3+B4*g4
        yerr=E1*g1+E2*g2+E4*g4
        sc = atol+rtol*np.maximum(np.abs(y),np.abs(ynew))
        e = np.max(np.abs(yerr)/sc) if norm=="max" else np.sqrt(np.mean((yerr/sc)**2))
        fac = min(5.0,max(0.2,0.9*e**-0.25)) if e>0 else 5.0
        if e<=1.0:
            y=ynew; t = t1 if last else t+hh
            if (not last) or fac<1: h=hh*fac
            else: h=max(h,hh*fac)
            if log is not None: log.append((t,hh,int(np.argmax(np.abs(yerr)/sc))))
        else:
            rej+=1; h=hh*fac
    return y,h,steps,rej

net=networks.jarm(); d=net.description(); f,jac=ode.rhs_jac(d)
like=jarm.jarm_loglikelihood(); proto=like.oracle_protocol()
base=np.array(d["param_values"],float)
def run(Y, rtol=1e-8, atol=1e-12, norm="max"):
    b=base.copy(); b[proto["sampled_idx"]]=10.0**Y
    tot=[]
    for c,ov in enumerate(proto["conditions"]):
        p=b.copy()
        for i,v in ov: p[i]=v
        y0=np.zeros(net.ns)
        for s,(k,v) in d["initials"].items(): y0[s]=p[v] if k=="p" else v
        pe=p.copy(); pe[proto["preeq"]["zero_idx"]]=0
        log=[]
        y,h,s1,r1=ros4(f,jac,y0,pe,0,100.0,0.1,rtol,atol,log,norm)
        stops=sorted(set(int(i) for r in proto["rows"] if r["cond"]==c for i in np.nonzero(r["mask"])[0]))
        ts=proto["tspan"]; t=0.0; h=min(1e-3*180,ts[stops[1]]-t); s2=0;r2=0; log2=[]
        for it in stops[1:]:
            y,h,s,r=ros4(f,jac,y,p,t,ts[it],h,rtol,atol,log2,norm); t=ts[it]; s2+=s; r2+=r
        tot.append((s1,r1,s2,r2,log,log2))
    return tot
Y0=np.log10(base[proto["sampled_idx"]])
for rtol,atol,norm in ((1e-8,1e-12,"max"),(1e-8,1e-12,"rms"),(1e-6,1e-12,"max"),(1e-8,1e-9,"max")):
    tot=run(Y0,rtol,atol,norm)
    print(rtol,atol,norm,[(a,b,c,dd) for a,b,c,dd,_,_ in tot])
tot=run(Y0)
for c in (0,1):
    log=tot[c][4]
    ts=np.array([l[0] for l in log]); 
    print("cond",c,"preeq accepted steps by decade of t:", np.histogram(np.log10(ts),bins=[-9,-6,-5,-4,-3,-2,-1,0,1,2.01])[0])
    from collections import Counter
    print(" limiting species (internal idx orig):", Counter(l[2] for l in log).most_common(5), [net.species[i] for i,_ in Counter(l[2] for l in log).most_common(5)])
    log2=tot[c][5]
    print(" main: steps per stop interval ~", len(log2), "hh range", min(l[1] for l in log2), max(l[1] for l in log2))

def rodas4(f, jac, y, p, t0, t1, h, rtol, atol, log=None, norm="max"):
    A21=0.1544000000000000e+01; A31=0.9466785280815826e+00; A32=0.2557011698983284e+00
    A41=0.3314825187068521e+01; A42=0.2896124015972201e+01; A43=0.9986419139977817e+00
    A51=0.1221224509226641e+01; A52=0.6019134481288629e+01; A53=0.1253708332932087e+02; A54=-0.6878860361058950e+00
    C21=-0.5668800000000000e+01; C31=-0.2430093356833875e+01; C32=-0.2063599157091915e+00
    C41=-0.1073529058151375e+00; C42=-0.9594562251023355e+01; C43=-0.2047028614809616e+02
    C51=0.7496443313967647e+01; C52=-0.1024680431464352e+02; C53=-0.3399990352819905e+02; C54=0.1170890893206160e+02
    C61=0.8083246795921522e+01; C62=-0.7981132988064893e+01; C63=-0.3152159432874371e+02; C64=0.1631930543123136e+02; C65=-0.6058818238834054e+01
    GAM=0.25
    t=t0; n=len(y); steps=0; rej=0; I=np.eye(n)
    while t < t1:
        rem=t1-t; last=h>=rem; hh=rem if last else h
        steps+=1
        J=jac(y,t,p); Minv=np.linalg.inv(I/(GAM*hh)-J); ih=1/hh
        k1=Minv@f(y,t,p)
        k2=Minv@(f(y+A21*k1,t,p)+C21*ih*k1)
        k3=Minv@(f(y+A31*k1+A32*k2,t,p)+ih*(C31*k1+C32*k2))
        k4=Minv@(f(y+A41*k1+A42*k2+A43*k3,t,p)+ih*(C41*k1+C42*k2+C43*k3))
        yn=y+A51*k1+A52*k2+A53*k3+A54*k4
        k5=Minv@(f(yn,t,p)+ih*(C51*k1+C52*k2+C53*k3+C54*k4))
        yn=yn+k5
        k6=Minv@(f(yn,t,p)+ih*(C61*k1+C62*k2+C63*k3+C64*k4+C65*k5))
        ynew=yn+k6; yerr=k6
        sc=atol+rtol*np.maximum(np.abs(y),np.abs(ynew))
        e=np.max(np.abs(yerr)/sc) if norm=="max" else np.sqrt(np.mean((yerr/sc)**2))
        fac=min(5.0,max(0.2,0.9*e**-0.25)) if e>0 else 5.0
        if e<=1.0:
            y=ynew; t=t1 if last else t+hh
            if (not last) or fac<1: h=hh*fac
            else: h=max(h,hh*fac)
            if log is not None: log.append((t,hh,int(np.argmax(np.abs(yerr)/sc))))
        else:
            rej+=1; h=hh*fac
    return y,h,steps,rej

# order check on a nonstiff nonlinear problem: fixed steps
def fixed(method, nsteps):
    ff=lambda y,t,p: np.array([-y[0]*y[1], y[0]*y[1]-0.5*y[1]])
    jj=lambda y,t,p: np.array([[-y[1],-y[0]],[y[1],y[0]-0.5]])
    y=np.array([1.0,0.5]); T=2.0; h=T/nsteps
    for i in range(nsteps):
        y,_,_,_=method(ff,jj,y,None,0.0,h,h,1e300,1e300)
    return y
from scipy.integrate import solve_ivp
ref=solve_ivp(lambda t,y: [-y[0]*y[1], y[0]*y[1]-0.5*y[1]],[0,2],[1.0,0.5],rtol=1e-13,atol=1e-14).y[:,-1]
for m in (ros4, rodas4):
    errs=[np.abs(fixed(m,n)-ref).max() for n in (10,20,40,80)]
    print(m.__name__, errs, "orders", [np.log2(errs[i]/errs[i+1]) for i in range(3)])

def run2(method, Y, rtol=1e-8, atol=1e-12, norm="max"):
    b=base.copy(); b[proto["sampled_idx"]]=10.0**Y
    out=[]; ys=[]
    for c,ov in enumerate(proto["conditions"]):
        p=b.copy()
        for i,v in ov: p[i]=v
        y0=np.zeros(net.ns)
        for s,(k,v) in d["initials"].items(): y0[s]=p[v] if k=="p" else v
        pe=p.copy(); pe[proto["preeq"]["zero_idx"]]=0
        y,h,s1,r1=method(f,jac,y0,pe,0,100.0,0.1,rtol,atol,None,norm)
        stops=sorted(set(int(i) for r in proto["rows"] if r["cond"]==c for i in np.nonzero(r["mask"])[0]))
        ts=proto["tspan"]; t=0.0; h=min(1e-3*180,ts[stops[1]]-t); s2=0
        for it in stops[1:]:
            y,h,s,r=method(f,jac,y,p,t,ts[it],h,rtol,atol,None,norm); t=ts[it]; s2+=s
        out.append((s1,s2)); ys.append(y)
    return out, ys
_, yref = run2(rodas4, Y0, 1e-11, 1e-14)
rng=np.random.default_rng(1)
for m in (ros4, rodas4):
    for rtol in (1e-8,1e-7,1e-6):
        o, ys = run2(m, Y0, rtol)
        err=max(np.max(np.abs(a-b)/(np.abs(b)+1e-9)) for a,b in zip(ys,yref))
        print(m.__name__, rtol, o, "final-state rel err %.2e"%err)
Y1 = Y0 + rng.uniform(-1,1,18)
for m in (ros4, rodas4):
    o, ys = run2(m, Y1, 1e-8); print(m.__name__, "perturbed", o)
