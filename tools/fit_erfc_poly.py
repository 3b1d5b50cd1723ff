"""Weighted-minimax fit of log2(erfc(z)) / z on [0, 4] (Lawson iteration) for the one-MUFU erf-GELU of the fc1 epilogue
(csrc/gemm_tc.cuh gelu_fast2), with a float32 emulation of the whole formula against the exact GELU."""
import numpy as np, struct
from scipy.special import erfc
zmax=4.0
z=np.linspace(1e-7,zmax,80001)
target=np.log2(erfc(z))
deg=7
w=erfc(z)*np.log(2); lw=np.ones_like(z)
A=np.stack([z**(k+1) for k in range(deg)],1)
for it in range(200):
    W=(w*lw)[:,None]
    c,*_=np.linalg.lstsq(A*W, target*w*lw, rcond=None)
    err=np.abs((A@c-target)*w)
    lw=lw*(1+(err/err.max())); lw/=lw.max()
cf=c.astype(np.float32)
print([float(x) for x in cf]); print([hex(struct.unpack('<I',struct.pack('<f',float(x)))[0]) for x in cf])
# full float32 emulation of gelu vs exact
x=np.concatenate([np.linspace(-12,12,2000001), np.random.default_rng(0).normal(size=1000000)*2]).astype(np.float32)
zf=np.minimum(np.abs(x)*np.float32(0.70710678118654752440),np.float32(zmax)).astype(np.float32)
p=np.full_like(zf,cf[-1])
for k in range(deg-2,-1,-1): p=(p*zf+cf[k]).astype(np.float32)
q=(p*zf).astype(np.float32)
e=np.exp2(q.astype(np.float64)).astype(np.float32)   # ex2.approx ~ 2 ulp; emulate exact then add noise below
relu=np.maximum(x,np.float32(0))
nhz=(zf*np.float32(-0.35355339059327379)).astype(np.float32)   # -0.5*|x| = -z*0.5*sqrt(2)... z=|x|/sqrt2 -> 0.5|x| = z*sqrt2/2=z*0.7071
nhz=(zf*np.float32(-0.70710678118654752440)).astype(np.float32)
g=(nhz*e+relu).astype(np.float32)
from scipy.special import erf
xe=x.astype(np.float64); ref=0.5*xe*(1+erf(xe/np.sqrt(2)))
d=np.abs(g-ref); print("max abs err",d.max(),"at x=",x[d.argmax()], "max rel-to-(|ref|+1e-3)",(d/(np.abs(ref)+1e-3)).max())
# old formula error for comparison
t=1/(1+0.3275911*np.abs(xe)/np.sqrt(2)); poly=((((1.061405429*t-1.453152027)*t+1.421413741)*t-0.284496736)*t+0.254829592)*t
erfo=np.sign(xe)*(1-poly*np.exp(-(xe**2)/2)); go=0.5*xe*(1+erfo); print("old max abs",np.abs(go-ref).max())
