#!/usr/bin/env python3
"""Offline fit of the texture unit's trilinear blend from gpurun_out/tex_weights.npz (written on the GPU box by tools/tex_weight_dump.py):
which integer corner weights does the hardware use?  Runs on the CPU.  Result: z -> x -> y hierarchical split of 256, see the output in
profiles/r02f_tex_weight_fit.txt and csrc/device/vpt_trace_brick.cuh."""
import os
import numpy as np, itertools
_root=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_full=os.path.join(_root, 'gpurun_out', 'tex_weights.npz')
# the full dump is scratch; a 20 000-point sample of the corner weights is committed under profiles/ (no random-texture part)
z=np.load(_full if os.path.exists(_full) else os.path.join(_root, 'profiles', 'r02f_tex_weights_sample.npz'))
HAVE_TEXTURE='data' in z.files
pts=z['pts'].astype(np.float64); W=np.round(z['w']*256).astype(np.int64)   # [8][n], corner c = (z<<2)|(y<<1)|x
n=pts.shape[0]
x=pts*2-0.5               # texel coords in [0,1]
F=x*256                   # exact weight*256 (real)
A=np.floor(F+0.5).astype(np.int64)   # per-axis 8-bit rounded weights (x,y,z)
print("axis marginal sums vs A:")
wx=W[[1,3,5,7]].sum(0); wy=W[[2,3,6,7]].sum(0); wz=W[[4,5,6,7]].sum(0)
for nm,m,a in (("x",wx,A[:,0]),("y",wy,A[:,1]),("z",wz,A[:,2])):
    print(nm, "marginal == round(f*256):", np.mean(m==a), "diff hist", np.unique(m-a, return_counts=True))
def rhu(v): return np.floor(v+0.5).astype(np.int64)
# hierarchical hypotheses: order perm of axes; split T by weight a: hi = round(T*a/256), lo = T-hi
def hier(order, rnd=rhu, use_real=False):
    parts={(): np.full(n,256,dtype=np.int64)}
    for ax in order:
        newp={}
        for key,T in parts.items():
            a = F[:,ax] if use_real else A[:,ax]
            hi = rnd(T*a/256.0); lo=T-hi
            newp[key+((ax,1),)]=hi; newp[key+((ax,0),)]=lo
        parts=newp
    out=np.zeros((8,n),dtype=np.int64)
    for key,T in parts.items():
        d=dict(key); c=(d[2]<<2)|(d[1]<<1)|d[0]
        out[c]=T
    return out
for order in itertools.permutations([0,1,2]):
    for ur in (False,True):
        H=hier(order,use_real=ur)
        print(order, "real" if ur else "8bit", "all 8 equal:", np.mean(np.all(H==W,axis=0)), "per-corner:", np.round(np.mean(H==W,axis=1),3))
print("---- mismatches for order z,x,y")
H=hier((2,0,1))
bad=np.where(~np.all(H==W,axis=0))[0]
for i in bad[:25]:
    Ax,Ay,Az=A[i]
    print("A(x,y,z)=",Ax,Ay,Az," F=",np.round(F[i],3)," hw",W[:,i]," mine",H[:,i])
print("---- hypothesis: z -> x -> y, ties up except the y split of the x=0 branch (ties down)")
def rhd(v): return np.ceil(v-0.5).astype(np.int64)
def model(A):
    Ax,Ay,Az=A[:,0],A[:,1],A[:,2]
    out=np.zeros((8,len(Ax)),dtype=np.int64)
    Z1=Az; Z0=256-Az
    for zb,T in ((0,Z0),(1,Z1)):
        X1=rhu(T*Ax/256.0); X0=T-X1
        y11=rhu(X1*Ay/256.0); y10=X1-y11
        y01=rhd(X0*Ay/256.0); y00=X0-y01
        out[(zb<<2)|0]=y00; out[(zb<<2)|1]=y10; out[(zb<<2)|2]=y01; out[(zb<<2)|3]=y11
    return out
H=model(A)
print("all 8 equal:", np.mean(np.all(H==W,axis=0)))
bad=np.where(~np.all(H==W,axis=0))[0]
for i in bad[:10]:
    print("A=",A[i]," F=",np.round(F[i],3)," hw",W[:,i]," mine",H[:,i])
if not HAVE_TEXTURE: raise SystemExit(0)
# and the random-valued 64^3 texture
data=z['data'].astype(np.float64); pr=z['pr'].astype(np.float64); hw=z['hw']
N=64
xx=pr*N-0.5; fl=np.floor(xx); Fr=(xx-fl)*256; Ar=np.floor(Fr+0.5).astype(np.int64); c=fl.astype(np.int64)
up=Ar>=256; c=np.where(up,c+1,c); Ar=np.where(up,0,Ar)
lo=c<0; Ar=np.where(lo,0,Ar); c=np.where(lo,0,c)
c0=np.clip(c,0,N-1); c1=np.clip(c+1,0,N-1)
Wm=model(Ar)
res=np.zeros(len(pr)); res32=np.zeros(len(pr),dtype=np.float32)
for cc in range(8):
    ix=np.where(cc&1,c1[:,0],c0[:,0]); iy=np.where((cc>>1)&1,c1[:,1],c0[:,1]); iz=np.where((cc>>2)&1,c1[:,2],c0[:,2])
    res+=Wm[cc]/256.0*data[iz,iy,ix]
d=np.abs(res-hw)
print("random 64^3 texture: max |d|",d.max()," mean",d.mean()," bit-equal as fp32:",np.mean(res.astype(np.float32)==hw), " within 1e-7:", np.mean(d<1e-7))
print("---- residual mismatches on the random texture")
bad=np.where(d>=1e-7)[0]
print(len(bad), "of", len(pr))
for i in bad[:14]:
    print("pr",pr[i]," F",np.round(Fr[i],4)," A",Ar[i]," cell",c[i]," d",d[i])
print("---- with the upper clamp zone mapped to weight 0")
hi=c>=N-1; Ar2=np.where(hi,0,Ar); c2=np.where(hi,N-1,c)
c0=np.clip(c2,0,N-1); c1=np.clip(c2+1,0,N-1)
Wm=model(Ar2)
res=np.zeros(len(pr))
for cc in range(8):
    ix=np.where(cc&1,c1[:,0],c0[:,0]); iy=np.where((cc>>1)&1,c1[:,1],c0[:,1]); iz=np.where((cc>>2)&1,c1[:,2],c0[:,2])
    res+=Wm[cc]/256.0*data[iz,iy,ix]
d=np.abs(res-hw)
print("random 64^3 texture: max |d|",d.max()," mean",d.mean()," bit-equal as fp32:",np.mean(res.astype(np.float32)==hw), " within 1e-7:", np.mean(d<1e-7))
bad=np.where(d>=1e-7)[0]; print(len(bad))
for i in bad[:10]: print("pr",pr[i]," F",np.round(Fr[i],4)," A",Ar2[i]," cell",c2[i]," d",d[i])
# fp32 blend orders
w32=(Wm/256.0).astype(np.float32)
vals=[]
for cc in range(8):
    ix=np.where(cc&1,c1[:,0],c0[:,0]); iy=np.where((cc>>1)&1,c1[:,1],c0[:,1]); iz=np.where((cc>>2)&1,c1[:,2],c0[:,2])
    vals.append(z['data'][iz,iy,ix])
import itertools
def seq(order):
    acc=np.zeros(len(pr),dtype=np.float32)
    for cc in order: acc=(acc.astype(np.float64)+w32[cc].astype(np.float64)*vals[cc].astype(np.float64)).astype(np.float32)  # fma chain
    return acc
for order in ([0,1,2,3,4,5,6,7],[7,6,5,4,3,2,1,0],[0,2,1,3,4,6,5,7],[0,4,1,5,2,6,3,7]):
    print(order, "fma chain bit-equal:", np.mean(seq(order)==hw))
