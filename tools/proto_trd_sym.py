"""NumPy statement of the index scheme of trd_panel_sym_kernel (csrc/eigh.cu): lower-triangle work units (I, J),
per-unit dot / row partials and their fixed-order summation in the next phase, checked against a dense matvec.
    python tools/proto_trd_sym.py"""
import numpy as np, math
RC=512; MAXNB=128
def sumfloor(J,q):
    a,b=divmod(J,q); return q*a*(a-1)//2+a*b
def cfg(n,j,i,G,force_tc=0):
    R0=(j+1)&~1; s=j+1-R0; mu=n-R0
    for TC in ((force_tc,) if force_tc else (64,32,16)):      # widest unit that still fills one round of the grid
        q=RC//TC; nJ=-(-mu//TC); nI=-(-mu//RC)
        UA=nJ*nI-sumfloor(nJ,q)
        nIW=-(-mu//(4*RC)); nsetW=-(-i//16); UW=2*nsetW*nIW
        U=UA+UW
        c=dict(R0=R0,s=s,mu=mu,TC=TC,q=q,nJ=nJ,nI=nI,UA=UA,nIW=nIW,nsetW=nsetW,UW=UW,U=U)
        if force_tc or U>=G: break
    return c
def decode(c,ka):
    lo,hi=0,c['nJ']-1
    pref=lambda J: J*c['nI']-sumfloor(J,c['q'])
    while lo<hi:
        mid=(lo+hi+1)//2
        if pref(mid)<=ka: lo=mid
        else: hi=mid-1
    J=lo; I=J//c['q']+(ka-pref(J)); return I,J
def phaseC(A,W,V,n,j,i,G,xfull,force_tc=0):
    """A full symmetric n x n (absolute), xfull[abs row] with zeros above j+1."""
    c=cfg(n,j,i,G,force_tc); R0,s,mu,TC,nI,nJ=c['R0'],c['s'],c['mu'],c['TC'],c['nI'],c['nJ']
    x=xfull[R0:]
    dotP=np.full((nJ,nI,TC),np.nan); zP=np.full((nI,nJ,RC),np.nan); pP=np.full((2,MAXNB,c['nIW']),np.nan)
    yhv=0.0; seen=set()
    for unit in range(c['U']):
        if unit<c['UW']:
            t,Iw=divmod(unit,c['nIW']); which,st=divmod(t,c['nsetW'])
            k0=st*16; nset=min(16,i-k0)
            M=(V if which else W)
            r0=Iw*4*RC; r1=min(r0+4*RC,mu)
            for q_ in range(nset):
                pP[which,k0+q_,Iw]=M[R0+r0:R0+r1,k0+q_]@x[r0:r1]
        else:
            I,J=decode(c,unit-c['UW']); assert (I,J) not in seen; seen.add((I,J))
            cbeg=J*TC; cend=min(cbeg+TC,mu); rbeg=max(I*RC,cbeg); rend=min((I+1)*RC,mu); z0=cbeg+TC
            assert rbeg<rend, (I,J)
            blk=np.tril(A)[R0+rbeg:R0+rend, R0+cbeg:R0+cend]      # the upper triangle is never read
            d=np.zeros(TC); d[:cend-cbeg]=blk.T@x[rbeg:rend]; dotP[J,I,:]=d
            yhv+=d[:cend-cbeg]@x[cbeg:cend]
            z=np.tril(A,-1)[R0+rbeg:R0+rend,R0+cbeg:R0+cend]@x[cbeg:cend]   # strictly lower: mirrored element
            zP[I,J,rbeg-I*RC:rend-I*RC]=z; yhv+=z@x[rbeg:rend]
    return c,dotP,zP,pP,yhv
def consume(c,dotP,zP,pP,ip):
    mu,TC,nI,nJ,q=c['mu'],c['TC'],c['nI'],c['nJ'],c['q']
    y=np.zeros(mu)
    for u in range(c['s'],mu):
        Ju=u//TC; Iu=u//RC
        acc=0.0
        for I in range(Ju//q,nI): acc+=dotP[Ju,I,u-Ju*TC]
        for J in range(0,Ju+1): acc+=zP[Iu,J,u-Iu*RC]
        y[u]=acc
    p1=pP[0,:ip,:].sum(1); p2=pP[1,:ip,:].sum(1)
    return y,p1,p2
def check(cases=None, verbose=True):
    """Runs the unit scheme against a dense matvec with a NaN-poisoned upper triangle; returns the largest relative error."""
    rng=np.random.default_rng(0)
    worst=0.0
    for n,j,i,G,ftc in cases or [(700,0,0,444,0),(700,1,1,444,0),(1100,64,0,444,16),(1100,65,1,444,32),(2300,130,2,444,64),(2300,131,3,30,64),(2300,2290,50,444,0),(2300,2297,57,444,0),(4700,7,7,444,0),(4700,6,6,296,0)]:
        A=rng.standard_normal((n,n)); A=A+A.T
        Aup=A.copy(); A=np.tril(A)+np.triu(np.full((n,n),np.nan),1)   # poison the upper triangle
        W=rng.standard_normal((n,MAXNB)); V=rng.standard_normal((n,MAXNB))
        x=np.zeros(n); x[j+1]=1.0; x[j+2:]=rng.standard_normal(n-j-2)
        # poison rows above j+1 of x-multiplied places with finite garbage: fine
        c,dotP,zP,pP,yhv=phaseC(A,W,V,n,j,i,G,x,ftc)
        y,p1,p2=consume(c,dotP,zP,pP,i)
        yref=Aup[j+1:,j+1:]@x[j+1:]
        R0=c['R0']
        err=np.abs(y[c['s']:]-yref).max()/np.abs(yref).max()
        e1=np.abs(p1-W[j+1:,:i].T@x[j+1:]).max() if i else 0; e2=np.abs(p2-V[j+1:,:i].T@x[j+1:]).max() if i else 0
        ey=abs(yhv-yref@x[j+1:])/abs(yref@x[j+1:])
        assert np.isfinite(err) and np.isfinite(ey)
        worst=max(worst,err,ey)
        if verbose: print(n,j,i,G,c['TC'],c['U'],'err',err,e1,e2,ey)
    return worst

if __name__ == '__main__':
    print('worst', check())
# efficiency table
for mu in (8192,6144,4096,2048,1024):
    c=cfg(mu,0,32,444); print(mu,c['TC'],c['U'],c['U']/444)
