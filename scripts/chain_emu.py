"""CPU mirror of the index logic of k_diag_factor2 / k_trsm2 (tulip.jl_b200/csrc/kernels_factor.cu), lane by lane.

Dev tool of round 2 (there is no GPU in the build container): every shared-memory address, fragment mapping
(mma.sync m8n8k4 f64: A[g][t4], B[t4][g], C[g][2 t4 + e], g = lane >> 2, t4 = lane & 3), tile / quad enumeration and
padding rule of the two kernels is restated here with the barriers as phase boundaries, so that a change of the schedule can be
checked on the CPU: every trailing tile must be updated exactly once, nothing that was never loaded (1e300 marks) may reach
the result, ragged widths and signed pivots must factor to rounding.  It caught a real race of an earlier variant (diagonal
rows written back before the other panel warps had read them).  tests/test_chain_emulation.py runs it on small blocks.
Not a performance model and not an oracle: parity of the CUDA kernels is tested on the GPU against oracle/.
"""
import math

import numpy as np

PIECE = 128; NBD = 8; LDC2 = 130; LDP2 = 132


def dmma(lanes_a, lanes_b, c0, c1):
    # lanes_*: arrays[32]; A[g][t4]=a(lane=4g+t4); B[t4][g]=b(lane=4g+t4); C[g][2*t4+e]
    A=np.zeros((8,4)); B=np.zeros((4,8))
    for lane in range(32):
        g,t4=lane>>2,lane&3
        A[g,t4]=lanes_a[lane]; B[t4,g]=lanes_b[lane]
    P=A@B
    for lane in range(32):
        g,t4=lane>>2,lane&3
        c0[lane]+=P[g,2*t4]; c1[lane]+=P[g,2*t4+1]

L=range(32)
def panel(Cs,Up,Wn,dd,rdd,sgn,j0,nb,nrows,badl,PW=5):
    # load all warps, then compute
    state=[]
    for wp in range(PW):
        ts=[(l if l<8 else 8+24*wp+(l-8)) for l in L]
        live=[t<nrows for t in ts]
        iis=[j0+(t if lv else 0) for t,lv in zip(ts,live)]
        u=[[Cs[(j0+j)*LDC2+i] for j in range(8)] for i in iis]
        state.append((ts,live,iis,u))
    for wp in range(PW):
        ts,live,iis,u=state[wp]
        an=[[0.0]*8 for _ in L]; bad=0; dm=[1.0]*32; rm=[1.0]*32
        for j in range(8):
            draw=u[j][j]; sj=sgn[j0+j]; ok=draw*sj>0
            r0=(1.0/draw if draw!=0 else float('inf')) if ok else sj
            d=draw if ok else sj
            if not ok: bad|=1<<j
            r=1.0/d
            dm[j]=d; rm[j]=r
            ukj=[u[k][j] for k in range(8)]
            for l in L:
                a=u[l][j]*r; an[l][j]=a
                for k in range(j+1,8): u[l][k]=-a*ukj[k]+u[l][k]
        for l in L:
            if not live[l]: continue
            i=iis[l]
            if l<8:
                if wp==0:
                    for j in range(8):
                        if j<l: Cs[(j0+j)*LDC2+i]=u[l][j]
                    Cs[(j0+l)*LDC2+i]=dm[l]; dd[j0+l]=dm[l]; rdd[j0+l]=rm[l]
                    if (bad>>l)&1 and l<nb: badl.append(j0+l)
            else:
                for j in range(8):
                    Cs[(j0+j)*LDC2+i]=u[l][j]; Up[j*LDP2+i]=u[l][j]; Wn[j*LDP2+i]=-an[l][j]
def quad(Cs,Up,Wn,ti,tk0,n,touched):
    a0=[Up[(l&3)*LDP2+8*ti+(l>>2)] for l in L]; a1=[Up[(4+(l&3))*LDP2+8*ti+(l>>2)] for l in L]
    pcs=[];c0=[];c1=[];b0=[];b1=[]
    for x in range(4):
        tk=tk0+min(x,n-1)
        b0.append([Wn[(l&3)*LDP2+8*tk+(l>>2)] for l in L]); b1.append([Wn[(4+(l&3))*LDP2+8*tk+(l>>2)] for l in L])
        pcs.append([(8*tk+2*(l&3))*LDC2+8*ti+(l>>2) for l in L])
        c0.append([Cs[a] for a in pcs[x]]); c1.append([Cs[a+LDC2] for a in pcs[x]])
    for x in range(4): dmma(a0,b0[x],c0[x],c1[x])
    for x in range(4): dmma(a1,b1[x],c0[x],c1[x])
    for x in range(4):
        if x<n:
            key=(ti,tk0+x); assert key not in touched,key; touched.add(key)
            for l in L: Cs[pcs[x][l]]=c0[x][l]; Cs[pcs[x][l]+LDC2]=c1[x][l]
def diag_factor4(Dm, sign, W=16, PW=5):
    w=Dm.shape[0]; nt=(w+7)//8; nt8=nt*8
    Cs=np.full(PIECE*LDC2, 1e300); Up=np.full(2*8*LDP2,np.nan); Wp=np.full(2*8*LDP2,np.nan)
    dd=np.full(128,np.nan); rdd=np.full(128,np.nan); sgn=np.ones(128); sgn[:w]=sign
    for k in range(nt8):
        for il in range(128):
            if (il|31)>=(k&~7): Cs[k*LDC2+il]= Dm[il,k] if (k<w and il<w and il>=k) else 0.0
    for t in range(w,nt8): Cs[t*LDC2+t]=1.0
    bad=[]
    panel(Cs,Up[:8*LDP2],Wp[:8*LDP2],dd,rdd,sgn,0,min(8,w),nt8,bad)
    for b in range(nt-1):
        o=(b&1)*8*LDP2; U=Up[o:o+8*LDP2]; Wn=Wp[o:o+8*LDP2]
        t1=b+1; nrem=nt-t1; o2=(t1&1)*8*LDP2
        touched=set()
        for warp in range(W):
            if warp<nrem: quad(Cs,U,Wn,t1+warp,t1,1,touched)
        assert touched=={(t1+x,t1) for x in range(nrem)}
        j0=t1*8
        panel(Cs,Up[o2:o2+8*LDP2],Wp[o2:o2+8*LDP2],dd,rdd,sgn,j0,min(8,w-j0),nt8-j0,bad)
        T=nrem-1; c0t=t1+1; touched=set()
        nq=0; qq=0
        while 4*qq<T: nq+=T-4*qq; qq+=1
        for warp in range(PW,W):
            idx=warp-PW
            while idx<nq:
                rem=idx; qq=0
                while rem>=T-4*qq: rem-=T-4*qq; qq+=1
                p=4*qq+rem
                quad(Cs,U,Wn,c0t+p,c0t+4*qq,min(4,p+1-4*qq),touched)
                idx+=W-PW
        assert touched=={(c0t+p,c0t+q) for p in range(T) for q in range(p+1)}, (b,T)
    Lm=np.zeros((w,w))
    for k in range(w):
        sk=sgn[k]; l=math.sqrt(dd[k]*sk); r=1.0/(sk*l)
        for il in range(k,w): Lm[il,k]= l if il==k else Cs[k*LDC2+il]*r
    return Lm,bad

LDB2=132; LDX2=68; LDCD=17; TRB=16
def trsm3(L11, sign, A21):
    w=L11.shape[0]; nr=A21.shape[0]; nblk=(w+15)//16; w16=nblk*16
    Bs=np.full(128*LDB2,np.nan); Xs=np.full(128*LDX2,np.nan); Cd=np.full(128*LDCD,np.nan); invd=np.full(128,np.nan); sgd=np.ones(128); sgd[:w]=sign
    for k in range(w16):
        for n in range(128): Bs[k*LDB2+n]= L11[n,k] if (k<w and n<w and n>=k) else 0.0
        for r in range(64): Xs[k*LDX2+r]= A21[r,k] if (k<w and r<nr) else 0.0
    for t in range(128): invd[t]= 1.0/(sgd[t]*Bs[t*LDB2+t]) if t<w else 1.0
    for e in range(w16*16):
        k=e>>4; j=e&15; nn=(k&~15)+j
        Cd[k*LDCD+j]= Bs[k*LDB2+nn]*sgd[k]*invd[k] if nn>k else 0.0
    out=np.zeros((nr,w)); L=range(32)
    for b in range(nblk):
        j0=b*16
        if j0>0:
            for warp in range(8):
                rr=[warp*8+(l>>2) for l in L]
                acc=[[[Xs[(j0+nj*8+(l&3)*2+e)*LDX2+rr[l]] for l in L] for e in range(2)] for nj in range(2)]
                acc2=[[[0.0]*32 for e in range(2)] for nj in range(2)]
                for k4 in range(0,j0,8):
                    a0=[Xs[(k4+(l&3))*LDX2+rr[l]] for l in L]; a1=[Xs[(k4+4+(l&3))*LDX2+rr[l]] for l in L]
                    b00=[Bs[(k4+(l&3))*LDB2+j0+(l>>2)] for l in L]; b01=[Bs[(k4+(l&3))*LDB2+j0+8+(l>>2)] for l in L]
                    b10=[Bs[(k4+4+(l&3))*LDB2+j0+(l>>2)] for l in L]; b11=[Bs[(k4+4+(l&3))*LDB2+j0+8+(l>>2)] for l in L]
                    dmma(a0,b00,acc[0][0],acc[0][1]); dmma(a0,b01,acc[1][0],acc[1][1]); dmma(a1,b10,acc2[0][0],acc2[0][1]); dmma(a1,b11,acc2[1][0],acc2[1][1])
                for nj in range(2):
                    for e in range(2):
                        for l in L: Xs[(j0+nj*8+(l&3)*2+e)*LDX2+rr[l]]=acc[nj][e][l]+acc2[nj][e][l]
        for tid in range(64):
            z=[Xs[(j0+j)*LDX2+tid] for j in range(16)]
            for k in range(16):
                for j in range(k+1,16): z[j]=-z[k]*Cd[(j0+k)*LDCD+j]+z[j]
            for j in range(16):
                x=z[j]*invd[j0+j]; Xs[(j0+j)*LDX2+tid]=-sgd[j0+j]*x
                if tid<nr and j0+j<w: out[tid,j0+j]=x
    return out


# ---- k_invert_diag2 (tulip.jl_b200/csrc/kernels.cu): block-doubling inverse of a 128x128 lower-triangular block ----
LDI2 = 132; LDT2 = 68


def _inv2_tiles(Cs, Tb, TPW, second, h, p, mi, nj0):
    a0 = 2 * p * h; c0 = a0 + h
    acc = [[[0.0] * 32, [0.0] * 32] for _ in range(TPW)]
    kbeg = 0 if second else 8 * nj0
    kend = 8 * (mi + 1) if second else h
    for k0 in range(kbeg, kend, 4):
        if second:
            a = [Cs[(c0 + k0 + (l & 3)) * LDI2 + c0 + 8 * mi + (l >> 2)] for l in L]
        else:
            a = [Cs[(a0 + k0 + (l & 3)) * LDI2 + c0 + 8 * mi + (l >> 2)] for l in L]
        for x in range(TPW):
            nj = nj0 + x
            if second:
                b = [Tb[(p * h + 8 * nj + (l >> 2)) * LDT2 + k0 + (l & 3)] for l in L]
            else:
                b = [Cs[(a0 + 8 * nj + (l >> 2)) * LDI2 + a0 + k0 + (l & 3)] for l in L]
            dmma(a, b, acc[x][0], acc[x][1])
    out = []
    for x in range(TPW):
        nj = nj0 + x
        for e in range(2):
            for l in L:
                g, t4 = l >> 2, l & 3
                if second:
                    out.append(('C', (a0 + 8 * nj + 2 * t4 + e) * LDI2 + c0 + 8 * mi + g, -acc[x][e][l]))
                else:
                    out.append(('T', (p * h + 8 * nj + 2 * t4 + e) * LDT2 + 8 * mi + g, acc[x][e][l]))
    return out


def invert2(Lm):
    """mirror of k_invert_diag2 up to the inverse in shared memory; returns X = L^-1 (nb x nb)"""
    nb = Lm.shape[0]
    Cs = np.full(128 * LDI2, 1e300); Tb = np.full(64 * LDT2, 1e300); rdiag = np.zeros(128)
    for k in range(128):
        for il in range(128):
            Cs[k * LDI2 + il] = Lm[il, k] if (k < nb and il < nb and il >= k) else (1.0 if (k == il and k >= nb) else 0.0)
    for t in range(128): rdiag[t] = 1.0 / Cs[t * LDI2 + t]
    for t in range(128):
        d0 = t & ~15; cc = t & 15; x = [0.0] * 16
        for i in range(16):
            a = 1.0 if i == cc else 0.0
            for k in range(i): a = -Cs[(d0 + k) * LDI2 + d0 + i] * x[k] + a
            x[i] = a * rdiag[d0 + i]
        for i in range(16): Tb[t * 16 + i] = x[i]
    for t in range(128):
        d0 = t & ~15
        for i in range(16): Cs[t * LDI2 + d0 + i] = Tb[t * 16 + i]
    for h in (16, 32, 64):
        TPW = h // 16; WPP = h // 4
        for second in (False, True):
            writes = []
            for warp in range(16):
                p = warp // WPP; widx = warp % WPP
                writes += _inv2_tiles(Cs, Tb, TPW, second, h, p, widx >> 1, (widx & 1) * TPW)
            seen = set()
            for kind, addr, val in writes:      # applied after the phase: the barrier
                assert (kind, addr) not in seen; seen.add((kind, addr))
                (Cs if kind == 'C' else Tb)[addr] = val
    X = np.zeros((nb, nb))
    for cc in range(nb):
        for r in range(cc, nb): X[r, cc] = Cs[cc * LDI2 + r]
    upper = max((abs(Cs[cc * LDI2 + r]) for cc in range(128) for r in range(cc)), default=0.0)
    assert upper == 0.0      # the outputs rely on a zero upper triangle
    return X
