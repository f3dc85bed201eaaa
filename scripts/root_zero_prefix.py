"""How much tcgen05 work is structurally zero inside the amalgamated root supernode? (CPU only, analyze_only handle)

Inside the root every row r is dense from first(r) -- the column where the elimination-tree path of one of its neighbours enters
the root -- to r, so the zero entries of a 128-row block form a PREFIX of the K range of its tcgen05 tasks.  Prints the structural
density of the root and the fraction of 128x128x128 tile products that survive when every task starts at
max(kstart[row block], kstart[column block]).  Measured: config 2 keeps 99.7 % (nothing to gain), config T 95.4 %.
usage: python scripts/root_zero_prefix.py [2|T]
"""
import sys, time, numpy as np, scipy.sparse as sp
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import tlpb200_loader
pkg = tlpb200_loader.load()
from tulip_jl_b200 import lpgen
cfg = sys.argv[1] if len(sys.argv)>1 else '2'
lp = lpgen.config('T' if cfg=='T' else int(cfg))
A = lp.A.tocsc(); m,n = A.shape
t=time.time()
k = pkg.setup(A, pkg.K1(), pkg.Backend(analyze_only=True))
sym = k.symbolic()
print('setup', round(time.time()-t,1),'s; keys', list(sym.keys()))
perm = np.asarray(sym['perm']); parent = np.asarray(sym['parent']).astype(np.int64)
sn_first = np.asarray(sym['sn_first'] if 'sn_first' in sym else sym['super'])
N = len(perm)
ncols = np.diff(sn_first)
big = int(np.argmax(ncols)); r0 = int(sn_first[big]); r1 = int(sn_first[big+1])
print('N',N,'largest supernode cols',r0,r1,r1-r0,'is last',r1==N)
# permuted pattern of A A'
iperm = np.empty(N,np.int64); iperm[perm]=np.arange(N)
P = sp.csr_matrix((np.ones(A.nnz,np.int8),(iperm[A.indices], np.repeat(np.arange(n),np.diff(A.indptr)))),shape=(m,n))
S = (P.astype(np.float32) @ P.T.astype(np.float32)).tocsr()
S.sort_indices()
# entry point of every node's etree path into the root supernode
entry = np.arange(N,dtype=np.int64)
par = parent.copy(); par[par<0] = N-1
for it in range(64):
    mask = entry < r0
    if not mask.any(): break
    entry[mask] = par[entry[mask]]
print('pointer steps',it)
first = np.full(N, N, np.int64)
indptr, ind = S.indptr, S.indices
rows = np.repeat(np.arange(N), np.diff(indptr))
lowmask = (ind <= rows) & (rows >= r0)
np.minimum.at(first, rows[lowmask], entry[ind[lowmask]])
first_root = first[r0:r1] - r0            # first nonzero column (relative) of every root row
rel = np.arange(r1-r0)
assert (first_root <= rel).all()
nnz_struct = int((rel - first_root + 1).sum()); full = (r1-r0)*(r1-r0+1)//2
print('structural nnz of the root rows inside the root', nnz_struct, 'of', full, '=', round(nnz_struct/full,4))
# block-level: kstart(rb) = min first over the 128-row block, in units of 128-column pieces
nbk = (r1-r0+127)//128
kstart = np.array([first_root[b*128:(b+1)*128].min()//128 for b in range(nbk)])
print('kstart quantiles', np.percentile(kstart,[0,10,25,50,75,90,100]))
# tcgen05 work: sum over rb >= cb > k (k < cb) of 1; with skipping: k >= max(kstart[rb], kstart[cb])
tot=0; kept=0
for rb in range(nbk):
    cb = np.arange(0, rb+1)
    ks = np.maximum(kstart[rb], kstart[cb])
    tot += int(cb.sum())                 # pieces k = 0..cb-1
    kept += int(np.maximum(cb - ks, 0).sum())
print('tile-products total', tot, 'kept with zero-prefix skipping', kept, 'fraction', round(kept/tot,4))
