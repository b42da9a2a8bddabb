import sys; sys.path.insert(0,'cnl-dyna_b200'); sys.path.insert(0,'oracle')
import numpy as np
from cnld import _lib as L
import pyoracle as po
rng = np.random.default_rng(0)
def rnd(m,n): return rng.standard_normal((m,n)) + 1j*rng.standard_normal((m,n))
for (m,n,k) in [(8,8,4),(37,1,37),(64,64,64),(130,70,9),(200,144,64),(300,300,300)]:
    A,B,C = rnd(m,k), rnd(k,n), rnd(m,n)
    R = L.zgemm(A,B,C,alpha=-1.0); ref = C - A@B
    R2 = L.zgemm(A,B); 
    print('zgemm',(m,n,k), np.abs(R-ref).max(), np.abs(R2-A@B).max())
for n in (37,64,100,200):
    G = rnd(n,n) + n*0.7*np.eye(n)
    LU = L.lrdecomp(G); LUo,_ = po.lrdecomp(G)
    print('lu', n, np.abs(LU-LUo).max())
    b = rnd(n,2)
    X = L.lrsolve(LUo, b); Xo = np.stack([po.lrsolve(LUo,b[:,j]) for j in range(2)],axis=1)
    print('solve', n, np.abs(X-Xo).max(), np.abs(Xo).max())
