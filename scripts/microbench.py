import sys; sys.path.insert(0,'cnl-dyna_b200')
from cnld import _lib
print('dfma TF', _lib.bench_dfma(5))
for (m,n,k) in [(4096,4096,4096),(8192,8192,64),(14736,14736,64),(14736,14736,128),(14736,144,64),(8192,8192,1024)]:
    print((m,n,k), 'zgemm TF', _lib.bench_zgemm(m,n,k,3))
