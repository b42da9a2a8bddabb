import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'cnl-dyna_b200'))
from cnld import _lib
L = _lib.lib()
print('dfma TF', _lib.bench_dfma(5))
for v in (0, 1):
    L.cnld_b200_set_zgemm_3m(v)
    for (m, n, k) in [(14736, 14736, 64), (14736, 14736, 256), (14736, 14736, 512), (8192, 8192, 4096),
                      (14736, 144, 64), (14736, 448, 64), (448, 14288, 64)]:
        print('3M' if v else '4M', (m, n, k), 'zgemm "4M-equivalent" TF %.2f' % _lib.bench_zgemm(m, n, k, 3))
