import sys, os, json
sys.path.insert(0, '.')
sys.argv = ['x']
import scripts.bench_methods as B
yeast, mtor = B.series('yeast'), B.series('mtor')
syn40, _, _ = B.make_synthetic(40, 400, 3, 4, 5)
B.run('yeast fp 64', yeast, 'fp_glob_coup_nh_dbn', 64, 20, 5)
B.run('mtor fp 256', mtor, 'fp_glob_coup_nh_dbn', 256, 10, 3)
B.run('syn40 fp 64', [syn40], 'fp_glob_coup_nh_dbn', 64, 5, 2)
