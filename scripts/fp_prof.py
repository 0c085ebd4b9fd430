import sys
sys.path.insert(0, '.')
import scripts.bench_methods as B
B.run('mtor fp 256', B.series('mtor'), 'fp_glob_coup_nh_dbn', 256, 5, 2)
