"""Method switch and the per-method literals of the reference's regressors.

`Network.fit` (network.py:154-281) selects a regressor by string and passes it a method-specific `num_samples`;
every hyper-parameter is a literal inside that regressor's `fit()`.  This table
reproduces them so that the CUDA path is a drop-in; keyword arguments of `Network` / `DbnEngine.chain_init` may
override any of them (BASELINE config 4 needs fan-in 4).
"""
H_DBN, NH_DBN, SEQ_DBN, GLOB_DBN, FP_GLOB_DBN, VV_GLOB_DBN, FP_NH_DBN, FP_H_DBN, FP_VV_DBN, FP_SEQ_DBN = 0, 1, 2, 3, 4, 5, 6, 7, 8, 9
FP_METHODS = (FP_GLOB_DBN, FP_NH_DBN, FP_H_DBN, FP_VV_DBN, FP_SEQ_DBN)   # full parents: design width k = n, run with dbn_run_fp

# method strings of network.py:154-281 and the CLI aliases of examples/yeast_tests.py:179-187
METHOD_IDS = {
    'h_dbn': H_DBN, 'h-dbn': H_DBN,
    'varying_nh_dbn': NH_DBN, 'nh-dbn': NH_DBN,
    'seq_coup_nh_dbn': SEQ_DBN, 'seq-dbn': SEQ_DBN,
    'glob_coup_nh_dbn': GLOB_DBN, 'glob-dbn': GLOB_DBN,
    'fp_glob_coup_nh_dbn': FP_GLOB_DBN,   # network.py:245-257; CLI 'glob-dbn' of examples/full_parents_test.py:165-179
    'var_glob_coup_nh_dbn': VV_GLOB_DBN,  # network.py:258-269 (vvglobCoup.py); CLI 'var-glob-dbn' of examples/yeast_tests.py:188-189
    'var-glob-dbn': VV_GLOB_DBN,
    'fp_varying_nh_dbn': FP_NH_DBN,       # network.py:165-177 (fullParentsBpwLinReg.py); CLI 'nh-dbn' of examples/full_parents_test.py
    'fp_var_glob_coup_nh_dbn': FP_VV_DBN,  # network.py:270-281 (fpvvGlobCoup.py)
    'fp_seq_coup_nh_dbn': FP_SEQ_DBN,      # network.py:221-233 (fpSeqCoupBpwlinReg.py)
    'fp_h_dbn': FP_H_DBN,                 # network.py:200-209 (fpBayesianLinearRegression.py); CLI 'h-dbn' of the same script
    'fixed_nh_dbn': NH_DBN,               # network.py:178-189: the nh regressor on predefined changepoints (FIXED_CP below)
}
FIXED_CP = ('fixed_nh_dbn',)              # no changepoint move (bayesianPwLinearRegression.py:117), num_samples - 1
METHOD_NAMES = {H_DBN: 'h_dbn', NH_DBN: 'varying_nh_dbn', SEQ_DBN: 'seq_coup_nh_dbn', GLOB_DBN: 'glob_coup_nh_dbn',
                FP_GLOB_DBN: 'fp_glob_coup_nh_dbn', VV_GLOB_DBN: 'var_glob_coup_nh_dbn',
                FP_NH_DBN: 'fp_varying_nh_dbn', FP_H_DBN: 'fp_h_dbn', FP_VV_DBN: 'fp_var_glob_coup_nh_dbn',
                FP_SEQ_DBN: 'fp_seq_coup_nh_dbn'}
NOT_YET = ()


def method_id(method):
    if isinstance(method, int):
        return method
    if method in METHOD_IDS:
        return METHOD_IDS[method]
    if method in NOT_YET:
        raise NotImplementedError(
            "method %r of network.py:154-281 is outside the first-pass hot path (SURVEY.md section 8f)" % method)
    raise ValueError('unknown method %r' % (method,))


def run_params(method, N, fan_in=3, burn_in=0, thin=10, with_tau=True, **over):
    """dbn_run_params for `method` on N regression rows.

    num_samples / T: network.py:160 (nh: N), :195 (h: N+1; its ML gets N, bayesianLinearRegression.py:134-135),
    :216 (seq: N-1), :240 (glob: N), :185 (fixed nh: N-1).  sigma^2 ~ IG(0.005, 0.005) for nh and seq
    (bayesianPwLinearRegression.py:69-70, seqCoupledBayesianPwLinReg.py:44-45), IG(0.01, 0.01) for h and glob
    (bayesianLinearRegression.py:110-111, globCoupBayesianPwLinReg.py:43-44); lambda^2, delta^2 ~ IG(2, 0.2);
    changepoint prior p = 0.125 (priors.py:20); thinning 10 (network.py:359).
    """
    m = method_id(method)
    fixed = method in FIXED_CP
    p = dict(method=m, fan_in=fan_in, with_tau=int(bool(with_tau) and m not in (H_DBN, FP_H_DBN) and not fixed), burn_in=burn_in, thin=thin,
             a_lambda=2.0, b_lambda=0.2, a_delta=2.0, b_delta=0.2, cp_p=0.125)
    if m == H_DBN:
        p.update(num_samples=N, T_ml=float(N), T_sigma=float(N + 1), a_sigma=0.01, b_sigma=0.01)
    elif m == NH_DBN and fixed:   # network.py:182-188 passes num_samples - 1
        p.update(num_samples=N - 1, T_ml=float(N - 1), T_sigma=float(N - 1), a_sigma=0.005, b_sigma=0.005)
    elif m in (NH_DBN, FP_NH_DBN):   # fp: network.py:168-174, fullParentsBpwLinReg.py:73-74
        p.update(num_samples=N, T_ml=float(N), T_sigma=float(N), a_sigma=0.005, b_sigma=0.005)
    elif m == FP_H_DBN:   # network.py:203-207 passes N + 1; fpBayesianLinearRegression.py:70,84-85,90-91
        p.update(num_samples=N, T_ml=float(N), T_sigma=float(N + 1), a_sigma=0.01, b_sigma=0.01)
    elif m in (SEQ_DBN, FP_SEQ_DBN):   # fp: network.py:224-230, fpSeqCoupBpwlinReg.py:50-51
        p.update(num_samples=N - 1, T_ml=float(N - 1), T_sigma=float(N - 1), a_sigma=0.005, b_sigma=0.005)
    else:  # glob, fp-glob, var-glob: network.py:237-243, :249-255, :261-267; IG(0.01, 0.01) in all three
        # (fpGlobCoupBpwLinReg.py:49-50, vvglobCoup.py:42-43); var-glob uses the segment lengths T_h instead of T
        p.update(num_samples=N, T_ml=float(N), T_sigma=float(N), a_sigma=0.01, b_sigma=0.01)
    for k, v in over.items():
        if k not in p:
            raise TypeError('unknown run parameter %r' % k)
        p[k] = v
    return p
