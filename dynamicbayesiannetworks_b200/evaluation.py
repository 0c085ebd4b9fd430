"""Evaluation / reporting helpers the reference's example drivers call on a finished `Network`
(examples/utils.py:24-131, :226-261): `transformResults`, `adjMatrixRoc` (ROC and precision-recall AUC) and
`save_chain` / `load_chain`.  Same names, argument order and return values; the two AUCs are computed with NumPy (same
definitions as sklearn's `roc_curve` / `precision_recall_curve` + trapezoidal `auc`, which the reference calls) and
returned instead of plotted -- there is no matplotlib on the GPU boxes.
"""
import os
import pickle
from pprint import pprint

import numpy as np


def transformResults(trueAdjMatrix, adjMatrixProp):
    """examples/utils.py:24-69: drop the diagonals, fold the lag-2 columns onto the lag-1 ones with an element-wise max
    when the proposed rows are wider than the true ones, flatten.  Returns (binary truth, scores)."""
    true_nd = [[v for j, v in enumerate(row) if j != i] for i, row in enumerate(trueAdjMatrix)]
    prop_nd = [[v for j, v in enumerate(row) if j != i] for i, row in enumerate(adjMatrixProp)]
    if len(true_nd[0]) < len(prop_nd[0]):
        d = len(trueAdjMatrix[0]) - 1
        prop_nd = [np.max([np.asarray(row[:d]), np.asarray(row[d:])], axis=0).tolist() for row in prop_nd]
    flattened_true = [1 if item else 0 for sub in true_nd for item in sub]
    flattened_scores = [item for sub in prop_nd for item in sub]
    return flattened_true, flattened_scores


def _binary_clf_curve(y_true, y_score):
    y_true = np.asarray(y_true, dtype=float)
    y_score = np.asarray(y_score, dtype=float)
    order = np.argsort(-y_score, kind='mergesort')
    y_true, y_score = y_true[order], y_score[order]
    idx = np.r_[np.nonzero(np.diff(y_score))[0], y_true.size - 1]       # last index of every distinct score
    tps = np.cumsum(y_true)[idx]
    fps = 1 + idx - tps
    return fps, tps, y_score[idx]


def roc_auc(y_true, y_score):
    """Area under the ROC curve (sklearn.metrics.roc_curve + auc, examples/utils.py:113-115)."""
    fps, tps, _ = _binary_clf_curve(y_true, y_score)
    fps, tps = np.r_[0, fps], np.r_[0, tps]
    fpr, tpr = fps / fps[-1], tps / tps[-1]
    return float(np.trapezoid(tpr, fpr)) if hasattr(np, 'trapezoid') else float(np.trapz(tpr, fpr))


def pr_auc(y_true, y_score):
    """Area under the precision-recall curve (sklearn.metrics.precision_recall_curve + auc, examples/utils.py:86-89)."""
    fps, tps, _ = _binary_clf_curve(y_true, y_score)
    precision = tps / (tps + fps)
    recall = tps / tps[-1]
    # sklearn reverses the curve and appends the (recall 0, precision 1) end point
    precision, recall = np.r_[precision[::-1], 1.0], np.r_[recall[::-1], 0.0]
    area = np.trapezoid(precision, recall) if hasattr(np, 'trapezoid') else np.trapz(precision, recall)
    return float(-area)


def adjMatrixRoc(adjMatrixProp, trueAdjMatrix, verbose):
    """examples/utils.py:81-90: takes the FLATTENED scores and truth (`transformResults`' outputs, in the reference's
    argument order); prints what the reference prints and returns {'roc_auc', 'pr_auc'} instead of saving figures."""
    if verbose:
        print('The true adj matrix is: ')
        pprint(trueAdjMatrix)
        print('The proposed adj matrix is: ')
        pprint(adjMatrixProp)
    out = {'roc_auc': roc_auc(trueAdjMatrix, adjMatrixProp), 'pr_auc': pr_auc(trueAdjMatrix, adjMatrixProp)}
    print('The AuC of the PR curve was: ', out['pr_auc'])
    return out


def save_chain(filename, network_object, folder='./output/'):
    """examples/utils.py:226-240: pickle the Network object into ./output/."""
    os.makedirs(folder, exist_ok=True)
    with open(os.path.join(folder, filename), 'wb') as f:
        pickle.dump(network_object, f)


def load_chain(filename, folder='./output/'):
    """examples/utils.py:242-261."""
    try:
        with open(os.path.join(folder, filename), 'rb') as f:
            network = pickle.load(f)
        print('Obj succesfully loaded.')
        return network
    except Exception:
        print('No file was loaded, please try again.')
        return None
