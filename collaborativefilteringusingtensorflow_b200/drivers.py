"""The reference's experiment drivers for the hot-path models, as one module.

Mirrors reference src/models/pl/testbprmf.py:19-125, testcml.py:19-102, testgbprmf.py:19-113, testprigp.py:17-109,
testcplr_u.py:17-126, src/models/basic/testwrmf.py:19-96, testicf.py:14-115, testucf.py:14-116, testpop.py:14-115 and (rating prediction, `mf` /
`svd`) src/models/basic/testmf.py:14-78, testsvd.py:14-79: the same module-level hyper-parameters, the same per-fold worker (load
``ratings__<fold>_tra.txt`` / ``_tst.txt``, binarise with ``rating > 3``, build the sampler and the model, train, print the
fold's scores) and the same ``ave`` / ``std`` summary.  The reference wraps every fold in a ``multiprocessing.Pool`` only
because ``tf.get_variable`` names collide (testbprmf.py:114-117); here folds run in-process.

    python -m collaborativefilteringusingtensorflow_b200.drivers bprmf <dataset_dir>/ 943 1682 [--folds 5] [--max-iter N]
"""
import argparse

import numpy as np
from scipy.sparse import lil_matrix

from .models.basic.models.itemcf import ItemCF
from .models.basic.models.mf import MF
from .models.basic.models.pop import PopRank
from .models.basic.models.svd import SVD
from .models.basic.models.usercf import UserCF
from .models.basic.models.wrmf import WRMF
from .models.pl.models.bprmf import BPRMF
from .models.pl.models.cml import CML
from .models.pl.models.cplr_u import CPLR
from .models.pl.models.gbprmf import GBPRMF
from .models.pl.models.prigp import PRIGP
from .samplers import sampler_gbpr, sampler_ranking, sampler_rating
from .utils.IOUtil import loadSparseR
from .utils.Util import matBinarize

binarize_threshold = 3
eval_metrics = ['pre', 'recall', 'map', 'mrr', 'ndcg']
split_method = 'cv'

# hyper-parameters exactly as the reference drivers set them
HYPER = {
    'bprmf': dict(reg=.1, topN=10, n_factors=100, batch_size=100, negSample=1),                                   # testbprmf.py:21-30
    'cml': dict(margin=1., reg_cov=1., use_rank_weight=True, clip_norm=1.0, topN=10, n_factors=50, batch_size=50,
                negSample=5),                                                                                     # testcml.py:22-34
    'gbprmf': dict(gsize=1, rho=.4, reg=.01, topN=100, n_factors=100, batch_size=100, negSample=5),               # testgbprmf.py:23-32
    'wrmf': dict(weight=2., reg=.1, topN=10, negRatio=1, n_factors=100, batch_size=100),                          # testwrmf.py:22-30
    'prigp': dict(topK=5, alpha=10, reg=.1, topN=100, n_factors=100, batch_size=1000),                            # testprigp.py:21-31
    'cplr': dict(topK=200, reg=.1, topN=100, alpha=1., beta=1., gamma=1., n_factors=100, batch_size=100),         # testcplr_u.py:21-33
    'itemcf': dict(topK=5, topN=100),                                                                             # testicf.py:18-22
    'usercf': dict(topK=5, topN=100),                                                                             # testucf.py:18-22
    'pop': dict(topN=100),                                                                                        # testpop.py:17
}


# basic/testmf.py:14-24 (rating prediction: raw ratings, no binarisation)
MF_HYPER = dict(eval_metrics=['rmse', 'mae', 'mse'], reg=.1, range_of_ratings=(1, 5), n_factors=100, batch_size=1000)
# basic/testsvd.py:14-24 (the reference runs one fold only, testsvd.py:66)
SVD_HYPER = dict(eval_metrics=['rmse', 'mae', 'mse'], reg=.1, range_of_ratings=(1, 5), n_factors=32, batch_size=100)


def _triads(sR):
    """testmf.py:32: (user, item, rating) rows of the non-zeros in row-major order."""
    coo = sR.tocoo()
    order = np.lexsort((coo.col, coo.row))
    return np.stack([coo.row[order], coo.col[order], coo.data[order]], 1).astype(np.float64)


def worker_mf(fold, n_users, n_items, dataset_dir, max_iter=None, seed=None, verbose=True, cls=MF, h=None):
    """testmf.py:27-50 / testsvd.py:27-51."""
    h = h or MF_HYPER
    trasR = loadSparseR(n_users, n_items, dataset_dir + 'ratings__' + str(fold + 1) + '_tra.txt')
    print(dataset_dir.split('/')[-2] + ':', trasR.shape, trasR.nnz, '%.2f' % (trasR.nnz / float(trasR.shape[0])))
    tra_tuple = _triads(trasR)
    tst_tuple = _triads(loadSparseR(n_users, n_items, dataset_dir + 'ratings__' + str(fold + 1) + '_tst.txt'))
    sampler = sampler_rating.Sampler(trasR=trasR, negRatio=.0, batch_size=h['batch_size'], seed=seed or 0)
    it = {} if max_iter is None else dict(max_iter=max_iter)
    mf = cls(n_users, n_items, h['eval_metrics'], h['range_of_ratings'], h['reg'], h['n_factors'], h['batch_size'],
             seed=seed, verbose=verbose, **it)
    scores = mf.train(fold + 1, tra_tuple, tst_tuple, sampler)
    print('fold=%d:' % fold, ','.join(['%s' % m for m in h['eval_metrics']]), '=', ','.join(['%.6f' % s for s in scores]))
    mf.close()
    return scores


def worker(model_name, fold, n_users, n_items, dataset_dir, max_iter=None, seed=None, verbose=True):
    if model_name == 'mf':
        return worker_mf(fold, n_users, n_items, dataset_dir, max_iter, seed, verbose)
    if model_name == 'svd':
        return worker_mf(fold, n_users, n_items, dataset_dir, max_iter, seed, verbose, cls=SVD, h=SVD_HYPER)
    h = HYPER[model_name]
    trasR = lil_matrix(matBinarize(loadSparseR(n_users, n_items, dataset_dir + 'ratings__' + str(fold + 1) + '_tra.txt'),
                                   binarize_threshold))
    print(dataset_dir.split('/')[-2] + '@%d:' % (fold + 1), trasR.shape, trasR.nnz, '%.2f' % (trasR.nnz / float(trasR.shape[0])))
    tstsR = lil_matrix(matBinarize(loadSparseR(n_users, n_items, dataset_dir + 'ratings__' + str(fold + 1) + '_tst.txt'),
                                   binarize_threshold))
    kw = dict(seed=seed, verbose=verbose)
    it = {} if max_iter is None else dict(max_iter=max_iter)
    if model_name == 'bprmf':
        sampler = sampler_ranking.Sampler(trasR=trasR, n_neg=h['negSample'], batch_size=h['batch_size'], seed=seed or 0)
        model = BPRMF(n_users, n_items, h['topN'], split_method, eval_metrics, h['reg'], h['n_factors'], h['batch_size'], **it, **kw)
    elif model_name == 'cml':
        sampler = sampler_ranking.Sampler(trasR, n_neg=h['negSample'], batch_size=h['batch_size'], seed=seed or 0)
        model = CML(n_users, n_items, h['topN'], split_method, eval_metrics, h['reg_cov'], h['margin'], h['use_rank_weight'],
                    h['clip_norm'], h['n_factors'], h['batch_size'], **it, **kw)
    elif model_name == 'gbprmf':
        sampler = sampler_gbpr.Sampler(trasR, h['gsize'], h['negSample'], h['batch_size'], seed=seed or 0)
        model = GBPRMF(n_users, n_items, h['topN'], h['rho'], h['gsize'], split_method, eval_metrics, h['reg'], h['n_factors'],
                       h['batch_size'], **it, **kw)
    elif model_name == 'wrmf':
        sampler = sampler_rating.Sampler(trasR, h['negRatio'], h['batch_size'], seed=seed or 0)
        model = WRMF(n_users, n_items, h['topN'], split_method, eval_metrics, h['weight'], h['reg'], h['n_factors'],
                     h['batch_size'], **it, **kw)
    elif model_name in ('itemcf', 'usercf', 'pop'):      # testicf.py:37-38, testpop.py:33-34: no sampler, no epochs
        if model_name == 'pop':
            model = PopRank(n_users, n_items, h['topN'], split_method, eval_metrics)
        else:
            model = (ItemCF if model_name == 'itemcf' else UserCF)(n_users, n_items, h['topK'], h['topN'], split_method, eval_metrics)
        scores = model.train(fold + 1, trasR, tstsR)
        print(dataset_dir.split('/')[-2] + '@%d:' % (fold + 1),
              ','.join(['%s' % m for m in eval_metrics]) + '@%d=' % h['topN'] + ','.join(['%.6f' % s for s in scores]))
        return scores
    elif model_name == 'prigp':                   # testprigp.py:44-48: the model builds its own sampler
        sampler = None
        model = PRIGP(n_users, n_items, h['topK'], h['topN'], split_method, eval_metrics, h['alpha'], h['reg'], h['n_factors'],
                      h['batch_size'], **it, **kw)
    elif model_name == 'cplr':                    # testcplr_u.py:49-53
        sampler = None
        model = CPLR(n_users, n_items, h['topK'], h['topN'], split_method, eval_metrics, h['alpha'], h['beta'], h['gamma'], h['reg'],
                     h['n_factors'], h['batch_size'], **it, **kw)
    else:
        raise ValueError('unknown model %r' % model_name)
    scores = model.train(fold + 1, trasR, tstsR, sampler)
    print(dataset_dir.split('/')[-2] + '@%d:' % (fold + 1),
          ','.join(['%s' % m for m in eval_metrics]) + '@%d=' % h['topN'] + ','.join(['%.6f' % s for s in scores]))
    model.close()
    return scores


def run(model_name, dataset_dir, n_users, n_items, folds=5, max_iter=None, seed=None, verbose=True):
    """testbprmf.py:55-125: all folds, then ``ave`` / ``std`` over the folds."""
    results = [worker(model_name, fold, n_users, n_items, dataset_dir, max_iter, seed, verbose) for fold in range(folds)]
    results = np.array(results)
    if model_name in ('mf', 'svd'):     # testmf.py:74-77
        aves = results.sum(0) / len(results)
        stds = np.sqrt(np.power(results - aves, 2).sum(0) / len(results))
        print('ave=[' + ','.join(['%.4f' % a for a in aves]) + ']', 'std=[' + ','.join(['%.4f' % d for d in stds]) + ']')
        return results
    topN = HYPER[model_name]['topN']
    print('ave:', ','.join(['%s' % m for m in eval_metrics]) + '@%d=' % topN + ','.join(['%.6f' % s for s in results.mean(0)]))
    print('std:', ','.join(['%s' % m for m in eval_metrics]) + '@%d=' % topN + ','.join(['%.6f' % s for s in results.std(0)]))
    return results


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('model', choices=sorted(HYPER) + ['mf', 'svd'])
    ap.add_argument('dataset_dir', help="directory with ratings__<k>_tra.txt / ratings__<k>_tst.txt (trailing '/')")
    ap.add_argument('n_users', type=int)
    ap.add_argument('n_items', type=int)
    ap.add_argument('--folds', type=int, default=5)
    ap.add_argument('--max-iter', type=int, default=None)
    ap.add_argument('--seed', type=int, default=None)
    a = ap.parse_args()
    run(a.model, a.dataset_dir if a.dataset_dir.endswith('/') else a.dataset_dir + '/', a.n_users, a.n_items, a.folds, a.max_iter, a.seed)
