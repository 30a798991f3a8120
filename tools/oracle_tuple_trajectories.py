#!/usr/bin/env python
"""CPU only: the oracle's PRIGP / CPLR end-to-end runs on ml-100k fold 1 (oracle/train_tuples.py, the reference drivers'
hyper-parameters, 50 epochs) beside the trajectories of the reference's own driver bodies on the TF-1.x stand-in
(tests/golden/e2e_{prigp,cplr}_refgraph_golden.json).  Prints, per checked epoch, ours / the reference's mean training loss
and the five metrics @100.

    python tools/oracle_tuple_trajectories.py [prigp|cplr] [seed] [epochs] > profiles/r5_oracle_tuple_trajectories.log
"""
import json
import os
import sys
import time

import numpy as np
from scipy.sparse import coo_matrix

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import train_tuples as tt  # noqa: E402

GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def ml100k():
    z = np.load(os.path.join(GOLDEN, 'ml100k_fold1.npz'))
    out = {}
    for part in ('tra', 'tst'):
        u, i, r = z[part + '_u'].astype(np.int64), z[part + '_i'].astype(np.int64), z[part + '_r']
        k = r > 3                                                                          # testprigp.py: matBinarize(R, 3)
        out[part] = coo_matrix((np.ones(int(k.sum()), dtype=np.float32), (u[k], i[k])), shape=(943, 1682)).tolil()
    return out


def main():
    which = sys.argv[1:2] and [sys.argv[1]] or ['prigp', 'cplr']
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    epochs = int(sys.argv[3]) if len(sys.argv) > 3 else 50
    data = ml100k()
    for w in which:
        gold = json.load(open(os.path.join(GOLDEN, 'e2e_%s_refgraph_golden.json' % w)))
        t0 = time.time()
        hist = tt.run(w, data['tra'], data['tst'], gold['hyper'], seed=seed, epochs=epochs, eval_epochs={1, 5, 10, 20, 30, 40, 50})
        ref = {x['epoch']: x for x in gold['history']}
        print('%s: oracle (seed %d) / reference driver body on the TF1 stand-in, %d epochs in %.0f s on the host'
              % (w.upper(), seed, epochs, time.time() - t0))
        for x in hist:
            r = ref[x['epoch']]
            print('  epoch %2d  TraLoss %9.3f / %9.3f (%+.2f %%)  ' % (x['epoch'], x['TraLoss'], r['TraLoss'], 100 * (x['TraLoss'] / r['TraLoss'] - 1))
                  + '  '.join('%s %.4f / %.4f' % (k, x[k], r[k]) for k in tt.NAMES))


if __name__ == '__main__':
    main()
