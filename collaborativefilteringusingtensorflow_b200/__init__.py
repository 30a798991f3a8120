"""B200-native pairwise-ranking training + full-catalog top-K evaluation behind the API of
BinFuPKU/CollaborativeFilteringUsingTensorflow (models/pl/models/{bprmf,cml,gbprmf}.py, models/basic/models/{wrmf,mf,svd,pop}.py,
samplers/sampler_{ranking,uij_ranking,gbpr,rating}.py, metrics/{ranking,rating}.py, utils/{IOUtil,Util}.py).

All arithmetic runs in hand-written sm_100a CUDA (libcf_b200.so, C ABI in include/cf_b200.h); there is no CPU fallback.
"""
__version__ = '0.1.0'


def __getattr__(name):   # lazy: importing the package must work on a box without a GPU (build / symbol checks)
    if name == 'BPRMF':
        from .models.pl.models.bprmf import BPRMF
        return BPRMF
    if name == 'CML':
        from .models.pl.models.cml import CML
        return CML
    if name == 'GBPRMF':
        from .models.pl.models.gbprmf import GBPRMF
        return GBPRMF
    if name == 'WRMF':
        from .models.basic.models.wrmf import WRMF
        return WRMF
    if name == 'PopRank':
        from .models.basic.models.pop import PopRank
        return PopRank
    if name == 'SVD':
        from .models.basic.models.svd import SVD
        return SVD
    if name == 'PRIGP':
        from .models.pl.models.prigp import PRIGP
        return PRIGP
    if name == 'CPLR':
        from .models.pl.models.cplr_u import CPLR
        return CPLR
    if name == 'ItemCF':
        from .models.basic.models.itemcf import ItemCF
        return ItemCF
    if name == 'UserCF':
        from .models.basic.models.usercf import UserCF
        return UserCF
    if name == 'MF':
        from .models.basic.models.mf import MF
        return MF
    raise AttributeError(name)
