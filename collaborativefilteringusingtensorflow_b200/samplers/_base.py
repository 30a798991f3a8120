"""Host side of the on-device samplers (csrc/cf_sampler.cu).

The reference samplers (reference src/samplers/sampler_*.py) run one numpy producer thread behind a bounded queue and
hand out one minibatch per ``next_batch()`` call.  Here a batch is a pure function of (seed, epoch, batch index): the
sampler object is only a cursor.  ``next_batch()`` keeps the reference's return types; ``next_chunk(n)`` is the fast
path the model classes use (n minibatches as CUDA tensors, no host round trip).
"""
from .. import _lib
from ..engine import resolve_device
from ..sparse import DeviceCSR, null_csr


class DeviceSamplerBase(object):
    _HOST_CHUNK = 64      # minibatches generated per launch when serving next_batch() one by one

    def __init__(self, trasR, batch_size, seed=0, device='GPU', with_values=False):
        self.torch = _lib.require_cuda()
        self.lib = _lib.lib()
        self.device = resolve_device(device)
        self.train = trasR if isinstance(trasR, DeviceCSR) else DeviceCSR.from_scipy(trasR, self.device, with_values)
        self.n_users, self.n_items = self.train.shape
        self.batch_size = int(batch_size)
        if self.batch_size <= 0:
            raise ValueError('batch_size must be positive')
        self.batches_per_epoch = int(self.train.nnz / self.batch_size)      # sampler_ranking.py:25 -- tail dropped
        if self.batches_per_epoch == 0:
            raise ValueError('fewer training pairs (%d) than batch_size (%d)' % (self.train.nnz, self.batch_size))
        self.seed = int(seed) & (2 ** 64 - 1)
        self.epoch, self.batch = 0, 0
        self.flags = self.torch.zeros(1, dtype=self.torch.int32, device=self.device)
        self._host_cache = []
        self.launches = 0

    # -- cursor ---------------------------------------------------------------------------------
    def seek(self, epoch, batch=0):
        self.epoch, self.batch = int(epoch), int(batch)
        self._host_cache = []

    def _segments(self, n):
        """Split n minibatches starting at the cursor into per-epoch (epoch, batch0, count) pieces; advances."""
        segs = []
        while n > 0:
            take = min(n, self.batches_per_epoch - self.batch)
            segs.append((self.epoch, self.batch, take))
            self.batch += take
            n -= take
            if self.batch == self.batches_per_epoch:
                self.epoch, self.batch = self.epoch + 1, 0
        return segs

    def _args(self, epoch, batch0, count):
        a = _lib.SampleArgs()
        a.train = self.train.as_c(True)
        a.train_t = null_csr()
        a.seed, a.epoch, a.batch0, a.n_batches, a.B = self.seed, epoch, batch0, count, self.batch_size
        a.flags = _lib.ptr(self.flags)
        return a

    def _stream(self):
        return self.torch.cuda.current_stream(self.device).cuda_stream

    def check_flags(self):
        if int(self.flags.item()) & _lib.FLAG_SAMPLER_GAVEUP:
            self.flags.zero_()
            raise RuntimeError('a user has (almost) every item as a positive: no negative could be drawn '
                               '(the reference sampler spins forever here, sampler_ranking.py:35)')

    # -- reference API ----------------------------------------------------------------------------
    def next_batch(self):
        if not self._host_cache:
            n = min(self._HOST_CHUNK, self.batches_per_epoch)
            self._host_cache = self._to_host_batches(self.next_chunk(n), n)
            self.check_flags()
        return self._host_cache.pop(0)
