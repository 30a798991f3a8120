"""Host side of the on-device samplers (csrc/cf_sampler.cu).

The reference samplers (reference src/samplers/sampler_*.py) run one numpy producer thread behind a bounded queue and
hand out one minibatch per ``next_batch()`` call.  Here a batch is a pure function of (seed, epoch, batch index): the
sampler object is only a cursor.  ``next_batch()`` keeps the reference's return types; ``next_chunk(n)`` is the fast
path the model classes use (n minibatches as CUDA tensors, no host round trip).
"""
import os

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
        self._pair_set = None         # hash set of the training pairs, built on first use (see _membership)

    # -- membership structure of the negatives' rejection test -----------------------------------------
    PAIR_SET_MIN_NNZ = 1 << 20        # below this the sampler is not worth 16 bytes per interaction
    PAIR_SET_MAX_BYTES = 16 << 30

    def _membership(self, a):
        """Builds (once) the open-addressing set of all training pairs and points the sampler at it: a negative's
        membership test is then one 32-byte sector instead of the 3-4 cold sectors at the bottom of a bisection of the user's
        CSR row.  Measured on configs[1]'s interactions, per minibatch of 2^20 pairs: W = 1 0.132 -> 0.076 ms, W = 5 0.229 -> 0.218 ms
        (five lanes of a pair share the top of the bisection, and at W = 5 the kernel is bound by its Philox / Feistel arithmetic).
        Same draws, same answers.  CF_SAMPLER_PAIR_SET=0 keeps the bisection."""
        if self._pair_set is None:
            self._pair_set = False
            nnz = int(self.train.nnz)
            bits = int(self.lib.cf_pair_set_bits(nnz))
            nbytes = 8 << bits
            if (os.environ.get('CF_SAMPLER_PAIR_SET', '1') != '0' and nnz >= self.PAIR_SET_MIN_NNZ
                    and nbytes <= self.PAIR_SET_MAX_BYTES):
                free = self.torch.cuda.mem_get_info(self.device)[0]
                if nbytes < free // 2:
                    table = self.torch.empty(1 << bits, dtype=self.torch.int64, device=self.device)
                    csr = self.train.as_c(True)
                    _lib.check(self.lib.cf_pair_set_build(csr, _lib.ptr(table), bits, self._stream()), 'cf_pair_set_build')
                    self.torch.cuda.current_stream(self.device).synchronize()    # (built once; later launches may come from any stream)
                    self.launches += 1
                    self._pair_set = (table, bits)
        if self._pair_set:
            a.pair_set, a.pair_set_bits = _lib.ptr(self._pair_set[0]), self._pair_set[1]

    # -- output buffers ---------------------------------------------------------------------------
    ring_slot = None      # set by the epoch loop of the model classes: next_chunk then writes into persistent buffers of
                          # that slot instead of fresh tensors (an allocation per minibatch on a side stream makes the
                          # caching allocator synchronise now and then: single minibatches came out 2-3x slower)

    def _empty(self, key, shape, dtype):
        if self.ring_slot is None:
            return self.torch.empty(*shape, dtype=dtype, device=self.device)
        pool = self.__dict__.setdefault('_ring', {})
        t = pool.get((self.ring_slot, key))
        if t is None or tuple(t.shape) != tuple(shape) or t.dtype != dtype:
            t = pool[(self.ring_slot, key)] = self.torch.empty(*shape, dtype=dtype, device=self.device)
        return t

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
        self._membership(a)
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
