"""kpop_b200 -- B200-native back end for KPop's k-mer spectrum counting stage (KPopCount).

The package holds only what that path needs:

* ``csrc/``            hand-written CUDA for sm_100a + the host engine + the C ABI (``include/kpopcount.h``),
                       built in-tree into ``libkpopcount_gpu.so`` and the ``bin/KPopCount`` executable;
* ``_native``          ctypes binding of the C ABI (fails loudly when the library or a GPU is missing);
* ``counter``          ``KMerCounter``: the host-side mirror of the reference's ``KMerCounter.compute`` functor
                       (bin/KPopCount.ml:20-64) on top of the C ABI;
* ``distributed``      one process per GPU: read-chunk sharding of one file / of a pair of files, the NCCL table
                       reduce, the sparse merge for large k, ``-L`` spectra gathered in record order.

There is no CPU fallback anywhere in this package.
"""
from .counter import Content, KMerCounter, KPopCountError, spectra_filename  # noqa: F401

__all__ = ["KMerCounter", "Content", "KPopCountError", "spectra_filename"]
