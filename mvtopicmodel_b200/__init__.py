"""mvtopicmodel_b200 -- B200-native collapsed-Gibbs sampling engine behind MVTopicModel's
FastQMVWVParallelTopicModel surface.  The compute path is libmvtm.so (hand-written sm_100a CUDA behind the
C ABI of include/mvtm.h); this package is the host-side mirror of the reference interface."""
from .engine import Engine, MvtmError  # noqa: F401
