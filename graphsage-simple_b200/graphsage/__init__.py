"""B200-native drop-in for zjzijielu/graphsage-simple's ``graphsage`` package.

Same import paths and call signatures as the reference for the sample-aggregate-update hot
path -- ``graphsage.aggregators.MeanAggregator``, ``graphsage.encoders.Encoder``,
``graphsage.model.SupervisedGraphSage`` -- with the work done by hand-written sm_100a CUDA
kernels behind the C ABI in include/gsage.h (libgsage_sm100.so).  There is no CPU path: the
modules raise if the extension or a CUDA device is missing.
"""
__all__ = ["aggregators", "encoders", "model", "graph", "engine"]
