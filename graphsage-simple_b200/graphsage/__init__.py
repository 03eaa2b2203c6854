"""B200-native drop-in for zjzijielu/graphsage-simple's ``graphsage`` package.

Same import paths and call signatures as the reference for the sample-aggregate-update hot
path -- ``graphsage.aggregators.MeanAggregator``, ``graphsage.encoders.Encoder``,
``graphsage.model.SupervisedGraphSage`` -- with the work done by hand-written sm_100a CUDA
kernels behind the C ABI in include/gsage.h (libgsage_sm100.so).  There is no CPU path: the
modules raise if the extension or a CUDA device is missing.

  aggregators / encoders / model   the reference's classes + ``run_model`` / ``main`` (python -m graphsage.model)
  data                             the reference's loaders and node-feature initialisers (host side)
  graph / sampling                 device CSR container, counter-based sampler state
  engine                           fused, CUDA-graph-captured, software-pipelined 2-layer train step
  dist / sharded                   data parallel (fused peer all-reduce + SGD) and partitioned table + CSR
  ops / functional / _native       tensor-level wrappers, autograd Functions, ctypes binding of the C ABI
"""
__all__ = ["aggregators", "encoders", "model", "graph", "engine", "data", "dist", "sharded", "sampling", "ops"]
