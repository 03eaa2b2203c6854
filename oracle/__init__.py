"""CPU oracle for the GraphSAGE sample-aggregate-update hot path.

THIS PACKAGE IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it, and there only as the checker (or as the
thing timed as "the reference's CPU path"), never as something the CUDA product
path falls back to.  The product package (``graphsage-simple_b200/graphsage``)
never imports ``oracle``.

Contents
  ref_path.py      torch-CPU restatement of the reference's algorithm
                   (graphsage/aggregators.py:34-76, graphsage/encoders.py:40-62,
                   graphsage/model.py:52-69, 237-250 of zjzijielu/graphsage-simple).
                   (Layer, TwoLayerModel, StackedModel for any depth).
                   PINNED: checked against outputs of the unmodified reference imported
                   from /root/reference (tests/golden/*.npz, made by
                   tests/golden/make_golden.py) -- see tests/test_oracle_golden.py.
  philox.py        Philox4x32-10 in numpy (Salmon et al., SC'11) -- the counter-based
                   generator the CUDA sampler uses.
  sampler_port.py  CPU restatement of OUR CSR sampler / dedup specification
                   (include/gsage.h); the reference samples with CPython's Mersenne
                   Twister (aggregators.py:42-46), which no GPU kernel can replay, so
                   the sampler is bit-exact against this port and property-checked
                   against the reference's semantics (take-all when deg<k, k distinct
                   members when deg>=k, uniform marginals).
  exchange_port.py CPU restatement of OUR owner-bucketing specification for the partitioned
                   feature table / CSR (gs_bucket_by_owner); the reference has no
                   distributed code, so the pin is the identity "exchange == table[ids]".
"""
