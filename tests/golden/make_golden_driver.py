#!/usr/bin/env python
"""Golden end-to-end numbers of the UNMODIFIED reference driver (`run_model`,
/root/reference/graphsage/model.py:184-259) on synthetic planted-partition datasets written in the
reference's file formats by graphsage.data.write_synthetic_dataset: validation F1 (micro / macro) per
seed.  The B200 driver (graphsage.model.run_model(as_run=True)) must land within run-to-run noise of
these (tests/test_gpu_driver.py).  Writes tests/golden/driver.json.

    python tests/golden/make_golden_driver.py        # build container only (needs /root/reference)
"""
import contextlib
import io
import json
import os
import random
import re
import sys
import tempfile
import time
import warnings

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "graphsage-simple_b200"))
warnings.filterwarnings("ignore")

RUNS = [("cora", "None", [1, 2, 3, 4, 5], 5), ("citeseer", "node_degree", [1, 2, 3], 5), ("cora", "random_normal", [1, 2, 3], 5)]
DATA_SEED = {"cora": 21, "citeseer": 22}


def main():
    import importlib.util
    import torch
    from graphsage import data as D
    _orig = random.sample
    random.sample = lambda pop, k: _orig(tuple(pop) if isinstance(pop, (set, frozenset)) else pop, k)
    for name in ("aggregators", "encoders"):
        spec = importlib.util.spec_from_file_location("graphsage_ref." + name, "/root/reference/graphsage/%s.py" % name)
        mod = importlib.util.module_from_spec(spec)
        sys.modules["graphsage_ref." + name] = mod
        spec.loader.exec_module(mod)
    src = open("/root/reference/graphsage/model.py").read()
    src = src.replace("from graphsage.encoders", "from graphsage_ref.encoders").replace(
        "from graphsage.aggregators", "from graphsage_ref.aggregators")
    ref = type(sys)("graphsage_ref.model")
    exec(compile(src, "/root/reference/graphsage/model.py", "exec"), ref.__dict__)
    out = {"data_seed": DATA_SEED, "runs": []}
    with tempfile.TemporaryDirectory() as tmp:
        for ds, seed in DATA_SEED.items():
            D.write_synthetic_dataset(ds, tmp, seed=seed)
        cwd = os.getcwd()
        os.chdir(tmp)
        try:
            for ds, init, seeds, epochs in RUNS:
                for seed in seeds:
                    torch.manual_seed(seed)              # the reference leaves torch unseeded (weight init)
                    buf = io.StringIO()
                    t0 = time.time()
                    with contextlib.redirect_stdout(buf):
                        ref.run_model(ds, init, seed, epochs)
                    txt = buf.getvalue()
                    f1 = [float(x) for x in re.findall(r"Validation F1 (?:micro|macro): ([0-9.eE+-]+)", txt)]
                    bt = float(re.findall(r"Average batch time: ([0-9.eE+-]+)", txt)[0])
                    out["runs"].append({"dataset": ds, "initializer": init, "seed": seed, "epochs": epochs,
                                        "f1_micro": f1[0], "f1_macro": f1[1], "avg_batch_time_cpu": bt})
                    print(ds, init, seed, f1, "%.1fs" % (time.time() - t0))
        finally:
            os.chdir(cwd)
    json.dump(out, open(os.path.join(HERE, "driver.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
