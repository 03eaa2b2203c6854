"""The TIMED path -- tcgen05 layer-1 GEMMs + fused head + the fused engine, hidden width 128/128 -- pinned on outputs
of the UNMODIFIED reference at the BASELINE widths (F = 602 / 500 / 3703): tests/golden/bs_*.npz, made by
tests/golden/make_golden_baseline_shapes.py from /root/reference (graphsage/encoders.py:47-61,
graphsage/model.py:245-250).  The replayed neighbour tiles go through the ENGINE (not the op-by-op path), and the
test asserts that the engine really took the tensor-core kernels and the fused head."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import baseline_shapes as S          # noqa: E402

from test_gpu_modules import REL, build_model, relerr, tiles_to_adj      # noqa: E402

pytestmark = pytest.mark.gpu


def _model(name, golden):
    g, p, x = golden(name), S.CASES[name], S.inputs(name)
    gg = dict(table=x["table"], w1=x["w1"], w2=x["w2"], wc=x["wc"])
    model, enc1, enc2 = build_model(gg, p["gcn"], tiles_to_adj(g["hop1"], g["idx1"], g["cnt1"]),
                                    tiles_to_adj(x["nodes"], g["idx2"], g["cnt2"]), None, None)
    return g, p, x, model, enc1, enc2


@pytest.mark.parametrize("name", list(S.CASES))
def test_engine_step_matches_reference_at_baseline_widths(golden, name):
    g, p, x, model, enc1, enc2 = _model(name, golden)
    nodes, labels = list(x["nodes"]), x["labels"][x["nodes"]]
    cols = slice(None, None, S.GW1_COLS) if name == "bs_citeseer" else slice(None)
    opt = torch.optim.SGD(filter(lambda q: q.requires_grad, model.parameters()), lr=0.7)      # model.py:237
    opt.zero_grad()
    loss = model.loss(nodes, torch.LongTensor(labels))                                        # model.py:247-248
    eng = model._engine
    assert eng is not None and eng.tc1 and eng.head, (eng.tc1, eng.head)      # tcgen05 GEMMs + fused head took the step
    loss.backward()
    assert abs(loss.item() - float(g["loss"])) / abs(float(g["loss"])) < REL
    assert relerr(eng.logits[:p["b"]].cpu().numpy(), g["scores"]) < REL       # scores of the engine's own forward
    assert relerr(model.weight.grad.cpu().numpy(), g["gwc"]) < REL
    assert relerr(enc2.weight.grad.cpu().numpy(), g["gw2"]) < REL
    assert relerr(enc1.weight.grad.cpu().numpy()[:, cols], g["gw1"]) < REL
    opt.step()
    assert relerr(model.weight.detach().cpu().numpy(), g["wc_new"]) < REL
    assert relerr(enc2.weight.detach().cpu().numpy(), g["w2_new"]) < REL
    assert relerr(enc1.weight.detach().cpu().numpy()[::8, cols], g["w1_new"]) < REL
    # layer-1 embeddings of the first hop-1 nodes with the UPDATED weights (op-by-op Encoder.forward: the same
    # tcgen05 forward kernel through functional.EncoderGemm)
    h1 = enc1(torch.LongTensor(g["hop1"][:32])).t()
    assert relerr(h1.detach().cpu().numpy(), g["h1_new_rows"]) < REL


@pytest.mark.parametrize("name", list(S.CASES))
def test_fused_train_step_matches_reference_at_baseline_widths(golden, name):
    """train_step (sample -> gather -> GEMMs -> head -> SGD as ONE captured graph) lands on the reference's weights."""
    g, p, x, model, enc1, enc2 = _model(name, golden)
    cols = slice(None, None, S.GW1_COLS) if name == "bs_citeseer" else slice(None)
    loss = model.train_step(list(x["nodes"]), x["labels"][x["nodes"]], lr=0.7)
    assert model._engine.tc1 and model._engine.head
    assert abs(loss - float(g["loss"])) / abs(float(g["loss"])) < REL
    assert relerr(model.weight.detach().cpu().numpy(), g["wc_new"]) < REL
    assert relerr(enc2.weight.detach().cpu().numpy(), g["w2_new"]) < REL
    assert relerr(enc1.weight.detach().cpu().numpy()[::8, cols], g["w1_new"]) < REL
    scores = model.forward(list(x["nodes"]))          # op-by-op forward with the updated weights stays finite / shaped
    assert tuple(scores.shape) == g["scores"].shape and torch.isfinite(scores).all()


@pytest.mark.parametrize("in_place", [True, False])
@pytest.mark.parametrize("name", ["bs_reddit", "bs_pubmed"])
def test_in_place_concat_and_materialised_tile_match_reference(golden, name, in_place, monkeypatch):
    """Default (GSAGE_SPLIT_SELF=1): the engine keeps only the neighbour-mean half of the layer-1 tile and the tcgen05
    GEMMs gather the self rows from the feature table themselves (gs_sage_encoder_fwd_tc / _wgrad_tc, encoders.py:53-61
    as one op).  GSAGE_SPLIT_SELF=0: the [self | mean] tile is materialised by the gather.  Same reference outputs."""
    monkeypatch.setenv("GSAGE_SPLIT_SELF", "1" if in_place else "0")
    g, p, x, model, enc1, enc2 = _model(name, golden)
    loss = model.train_step(list(x["nodes"]), x["labels"][x["nodes"]], lr=0.7)
    eng = model._engine
    assert eng.split_self == in_place and eng.sets[0].comb1.shape[1] == (p["f"] if in_place else 2 * p["f"])
    assert abs(loss - float(g["loss"])) / abs(float(g["loss"])) < REL
    assert relerr(eng.gw1.cpu().numpy(), g["gw1"]) < REL
    assert relerr(enc2.weight.detach().cpu().numpy(), g["w2_new"]) < REL
    assert relerr(enc1.weight.detach().cpu().numpy()[::8], g["w1_new"]) < REL
