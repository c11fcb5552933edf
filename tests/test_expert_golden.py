"""Multi-task model (model_expert_s / main_11) against golden vectors from the reference code.
The trust-path branch is plain PyTorch and is checked on CPU; the recommendation branch needs the GPU."""
import os
import pickle
import re
import shutil

import numpy as np
import pytest
import torch

from helpers import make_args, rel_err

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def E():
    return dict(np.load(os.path.join(GOLD, "expert_tiny.npz")))


@pytest.fixture()
def tiny_root(tmp_path):
    shutil.copytree(os.path.join(GOLD, "tiny"), tmp_path / "tiny")
    return str(tmp_path)


def _args(root):
    return make_args(dataset="tiny", data_path=root, hiddenSize=64, batchSize=32, nonhybrid=False, nb_heads=3)


def _model(E, root):
    from spex_b200.dataloader import Loader
    from spex_b200.model_expert_s import LightGCN

    args = _args(root)
    ds = Loader(args)
    model = LightGCN(args, ds)
    sd = {k[3:]: torch.from_numpy(v) for k, v in E.items() if k.startswith("sd.")}
    missing = model.load_state_dict(sd, strict=True)
    return args, ds, model


def _paths(root, n_users):
    from spex_b200.path_data import Data

    tr = pickle.load(open(os.path.join(root, "tiny", "trust", "train.txt"), "rb"))
    te = pickle.load(open(os.path.join(root, "tiny", "trust", "test2.txt"), "rb"))
    return Data(tr, n_users, shuffle=False), Data(te, n_users, shuffle=False, test=True)


def test_same_seed_gives_reference_initial_weights(E, tiny_root):
    from spex_b200 import utils
    from spex_b200.dataloader import Loader
    from spex_b200.model_expert_s import LightGCN

    args = _args(tiny_root)
    ds = Loader(args)
    utils.set_seed(2020)
    model = LightGCN(args, ds)
    for k, v in model.state_dict().items():
        assert np.array_equal(v.numpy(), E["sd." + k]), k


def test_path_data_matches_reference_padding(tiny_root):
    tr, te = _paths(tiny_root, 60)
    assert tr.inputs.shape == (180, 5) and te.neg.shape == (60, 50)
    assert (tr.inputs[tr.mask == 0] == 60).all() and (tr.mask.sum(1) >= 1).all()
    sl = te.generate_batch(32)
    assert [len(s) for s in sl] == [32, 28]


def test_trust_branch_matches_reference_on_cpu(E, tiny_root):
    args, ds, model = _model(E, tiny_root)
    tr, te = _paths(tiny_root, ds.n_users)
    model.eval()
    with torch.no_grad():
        inp, mask, tgt, neg = te.get_slice(np.arange(10))
        sc = model._trust_scores(torch.from_numpy(inp), torch.from_numpy(mask))
    assert rel_err(sc, torch.from_numpy(E["trust_scores"])) < 1e-5
    model.train()
    inp, mask, tgt = tr.get_slice(E["slice"])
    loss2 = model.loss_function(model._trust_scores(torch.from_numpy(inp), torch.from_numpy(mask)),
                                torch.from_numpy(tgt))
    loss2.backward()
    assert abs(float(loss2.detach()) - float(E["loss2"])) < 1e-5 * abs(float(E["loss2"]))
    for name in ("w", "linear_one.weight", "linear_two.bias", "linear_three.weight", "linear_transform.weight",
                 "attention_0.a", "attention_2.a", "out_att.a", "att_t"):
        g = dict(model.named_parameters())[name].grad
        assert rel_err(g, torch.from_numpy(E["grad." + name])) < 1e-4, name


@pytest.mark.gpu
def test_multitask_forward_backward_matches_reference(E, tiny_root, cuda_device):
    from spex_b200 import batch_test
    from spex_b200.batch_test_gnn import trust_test5

    args, ds, model = _model(E, tiny_root)
    model = model.to(cuda_device)
    tr, te = _paths(tiny_root, ds.n_users)
    users, items, labels = (torch.from_numpy(E[k]).to(cuda_device) for k in ("users", "items", "labels"))
    model.train()
    l1, l2 = model(users, items, labels, E["slice"], tr, flag=0)
    (l1 + l2).backward()
    assert abs(float(l1.detach()) - float(E["loss1"])) < 1e-5 * abs(float(E["loss1"]))
    assert abs(float(l2.detach()) - float(E["loss2"])) < 1e-5 * abs(float(E["loss2"]))
    for name, p in model.named_parameters():
        key = "grad." + name
        if key in E:
            assert p.grad is not None, name
            assert rel_err(p.grad, torch.from_numpy(E[key])) < 1e-4, name
    model.eval()
    with torch.no_grad():
        gamma = model(users, items, None, None, None, flag=1)
    assert rel_err(gamma, torch.from_numpy(E["gamma"])) < 1e-5
    ret = batch_test.rec_test(model, ds.testRatings, ds.testNegatives)
    assert np.array_equal(ret["recall"], E["rec_recall"])
    assert np.allclose(ret["ndcg"], E["rec_ndcg"], rtol=0, atol=1e-12)
    tm = np.array(trust_test5(model, te))
    assert np.allclose(tm, E["trust_metrics"], rtol=0, atol=1e-6)


@pytest.mark.gpu
def test_expert_gate_backward_against_autograd(cuda_device):
    from spex_b200 import ops

    torch.manual_seed(3)
    n, D = 5000, 64
    e0 = (torch.randn(n, D) * 0.3).requires_grad_(True)
    e1 = (torch.randn(n, D) * 0.3).requires_grad_(True)
    W = (torch.randn(2 * D, 2) * 0.2).requires_grad_(True)
    G = torch.randn(n, D)
    att = torch.softmax(torch.cat([e0, e1], 1) @ W, 1)
    (( e0 * att[:, 0:1] + e1 * att[:, 1:2]) * G).sum().backward()
    a, b, w = (t.detach().to(cuda_device).requires_grad_(True) for t in (e0, e1, W))
    out = ops.expert_gate(a, b, w)
    (out * G.to(cuda_device)).sum().backward()
    assert rel_err(a.grad, e0.grad) < 1e-5 and rel_err(b.grad, e1.grad) < 1e-5
    assert rel_err(w.grad, W.grad) < 1e-4
    # deterministic dW (fixed-order reduction, no atomics)
    a2, b2, w2 = (t.detach().clone().requires_grad_(True) for t in (a, b, w))
    (ops.expert_gate(a2, b2, w2) * G.to(cuda_device)).sum().backward()
    assert torch.equal(w2.grad, w.grad)


@pytest.mark.gpu
def test_main_11_entry_point(tiny_root, capsys):
    from spex_b200 import main_11

    main_11.main(["--dataset", "tiny", "--data_path", tiny_root, "--epochs", "2", "--batchSize", "32"])
    out = capsys.readouterr().out.splitlines()
    train = [l for l in out if re.fullmatch(r"\d+,\d\.\d{5},\d\.\d{5},\d+\.\d{5},\d+\.\d{5}", l)]
    assert len(train) == 2
    assert len([l for l in out if l.startswith("Rec:  Epoch ")]) == 2
    assert len([l for l in out if l.startswith("Trust:Epoch ")]) == 2
    t0 = [float(x) for x in train[0].split(",")[3:]]
    t1 = [float(x) for x in train[1].split(",")[3:]]
    assert sum(t1) < sum(t0)
