"""GPU: the reference's own call sites driving the drop-ins — `_evaluate_with_loader` (experiments/utils.py:330-377, the
S-sample MC loop + metric container) and `Trainer._step` / `Trainer.infer` (src/trainer.py:87-174) run UNMODIFIED from
oracle/_ref (or /root/reference), with the two import redirections INTEGRATION.md describes: the model comes from
qbn_b200.zoo (drop-in layers) and `ClassificationMetric` from qbn_b200.metrics.  The only stand-in is the plotting module
(matplotlib is not installed).  Skipped where the reference is absent."""
import sys
import types

import numpy as np
import pytest
import torch

from oracle import ref_harness

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not ref_harness.reference_available(), reason="reference sources not present (oracle/_ref)")]


def _reference_callsites():
    import __graft_entry__ as ge
    ge.build()
    ref_harness.import_reference()
    if "experiments.presentation.plot_settings" not in sys.modules:      # plotting only (matplotlib absent); nothing on the path under test
        for name in ("experiments.presentation", "experiments.presentation.plot_settings"):
            sys.modules.setdefault(name, types.ModuleType(name))
        sys.modules["experiments.presentation.plot_settings"].PLT = None
    import experiments.utils as eu
    import src.trainer as tr
    return eu, tr


def _lenet(args):
    from qbn_b200 import zoo
    torch.manual_seed(0)
    net = zoo.ConvNetwork_LeNet([1, 1, 28, 28], 10, False, args)
    with torch.no_grad():
        for m in net.modules():
            if hasattr(m, "std") and hasattr(m, "weight"):
                m.weight.normal_(0, 1.0 / m.weight[0].numel() ** 0.5)
                m.std.fill_(-4.0)
    return net.cuda()


def test_reference_evaluate_with_loader_runs_on_the_dropins(monkeypatch):
    eu, _ = _reference_callsites()
    from qbn_b200 import metrics as qm, noise, zoo
    import src.metrics as ref_metrics
    monkeypatch.setattr(eu, "ClassificationMetric", qm.ClassificationMetric)          # one import line
    args = zoo.Args(sigma_prior=0.1, model="conv_lenet_bbb", task="classification", samples=4, q=False, debug=False, output_size=10)
    model = _lenet(args).eval()
    g = torch.Generator().manual_seed(1)
    loader = [(torch.rand(16, 1, 28, 28, generator=g), torch.randint(0, 10, (16,), generator=g)) for _ in range(3)]
    noise.manual_seed(3)
    with torch.no_grad():
        error, ece, entropy, nll, output, target = eu._evaluate_with_loader(loader, model, args)
    assert output.shape == (48, 10) and target.shape == (48,)
    np.testing.assert_allclose(output.sum(-1).numpy(), np.ones(48), atol=1e-5)
    # the same outputs through the reference's own metric container (torchmetrics stand-in for the ECE) give the same numbers
    ref = ref_metrics.ClassificationMetric(output_size=10)
    for i in range(3):
        ref.update(output[16 * i:16 * (i + 1)], target[16 * i:16 * (i + 1)])
    np.testing.assert_allclose(error, ref.error.compute().item(), atol=1e-6)
    np.testing.assert_allclose(nll, ref.nll.compute().item(), rtol=1e-5)
    np.testing.assert_allclose(entropy, ref.entropy.compute().item(), rtol=1e-5)
    np.testing.assert_allclose(ece, ref.ece.compute().item(), atol=1e-5)


def test_reference_trainer_step_and_infer_run_on_the_dropins(monkeypatch):
    _, tr = _reference_callsites()
    from qbn_b200 import metrics as qm, noise, zoo
    import src.losses as ref_losses
    monkeypatch.setattr(tr, "ClassificationMetric", qm.ClassificationMetric)
    args = zoo.Args(sigma_prior=0.1, model="conv_lenet_bbb", task="classification", samples=1, q=False, debug=False, output_size=10,
                    gamma=0.1, loss_multiplier=1.0, report_freq=1000, epochs=1, learning_rate=1e-3, save_last=False)
    model = _lenet(args)
    crit = ref_losses.LOSS_FACTORY["classification"](args, "batch")                  # the reference's own ELBO (losses.py:14-29)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    trainer = tr.Trainer(model, crit, opt, None, args)
    g = torch.Generator().manual_seed(2)
    x, t = torch.rand(32, 1, 28, 28, generator=g), torch.randint(0, 10, (32,), generator=g)
    before = [p.detach().clone() for p in model.parameters()]
    noise.manual_seed(4)
    model.train()
    for _ in range(3):
        trainer._step(x, t, opt, 10, 320, True)                                      # trainer.py:87-132, unmodified
    changed = [not torch.equal(a, b) for a, b in zip(before, model.parameters()) if b.requires_grad]
    assert all(changed), "every trainable parameter (mu and rho of every layer) must have moved"
    packed = trainer.train_metrics.get_packed()
    assert set(packed) == {"obj", "main_obj", "kl", "nll", "error", "entropy", "brier", "ece"}
    assert all(np.isfinite(v) for v in packed.values()) and packed["kl"] > 0 and packed["obj"] > packed["main_obj"]
    assert "Objective" in trainer.train_metrics.get_str()
    loader = [(x, t), (x, t)]

    class _DS(list):
        dataset = list(range(64))
    trainer.infer(_DS(loader))                                                       # trainer.py:154-174: eval-mode single-sample pass
    assert trainer.valid_metrics.count == 64 and 0.0 <= trainer.valid_metrics.get_key_metric().item() <= 1.0
