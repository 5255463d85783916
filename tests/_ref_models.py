"""TEST INFRASTRUCTURE: builds the reference's OWN stock-quantised model families (MC-Dropout ResNet, SGHMC ensemble) from the
unmodified reference (oracle/ref_harness: /root/reference here, oracle/_ref on the GPU box), runs its QAT -> convert lifecycle
on the CPU (torch's FBGEMM kernels) and returns the converted model.  The parity tests then re-house a deep copy on the GPU
kernels (quant_utils.to_device_int8) and compare with the reference's CPU result on the same state, input and noise."""
import copy

import torch

from oracle import ref_harness


def available():
    return ref_harness.reference_available()


def _randomise(net, g):
    with torch.no_grad():
        for m in net.modules():
            if isinstance(m, (torch.nn.Conv2d, torch.nn.Linear)):
                fan_in = m.weight[0].numel()
                m.weight.copy_(torch.randn(m.weight.shape, generator=g) / fan_in ** 0.5)
                if m.bias is not None:
                    m.bias.copy_(torch.randn(m.bias.shape, generator=g) * 0.1)
            if isinstance(m, torch.nn.BatchNorm2d):
                m.running_mean.normal_(0, 0.1, generator=g)
                m.running_var.uniform_(0.5, 1.5, generator=g)
                m.weight.uniform_(0.5, 1.5, generator=g)
                m.bias.normal_(0, 0.1, generator=g)


def _lifecycle(net, args, forwards, seed):
    """prepare_model (stock prepare_qat for non-BBB families, src/quant_utils.py:140-141) -> calibration forwards in train and
    eval mode -> the reference's convert."""
    import src.quant_utils as qu
    g = torch.Generator().manual_seed(seed)
    xc = torch.randn(8, 3, 32, 32, generator=g)
    net.train()
    qu.prepare_model(net, args)
    for m in net.modules():
        if hasattr(m, "freeze_bn_stats"):
            m.freeze_bn_stats()
    torch.manual_seed(seed)
    for _ in range(forwards):
        net(xc)
    net.eval()
    with torch.no_grad():
        for _ in range(forwards):
            net(xc)
    qu.convert(net)
    return net.eval()


def sgld_ensemble(n_members=3, seed=41):
    """`Network(training_mode=False)` of models_sgld.py:216-288 with `n_members` independently initialised members."""
    ref_harness.import_reference()
    from src.models.stochastic.sgld.models_sgld import Network
    args = ref_harness.Args(model="conv_resnet_sgld", q=True, at=True, samples=n_members, task="classification",
                            activation_precision=7, weight_precision=8)
    torch.manual_seed(seed)
    net = Network([1, 3, 32, 32], 10, True, args, training_mode=False)
    g = torch.Generator().manual_seed(seed + 1)
    for member in net.ensemble:
        _randomise(member, g)
    return _lifecycle(net, args, n_members, seed + 2), args


def mc_dropout_resnet(p=0.15, seed=51):
    """conv_resnet_mc (models_mc.py:159-226) quantised A7/W8 by the reference's own lifecycle."""
    ref_harness.import_reference()
    from src.models.stochastic.mcdropout.models_mc import ConvNetwork_ResNet
    args = ref_harness.Args(model="conv_resnet_mc", q=True, at=True, p=p, samples=2, task="classification",
                            activation_precision=7, weight_precision=8)
    torch.manual_seed(seed)
    net = ConvNetwork_ResNet([1, 3, 32, 32], 10, True, args)
    _randomise(net, torch.Generator().manual_seed(seed + 1))
    return _lifecycle(net, args, 1, seed + 2), args


def dropout_sites(net):
    return [m for m in net.modules() if type(m).__name__ == "BernoulliDropout" and float(m.p) > 0]


def clone(net):
    return copy.deepcopy(net)
