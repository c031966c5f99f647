"""CPU suite for the host side of the training path (no compute calls): the flat 64 B-aligned parameter groups, the
gradient sink, the trainer's option handling (reference keys of options/train_restoration/*.yml `train:`), and the
autograd bridge's refusal to run without CUDA."""
import pytest
import torch
import torch.nn as nn

from textualdegremoval_b200 import TdrError, define_network
from textualdegremoval_b200.archs.restormer_train import Grads
from textualdegremoval_b200.ddp import FlatGroup, RefGuidedTrainer, split_param_groups

TINY = dict(dim=16, num_blocks=[1, 1, 1, 1], num_refinement_blocks=1, heads=[1, 2, 4, 8], nf=16,
            ext_n_blocks=[1, 1, 1, 1], reffusion_n_blocks=[1, 1, 1, 1], LayerNorm_type="WithBias")


def test_flat_group_alignment_and_views():
    ps = [nn.Parameter(torch.randn(*s)) for s in ((3,), (5, 7), (1, 1, 1), (33,))]
    before = [p.detach().clone() for p in ps]
    g = FlatGroup(ps, lr=1e-3)
    assert g.n % FlatGroup.ALIGN == 0 and g.n_params == sum(p.numel() for p in ps)
    for p, b, off in zip(ps, before, g.offsets):
        assert off % FlatGroup.ALIGN == 0 and torch.equal(p.detach(), b)          # values preserved, 64 B-aligned start
        assert p.data_ptr() == g.flat.data_ptr() + 4 * off and p.grad.data_ptr() == g.grad.data_ptr() + 4 * off
    ps[1].grad.fill_(2.0)                                                          # .grad is a view of the flat buffer
    assert g.grad[g.offsets[1]:g.offsets[1] + 35].eq(2.0).all() and g.grad.sum().item() == 70.0
    assert [v.shape for v in g.views(g.m)] == [p.shape for p in ps]


def test_param_groups_split_on_masa_like_the_reference():
    """image_restoration_ref_model.py:149-158: names containing 'masa' use ref_lr."""
    net = define_network(dict(type="RestormerRefFusion", **TINY))
    groups = split_param_groups(list(net.named_parameters()), lr=3e-4, ref_lr=1e-4)
    n_masa = sum(p.numel() for n, p in net.named_parameters() if "masa" in n)
    assert [g.lr for g in groups] == [3e-4, 1e-4]
    assert groups[1].n_params == n_masa and groups[0].n_params + n_masa == sum(p.numel() for p in net.parameters())


def test_grads_sink_direct_and_allocating():
    p = nn.Parameter(torch.zeros(4, 3))
    q = nn.Parameter(torch.zeros(5))
    FlatGroup([p], lr=1.0)                         # p.grad becomes a flat view
    G = Grads(direct=True)
    assert G(p).data_ptr() == p.grad.data_ptr() and id(p) in G.in_place
    t = G(q)                                        # no .grad yet -> fresh zero buffer
    assert t.shape == q.shape and id(q) not in G.in_place and G(q) is t and G(None) is None
    assert Grads(direct=False)(p).data_ptr() != p.grad.data_ptr()


def test_trainer_rejects_unsupported_options():
    net = define_network(dict(type="Restormer", dim=16, num_blocks=[1, 1, 1, 1], num_refinement_blocks=1))
    with pytest.raises(TdrError):
        RefGuidedTrainer(net, dict(optim_g=dict(type="Adam", lr=1e-4)))
    with pytest.raises(TdrError):
        RefGuidedTrainer(net, dict(optim_g=dict(type="AdamW", lr=1e-4), pixel_opt=dict(type="MSELoss")))


def test_training_forward_refuses_cpu_tensors():
    """No CPU fallback on the training path either."""
    net = define_network(dict(type="Restormer", dim=16, num_blocks=[1, 1, 1, 1], num_refinement_blocks=1)).train()
    with pytest.raises(TdrError):
        net(torch.rand(1, 3, 64, 64))
    naf = define_network(dict(type="NAFNet", img_channel=3, width=16, middle_blk_num=1, enc_blk_nums=[1], dec_blk_nums=[1])).train()
    with pytest.raises(TdrError):
        naf(torch.rand(1, 3, 32, 32))


def test_ema_state_dict_views():
    """EMA copies are exposed under the reference's parameter names (what `save_network(..., 'params_ema')` writes)."""
    from textualdegremoval_b200.ddp import DDPStep
    net = define_network(dict(type="Restormer", dim=16, num_blocks=[1, 1, 1, 1], num_refinement_blocks=1))
    tr = RefGuidedTrainer.__new__(RefGuidedTrainer)
    tr.net_g = net
    tr.engine = DDPStep.__new__(DDPStep)
    tr.engine.groups = split_param_groups(list(net.named_parameters()), 1e-4, 1e-4)
    tr.engine.ema = [g.flat.clone() for g in tr.engine.groups]
    sd = tr.ema_state_dict()
    assert set(sd) == set(net.state_dict()) and all(sd[n].shape == p.shape for n, p in net.named_parameters())
    assert all(torch.equal(sd[n], p.detach()) for n, p in net.named_parameters())
    tr.engine.ema = None
    assert tr.ema_state_dict() is None
