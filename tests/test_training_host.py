"""Host-side logic of the training path that needs no GPU: routing of `BaseUNetND.forward` to the differentiable graph,
which denoiser variants have a training path, the pack plan's notion of a stable weight source, and the loud refusal to
optimise CPU parameters (no CPU implementation)."""
import pytest
import torch

from fmdm_b200.models.generators import DiffusionUNetFactory
from fmdm_b200.training import FusedAdamW, graph
from fmdm_b200.training.packplan import PackPlan, _stable_source

SMALL = {"unet_impl": "diffusers_nd", "in_channels": 1, "out_channels": 1, "layers_per_block": 1,
         "block_out_channels": [32, 64], "down_block_types": ["DownBlock2D", "AttnDownBlock2D"],
         "up_block_types": ["AttnUpBlock2D", "UpBlock2D"]}


def test_which_variants_have_a_training_path():
    f = DiffusionUNetFactory()
    assert graph.supported(f.build(SMALL, "concatenate", 1))
    assert graph.supported(f.build({"in_channels": 1, "out_channels": 1, "num_res_blocks": 1, "channel_mult": [1, 2],
                                    "model_channels": 32, "attention_resolutions": []}, "concatenate", 1))
    ca = dict(SMALL, cross_attention_dim=4, down_block_types=["DownBlock2D", "CrossAttnDownBlock2D"])
    assert graph.supported(f.build(ca, "attention", 1))              # cross-attention trains at head_dim 8
    assert not graph.supported(f.build(dict(ca, attention_head_dim=16), "attention", 1))
    ce = {"unet_impl": "efficient_nd", "in_channels": 1, "out_channels": 1, "num_res_blocks": 1, "channel_mult": [1, 2],
          "model_channels": 64, "attention_resolutions": [2], "cross_attention_resolutions": [2],
          "cross_attention_in_middle": True, "cross_attention_dim": 4}
    assert graph.supported(f.build(ce, "attention", 1))              # CompVis cross- / linear attention train too
    assert not graph.supported(f.build(dict(ce, cross_attention_dim=32), "attention", 1))   # context of > 16 channels
    assert not graph.supported(torch.nn.Linear(2, 2))


def test_forward_routes_to_the_differentiable_graph_only_when_training_with_grad(monkeypatch):
    model = DiffusionUNetFactory().build(SMALL, "concatenate", 1)
    x = torch.zeros(1, 1, 8, 8)
    # CPU tensors never take the training route (and the inference route refuses CPU tensors loudly)
    assert not model._wants_autograd(x, None)

    class Probe:  # the predicate only looks at `.is_cuda`
        is_cuda = True

    model.train()
    assert model._wants_autograd(Probe(), None)
    with torch.no_grad():
        assert not model._wants_autograd(Probe(), None)
    model.eval()
    assert not model._wants_autograd(Probe(), None)
    model.train()
    assert not model._wants_autograd(Probe(), torch.zeros(1))        # graph-replay form (t_table) is inference
    for p in model.parameters():
        p.requires_grad_(False)
    assert not model._wants_autograd(Probe(), None)


def test_pack_plan_only_plans_stable_sources():
    p = torch.nn.Parameter(torch.randn(8, 4, 3, 3))
    lin = torch.nn.Parameter(torch.randn(8, 4, 1, 1))
    assert _stable_source(p)
    assert _stable_source(lin.reshape(8, 4))                          # a view of a Parameter keeps its address
    assert not _stable_source(torch.cat([p, p], 0))                   # rebuilt every step: packed per call
    assert not _stable_source(p.detach().clone())
    assert not _stable_source(torch.nn.Parameter(torch.randn(4, 4).half()))
    plan = PackPlan()
    plan.begin_step(torch.device("cpu"))                              # nothing recorded: nothing to launch
    assert not plan.ready and not plan.entries


def test_fused_adamw_refuses_cpu_parameters():
    net = torch.nn.Linear(3, 3)
    before = [p.data_ptr() for p in net.parameters()]
    with pytest.raises(RuntimeError, match="no CPU implementation"):
        FusedAdamW(net.parameters(), lr=1e-3)
    assert [p.data_ptr() for p in net.parameters()] == before         # parameters were not re-seated
    with pytest.raises(ValueError):
        FusedAdamW([{"params": list(net.parameters())}])


def test_flat_param_order_makes_fused_views_possible():
    """`flat_param_order` puts the per-stage `emb_layers` weights and each attention block's q/k/v weights back to back
    in the flat buffers; `fused_param` then views them (and their gradients) as one matrix, and refuses anything that
    is not adjacent, not direct-gradient mode, or not a parameter."""
    from fmdm_b200.models.generators import DiffusionUNetFactory
    from fmdm_b200.nn.blocks.attention import DiffusersAttentionND
    from fmdm_b200.training import functions as F
    from fmdm_b200.training.graph import _cut_groups, _embedding_blocks, flat_param_order
    from fmdm_b200.training.optim import FlatBuffers

    cfg = {"unet_impl": "diffusers_nd", "in_channels": 1, "out_channels": 1, "layers_per_block": 1,
           "block_out_channels": [32, 64, 64], "down_block_types": ["DownBlock2D", "AttnDownBlock2D", "DownBlock2D"],
           "up_block_types": ["UpBlock2D", "AttnUpBlock2D", "UpBlock2D"]}
    for cut in (None, 1):
        torch.manual_seed(0)
        model = DiffusionUNetFactory().build(cfg, "concatenate", 1)
        before = {k: v.detach().clone() for k, v in model.state_dict().items()}
        order = flat_param_order(model, cut)
        assert len(order) == len(list(model.parameters())) and len({id(p) for p in order}) == len(order)
        flat = FlatBuffers(model.parameters(), order=order)
        assert all(torch.equal(v, before[k]) for k, v in model.state_dict().items())      # re-seating keeps the values
        early, late = _cut_groups(model, cut)
        old = F.DIRECT_PARAM_GRADS
        try:
            F.DIRECT_PARAM_GRADS = False
            assert F.fused_param([b.emb_layers.weight for b in _embedding_blocks(late)]) is None
            F.DIRECT_PARAM_GRADS = True
            for mods in (early, late):
                blks = _embedding_blocks(mods)
                if not blks:
                    continue
                for plist in ([b.emb_layers.weight for b in blks], [b.emb_layers.bias for b in blks]):
                    w = F.fused_param(plist)
                    assert w is not None and w.shape[0] == sum(p.shape[0] for p in plist)
                    assert torch.equal(w, torch.cat([p.detach() for p in plist], 0))
                    w._fm_grad_view.fill_(3.0)
                    assert all(bool((p.grad == 3.0).all()) for p in plist)
                    assert w._fm_params == tuple(plist)
            att = next(m for m in model.modules() if isinstance(m, DiffusersAttentionND))
            qkv = F.fused_param([att.to_q.weight, att.to_k.weight, att.to_v.weight])
            assert qkv is not None and torch.equal(qkv[att.to_q.weight.shape[0]:2 * att.to_q.weight.shape[0]],
                                                   att.to_k.weight.detach())
            assert F.fused_param([att.to_k.weight, att.to_q.weight]) is None          # not adjacent in this order
            assert F.fused_param([att.to_q.weight.detach()]) is None                  # not a parameter
            if cut is not None:    # early and late stages are separate runs of the buffer
                assert F.fused_param([b.emb_layers.weight for b in _embedding_blocks(early + late)]) is None
        finally:
            F.DIRECT_PARAM_GRADS = old
        with pytest.raises(ValueError):
            FlatBuffers(model.parameters(), order=order[:-1])
