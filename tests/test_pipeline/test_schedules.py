"""Pipeline schedules (1F1B / interleaved / ZB-V) through HybridParallelPlugin on gloo, PP=2, vs single-process oracle
(reference pattern: tests/test_pipeline/test_schedule/test_oneF_oneB.py, test_interleaved.py, test_zerobubble_pp.py)."""
import copy

import pytest
import torch
import torch.distributed as dist

import colossalai_b200
from colossalai_b200.booster import Booster
from colossalai_b200.booster.plugin import HybridParallelPlugin
from colossalai_b200.models import build_model
from colossalai_b200.nn.optimizer import FusedAdam
from colossalai_b200.pipeline.schedule.v_schedule import (PipelineGraph, interleaved_1f1b_schedule,
                                                         one_f_one_b_schedule)
from colossalai_b200.testing import rerun_if_address_is_in_use, spawn


def test_schedule_graphs_are_complete():
    for n_stage, n_micro in [(2, 4), (4, 8), (4, 5), (8, 16)]:
        sched = PipelineGraph(n_stage, n_micro, 2, 2, 2, 1, 1.0, -0.5, -0.5).get_v_schedule()
        for s in range(n_stage):
            for typ in "FBW":
                got = sorted((n.chunk, n.minibatch) for n in sched[s] if n.type == typ)
                assert got == sorted((c, m) for c in range(2) for m in range(n_micro)), (n_stage, n_micro, s, typ)
        sched = interleaved_1f1b_schedule(n_stage, n_micro, 2)
        for s in range(n_stage):
            assert sum(1 for n in sched[s] if n.type == "F") == 2 * n_micro
            assert sum(1 for n in sched[s] if n.type == "B") == 2 * n_micro


def test_one_f_one_b_node_list_is_classic_1f1b():
    """The generated 1F1B node list: every micro-batch forward before its backward, warm-up depth n_stage - stage - 1,
    at most n_stage - stage forwards in flight, strict F/B alternation in the steady state, FIFO channels."""
    for n_stage, n_micro in [(2, 4), (4, 8), (4, 3), (8, 16), (3, 1)]:
        sched = one_f_one_b_schedule(n_stage, n_micro)
        for s, nodes in enumerate(sched):
            comp = [(n.type, n.minibatch) for n in nodes if n.type in ("F", "B")]
            assert sorted(m for t, m in comp if t == "F") == list(range(n_micro))
            assert sorted(m for t, m in comp if t == "B") == list(range(n_micro))
            assert [m for t, m in comp if t == "F"] == list(range(n_micro))      # micro-batches in order
            assert [m for t, m in comp if t == "B"] == list(range(n_micro))
            inflight = peak = 0
            for t, m in comp:
                inflight += 1 if t == "F" else -1
                assert inflight >= 0
                peak = max(peak, inflight)
            assert peak == min(n_stage - s, n_micro), (n_stage, n_micro, s, peak)
            warm = 0
            while warm < len(comp) and comp[warm][0] == "F":
                warm += 1
            assert warm == min(n_stage - s, n_micro)                              # warm-up forwards + the first steady F
            steady = comp[warm:len(comp) - (warm - 1)] if warm > 1 else comp[warm:]
            for a, b in zip(steady, steady[1:]):
                assert a[0] != b[0], (n_stage, n_micro, s, comp)                  # B F B F ... alternation
            # every send has exactly one matching receive on the neighbour, in the same order
            if s + 1 < n_stage:
                sent = [n.minibatch for n in nodes if n.type == "SEND_FORWARD"]
                recv = [n.minibatch for n in sched[s + 1] if n.type == "RECV_FORWARD"]
                assert sent == recv == list(range(n_micro))
                sent_b = [n.minibatch for n in sched[s + 1] if n.type == "SEND_BACKWARD"]
                recv_b = [n.minibatch for n in nodes if n.type == "RECV_BACKWARD"]
                assert sent_b == recv_b == list(range(n_micro))


def _run(pp_style, num_model_chunks, tied=False, **plugin_kw):
    torch.manual_seed(7)
    base = build_model("gpt2-tiny" if tied else "llama-tiny")
    model = copy.deepcopy(base)
    ref_opt = torch.optim.AdamW(base.parameters(), lr=1e-2, weight_decay=0.0)
    opt = FusedAdam(model.parameters(), lr=1e-2, weight_decay=0.0)
    plugin = HybridParallelPlugin(tp_size=1, pp_size=2, precision="fp32", num_microbatches=4, pp_style=pp_style,
                                  num_model_chunks=num_model_chunks, **plugin_kw)
    booster = Booster(plugin=plugin)
    model, opt, *_ = booster.boost(model, opt)
    torch.manual_seed(11)
    ids = torch.randint(0, 512, (4, 16))
    for it in range(2):
        out = booster.execute_pipeline(iter([{"input_ids": ids, "labels": ids}]), model,
                                       lambda o, b: o["loss"], opt, return_loss=True)
        opt.step()
        opt.zero_grad()
        # oracle: mean over 4 micro-batches of size 1
        total = 0.0
        for i in range(4):
            l = base(input_ids=ids[i:i + 1], labels=ids[i:i + 1])["loss"] / 4
            l.backward()
            total += l.item()
        ref_opt.step()
        ref_opt.zero_grad()
        if out["loss"] is not None:
            assert abs(out["loss"].item() - total) < 2e-4, (pp_style, out["loss"].item(), total)
    ref_params = dict(base.named_parameters())
    n = 0
    for name, p in model.unwrap().named_parameters():
        if p is None:
            continue
        ref_name = name if name in ref_params else "model.embed_tokens.weight"   # tied head lives on the last stage
        torch.testing.assert_close(p.detach(), ref_params[ref_name].detach(), atol=3e-4, rtol=3e-3,
                                   msg=lambda m: f"{pp_style} {name}: {m}")
        n += 1
    assert n > 3
    if "num_layers_per_stage" in plugin_kw:
        held = sum(1 for name, _ in model.unwrap().named_parameters() if name.endswith("input_layernorm.weight"))
        assert held == plugin_kw["num_layers_per_stage"][dist.get_rank()], (held, plugin_kw)
    del plugin


def _worker(rank, world_size, port):
    colossalai_b200.launch(rank, world_size, "127.0.0.1", port, backend="gloo", verbose=False)
    _run("1f1b", 1)
    _run("1f1b", 1, tied=True)
    _run("interleaved", 2)
    _run("zbv", 2)
    # uneven stages with per-stage activation checkpointing (the Llama-3-70B headline layout in miniature)
    from colossalai_b200.shardformer import PipelineGradientCheckpointConfig

    _run("1f1b", 1, num_layers_per_stage=[1, 3],
         gradient_checkpoint_config=PipelineGradientCheckpointConfig(num_ckpt_layers_per_stage=[1, 2]))
    _run("1f1b", 1, gradient_checkpoint_config=PipelineGradientCheckpointConfig(gradient_checkpointing_ratio=0.5))
    dist.destroy_process_group()


@pytest.mark.dist
@rerun_if_address_is_in_use()
def test_pipeline_schedules_cpu():
    spawn(_worker, 2)


if __name__ == "__main__":
    test_schedule_graphs_are_complete()
    test_pipeline_schedules_cpu()


def _variable_shape_worker(rank, world_size, port):
    """Sequence length and batch size change from step to step: the cached P2P metadata is re-exchanged when the
    micro-batch shape changes, and the micro-batch size follows the batch (fixed number of micro-batches) - every
    step's loss equals the single-process loss on the same data."""
    import copy

    import torch.distributed as dist

    import colossalai_b200
    from colossalai_b200.booster import Booster
    from colossalai_b200.booster.plugin import HybridParallelPlugin
    from colossalai_b200.models import build_model

    colossalai_b200.launch(rank, world_size, "127.0.0.1", port, backend="gloo", verbose=False)
    torch.manual_seed(0)
    base = build_model("llama-tiny").float()
    model = copy.deepcopy(base)
    opt, ref_opt = torch.optim.SGD(model.parameters(), lr=0.05), torch.optim.SGD(base.parameters(), lr=0.05)
    booster = Booster(plugin=HybridParallelPlugin(tp_size=1, pp_size=2, precision="fp32", num_microbatches=2))
    model, opt, *_ = booster.boost(model, opt)
    g = torch.Generator().manual_seed(1)
    for B, S in [(2, 16), (2, 24), (4, 8), (2, 16), (6, 12)]:
        ids = torch.randint(0, 256, (B, S), generator=g)
        out = booster.execute_pipeline(iter([{"input_ids": ids, "labels": ids}]), model, lambda o, b: o["loss"], opt,
                                       return_loss=True)
        opt.step()
        opt.zero_grad()
        total, mb = 0.0, B // 2
        for i in range(2):
            l = base(input_ids=ids[i * mb:(i + 1) * mb], labels=ids[i * mb:(i + 1) * mb])["loss"] / 2
            l.backward()
            total += l.item()
        ref_opt.step()
        ref_opt.zero_grad()
        if out["loss"] is not None:
            assert abs(out["loss"].item() - total) < 1e-4, ((B, S), out["loss"].item(), total)
    dist.destroy_process_group()


@pytest.mark.dist
def test_pipeline_variable_batch_shapes():
    from colossalai_b200.testing import spawn

    spawn(_variable_shape_worker, 2)
