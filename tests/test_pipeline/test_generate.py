"""Pipeline-parallel greedy generation == single-process greedy decoding (reference: GenerateSchedule usage in the
legacy pipeline inference engine)."""
import copy

import torch
import torch.distributed as dist

import colossalai_b200
from colossalai_b200.cluster import DeviceMesh
from colossalai_b200.models import build_model
from colossalai_b200.pipeline.schedule.generate import GenerateSchedule, MicroBatchManager
from colossalai_b200.pipeline.stage_manager import PipelineStageManager
from colossalai_b200.shardformer import ShardConfig, ShardFormer
from colossalai_b200.testing import rerun_if_address_is_in_use, spawn


def _naive(model, ids, n_new):
    out = ids.clone()
    for _ in range(n_new):
        with torch.no_grad():
            logits = model(input_ids=out)["logits"].view(out.shape[0], out.shape[1], -1)
        out = torch.cat([out, logits[:, -1, : model.cfg.vocab_size].argmax(-1, keepdim=True)], dim=1)
    return out[:, ids.shape[1]:]


def _worker(rank, world_size, port, family="llama-tiny"):
    colossalai_b200.launch(rank, world_size, "127.0.0.1", port, backend="gloo", verbose=False)
    torch.manual_seed(3)
    full = build_model(family).float().eval()
    mesh = DeviceMesh(pp=world_size)
    sm = PipelineStageManager(mesh, pipeline_axis=0)
    sc = ShardConfig(pipeline_stage_manager=sm, enable_tensor_parallelism=False)
    staged, _ = ShardFormer(sc).optimize(copy.deepcopy(full))
    staged = staged.float().eval()
    torch.manual_seed(9)
    prompts = torch.randint(3, 500, (4, 7))
    sched = GenerateSchedule(sm, MicroBatchManager(sm.stage, new_length=5, micro_batch_size=2, micro_batch_buffer_size=2))
    outs = sched.generate_step(staged, iter([{"input_ids": prompts}]))
    if sm.is_last_stage():
        got = torch.cat(outs, dim=0)
        ref = _naive(full, prompts, 5)
        assert torch.equal(got, ref), (got, ref)
    else:
        assert outs == []
    dist.barrier()
    dist.destroy_process_group()


@rerun_if_address_is_in_use()
def test_pipeline_generate_pp2():
    spawn(_worker, 2)


@rerun_if_address_is_in_use()
def test_pipeline_generate_alibi_pp2():
    """ALiBi family (bloom): the dense stage cache applies the per-head linear bias."""
    spawn(_worker, 2, family="bloom-tiny")
