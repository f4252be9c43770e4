"""Sharding a user's HuggingFace module in place (sub-module + method replacement through the generic ModelSharder):
a `transformers` Llama / Mistral / Qwen2 built by the USER is tensor-parallelised over 2 ranks and must reproduce the
single-process logits, loss and (gathered) gradients.  Reference pattern: tests/test_shardformer/test_model/
test_shard_llama.py:29-170 (org model vs sharded model)."""
import copy

import pytest
import torch
import torch.distributed as dist

import colossalai_b200
from colossalai_b200.parallel import comm
from colossalai_b200.shardformer import ShardConfig, ShardFormer
from colossalai_b200.testing import rerun_if_address_is_in_use, spawn


def _build(family):
    import transformers

    kw = dict(vocab_size=320, hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=4,
              num_key_value_heads=2, max_position_embeddings=64, tie_word_embeddings=False)
    cfg_cls = {"llama": transformers.LlamaConfig, "mistral": transformers.MistralConfig, "qwen2": transformers.Qwen2Config,
               "cohere": transformers.CohereConfig, "qwen3": transformers.Qwen3Config,
               "glm": transformers.GlmConfig}[family]
    model_cls = {"llama": transformers.LlamaForCausalLM, "mistral": transformers.MistralForCausalLM,
                 "qwen2": transformers.Qwen2ForCausalLM, "cohere": transformers.CohereForCausalLM,
                 "qwen3": transformers.Qwen3ForCausalLM, "glm": transformers.GlmForCausalLM}[family]
    if family == "glm":
        kw.update(head_dim=16, pad_token_id=0)               # (GLM: fused gate|up MLP, partial interleaved rotary)
    cfg = cfg_cls(**kw)
    cfg._attn_implementation = "eager"
    torch.manual_seed(0)
    return model_cls(cfg).float()


def _gather_grad(p):
    """Full gradient of a (possibly sharded) parameter, using the parameter's own sharding annotation."""
    g = p.grad.detach()
    if hasattr(p, "dist_shard"):
        dim, group = p.dist_shard
        return comm.all_gather(g, dim, group)
    if hasattr(p, "gather_fn"):
        return p.gather_fn(g)
    return g


def _check(family):
    org = _build(family)
    sharded = copy.deepcopy(org)
    sc = ShardConfig(tensor_parallel_process_group=dist.group.WORLD, enable_tensor_parallelism=True,
                     enable_fused_normalization=True)
    sharded, _ = ShardFormer(sc).optimize(sharded)          # auto policy: looked up by the HF class's qualified name
    # the policy really rewrote the user's module
    layer = sharded.model.layers[0]
    assert type(layer.self_attn.q_proj).__name__ == "Linear1D_Col" and type(layer.mlp.down_proj).__name__ == "Linear1D_Row"
    assert layer.self_attn.q_proj.weight.shape[0] == org.model.layers[0].self_attn.q_proj.weight.shape[0] // 2
    assert type(sharded.model.embed_tokens).__name__ == "VocabParallelEmbedding1D"
    if family != "cohere":                                   # Cohere uses a (bias-free) LayerNorm, left to PyTorch
        assert layer.input_layernorm.forward.__func__.__name__ == "_fused_rmsnorm_forward"      # method replacement
    torch.manual_seed(5)
    ids = torch.randint(0, 320, (2, 16))
    ref = org(input_ids=ids, labels=ids)
    out = sharded(input_ids=ids, labels=ids)
    torch.testing.assert_close(out.logits, ref.logits, atol=2e-4, rtol=2e-4)
    torch.testing.assert_close(out.loss, ref.loss, atol=1e-5, rtol=1e-5)
    ref.loss.backward()
    out.loss.backward()
    ref_grads = {n: p.grad for n, p in org.named_parameters()}
    n = 0
    for name, p in sharded.named_parameters():
        full = _gather_grad(p)
        r = ref_grads[name]
        if full.shape != r.shape:
            full = full[: r.shape[0]]
        torch.testing.assert_close(full, r, atol=2e-4, rtol=2e-3, msg=lambda m: f"{family} {name}: {m}")
        n += 1
    assert n > 10


def _build_tied(family, layers=2):
    import transformers

    torch.manual_seed(0)
    if family.startswith("falcon"):
        new_arch = family == "falcon-new"
        cfg = transformers.FalconConfig(vocab_size=320, hidden_size=64, num_hidden_layers=layers, num_attention_heads=4,
                                        num_kv_heads=2 if new_arch else None, new_decoder_architecture=new_arch,
                                        multi_query=False, parallel_attn=True, bias=False, alibi=False,
                                        hidden_dropout=0.0, attention_dropout=0.0, max_position_embeddings=64)
        cfg._attn_implementation = "eager"
        org = transformers.FalconForCausalLM(cfg).float()
    elif family == "bloom":
        cfg = transformers.BloomConfig(vocab_size=320, hidden_size=64, n_layer=layers, n_head=4, hidden_dropout=0.0,
                                       attention_dropout=0.0)
        cfg._attn_implementation = "eager"
        org = transformers.BloomForCausalLM(cfg).float()
    elif family == "gptj":
        cfg = transformers.GPTJConfig(vocab_size=320, n_positions=64, n_embd=64, n_layer=layers, n_head=4, rotary_dim=8,
                                      resid_pdrop=0.0, embd_pdrop=0.0, attn_pdrop=0.0)
        cfg._attn_implementation = "eager"
        org = transformers.GPTJForCausalLM(cfg).float()
    elif family == "gpt2":
        cfg = transformers.GPT2Config(vocab_size=320, n_positions=64, n_embd=64, n_layer=layers, n_head=4, resid_pdrop=0.0,
                                      embd_pdrop=0.0, attn_pdrop=0.0)
        cfg._attn_implementation = "eager"
        org = transformers.GPT2LMHeadModel(cfg).float()
    else:
        cfg = transformers.OPTConfig(vocab_size=320, hidden_size=64, ffn_dim=128, num_hidden_layers=layers, num_attention_heads=4,
                                     max_position_embeddings=64, word_embed_proj_dim=64, dropout=0.0)
        cfg._attn_implementation = "eager"
        org = transformers.OPTForCausalLM(cfg).float()
    return org


def _check_tied(family):
    """GPT-2 (fused Conv1D q|k|v, `split_size` attribute) and OPT (`num_heads` attribute), both with tied LM heads."""
    org = _build_tied(family)
    sharded = copy.deepcopy(org)
    sc = ShardConfig(tensor_parallel_process_group=dist.group.WORLD, enable_tensor_parallelism=True)
    sharded, _ = ShardFormer(sc).optimize(sharded)
    if family.startswith("falcon"):
        blk = sharded.transformer.h[0]
        assert type(blk.self_attention.query_key_value).__name__ == "Linear1D_Col" and blk.self_attention.num_heads == 2
        assert blk.self_attention.num_kv_heads == (1 if family == "falcon-new" else 2)
        assert type(blk.mlp.dense_4h_to_h).__name__ == "Linear1D_Row"
    elif family == "bloom":
        blk = sharded.transformer.h[0]
        assert type(blk.self_attention.query_key_value).__name__ == "Linear1D_Col" and blk.self_attention.num_heads == 2
        assert blk.self_attention.forward.__func__.__name__ == "_bloom_attention_forward"
        assert sharded.lm_head.weight is sharded.transformer.word_embeddings.weight
    elif family == "gptj":
        blk = sharded.transformer.h[0]
        assert type(blk.attn.q_proj).__name__ == "Linear1D_Col" and blk.attn.num_attention_heads == 2
        assert type(blk.mlp.fc_out).__name__ == "Linear1D_Row" and type(sharded.lm_head).__name__ == "VocabParallelLMHead1D"
    elif family == "gpt2":
        blk = sharded.transformer.h[0]
        assert type(blk.attn.c_attn).__name__ == "GPT2FusedLinearConv1D_Col" and blk.attn.split_size == 32
        assert type(blk.mlp.c_proj).__name__ == "GPT2FusedLinearConv1D_Row"
        assert sharded.lm_head.weight is sharded.transformer.wte.weight and sharded.lm_head.weight.shape[0] == 192   # 320 -> 384 / 2
    else:
        layer = sharded.model.decoder.layers[0]
        assert type(layer.self_attn.q_proj).__name__ == "Linear1D_Col" and layer.self_attn.num_heads == 2
        assert type(layer.fc2).__name__ == "Linear1D_Row"
        assert sharded.lm_head.weight is sharded.model.decoder.embed_tokens.weight
    torch.manual_seed(6)
    ids = torch.randint(0, 320, (2, 16))
    ref = org(input_ids=ids, labels=ids)
    out = sharded(input_ids=ids, labels=ids)
    torch.testing.assert_close(out.logits, ref.logits, atol=2e-4, rtol=2e-4)
    torch.testing.assert_close(out.loss, ref.loss, atol=1e-5, rtol=1e-5)
    ref.loss.backward()
    out.loss.backward()
    # replicated parameters (LayerNorms, position embeddings) must carry the full gradient on every rank
    ref_grads = {n: p.grad for n, p in org.named_parameters()}
    checked = 0
    for name, p in sharded.named_parameters():
        if hasattr(p, "dist_shard") or hasattr(p, "gather_fn") or p.shape != ref_grads.get(name, p).shape:
            continue
        torch.testing.assert_close(p.grad, ref_grads[name], atol=2e-4, rtol=2e-3, msg=lambda m: f"{family} {name}: {m}")
        checked += 1
    assert checked >= 6
    # the tied vocabulary shard: this rank's rows of the reference gradient
    emb = sharded.get_input_embeddings().weight
    r = dist.get_rank()
    full = ref_grads[{"opt": "model.decoder.embed_tokens.weight", "bloom": "transformer.word_embeddings.weight",
                      "falcon-new": "transformer.word_embeddings.weight",
                      "falcon-mha": "transformer.word_embeddings.weight"}.get(family, "transformer.wte.weight")]
    rows = full[r * 192:(r + 1) * 192]                       # the vocabulary is padded to 384 rows: the tail shard is short
    torch.testing.assert_close(emb.grad[: rows.shape[0]], rows, atol=2e-4, rtol=2e-3)
    assert emb.grad[rows.shape[0]:].abs().max() < 1e-6 if rows.shape[0] < 192 else True


def _check_bert():
    import transformers

    torch.manual_seed(0)
    cfg = transformers.BertConfig(vocab_size=320, hidden_size=64, num_hidden_layers=2, num_attention_heads=4,
                                  intermediate_size=128, max_position_embeddings=64, hidden_dropout_prob=0.0,
                                  attention_probs_dropout_prob=0.0, num_labels=3)
    cfg._attn_implementation = "eager"
    org = transformers.BertForSequenceClassification(cfg).float()
    sharded = copy.deepcopy(org)
    sc = ShardConfig(tensor_parallel_process_group=dist.group.WORLD, enable_tensor_parallelism=True)
    sharded, _ = ShardFormer(sc).optimize(sharded)
    layer = sharded.bert.encoder.layer[0]
    assert type(layer.attention.self.query).__name__ == "Linear1D_Col" and layer.attention.self.num_attention_heads == 2
    assert type(layer.output.dense).__name__ == "Linear1D_Row"
    assert type(sharded.bert.embeddings.word_embeddings).__name__ == "VocabParallelEmbedding1D"
    torch.manual_seed(8)
    ids = torch.randint(0, 320, (3, 16))
    mask = torch.ones(3, 16, dtype=torch.long)
    mask[0, 10:] = 0
    labels = torch.tensor([0, 2, 1])
    ref = org(input_ids=ids, attention_mask=mask, labels=labels)
    out = sharded(input_ids=ids, attention_mask=mask, labels=labels)
    torch.testing.assert_close(out.logits, ref.logits, atol=2e-4, rtol=2e-4)
    torch.testing.assert_close(out.loss, ref.loss, atol=1e-5, rtol=1e-5)
    ref.loss.backward()
    out.loss.backward()
    ref_grads = {n: p.grad for n, p in org.named_parameters()}
    n = 0
    for name, p in sharded.named_parameters():
        full = _gather_grad(p)
        r = ref_grads[name]
        if full.shape != r.shape:
            full = full[: r.shape[0]]
        torch.testing.assert_close(full, r, atol=2e-4, rtol=2e-3, msg=lambda m: f"bert {name}: {m}")
        n += 1
    assert n > 20


def _check_bert_pipeline():
    """BERT in place as two 1F1B stages: embeddings on the first, pooler + classifier + loss (the model's own forward on
    `inputs_embeds`) on the last; padded batch, two micro-batches, two SGD steps vs the single-process model."""
    import re

    import transformers

    from colossalai_b200.booster import Booster
    from colossalai_b200.booster.plugin import HybridParallelPlugin

    torch.manual_seed(0)
    cfg = transformers.BertConfig(vocab_size=320, hidden_size=64, num_hidden_layers=4, num_attention_heads=4,
                                  intermediate_size=128, max_position_embeddings=64, hidden_dropout_prob=0.0,
                                  attention_probs_dropout_prob=0.0, num_labels=3)
    cfg._attn_implementation = "eager"
    org = transformers.BertForSequenceClassification(cfg).float()
    model = copy.deepcopy(org)
    opt = torch.optim.SGD(model.parameters(), lr=0.05)
    ref_opt = torch.optim.SGD(org.parameters(), lr=0.05)
    plugin = HybridParallelPlugin(tp_size=1, pp_size=2, precision="fp32", num_microbatches=2)
    booster = Booster(plugin=plugin, convert_hf_models=False)
    model, opt, *_ = booster.boost(model, opt)
    inner = model.unwrap()
    r = dist.get_rank()
    assert len(inner.bert.encoder.layer) == 2 and (inner.bert.pooler is None) == (r == 0)
    torch.manual_seed(8)
    ids = torch.randint(0, 320, (4, 16))
    mask = torch.ones(4, 16, dtype=torch.long)
    mask[0, 10:] = 0
    types = torch.randint(0, 2, (4, 16))
    labels = torch.tensor([0, 2, 1, 1])
    for _ in range(2):
        batch = {"input_ids": ids, "attention_mask": mask, "token_type_ids": types, "labels": labels}
        out = booster.execute_pipeline(iter([batch]), model, lambda o, b: o["loss"], opt, return_loss=True)
        opt.step()
        opt.zero_grad()
        total = 0.0
        for i in range(2):
            sl = slice(2 * i, 2 * i + 2)
            l = org(input_ids=ids[sl], attention_mask=mask[sl], token_type_ids=types[sl], labels=labels[sl]).loss / 2
            l.backward()
            total += l.item()
        ref_opt.step()
        ref_opt.zero_grad()
        if out["loss"] is not None:
            assert abs(out["loss"].item() - total) < 1e-4, (out["loss"].item(), total)
    ref_params = dict(org.named_parameters())
    n = 0
    for name, p in inner.named_parameters():
        if p is None:
            continue
        ref_name = re.sub(r"\.layer\.(\d+)\.", lambda m: f".layer.{int(m.group(1)) + 2 * r}.", name)
        torch.testing.assert_close(p.detach(), ref_params[ref_name].detach(), atol=2e-5, rtol=1e-4,
                                   msg=lambda m: f"pp bert {name}: {m}")
        n += 1
    assert n >= 30, n
    del plugin


def _check_vit():
    import transformers

    torch.manual_seed(0)
    cfg = transformers.ViTConfig(hidden_size=64, num_hidden_layers=2, num_attention_heads=4, intermediate_size=128,
                                 image_size=32, patch_size=8, num_channels=3, hidden_dropout_prob=0.0,
                                 attention_probs_dropout_prob=0.0, num_labels=5)
    cfg._attn_implementation = "eager"
    org = transformers.ViTForImageClassification(cfg).float()
    sharded = copy.deepcopy(org)
    sc = ShardConfig(tensor_parallel_process_group=dist.group.WORLD, enable_tensor_parallelism=True)
    sharded, _ = ShardFormer(sc).optimize(sharded)
    layer = sharded.vit.encoder.layer[0]
    assert type(layer.attention.attention.key).__name__ == "Linear1D_Col" and layer.attention.attention.all_head_size == 32
    torch.manual_seed(9)
    px = torch.randn(2, 3, 32, 32)
    labels = torch.tensor([1, 4])
    ref = org(pixel_values=px, labels=labels)
    out = sharded(pixel_values=px, labels=labels)
    torch.testing.assert_close(out.logits, ref.logits, atol=2e-4, rtol=2e-4)
    ref.loss.backward()
    out.loss.backward()
    ref_grads = {n: p.grad for n, p in org.named_parameters()}
    n = 0
    for name, p in sharded.named_parameters():
        torch.testing.assert_close(_gather_grad(p), ref_grads[name], atol=2e-4, rtol=2e-3, msg=lambda m: f"vit {name}: {m}")
        n += 1
    assert n > 20


def _check_whisper():
    import transformers

    torch.manual_seed(0)
    cfg = transformers.WhisperConfig(d_model=64, encoder_layers=2, decoder_layers=2, encoder_attention_heads=4,
                                     decoder_attention_heads=4, encoder_ffn_dim=128, decoder_ffn_dim=128, vocab_size=320,
                                     num_mel_bins=8, max_source_positions=16, max_target_positions=16, pad_token_id=0,
                                     bos_token_id=1, eos_token_id=2, decoder_start_token_id=1, suppress_tokens=None,
                                     begin_suppress_tokens=None, dropout=0.0, attention_dropout=0.0, activation_dropout=0.0)
    cfg._attn_implementation = "eager"
    org = transformers.WhisperForConditionalGeneration(cfg).float()
    sharded = copy.deepcopy(org)
    sc = ShardConfig(tensor_parallel_process_group=dist.group.WORLD, enable_tensor_parallelism=True)
    sharded, _ = ShardFormer(sc).optimize(sharded)
    dec = sharded.model.decoder.layers[0]
    assert type(dec.encoder_attn.k_proj).__name__ == "Linear1D_Col" and dec.encoder_attn.num_heads == 2
    assert type(sharded.model.encoder.layers[1].fc2).__name__ == "Linear1D_Row"
    assert sharded.proj_out.weight is sharded.model.decoder.embed_tokens.weight
    torch.manual_seed(10)
    feats = torch.randn(2, 8, 32)
    ids = torch.randint(3, 320, (2, 6))
    ref = org(input_features=feats, decoder_input_ids=ids, labels=ids)
    out = sharded(input_features=feats, decoder_input_ids=ids, labels=ids)
    torch.testing.assert_close(out.logits, ref.logits, atol=2e-4, rtol=2e-4)
    torch.testing.assert_close(out.loss, ref.loss, atol=1e-5, rtol=1e-5)
    ref.loss.backward()
    out.loss.backward()
    ref_grads = {n: p.grad for n, p in org.named_parameters()}
    n = 0
    for name, p in sharded.named_parameters():
        if p.grad is None:                                   # the encoder's sinusoidal positions are frozen
            assert ref_grads[name] is None, name
            continue
        full = _gather_grad(p)
        r = ref_grads[name]
        if full.shape != r.shape:
            full = full[: r.shape[0]]
        torch.testing.assert_close(full, r, atol=2e-4, rtol=2e-3, msg=lambda m: f"whisper {name}: {m}")
        n += 1
    assert n > 40


def _check_t5(gated, tied):
    import transformers

    torch.manual_seed(0)
    cfg = transformers.T5Config(vocab_size=320, d_model=64, d_kv=16, d_ff=128, num_layers=2, num_decoder_layers=2,
                                num_heads=4, relative_attention_num_buckets=8, relative_attention_max_distance=16,
                                dropout_rate=0.0, feed_forward_proj="gated-gelu" if gated else "relu",
                                tie_word_embeddings=tied, decoder_start_token_id=0, pad_token_id=0, eos_token_id=1)
    cfg._attn_implementation = "eager"
    org = transformers.T5ForConditionalGeneration(cfg).float()
    sharded = copy.deepcopy(org)
    sc = ShardConfig(tensor_parallel_process_group=dist.group.WORLD, enable_tensor_parallelism=True)
    sharded, _ = ShardFormer(sc).optimize(sharded)
    sa = sharded.encoder.block[0].layer[0].SelfAttention
    assert type(sa.q).__name__ == "Linear1D_Col" and sa.n_heads == 2 and sa.relative_attention_bias.weight.shape == (8, 2)
    assert type(sharded.decoder.block[1].layer[1].EncDecAttention.o).__name__ == "Linear1D_Row"
    assert sharded.encoder.embed_tokens is sharded.shared and sharded.decoder.embed_tokens is sharded.shared
    assert (sharded.lm_head.weight is sharded.shared.weight) == (org.lm_head.weight is org.shared.weight)
    torch.manual_seed(11)
    ids = torch.randint(2, 320, (2, 12))
    dec = torch.randint(2, 320, (2, 7))
    ref = org(input_ids=ids, decoder_input_ids=dec, labels=dec)
    out = sharded(input_ids=ids, decoder_input_ids=dec, labels=dec)
    torch.testing.assert_close(out.logits, ref.logits, atol=3e-4, rtol=3e-4)
    torch.testing.assert_close(out.loss, ref.loss, atol=1e-5, rtol=1e-5)
    ref.loss.backward()
    out.loss.backward()
    ref_grads = {n: p.grad for n, p in org.named_parameters()}
    n = 0
    for name, p in sharded.named_parameters():
        full = _gather_grad(p)
        r = ref_grads[name]
        if full.shape != r.shape:
            full = full[: r.shape[0]]
        torch.testing.assert_close(full, r, atol=3e-4, rtol=3e-3, msg=lambda m: f"t5 {name}: {m}")
        n += 1
    assert n > 30


def _compare_grads(tag, sharded, org, atol=2e-4, rtol=2e-3, min_n=10, relative_to_max=False):
    ref = {n: p.grad for n, p in org.named_parameters()}
    n = 0
    for name, p in sharded.named_parameters():
        if p.grad is None:
            assert ref[name] is None, f"{tag} {name}: no gradient on the sharded model"
            continue
        full = _gather_grad(p)
        r = ref[name]
        if full.shape != r.shape:
            full = full[: r.shape[0]]
        a = atol * max(1.0, float(r.abs().max())) if relative_to_max else atol
        torch.testing.assert_close(full, r, atol=a, rtol=rtol, msg=lambda m: f"{tag} {name}: {m}")
        n += 1
    assert n >= min_n, (tag, n)


def _check_sam():
    """SAM in place: windowed + global vision layers (fused qkv split per projection, decomposed relative-position
    tables with their gradient summed over the group), two-way mask decoder attention, masks + IoU head vs unsharded."""
    import transformers
    from transformers.models.sam.configuration_sam import SamMaskDecoderConfig, SamPromptEncoderConfig, SamVisionConfig

    v = SamVisionConfig(hidden_size=32, output_channels=16, num_hidden_layers=2, num_attention_heads=4, image_size=32,
                        patch_size=8, window_size=2, global_attn_indexes=[1], mlp_dim=64, num_pos_feats=8)
    pe = SamPromptEncoderConfig(hidden_size=16, image_size=32, patch_size=8, num_point_embeddings=4)
    d = SamMaskDecoderConfig(hidden_size=16, num_hidden_layers=2, num_attention_heads=4, mlp_dim=32,
                             iou_head_hidden_dim=16, attention_downsample_rate=2)
    for impl in ("eager", "sdpa"):
        cfg = transformers.SamConfig(vision_config=v, prompt_encoder_config=pe, mask_decoder_config=d)
        cfg._attn_implementation = impl
        cfg.vision_config._attn_implementation = impl
        torch.manual_seed(0)
        org = transformers.SamModel(cfg).float()
        with torch.no_grad():                                  # (the relative-position tables are zero-initialised)
            for n, p in org.named_parameters():
                if "rel_pos" in n:
                    p.normal_(0, 0.5)
        sharded = copy.deepcopy(org)
        sc = ShardConfig(tensor_parallel_process_group=dist.group.WORLD, enable_tensor_parallelism=True)
        sharded, _ = ShardFormer(sc).optimize(sharded)
        attn = sharded.vision_encoder.layers[0].attn
        assert type(attn.qkv).__name__ == "FusedLinear1D_Col" and attn.qkv.weight.shape[0] == 48 and attn.num_attention_heads == 2
        blk = sharded.mask_decoder.transformer.layers[0]
        assert type(blk.self_attn.q_proj).__name__ == "Linear1D_Col" and type(blk.mlp.lin2).__name__ == "Linear1D_Row"
        torch.manual_seed(6)
        px, pts = torch.randn(2, 3, 32, 32), torch.rand(2, 1, 2, 2) * 32
        outs = []
        for m in (org, sharded):
            o = m(pixel_values=px, input_points=pts, multimask_output=True)
            outs.append(o)
            (o.pred_masks.square().mean() + o.iou_scores.sum()).backward()
        torch.testing.assert_close(outs[1].pred_masks, outs[0].pred_masks, atol=2e-4, rtol=2e-4)
        torch.testing.assert_close(outs[1].iou_scores, outs[0].iou_scores, atol=2e-4, rtol=2e-4)
        _compare_grads(f"sam/{impl}", sharded, org, atol=1e-4, rtol=2e-3)


def _check_blip2():
    """BLIP-2 in place: vision tower (rebound attention forward), Q-Former (self + cross attention), OPT language model
    through its own in-place policy (tied, vocab-parallel head): LM loss and gradients vs unsharded."""
    import transformers
    from transformers.models.blip_2.configuration_blip_2 import Blip2QFormerConfig, Blip2VisionConfig

    vis = Blip2VisionConfig(hidden_size=32, intermediate_size=64, num_hidden_layers=2, num_attention_heads=4,
                            image_size=16, patch_size=8, attention_dropout=0.0)
    qf = Blip2QFormerConfig(vocab_size=64, hidden_size=32, num_hidden_layers=2, num_attention_heads=4,
                            intermediate_size=64, encoder_hidden_size=32, cross_attention_frequency=1,
                            hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0, max_position_embeddings=32)
    txt = transformers.OPTConfig(vocab_size=320, hidden_size=64, ffn_dim=128, num_hidden_layers=2, num_attention_heads=4,
                                 max_position_embeddings=64, word_embed_proj_dim=64, dropout=0.0, attention_dropout=0.0,
                                 activation_dropout=0.0, layerdrop=0.0)
    cfg = transformers.Blip2Config(vision_config=vis, qformer_config=qf, text_config=txt, num_query_tokens=4,
                                   image_token_index=319)
    cfg._attn_implementation = "eager"
    torch.manual_seed(0)
    org = transformers.Blip2ForConditionalGeneration(cfg).float().eval()
    sharded = copy.deepcopy(org)
    sc = ShardConfig(tensor_parallel_process_group=dist.group.WORLD, enable_tensor_parallelism=True)
    sharded, _ = ShardFormer(sc).optimize(sharded)
    lay = sharded.vision_model.encoder.layers[0]
    assert lay.self_attn.forward.__func__.__name__ == "_blip2_attention_forward" and lay.self_attn.num_heads == 2
    ql = sharded.qformer.encoder.layer[0]
    assert type(ql.attention.attention.query).__name__ == "Linear1D_Col" and \
        type(ql.crossattention.output.dense).__name__ == "Linear1D_Row" and ql.attention.attention.all_head_size == 16
    assert type(sharded.language_model.model.decoder.layers[0].fc1).__name__ == "Linear1D_Col"
    torch.manual_seed(8)
    px = torch.randn(2, 3, 16, 16)
    ids = torch.randint(0, 300, (2, 12))
    ids[:, :4] = 319                                             # the placeholders the query outputs are scattered into
    ref = org(pixel_values=px, input_ids=ids, labels=ids)
    out = sharded(pixel_values=px, input_ids=ids, labels=ids)
    torch.testing.assert_close(out.loss, ref.loss, atol=1e-5, rtol=1e-5)
    torch.testing.assert_close(out.logits[..., :320], ref.logits, atol=2e-4, rtol=2e-4)
    ref.loss.backward()
    out.loss.backward()
    # (the random tiny tower has gradients in the hundreds on the patch embedding: tolerance relative to the tensor's max)
    _compare_grads("blip2", sharded, org, min_n=40, relative_to_max=True)


def _moe_model(family):
    import transformers

    torch.manual_seed(0)
    if family == "mixtral":
        cfg = transformers.MixtralConfig(vocab_size=320, hidden_size=64, intermediate_size=96, num_hidden_layers=2,
                                         num_attention_heads=4, num_key_value_heads=2, num_local_experts=4,
                                         num_experts_per_tok=2, max_position_embeddings=64, router_aux_loss_coef=0.0,
                                         output_router_logits=False)
        cls = transformers.MixtralForCausalLM
    elif family == "qwen3_moe":
        cfg = transformers.Qwen3MoeConfig(vocab_size=320, hidden_size=64, intermediate_size=96, moe_intermediate_size=48,
                                          num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=2, num_experts=4,
                                          num_experts_per_tok=2, max_position_embeddings=64, head_dim=16,
                                          router_aux_loss_coef=0.0)
        cls = transformers.Qwen3MoeForCausalLM
    elif family == "qwen2_moe":
        cfg = transformers.Qwen2MoeConfig(vocab_size=320, hidden_size=64, intermediate_size=96, moe_intermediate_size=48,
                                          shared_expert_intermediate_size=64, num_hidden_layers=2, num_attention_heads=4,
                                          num_key_value_heads=2, num_experts=4, num_experts_per_tok=2,
                                          max_position_embeddings=64, router_aux_loss_coef=0.0, decoder_sparse_step=1)
        cls = transformers.Qwen2MoeForCausalLM
    elif family == "deepseek_v2":
        cfg = transformers.DeepseekV2Config(vocab_size=320, hidden_size=64, intermediate_size=96, moe_intermediate_size=48,
                                            num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=4,
                                            n_routed_experts=4, n_shared_experts=1, num_experts_per_tok=2,
                                            first_k_dense_replace=1, kv_lora_rank=16, q_lora_rank=24, qk_rope_head_dim=8,
                                            qk_nope_head_dim=16, v_head_dim=16, n_group=2, topk_group=1,
                                            max_position_embeddings=64)
        cls = transformers.DeepseekV2ForCausalLM
    else:
        cfg = transformers.DeepseekV3Config(vocab_size=320, hidden_size=64, intermediate_size=96, moe_intermediate_size=48,
                                            num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=4,
                                            n_routed_experts=4, n_shared_experts=1, num_experts_per_tok=2,
                                            first_k_dense_replace=1, kv_lora_rank=16, q_lora_rank=24, qk_rope_head_dim=8,
                                            qk_nope_head_dim=16, v_head_dim=16, n_group=2, topk_group=1,
                                            max_position_embeddings=64)
        cls = transformers.DeepseekV3ForCausalLM
    cfg._attn_implementation = "eager"
    return cls(cfg).float()


def _check_moe_ep(family):
    """Expert parallelism of a user's HF MoE model: every rank keeps half of the experts and runs ITS OWN batch; outputs
    must equal the unsharded model on that batch, and the local experts' gradients the sum over both ranks' batches."""
    org = _moe_model(family)
    sharded = copy.deepcopy(org)
    sc = ShardConfig(enable_tensor_parallelism=False, ep_group=dist.group.WORLD)
    sharded, _ = ShardFormer(sc).optimize(sharded)
    ex = sharded.model.layers[-1].mlp.experts
    assert ex.gate_up_proj.shape[0] == 2 and ex.forward.__func__.__name__ == "_ep_experts_forward"
    r = dist.get_rank()
    torch.manual_seed(20 + r)                                 # different data on every rank
    ids = torch.randint(0, 320, (2, 12))
    ref = org(input_ids=ids, labels=ids)
    out = sharded(input_ids=ids, labels=ids)
    torch.testing.assert_close(out.logits, ref.logits, atol=3e-4, rtol=3e-4)
    torch.testing.assert_close(out.loss, ref.loss, atol=1e-5, rtol=1e-5)
    ref.loss.backward()
    out.loss.backward()
    ref_grads = {n: (p.grad.clone() if p.grad is not None else None) for n, p in org.named_parameters()}
    n_expert = 0
    for name, p in sharded.named_parameters():
        if p.grad is None:
            assert ref_grads[name] is None, name
            continue
        if ".experts." in name:
            total = ref_grads[name].clone()
            dist.all_reduce(total)                            # tokens of both ranks reach the owner
            torch.testing.assert_close(p.grad, total[r * 2:(r + 1) * 2], atol=3e-4, rtol=3e-3,
                                       msg=lambda m: f"{family} {name}: {m}")
            n_expert += 1
        else:
            torch.testing.assert_close(p.grad, ref_grads[name], atol=3e-4, rtol=3e-3, msg=lambda m: f"{family} {name}: {m}")
    assert n_expert >= 2


def _check_booster_in_place():
    """`Booster(convert_hf_models=False)`: the plugin shards the user's HF module itself (no conversion to the native
    zoo) and two optimizer steps under TP2 track the single-process model exactly."""
    from colossalai_b200.booster import Booster
    from colossalai_b200.booster.plugin import HybridParallelPlugin

    org = _build("llama")
    model = copy.deepcopy(org)
    opt = torch.optim.AdamW(model.parameters(), lr=1e-2)
    ref_opt = torch.optim.AdamW(org.parameters(), lr=1e-2)
    booster = Booster(plugin=HybridParallelPlugin(tp_size=2, pp_size=1, precision="fp32"), convert_hf_models=False)
    model, opt, *_ = booster.boost(model, opt)
    inner = model.unwrap()
    assert type(inner).__name__ == "LlamaForCausalLM" and type(inner.model.layers[0].mlp.up_proj).__name__ == "Linear1D_Col"
    torch.manual_seed(3)
    ids = torch.randint(0, 320, (2, 16))
    for _ in range(2):
        loss = model(input_ids=ids, labels=ids).loss
        booster.backward(loss, opt)
        opt.step()
        opt.zero_grad()
        ref = org(input_ids=ids, labels=ids).loss
        ref.backward()
        ref_opt.step()
        ref_opt.zero_grad()
        torch.testing.assert_close(loss.detach(), ref.detach(), atol=1e-5, rtol=1e-5)
    w = inner.model.layers[1].self_attn.o_proj.weight                      # row-parallel: columns of the full weight
    full = org.model.layers[1].self_attn.o_proj.weight
    r = dist.get_rank()
    torch.testing.assert_close(w.detach(), full[:, r * 32:(r + 1) * 32].detach(), atol=1e-5, rtol=1e-4)
    # a sharded save of the in-place sharded module is the plain HuggingFace state dict again (TP shards gathered,
    # vocabulary padding stripped): it loads straight back into `transformers`
    import glob
    import os
    import tempfile

    path = [tempfile.mkdtemp() if r == 0 else None]
    dist.broadcast_object_list(path, src=0)
    booster.save_model(model, path[0], shard=True)
    dist.barrier()
    if r == 0:
        sd = {}
        for f in glob.glob(os.path.join(path[0], "*.bin")):
            sd.update(torch.load(f, weights_only=True))
        ref_sd = org.state_dict()
        assert set(sd) == set(ref_sd)
        for k, v in ref_sd.items():
            torch.testing.assert_close(sd[k].float(), v.float(), atol=5e-4, rtol=1e-3, msg=lambda m: f"checkpoint {k}: {m}")
        fresh = _build("llama")
        fresh.load_state_dict(sd)
    dist.barrier()


def _check_sequence_parallel_tied(family):
    """`split_gather` sequence parallelism of the GPT-style HF decoders through the plugin: blocks run on sequence
    shards, two SGD steps track the single-process model (loss, block norms, final norm, position embeddings)."""
    from colossalai_b200.booster import Booster
    from colossalai_b200.booster.plugin import HybridParallelPlugin

    org = _build_tied(family)
    model = copy.deepcopy(org)
    opt = torch.optim.SGD(model.parameters(), lr=0.05)
    ref_opt = torch.optim.SGD(org.parameters(), lr=0.05)
    plugin = HybridParallelPlugin(tp_size=2, pp_size=1, precision="fp32", enable_sequence_parallelism=True,
                                  sequence_parallelism_mode="split_gather")
    booster = Booster(plugin=plugin, convert_hf_models=False)
    model, opt, *_ = booster.boost(model, opt)
    inner = model.unwrap()
    blocks = inner.model.decoder.layers if family == "opt" else inner.transformer.h
    seen = {}
    blocks[0].register_forward_hook(
        lambda m, a, o: seen.update(block=tuple((o[0] if isinstance(o, tuple) else o).shape)))
    torch.manual_seed(3)
    ids = torch.randint(0, 320, (2, 16))
    for step in range(2):
        loss = model(input_ids=ids, labels=ids).loss
        booster.backward(loss, opt)
        opt.step()
        opt.zero_grad()
        ref = org(input_ids=ids, labels=ids).loss
        ref.backward()
        ref_opt.step()
        ref_opt.zero_grad()
        torch.testing.assert_close(loss.detach(), ref.detach(), atol=1e-5, rtol=1e-5,
                                   msg=lambda m: f"sp {family} step {step}: {m}")
    assert seen["block"][1] == (16 if len(blocks) == 1 else 8), seen       # first block: sequence shard out
    ref_params = dict(org.named_parameters())
    n = 0
    for name, p in inner.named_parameters():
        if hasattr(p, "dist_shard") or hasattr(p, "gather_fn") or p.shape != ref_params[name].shape:
            continue                                             # replicated parameters: norms, positions, row biases
        torch.testing.assert_close(p.detach(), ref_params[name].detach(), atol=1e-6, rtol=1e-5,
                                   msg=lambda m: f"sp {family} {name}: {m}")
        n += 1
    assert n >= 5, (family, n)
    del plugin


def _check_sequence_parallel_in_place(family):
    """`split_gather` sequence parallelism of a user's HF decoder (TP2 + SP inside the TP group): the layers run on
    sequence shards (checked with a hook), the norm-weight gradients are summed over the group by the plugin, and two
    SGD steps track the single-process model."""
    from colossalai_b200.booster import Booster
    from colossalai_b200.booster.plugin import HybridParallelPlugin

    org = _build(family)
    model = copy.deepcopy(org)
    opt = torch.optim.SGD(model.parameters(), lr=0.05)      # (SGD: Adam turns rounding noise on ~0 gradients into +-lr)
    ref_opt = torch.optim.SGD(org.parameters(), lr=0.05)
    plugin = HybridParallelPlugin(tp_size=2, pp_size=1, precision="fp32", enable_sequence_parallelism=True,
                                  sequence_parallelism_mode="split_gather")
    booster = Booster(plugin=plugin, convert_hf_models=False)
    model, opt, *_ = booster.boost(model, opt)
    inner = model.unwrap()
    seen = {}
    inner.model.layers[1].mlp.register_forward_hook(lambda m, a, o: seen.update(mlp_out=tuple(o.shape)))
    inner.model.norm.register_forward_hook(lambda m, a, o: seen.update(final_norm=tuple(o.shape)))
    torch.manual_seed(3)
    ids = torch.randint(0, 320, (2, 16))
    for step in range(2):
        loss = model(input_ids=ids, labels=ids).loss
        booster.backward(loss, opt)
        opt.step()
        opt.zero_grad()
        ref = org(input_ids=ids, labels=ids).loss
        ref.backward()
        ref_opt.step()
        ref_opt.zero_grad()
        torch.testing.assert_close(loss.detach(), ref.detach(), atol=1e-5, rtol=1e-5,
                                   msg=lambda m: f"sp {family} step {step}: {m}")
    assert seen["mlp_out"][1] == 8 and seen["final_norm"][1] == 16, seen      # shards inside, full sequence outside
    ref_params = dict(org.named_parameters())
    mine = dict(inner.named_parameters())
    for name in ("model.layers.0.input_layernorm.weight", "model.layers.1.post_attention_layernorm.weight",
                 "model.norm.weight"):
        if name not in mine:                                     # (Cohere: one norm per parallel block)
            continue
        p = mine[name]
        torch.testing.assert_close(p.detach(), ref_params[name].detach(), atol=1e-6, rtol=1e-5,
                                   msg=lambda m: f"sp {family} {name}: {m}")
    del plugin


def _check_zero_and_ddp_keep_hf_module():
    """Data-parallel plugins need no policy: with `convert_hf_models=False` the user's module is wrapped as it is
    (ZeRO-1 in bf16, torch DDP in fp32, Gemini chunks in bf16) and trains."""
    import transformers

    from colossalai_b200.booster import Booster
    from colossalai_b200.booster.plugin import GeminiPlugin, LowLevelZeroPlugin, TorchDDPPlugin

    for plugin, steps in ((LowLevelZeroPlugin(stage=1, precision="bf16"), 4), (TorchDDPPlugin(), 4),
                          (GeminiPlugin(precision="bf16", placement_policy="static", initial_scale=1), 4)):
        torch.manual_seed(0)
        cfg = transformers.GPT2Config(vocab_size=320, n_positions=64, n_embd=64, n_layer=2, n_head=4, resid_pdrop=0.0,
                                      embd_pdrop=0.0, attn_pdrop=0.0)
        model = transformers.GPT2LMHeadModel(cfg)
        opt = torch.optim.AdamW(model.parameters(), lr=3e-3)
        booster = Booster(plugin=plugin, convert_hf_models=False)
        model, opt, *_ = booster.boost(model, opt)
        assert type(model.unwrap()).__name__ == "GPT2LMHeadModel"
        torch.manual_seed(4 + dist.get_rank())
        ids = torch.randint(0, 320, (4, 16))
        losses = []
        for _ in range(steps):
            loss = model(input_ids=ids, labels=ids).loss
            booster.backward(loss, opt)
            opt.step()
            opt.zero_grad()
            losses.append(float(loss))
        assert all(l == l for l in losses) and losses[-1] < losses[0], (type(plugin).__name__, losses)
        if isinstance(plugin, GeminiPlugin):                      # parameters live in chunks, sharded over the ranks
            continue
        # replicas stay identical across the data-parallel ranks
        w = model.unwrap().transformer.h[0].mlp.c_fc.weight.detach().float().clone()
        other = w.clone()
        dist.broadcast(other, src=0)
        torch.testing.assert_close(w, other, atol=1e-6, rtol=1e-6)


def _check_pipeline_tied(family):
    """1F1B pipeline stages of the GPT-style HF decoders in place (4 blocks over 2 stages, 2 micro-batches): learned
    position embeddings / embedding dropout / embedding LayerNorm act on the first stage only, the tied head on the last
    stage keeps its gradient in step with the first stage's embedding; two SGD steps track the single-process model."""
    import re

    from colossalai_b200.booster import Booster
    from colossalai_b200.booster.plugin import HybridParallelPlugin

    org = _build_tied(family, layers=4)
    tied = org.get_output_embeddings().weight is org.get_input_embeddings().weight
    model = copy.deepcopy(org)
    opt = torch.optim.SGD(model.parameters(), lr=0.05)
    ref_opt = torch.optim.SGD(org.parameters(), lr=0.05)
    plugin = HybridParallelPlugin(tp_size=1, pp_size=2, precision="fp32", num_microbatches=2)
    booster = Booster(plugin=plugin, convert_hf_models=False)
    model, opt, *_ = booster.boost(model, opt)
    inner = model.unwrap()
    r = dist.get_rank()
    blocks = inner.model.decoder.layers if family == "opt" else inner.transformer.h
    assert len(blocks) == 2
    torch.manual_seed(12)
    ids = torch.randint(0, 320, (2, 16))
    for _ in range(2):
        out = booster.execute_pipeline(iter([{"input_ids": ids, "labels": ids}]), model, lambda o, b: o["loss"], opt,
                                       return_loss=True)
        opt.step()
        opt.zero_grad()
        total = 0.0
        for i in range(2):
            l = org(input_ids=ids[i:i + 1], labels=ids[i:i + 1]).loss / 2
            l.backward()
            total += l.item()
        ref_opt.step()
        ref_opt.zero_grad()
        if out["loss"] is not None:
            assert abs(out["loss"].item() - total) < 1e-4, (family, out["loss"].item(), total)
    ref_params = dict(org.named_parameters())
    emb_name = [n for n, p in org.named_parameters() if p is org.get_input_embeddings().weight][0]
    n = 0
    for name, p in inner.named_parameters():
        if p is None:
            continue
        # the stage's blocks are renumbered from 0: map back to the global block index
        ref_name = re.sub(r"\.(h|layers)\.(\d+)\.", lambda m: f".{m.group(1)}.{int(m.group(2)) + 2 * r}.", name)
        if ref_name not in ref_params and "lm_head" in ref_name and tied:
            ref_name = emb_name
        torch.testing.assert_close(p.detach(), ref_params[ref_name].detach(), atol=2e-5, rtol=1e-4,
                                   msg=lambda m: f"pp {family} {name}: {m}")
        n += 1
    assert n >= 10, (family, n)
    del plugin


def _check_pipeline_in_place(family, tied):
    """Pipeline parallelism of a user's HF decoder without converting it: stage 0 keeps the embedding + first layers,
    stage 1 the rest + norm + head (tied head: gradients of the two copies are synchronised); two 1F1B steps with two
    micro-batches track the single-process model."""
    import transformers

    from colossalai_b200.booster import Booster
    from colossalai_b200.booster.plugin import HybridParallelPlugin

    kw = dict(vocab_size=320, hidden_size=64, intermediate_size=128, num_hidden_layers=4, num_attention_heads=4,
              num_key_value_heads=2, max_position_embeddings=64, tie_word_embeddings=tied)
    cfg = {"llama": transformers.LlamaConfig, "qwen2": transformers.Qwen2Config}[family](**kw)
    cfg._attn_implementation = "eager"
    torch.manual_seed(0)
    org = {"llama": transformers.LlamaForCausalLM, "qwen2": transformers.Qwen2ForCausalLM}[family](cfg).float()
    model = copy.deepcopy(org)
    opt = torch.optim.AdamW(model.parameters(), lr=1e-2, weight_decay=0.0)
    ref_opt = torch.optim.AdamW(org.parameters(), lr=1e-2, weight_decay=0.0)
    plugin = HybridParallelPlugin(tp_size=1, pp_size=2, precision="fp32", num_microbatches=2)
    booster = Booster(plugin=plugin, convert_hf_models=False)
    model, opt, *_ = booster.boost(model, opt)
    inner = model.unwrap()
    r = dist.get_rank()
    assert len(inner.model.layers) == 2
    assert (inner.model.embed_tokens.weight is not None) == (r == 0) or tied
    torch.manual_seed(12)
    ids = torch.randint(0, 320, (2, 16))
    for _ in range(2):
        out = booster.execute_pipeline(iter([{"input_ids": ids, "labels": ids}]), model, lambda o, b: o["loss"], opt,
                                       return_loss=True)
        opt.step()
        opt.zero_grad()
        total = 0.0
        for i in range(2):
            l = org(input_ids=ids[i:i + 1], labels=ids[i:i + 1]).loss / 2
            l.backward()
            total += l.item()
        ref_opt.step()
        ref_opt.zero_grad()
        if out["loss"] is not None:
            assert abs(out["loss"].item() - total) < 2e-4, (family, out["loss"].item(), total)
    ref_params = dict(org.named_parameters())
    n = 0
    for name, p in inner.named_parameters():
        if p is None:
            continue
        # the stage's layers are renumbered from 0: map back to the global layer index
        ref_name = name
        if ".layers." in name and r == 1:
            head, rest = name.split(".layers.")
            idx, tail = rest.split(".", 1)
            ref_name = f"{head}.layers.{int(idx) + 2}.{tail}"
        if ref_name not in ref_params and "lm_head" in ref_name:
            ref_name = "model.embed_tokens.weight"                      # tied head
        torch.testing.assert_close(p.detach(), ref_params[ref_name].detach(), atol=3e-4, rtol=3e-3,
                                   msg=lambda m: f"pp {family} {name}: {m}")
        n += 1
    assert n >= 10


def _worker(rank, world_size, port):
    colossalai_b200.launch(rank, world_size, "127.0.0.1", port, backend="gloo", verbose=False)
    for family in ("llama", "mistral", "qwen2", "cohere", "qwen3", "glm"):
        _check(family)
    for family in ("gpt2", "opt", "gptj", "bloom", "falcon-new", "falcon-mha"):
        _check_tied(family)
    _check_bert()
    _check_vit()
    _check_whisper()
    _check_sam()
    _check_blip2()
    for gated, tied in ((False, True), (True, False)):
        _check_t5(gated, tied)
    for family in ("mixtral", "qwen3_moe", "qwen2_moe", "deepseek_v2", "deepseek_v3"):
        _check_moe_ep(family)
    _check_booster_in_place()
    for family in ("llama", "qwen3", "cohere", "mistral", "glm"):
        _check_sequence_parallel_in_place(family)
    for family in ("gpt2", "opt", "gptj", "bloom", "falcon-new", "falcon-mha"):
        _check_sequence_parallel_tied(family)
    _check_zero_and_ddp_keep_hf_module()
    for family, tied in (("llama", False), ("qwen2", True)):
        _check_pipeline_in_place(family, tied)
    for family in ("gpt2", "opt", "gptj", "bloom", "falcon-new"):
        _check_pipeline_tied(family)
    _check_bert_pipeline()
    dist.destroy_process_group()


@pytest.mark.dist
@rerun_if_address_is_in_use()
def test_shard_user_hf_modules_tp2():
    pytest.importorskip("transformers")
    spawn(_worker, 2)
