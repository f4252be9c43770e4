"""Fine-tune a `transformers` model WITHOUT converting it: the plugin shards the user's module in place
(`Booster(convert_hf_models=False)` -> `shardformer/policies/hf_*.py`) - tensor parallelism for the dense families,
expert parallelism for the MoE families - and the sharded checkpoint is the plain HuggingFace state dict again.

    torchrun --nproc-per-node 2 --master-addr 127.0.0.1 examples/language/hf_inplace/finetune_hf.py --family llama --tp 2
    torchrun --nproc-per-node 2 --master-addr 127.0.0.1 examples/language/hf_inplace/finetune_hf.py --family mixtral --ep 2
    torchrun --nproc-per-node 4 --master-addr 127.0.0.1 examples/language/hf_inplace/finetune_hf.py --family llama --tp 2 --pp 2

Runs on CPU (gloo) with the tiny configs below; pass `--pretrained <dir>` for real weights on GPUs.
Reference counterpart: `examples/language/llama/benchmark.py` with a HF model handed to `booster.boost`."""
import argparse
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", ".."))
import colossalai_b200  # noqa: E402
from colossalai_b200.booster import Booster  # noqa: E402
from colossalai_b200.booster.plugin import HybridParallelPlugin, MoeHybridParallelPlugin  # noqa: E402


def tiny(family: str):
    import transformers as tf

    common = dict(vocab_size=512, hidden_size=64, num_hidden_layers=2, num_attention_heads=4, max_position_embeddings=128)
    if family == "llama":
        return tf.LlamaForCausalLM(tf.LlamaConfig(intermediate_size=128, num_key_value_heads=2, **common))
    if family == "gpt2":
        return tf.GPT2LMHeadModel(tf.GPT2Config(vocab_size=512, n_embd=64, n_layer=2, n_head=4, n_positions=128))
    if family == "opt":
        return tf.OPTForCausalLM(tf.OPTConfig(ffn_dim=128, word_embed_proj_dim=64, **common))
    if family == "bloom":
        return tf.BloomForCausalLM(tf.BloomConfig(vocab_size=512, hidden_size=64, n_layer=2, n_head=4))
    if family == "mixtral":
        return tf.MixtralForCausalLM(tf.MixtralConfig(intermediate_size=96, num_key_value_heads=2, num_local_experts=4,
                                                      num_experts_per_tok=2, **common))
    raise ValueError(family)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--family", default="llama", choices=["llama", "gpt2", "opt", "bloom", "mixtral"])
    ap.add_argument("--pretrained", default=None, help="HuggingFace checkpoint directory (AutoModelForCausalLM)")
    ap.add_argument("--tp", type=int, default=1)
    ap.add_argument("--ep", type=int, default=1)
    ap.add_argument("--pp", type=int, default=1, help="pipeline stages (decoder families)")
    ap.add_argument("--sp", action="store_true", help="split_gather sequence parallelism inside the TP group (decoder families)")
    ap.add_argument("--zero", type=int, default=0)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--save", default=None)
    args = ap.parse_args()
    colossalai_b200.launch_from_torch()
    cuda = torch.cuda.is_available()
    precision = "bf16" if (cuda or args.zero) else "fp32"         # (ZeRO keeps fp32 masters of low-precision params)
    if args.pretrained:
        import transformers as tf

        model = tf.AutoModelForCausalLM.from_pretrained(args.pretrained, torch_dtype=torch.bfloat16 if cuda else torch.float32)
    else:
        torch.manual_seed(0)
        model = tiny(args.family)
    if args.ep > 1:
        plugin = MoeHybridParallelPlugin(tp_size=args.tp, pp_size=args.pp, ep_size=args.ep, zero_stage=args.zero,
                                         precision=precision, num_microbatches=2 if args.pp > 1 else None)
    else:
        sp = dict(enable_sequence_parallelism=True, sequence_parallelism_mode="split_gather") if args.sp else {}
        plugin = HybridParallelPlugin(tp_size=args.tp, pp_size=args.pp, zero_stage=args.zero, precision=precision,
                                      num_microbatches=2 if args.pp > 1 else None, **sp)
    booster = Booster(plugin=plugin, convert_hf_models=False)          # keep the user's module, shard it in place
    optimizer = torch.optim.AdamW(model.parameters(), lr=1e-3)
    model, optimizer, *_ = booster.boost(model, optimizer)
    dev = colossalai_b200.accelerator.get_accelerator().get_current_device()
    vocab = model.unwrap().config.vocab_size
    g = torch.Generator().manual_seed(dist.get_rank() // max(args.tp * args.pp, 1))   # same data inside a model replica
    for step in range(args.steps):
        ids = torch.randint(0, vocab, (2, 32), generator=g).to(dev)
        if args.pp > 1:
            out = booster.execute_pipeline(iter([{"input_ids": ids, "labels": ids}]), model, lambda o, b: o["loss"],
                                           optimizer, return_loss=True)
            loss = out["loss"]
            lt = torch.full((1,), float("-inf"), device=dev) if loss is None else loss.detach().float().reshape(1)
            dist.all_reduce(lt, op=dist.ReduceOp.MAX)            # the loss lives on the last stage
            loss = lt[0]
        else:
            loss = model(input_ids=ids, labels=ids).loss
            booster.backward(loss, optimizer)
        optimizer.step()
        optimizer.zero_grad()
        if dist.get_rank() == 0:
            print(f"step {step}: loss {loss.item():.4f} ({type(model.unwrap()).__name__}, tp={args.tp}, pp={args.pp}, ep={args.ep})")
    if args.save:
        booster.save_model(model, args.save, shard=True)               # plain HF names / shapes on disk
    dist.barrier()
    colossalai_b200.initialize.shutdown()


if __name__ == "__main__":
    main()
