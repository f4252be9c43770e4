"""The example scripts run end to end on the CPU tier (gloo, 2 processes, tiny configs): guards them against rot.
Reference: examples/*/test_ci.sh."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _torchrun(script, *args, nproc=2, port=29750, timeout=240):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, script), *args]
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""), CUDA_VISIBLE_DEVICES="")
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    return out.stdout


@pytest.mark.dist
@pytest.mark.parametrize("args,port,nproc", [(("--family", "gpt2", "--tp", "2", "--steps", "2"), 29751, 2),
                                             (("--family", "mixtral", "--ep", "2", "--steps", "2"), 29752, 2),
                                             (("--family", "llama", "--tp", "2", "--pp", "2", "--steps", "2"), 29754, 4),
                                             (("--family", "bloom", "--tp", "2", "--sp", "--steps", "2"), 29767, 2),
                                             (("--family", "opt", "--pp", "2", "--steps", "2"), 29768, 2)])
def test_hf_inplace_example(args, port, nproc):
    pytest.importorskip("transformers")
    out = _torchrun("examples/language/hf_inplace/finetune_hf.py", *args, port=port, nproc=nproc)
    lines = [l for l in out.splitlines() if l.startswith("step ")]
    assert len(lines) == 2 and all("loss" in l and "nan" not in l for l in lines)


@pytest.mark.dist
def test_long_context_ring_attention_example():
    out = _torchrun("examples/language/long_context/train_ring_attention.py", "--steps", "2", "--seq", "128", port=29753)
    lines = [l for l in out.splitlines() if l.startswith("step ")]
    assert len(lines) == 2 and "sp 2" in lines[0] and "nan" not in out


def _python(script, *args, timeout=240):
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""), CUDA_VISIBLE_DEVICES="")
    out = subprocess.run([sys.executable, os.path.join(ROOT, script), *args], capture_output=True, text=True,
                         timeout=timeout, env=env, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    return out.stdout


@pytest.mark.dist
def test_sequence_parallel_tutorial_matches_single_process():
    """Every SP mode prints the loss curve of the single-process run (same seed, same data)."""
    def losses(out):
        return [l.split("loss ")[1].split(" ")[0] for l in out.splitlines() if l.startswith("step ")]

    script = "examples/tutorial/sequence_parallel/train.py"
    ref = losses(_python(script, "--mode", "none", "--steps", "3"))
    assert len(ref) == 3
    for i, mode in enumerate(("split_gather", "all_to_all")):
        assert losses(_torchrun(script, "--mode", mode, "--steps", "3", port=29755 + i)) == ref, mode
    ref = losses(_python(script, "--mode", "none", "--model", "llama-tiny", "--steps", "3"))
    got = losses(_torchrun(script, "--mode", "ring_attn", "--model", "llama-tiny", "--steps", "3", port=29757))
    assert got == ref


@pytest.mark.dist
def test_large_batch_optimizer_tutorial():
    out = _torchrun("examples/tutorial/large_batch_optimizer/train.py", "--optimizer", "lars", "--plugin", "zero1",
                    "--steps", "21", "--batch", "64", port=29758)
    lines = [l for l in out.splitlines() if l.startswith("step ")]
    first, last = float(lines[0].split("loss ")[1].split(" ")[0]), float(lines[-1].split("loss ")[1].split(" ")[0])
    assert "global batch 64" in lines[0] and last < first


def test_inference_examples_cpu():
    out = _python("examples/inference/stable_diffusion/sd3_generation.py", "--model", "sd3-tiny", "--steps", "2")
    assert "generated (1, 3, 64, 64)" in out and "finite True" in out
    out = _python("examples/inference/benchmark_ops/benchmark_ops.py")
    assert out.count(" ok") >= 8                                   # every op agrees with its PyTorch formulation
    out = _python("applications/ColossalQA/examples/retrieval_conversation_universal.py", "--ask",
                  "What is the warranty period?", "保修期是多久？", "--max_new_tokens", "3")
    assert "[en] What is the warranty" in out and "[zh]" in out and out.count("source :") >= 2
    pytest.importorskip("uvicorn")
    out = _python("examples/inference/client/run_client.py", "--self-host", "--concurrency", "4")
    assert "Healthy" in out and "16/16 ok" in out


@pytest.mark.dist
def test_colossal_llama_train_and_resume(tmp_path):
    """The continual-pretraining driver: spliced data, ZeRO-2, checkpoint at step 3, resume reproduces step 4's loss."""
    script = "applications/Colossal-LLaMA/train.py"
    common = ("--model", "llama-tiny", "--plugin", "zero2", "--max_length", "64", "--steps", "5", "--synthetic_docs", "120")
    out = _torchrun(script, *common, "--save_dir", str(tmp_path), "--save_interval", "3", "--log_file",
                    str(tmp_path / "log.jsonl"), port=29759)
    first = {l.split()[1]: l.split()[3] for l in out.splitlines() if l.startswith("step ")}
    assert len(first) == 5 and float(first["5"]) < float(first["1"]) and (tmp_path / "epoch-0_step-3").is_dir()
    assert len((tmp_path / "log.jsonl").read_text().splitlines()) == 5
    out = _torchrun(script, *common, "--load_checkpoint", str(tmp_path / "epoch-0_step-3"), port=29760)
    assert "resumed from" in out
    resumed = {l.split()[1]: l.split()[3] for l in out.splitlines() if l.startswith("step ")}
    assert sorted(resumed) == ["4", "5"] and abs(float(resumed["4"]) - float(first["4"])) < 2e-3


@pytest.mark.dist
def test_colossal_moe_train_then_infer(tmp_path):
    """Expert-parallel training with load monitoring + checkpoint; the expert-parallel decode loop and the single-rank
    paged-KV engine produce the same greedy completion from that checkpoint."""
    out = _torchrun("applications/ColossalMoE/train.py", "--model", "mixtral-tiny", "--ep", "2", "--steps", "4",
                    "--log_interval", "2", "--max_length", "32", "--save_dir", str(tmp_path), "--save_interval", "4",
                    port=29764)
    lines = [l for l in out.splitlines() if l.startswith("step ")]
    assert len(lines) == 2 and "expert load" in lines[0] and "imbalance" in lines[0] and "nan" not in lines[0]
    ckpt = str(tmp_path / "epoch-0_step-4")
    prompt = ("--prompt", "The capital of France is", "--max_new_tokens", "5")
    ep = _torchrun("applications/ColossalMoE/infer.py", "--model", "mixtral-tiny", "--ep", "2", "--checkpoint", ckpt,
                   *prompt, port=29765)
    single = _python("applications/ColossalMoE/infer.py", "--model", "mixtral-tiny", "--engine", "--checkpoint", ckpt, *prompt)
    pick = lambda text: [l for l in text.splitlines() if l.startswith("[output]")]      # noqa: E731
    assert pick(ep) and pick(ep) == pick(single)


@pytest.mark.dist
def test_colossal_eval_two_phase_example(tmp_path):
    cfg = "applications/ColossalEval/examples/dataset_evaluation/config.json"
    out = _torchrun("applications/ColossalEval/examples/dataset_evaluation/inference.py", "--config", cfg, "--out_dir",
                    str(tmp_path / "answers"), port=29766)
    assert out.count("_inference.json") == 3
    table = _python("applications/ColossalEval/examples/dataset_evaluation/eval_dataset.py", "--inference_dir",
                    str(tmp_path / "answers"), "--config", cfg, "--out", str(tmp_path / "results.json"))
    assert "accuracy" in table and "perplexity" in table and "rouge_l" in table and (tmp_path / "results.json").exists()
