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
                                             (("--family", "llama", "--tp", "2", "--pp", "2", "--steps", "2"), 29754, 4)])
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
