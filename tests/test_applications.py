"""ColossalChat (coati) on tiny models: objectives against hand-computed values, every trainer makes progress, the
producer -> consumer GRPO loop learns a verifiable reward (reference: applications/ColossalChat/tests/*)."""
import math
import os
import sys

import pytest
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "applications", "ColossalChat"))

from coati.dataset import (Conversation, DataCollatorForKTODataset, DataCollatorForPreferenceDataset,  # noqa: E402
                           DataCollatorForPromptDataset, DataCollatorForSupervisedDataset, ListDataset, tokenize_kto,
                           tokenize_preference, tokenize_prompt, tokenize_sft)
from coati.distributed import (GRPOConsumer, ModelRolloutBackend, Producer, boxed_math_reward, extract_boxed,  # noqa: E402
                               format_reward, launch_distributed)
from coati.experience import NaiveExperienceMaker  # noqa: E402
from coati.models import (Critic, DpoLoss, KTOLoss, LogSigLoss, OddsRatioLoss, PolicyLoss, RewardModel, ValueLoss,  # noqa: E402
                          calc_action_log_probs, compute_reward, generate, get_logits)
from coati.trainer import (DPOTrainer, GRPOTrainer, KTOTrainer, ORPOTrainer, PPOTrainer, RewardModelTrainer,  # noqa: E402
                           SFTTrainer)
from coati.trainer.grpo import group_advantages  # noqa: E402
from colossalai_b200.models import build_model, get_config  # noqa: E402


def _tok(text):           # byte-level toy tokenizer (ids 3..258), 0 pad / 1 bos / 2 eos
    return [3 + b for b in text.encode()]


def _tiny(seed=0, **kw):
    torch.manual_seed(seed)
    return build_model(get_config("llama-tiny", num_hidden_layers=2, **kw)).float()


def test_losses_match_formulas():
    lp, old = torch.tensor([[-1.0, -2.0]]), torch.tensor([[-1.2, -1.5]])
    adv = torch.tensor([[1.0, -1.0]])
    loss, skipped, _ = PolicyLoss(0.2)(lp, old, adv, torch.ones(1, 2))
    r = (lp - old).exp()
    expect = -torch.min(r * adv, r.clamp(0.8, 1.2) * adv).mean()
    assert not skipped and torch.allclose(loss, expect)
    assert PolicyLoss(0.2, skip_threshold=1.1)(lp, old - 5, adv, torch.ones(1, 2))[1]          # off-policy batch skipped
    v = ValueLoss(0.2)(torch.tensor([[1.0]]), torch.tensor([[0.5]]), torch.tensor([[2.0]]))
    assert torch.allclose(v, torch.tensor(0.5 * max((0.7 - 2) ** 2, (1.0 - 2) ** 2)))
    m = torch.ones(2, 3)
    pc, pr, rc, rr = torch.full((2, 3), -1.0), torch.full((2, 3), -2.0), torch.full((2, 3), -1.5), torch.full((2, 3), -1.5)
    d, cw, rw = DpoLoss(0.1)(pc, pr, rc, rr, m, m)
    assert torch.allclose(d, -F.logsigmoid(torch.tensor(0.1 * 3.0))) and (cw > rw).all()
    assert torch.allclose(LogSigLoss()(torch.tensor([2.0]), torch.tensor([0.0])), -F.logsigmoid(torch.tensor(2.0)))
    o, lo = OddsRatioLoss()(torch.full((1, 2), -0.1), torch.full((1, 2), -2.0), torch.ones(1, 2), torch.ones(1, 2))
    assert lo.item() > 0 and o.item() < math.log(2)
    k, *_ = KTOLoss(0.1)(torch.tensor([-1.0]), torch.tensor([-3.0]), torch.tensor([-2.0, -2.0]), torch.tensor([-2.0]),
                         torch.tensor([-2.0]), torch.tensor([-2.0, -2.0]))
    assert 0 < k.item() < 1
    g = group_advantages(torch.tensor([1.0, 0.0, 0.0, 1.0, 5.0, 5.0, 5.0, 5.0]), 4)
    assert torch.allclose(g[:4].sum(), torch.tensor(0.0), atol=1e-5) and torch.all(g[4:] == 0)
    r, kl = compute_reward(torch.tensor([1.0]), 0.1, torch.tensor([[-1.0, -1.0, -1.0]]), torch.tensor([[-1.5, -1.5, -1.5]]),
                           torch.tensor([[1.0, 1.0, 0.0]]))
    assert torch.allclose(r, torch.tensor([[-0.05, 0.95, 0.0]])) and torch.allclose(kl, torch.tensor([0.5]))


def test_dataset_tokenisation_and_rewards():
    msgs = [{"role": "user", "content": "hi"}, {"role": "assistant", "content": "yo"}]
    s = tokenize_sft(msgs, _tok)
    trained = bytes(t - 3 for t, l in zip(s["input_ids"], s["labels"]) if l != -100).decode()
    assert trained == "yo<|end|>\n" and len(s["input_ids"]) == len(s["labels"])
    p = tokenize_prompt(msgs[:1], _tok)
    assert bytes(t - 3 for t in p["input_ids"]).decode().endswith("<|assistant|>\n")
    pref = tokenize_preference(msgs[:1], "good", "bad!", _tok)
    batch = DataCollatorForPreferenceDataset()([pref, pref])
    assert batch["chosen_input_ids"].shape[0] == 2 and batch["rejected_loss_mask"].sum() == 2 * len(_tok("bad!<|end|>\n"))
    kb = DataCollatorForKTODataset()([tokenize_kto(msgs[:1], "a", True, _tok), tokenize_kto(msgs[:1], "bcd", False, _tok)])
    assert kb["label"].tolist() == [True, False] and kb["kl_input_ids"].shape[0] == 2
    pb = DataCollatorForPromptDataset()([{"input_ids": [5, 6, 7]}, {"input_ids": [8]}])
    assert pb["input_ids"].tolist() == [[5, 6, 7], [0, 0, 8]] and pb["attention_mask"].tolist() == [[1, 1, 1], [0, 0, 1]]
    assert extract_boxed("so \\boxed{\\frac{1}{2}} and \\boxed{42}.") == "42"
    assert boxed_math_reward("answer \\boxed{42.0}", "42") == 1.0 and boxed_math_reward("no box", "42") == 0.0
    assert boxed_math_reward("\\boxed{7}", "42", format_score=0.1) == 0.1
    assert format_reward("<think>a</think><answer>b</answer>") == 1.0 and format_reward("<answer>b</answer>") == 0.0


def _pref_loader(n=8):
    torch.manual_seed(0)
    items = [tokenize_preference([{"role": "user", "content": f"q{i}"}], "yes " * 3, "no " * 3, _tok) for i in range(n)]
    return torch.utils.data.DataLoader(ListDataset(items), batch_size=4, collate_fn=DataCollatorForPreferenceDataset())


def test_supervised_style_trainers_learn():
    sft_items = [tokenize_sft([{"role": "user", "content": "ping"}, {"role": "assistant", "content": "pong"}], _tok)] * 8
    dl = torch.utils.data.DataLoader(ListDataset(sft_items), batch_size=4, collate_fn=DataCollatorForSupervisedDataset())
    m = _tiny()
    h = SFTTrainer(m, None, torch.optim.AdamW(m.parameters(), lr=3e-3), max_epochs=6).fit(dl)
    assert h[-1]["loss"] < 0.5 * h[0]["loss"]
    rm = RewardModel(_tiny(1))
    h = RewardModelTrainer(rm, None, torch.optim.AdamW(rm.parameters(), lr=3e-3), max_epochs=6).fit(_pref_loader())
    assert h[-1]["accuracy"] == 1.0 and h[-1]["loss"] < h[0]["loss"]
    actor, ref = _tiny(2), _tiny(2)
    h = DPOTrainer(actor, ref, None, torch.optim.AdamW(actor.parameters(), lr=2e-3), beta=0.5, max_epochs=5).fit(_pref_loader())
    assert h[-1]["loss"] < h[0]["loss"] and h[-1]["accuracy"] == 1.0
    actor = _tiny(3)      # SimPO flavour: no reference, length-normalised, margin
    h = DPOTrainer(actor, None, None, torch.optim.AdamW(actor.parameters(), lr=2e-3), beta=2.0, gamma=0.5,
                   length_normalization=True, max_epochs=4).fit(_pref_loader())
    assert h[-1]["loss"] < h[0]["loss"]
    actor = _tiny(4)
    h = ORPOTrainer(actor, None, torch.optim.AdamW(actor.parameters(), lr=2e-3), lam=0.5, max_epochs=4).fit(_pref_loader())
    assert h[-1]["loss"] < h[0]["loss"] and h[-1]["log_odds_ratio"] > h[0]["log_odds_ratio"]
    kto_items = [tokenize_kto([{"role": "user", "content": f"q{i}"}], "yes yes" if i % 2 == 0 else "no no", i % 2 == 0, _tok)
                 for i in range(8)]
    kdl = torch.utils.data.DataLoader(ListDataset(kto_items), batch_size=4, collate_fn=DataCollatorForKTODataset())
    actor, ref = _tiny(5), _tiny(5)
    h = KTOTrainer(actor, ref, None, torch.optim.AdamW(actor.parameters(), lr=2e-3), beta=0.5, max_epochs=5).fit(kdl)
    assert h[-1]["loss"] < h[0]["loss"] and h[-1]["chosen_reward"] > h[-1]["rejected_reward"]


TARGET = 7      # the "verifiable" task: emit token 7 as often as possible


def _count_reward(seq, prompt_len, **_):
    return (seq[:, prompt_len:] == TARGET).float().mean(-1)


def _prompt_loader():
    items = [{"input_ids": [1, 10 + i, 11 + i]} for i in range(4)]
    return torch.utils.data.DataLoader(ListDataset(items), batch_size=4, collate_fn=DataCollatorForPromptDataset())


def test_gae_and_ppo_improves_reward():
    v = torch.zeros(1, 3)
    r = torch.tensor([[0.0, 0.0, 1.0]])
    adv = NaiveExperienceMaker.gae(v, r, torch.ones(1, 3), gamma=1.0, lam=1.0)
    assert torch.allclose(adv, torch.ones(1, 3))
    actor, init = _tiny(0, vocab_size=32), _tiny(0, vocab_size=32)
    critic = Critic(_tiny(1, vocab_size=32))
    tr = PPOTrainer(None, None, actor, critic, None, init, torch.optim.AdamW(actor.parameters(), lr=3e-3),
                    torch.optim.AdamW(critic.parameters(), lr=3e-3), kl_coef=0.0, train_batch_size=16,
                    reward_fn=_count_reward, generate_kwargs=dict(max_new_tokens=6))
    torch.manual_seed(0)
    h = tr.fit(_prompt_loader(), num_episodes=12, num_collect_steps=4, num_update_steps=3)
    first = sum(x["reward"] for x in h[:6]) / 6
    last = sum(x["reward"] for x in h[-6:]) / 6
    assert last > first + 0.05, (first, last)


def test_grpo_trainer_and_producer_consumer_loop():
    actor, init = _tiny(0, vocab_size=32), _tiny(0, vocab_size=32)
    tr = GRPOTrainer(None, actor, init, torch.optim.AdamW(actor.parameters(), lr=3e-3), _count_reward,
                     num_generations=8, beta=0.01, clip_eps_low=0.2, clip_eps_high=0.28, loss_variation="token_level",
                     filter_uniform_groups=True, generate_kwargs=dict(max_new_tokens=6))
    torch.manual_seed(0)
    h = tr.fit(_prompt_loader(), num_episodes=15, num_collect_steps=1, num_update_steps=1)
    assert sum(x["reward"] for x in h[-4:]) / 4 > sum(x["reward"] for x in h[:4]) / 4 + 0.05
    # producer / consumer with periodic weight sync (single-process launch)
    policy, sampler = _tiny(0, vocab_size=32), _tiny(0, vocab_size=32)
    prod = Producer(ModelRolloutBackend(sampler, dict(max_new_tokens=6)), _prompt_loader(), num_generations=8)
    cons = GRPOConsumer(policy, torch.optim.AdamW(policy.parameters(), lr=3e-3), _count_reward, num_generations=8,
                        minibatch_size=16)
    torch.manual_seed(0)
    h = launch_distributed(prod, cons, num_steps=15, sync_every=2)
    assert sum(x["reward"] for x in h[-4:]) / 4 > sum(x["reward"] for x in h[:4]) / 4 + 0.05
    assert prod.model_version >= 14 and max(x["staleness"] for x in h) <= 2
    for a, b in zip(policy.parameters(), sampler.parameters()):       # last sync landed on step 14 of 15
        assert a.shape == b.shape


def test_zero_bubble_loop_buffer_and_rewards(tmp_path):
    """Producer thread + bounded buffer + latest-wins weight mailbox; rollout (de)serialisation; extra verifiers."""
    from coati.distributed import (RolloutBuffer, StepProfiler, WeightMailbox, code_reward, combine_rewards,
                                   deserialize_rollout, launch_zero_bubble, length_penalty, serialize_rollout)

    policy, sampler = _tiny(0, vocab_size=32), _tiny(0, vocab_size=32)
    prod = Producer(ModelRolloutBackend(sampler, dict(max_new_tokens=6)), _prompt_loader(), num_generations=8)
    cons = GRPOConsumer(policy, torch.optim.AdamW(policy.parameters(), lr=3e-3), _count_reward, num_generations=8,
                        minibatch_size=16)
    prof = StepProfiler(log_file=str(tmp_path / "prof.log"))
    torch.manual_seed(0)
    h = launch_zero_bubble(prod, cons, num_steps=16, sync_every=1, buffer_capacity=2, max_staleness=3, profiler=prof)
    assert len(h) == 16 and cons.version == 16
    assert sum(x["reward"] for x in h[-4:]) / 4 > sum(x["reward"] for x in h[:4]) / 4 + 0.03
    assert max(x["staleness"] for x in h) <= 3 and prod.model_version >= 10       # weights kept flowing to the producer
    s = prof.summary()
    assert s["rollout"]["calls"] >= 16 and s["train"]["calls"] == 16 and 0.0 <= prof.overlap_fraction() <= 1.0
    assert (tmp_path / "prof.log").read_text().count("train") == 16
    # staleness filter + back-pressure
    buf = RolloutBuffer(capacity=2, max_staleness=1)
    assert buf.push({"model_version": 0}) and buf.push({"model_version": 3})
    assert buf.push({"model_version": 9}, timeout=0.05) is False                  # full
    assert buf.pop(current_version=4)["model_version"] == 3 and buf.stats["dropped_stale"] == 1
    buf.close()
    assert buf.pop(0) is None
    box = WeightMailbox()
    box.publish({"w": torch.ones(2)}, 1)
    box.publish({"w": torch.full((2,), 2.0)}, 2)
    sd, ver = box.take()
    assert ver == 2 and float(sd["w"][0]) == 2.0 and box.take() is None
    r = {"sequences": torch.arange(12).view(3, 4), "mask": torch.ones(3, 4, dtype=torch.bool), "prompt_len": 2,
         "gt_answer": ["1", "2", "3"], "model_version": 7, "lp": torch.randn(3, 2).bfloat16()}
    back = deserialize_rollout(serialize_rollout(r))
    assert back["gt_answer"] == r["gt_answer"] and back["prompt_len"] == 2 and back["lp"].dtype == torch.bfloat16
    for k in ("sequences", "mask", "lp"):
        assert torch.equal(back[k], r[k])
    assert length_penalty(90, 100, soft_cache=20) == -0.5 and length_penalty(50, 100, soft_cache=20) == 0.0
    assert length_penalty(100, 100) == -1.0
    good = "Here:\n```python\ndef add(a, b):\n    return a + b\n```"
    assert code_reward(good, "assert add(2, 3) == 5") == 1.0
    assert code_reward(good, "assert add(2, 3) == 6") == 0.0 and code_reward("no code", "assert True") == 0.0
    sc = combine_rewards((boxed_math_reward, 1.0), (format_reward, 0.5))
    assert sc("<think>x</think><answer>\\boxed{4}</answer>", "4") == 1.5


def test_colossal_llama_eval_and_qa_utilities(tmp_path):
    for sub in ("Colossal-LLaMA", "ColossalEval", "ColossalQA"):
        sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "applications", sub))
    from colossal_eval import Evaluator, exact_match, f1_score, perplexity, rouge_l
    from colossal_llama import ClosedToConstantLengthSplicedDataset, expand_vocab, supervised_tokenize_pretrain
    from colossalqa import EmbeddingIndex, RetrievalQA, split_text

    # spliced dataset: every sample appears once, items are exactly max_length, little padding
    samples = [supervised_tokenize_pretrain({"source": "s" * (i % 3), "target": "t" * (5 + 3 * i)}, _tok) for i in range(12)]
    ds = list(ClosedToConstantLengthSplicedDataset(samples, max_length=64, num_packed_sequences=6))
    assert all(d["input_ids"].shape == (64,) for d in ds)
    assert sum(int(d["attention_mask"].sum()) for d in ds) == sum(s["seq_length"] for s in samples)
    assert sum(int(d["attention_mask"].sum()) for d in ds) / (64 * len(ds)) > 0.75
    assert all(int(d["seq_boundaries"][-1]) == int(d["attention_mask"].sum()) for d in ds)
    assert samples[1]["labels"][:2] == [-100, -100] and samples[1]["labels"][-1] == 2
    # vocabulary expansion keeps old logits and mean-initialises the new rows
    m = _tiny(0, vocab_size=32)
    ids = torch.randint(0, 32, (1, 6))
    before = m(input_ids=ids)["logits"][:, :32].clone()
    m = expand_vocab(m, 40, {32: [3, 4]})
    after = m(input_ids=ids)["logits"]
    assert after.shape[-1] == 40 and torch.allclose(after[:, :32], before, atol=1e-5)
    w = m.model.embed_tokens.weight
    assert torch.allclose(w[32], (w[3] + w[4]) / 2, atol=1e-6)
    # metrics + evaluator
    assert exact_match("The Cat.", "cat") == 1.0 and 0 < f1_score("a big cat", "big dog") < 1 and rouge_l("x y z", "x z") > 0.7
    ev = Evaluator(_tiny(1), _tok, max_new_tokens=2)
    res = ev.evaluate({"mc": [{"instruction": "2+2=", "choices": ["4", "5"], "answer": 0}],
                       "gen": [{"instruction": "hi", "target": "zzz"}]}, {"gen": ["exact_match", "f1"]})
    assert set(res) == {"mc", "gen"} and res["mc"]["accuracy"] in (0.0, 1.0) and perplexity(_tiny(1), _tok, ["abc"]) > 1
    # retrieval QA
    doc = ("The B200 has 148 streaming multiprocessors. Its HBM3e capacity is 180 gigabytes. "
           "NVLink 5 gives every GPU 900 gigabytes per second to each peer. Bananas are yellow.")
    chunks = split_text(doc, chunk_size=70, chunk_overlap=10)
    assert len(chunks) >= 3 and all(len(c) <= 70 for c in chunks)
    index = EmbeddingIndex()
    assert index.add_documents(chunks, source="b200.txt") == len(chunks) and index.add_documents(chunks, source="b200.txt") == 0
    qa = RetrievalQA(index, generate=lambda prompt: prompt.split("context:\n")[1].split("\n")[0])
    ans, src = qa.run("How many streaming multiprocessors does the B200 have?")
    assert "148" in ans and src and "question:" in qa.build_prompt("x")[0] and len(qa.memory.turns) == 1


def test_eval_dataset_adapters_and_category_report(tmp_path):
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "applications", "ColossalEval"))
    from colossal_eval import (Evaluator, bleu, cloze_items, extract_last_number, few_shot_prompt, gsm8k_items,
                               load_mmlu_csv, mmlu_items)

    rows = [{"question": "2+2?", "choices": ["3", "4", "5", "6"], "answer": "B", "subject": "math"},
            {"question": "Capital of France?", "A": "Paris", "B": "Rome", "C": "Oslo", "D": "Bern", "answer": 0,
             "subject": "geo"}]
    items = mmlu_items(rows)
    assert items[0]["answer"] == 1 and items[0]["choices"] == [" A", " B", " C", " D"] and "B. 4" in items[0]["instruction"]
    assert items[1]["answer"] == 0 and items[1]["category"] == "geo" and items[1]["instruction"].endswith("Answer:")
    csv_path = tmp_path / "astronomy_test.csv"
    csv_path.write_text('"Which is a planet?",Sun,Mars,Moon,Vega,B\n')
    (csv_item,) = load_mmlu_csv(csv_path)
    assert csv_item["category"] == "astronomy" and csv_item["answer"] == 1
    g = gsm8k_items([{"question": "3 apples + 4 apples?", "answer": "3 + 4 = 7\n#### 7"}])[0]
    assert g["target"] == "7" and g["postprocess"] == "last_number"
    assert extract_last_number("so we get 1,234.0 apples.") == "1234" and extract_last_number("none") is None
    shot = few_shot_prompt(items[0], [items[1], g], header="The following are questions.")
    assert shot["instruction"].startswith("The following are questions.") and "Answer: A" in shot["instruction"]
    assert "#### 7" not in shot["instruction"] and shot["instruction"].rstrip().endswith("Answer:")
    assert cloze_items([{"context": "He opened the", "endings": ["door", "sky"], "label": 0}])[0]["choices"] == [" door", " sky"]
    assert bleu("the cat sat on the mat", "the cat sat on the mat") > 0.99 and bleu("a b", "c d") == 0.0
    assert 0 < bleu("the cat slept on the mat today", "the cat sat on the mat today") < 1
    ev = Evaluator(_tiny(2), _tok, max_new_tokens=2)
    rep = ev.evaluate_by_category(items, shots=[items[1]])
    assert set(rep) == {"math", "geo", "macro_avg", "micro_avg"} and 0.0 <= rep["macro_avg"]["accuracy"] <= 1.0
    ev.save(rep, tmp_path / "out" / "mmlu.json")
    assert (tmp_path / "out" / "mmlu.json").exists()


def test_qa_bm25_hybrid_and_loaders(tmp_path):
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "applications", "ColossalQA"))
    from colossalqa import (BM25Index, ConversationMemory, EmbeddingIndex, HybridRetriever, RetrievalQA, load_documents,
                            rewrite_follow_up, tokenize)

    assert tokenize("NVLink-5 给每个GPU 900 GB/s!") == ["nvlink", "5", "给", "每", "个", "gpu", "900", "gb", "s"]
    (tmp_path / "b200.md").write_text("# Compute\nThe B200 has 148 streaming multiprocessors.\n"
                                      "# Memory\nIts HBM3e capacity is 180 gigabytes.\n")
    (tmp_path / "misc.jsonl").write_text('{"text": "Bananas are yellow and rich in potassium."}\n'
                                         '{"text": "NVLink 5 gives every GPU 900 gigabytes per second to each peer."}\n')
    docs = load_documents([tmp_path / "b200.md", tmp_path / "misc.jsonl"], chunk_size=120, chunk_overlap=0)
    assert {d["source"] for d in docs} == {"b200.md", "misc.jsonl"} and len(docs) == 4
    assert not any("148" in d["text"] and "180" in d["text"] for d in docs)      # sections are not merged
    bm = BM25Index()
    for src in ("b200.md", "misc.jsonl"):
        bm.add_documents([d["text"] for d in docs if d["source"] == src], source=src)
    assert bm.add_documents([docs[0]["text"]], source=docs[0]["source"]) == 0       # deduplicated
    top = bm.search("how many gigabytes of HBM3e capacity", k=2)
    assert "180" in top[0][0]["text"] and top[0][1] > 0
    assert bm.search("potassium", k=3, source="b200.md") == []
    emb = EmbeddingIndex()
    for src in ("b200.md", "misc.jsonl"):
        emb.add_documents([d["text"] for d in docs if d["source"] == src], source=src)
    hy = HybridRetriever([emb, bm], weights=[1.0, 1.0], k_per_source=1)
    res = hy.search("NVLink gigabytes per second", k=3)
    assert "900" in res[0][0]["text"] and len({d["source"] for d, _ in res}) == len(res) <= 2   # one chunk per source
    # follow-up rewriting + the QA chain on top of the hybrid retriever
    mem = ConversationMemory()
    mem.add("How many SMs does the B200 have?", "148.")
    calls = []

    def gen(prompt):
        calls.append(prompt)
        if prompt.startswith("Rewrite"):
            return "What is the HBM3e capacity of the B200?"
        return prompt.split("context:\n")[1].split("\n")[0]

    assert rewrite_follow_up("And its memory?", mem, gen) == "What is the HBM3e capacity of the B200?"
    assert rewrite_follow_up("And its memory?", ConversationMemory(), gen) == "And its memory?"
    qa = RetrievalQA(HybridRetriever([emb, bm], weights=[0.5, 1.0]), gen, k=2, min_score=0.0, memory=mem, rewrite=True)
    ans, src = qa.run("And its memory?")
    assert any("180" in d["text"] for d in src) and ans and any(c.startswith("Rewrite") for c in calls)
    assert "HBM3e capacity" not in calls[-1].split("question:")[1]      # the user's wording goes into the prompt


def test_trainer_callbacks_performance_checkpoint_and_metrics(tmp_path):
    from coati.trainer.callbacks import Callback, MetricsLogger, PerformanceEvaluator, SaveCheckpoint

    sft_items = [tokenize_sft([{"role": "user", "content": "ping"}, {"role": "assistant", "content": "pong"}], _tok)] * 8
    dl = torch.utils.data.DataLoader(ListDataset(sft_items), batch_size=4, collate_fn=DataCollatorForSupervisedDataset())
    m = _tiny()
    n_params = sum(p.numel() for p in m.parameters())
    events = []

    class Recorder(Callback):
        def on_fit_start(self, trainer):
            events.append("fit_start")

        def on_epoch_end(self, trainer, epoch):
            events.append(f"epoch_end_{epoch}")

        def on_batch_end(self, trainer, batch, metrics):
            events.append("batch_end")
            assert "loss" in metrics and torch.is_tensor(batch["input_ids"])

    perf = PerformanceEvaluator(n_params, num_layers=m.cfg.num_hidden_layers, hidden_size=m.cfg.hidden_size, ignore_steps=1)
    tr = SFTTrainer(m, None, torch.optim.AdamW(m.parameters(), lr=1e-3), max_epochs=2)
    tr.add_callbacks(Recorder(), perf, SaveCheckpoint(str(tmp_path / "ckpt"), interval=2), MetricsLogger(str(tmp_path / "log.jsonl")))
    tr.fit(dl)
    assert events == ["fit_start", "batch_end", "batch_end", "epoch_end_0", "batch_end", "batch_end", "epoch_end_1"]
    s = tr.performance
    assert s["steps"] == 4 and s["train_tokens_per_s"] > 0 and s["train_samples_per_s"] > 0 and s["train_tflops_per_device"] > 0
    # the flop model: 3 x (2 N + 4 L H S) per token for a training step
    ids = next(iter(dl))["input_ids"]
    per_tok = 3 * (2 * n_params + 4 * m.cfg.num_hidden_layers * m.cfg.hidden_size * ids.shape[1])
    assert abs(perf.train_flops / perf.train_tokens - per_tok) / per_tok < 0.35     # masks make tokens < B x S
    assert (tmp_path / "ckpt" / "epoch_1" / "model.pt").exists() and not (tmp_path / "ckpt" / "epoch_0").exists()
    lines = (tmp_path / "log.jsonl").read_text().strip().splitlines()
    assert len(lines) == len(tr.history) == 4
    # online loop: collect / update hooks
    actor, init = _tiny(0, vocab_size=32), _tiny(0, vocab_size=32)
    g = GRPOTrainer(None, actor, init, torch.optim.AdamW(actor.parameters(), lr=1e-3), _count_reward, num_generations=4,
                    generate_kwargs=dict(max_new_tokens=4))
    perf2 = PerformanceEvaluator(sum(p.numel() for p in actor.parameters()))
    g.add_callbacks(perf2)
    g.fit(_prompt_loader(), num_episodes=1, num_collect_steps=2, num_update_steps=1)
    assert g.performance["generate_tokens_per_s"] > 0 and g.performance["steps"] == 1


def test_generation_controls():
    from coati.models.generation import _apply_repetition_penalty, generate

    m = _tiny(3, vocab_size=32)
    ids = torch.tensor([[5, 6, 7], [8, 9, 10]])
    torch.manual_seed(0)
    greedy = generate(m, ids, max_new_tokens=6, do_sample=False)
    assert greedy.shape == (2, 9) and torch.equal(greedy[:, :3], ids)
    # EOS = the first greedy token: without min_new_tokens the rows stop at once, with it they run on
    eos = int(greedy[0, 3])
    short, amask = generate(m, ids[:1], max_new_tokens=6, do_sample=False, eos_token_id=eos, pad_token_id=0,
                            return_action_mask=True)
    assert int(short[0, 3]) == eos and amask[0].tolist() == [True] + [False] * 5 and (short[0, 4:] == 0).all()
    longer, amask2 = generate(m, ids[:1], max_new_tokens=6, do_sample=False, eos_token_id=eos, min_new_tokens=3,
                              return_action_mask=True)
    assert (longer[0, 3:6] != eos).all() and amask2[0, :3].all()
    # a two-token stop sequence taken from the greedy continuation ends the row right after it
    stop = greedy[1, 4:6].tolist()
    stopped, am = generate(m, ids[1:], max_new_tokens=6, do_sample=False, stop_sequences=[stop], return_action_mask=True)
    assert stopped[0, 4:6].tolist() == stop and am[0].tolist() == [True, True, True, False, False, False]
    # repetition penalty: seen tokens are pushed down, unseen ones untouched, padding positions do not count as seen
    lg = torch.tensor([[2.0, -2.0, 1.0, 0.5]])
    out = _apply_repetition_penalty(lg, torch.tensor([[0, 1, 3]]), torch.tensor([[1, 1, 0]]), 2.0)
    assert out.tolist() == [[1.0, -4.0, 1.0, 0.5]]
    rep = generate(m, ids, max_new_tokens=6, do_sample=False, repetition_penalty=50.0)
    assert all(len(set(r[3:].tolist())) == 6 for r in rep)        # a huge penalty forbids repeats


def test_colossal_llama_tuning_helpers():
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "applications", "Colossal-LLaMA"))
    from colossal_llama import (activate_neftune, deactivate_neftune, expand_vocab, format_numel_str,
                                freeze_non_embeds_parameters, get_model_numel, plan_vocab_expansion, unfreeze_parameters)

    m = _tiny(0, vocab_size=32)
    ids = torch.randint(0, 32, (2, 8))
    clean = m.model.embed_tokens(ids)
    activate_neftune(m, neftune_noise_alpha=5.0, generator=torch.Generator().manual_seed(0))
    m.train()
    noisy = m.model.embed_tokens(ids)
    mag = 5.0 / (8 * clean.shape[-1]) ** 0.5
    diff = (noisy - clean).abs()
    assert 0 < diff.max() <= mag + 1e-6 and diff.mean() > 0.3 * mag          # uniform in [-mag, mag]
    m.eval()
    assert torch.equal(m.model.embed_tokens(ids), clean)                       # no noise at evaluation time
    m.train()
    deactivate_neftune(m)
    assert torch.equal(m.model.embed_tokens(ids), clean)
    # stage 1 of vocabulary expansion: only embeddings + head train
    tokens, sources = plan_vocab_expansion({"hello": 50, "hi": 500, "a": 1000, "worlds": 10},
                                           encode=lambda t: [ord(c) % 32 for c in t], old_vocab_size=32, max_new_tokens=2)
    assert tokens == ["hi", "hello"] and sources == {32: [ord("h") % 32, ord("i") % 32],
                                                     33: [ord(c) % 32 for c in "hello"]}     # gain 500 > 200 > 50
    m = expand_vocab(m, 34, sources)
    kept = freeze_non_embeds_parameters(m)
    assert kept and all(("embed_tokens" in n) or ("lm_head" in n) for n in kept)
    assert get_model_numel(m, trainable_only=True) < get_model_numel(m)
    loss = m(input_ids=torch.randint(0, 34, (1, 6)), labels=torch.randint(0, 34, (1, 6)))["loss"]
    loss.backward()
    assert m.model.embed_tokens.weight.grad is not None and all(
        p.grad is None for n, p in m.named_parameters() if "embed_tokens" not in n and "lm_head" not in n)
    unfreeze_parameters(m)
    assert get_model_numel(m, trainable_only=True) == get_model_numel(m)
    assert format_numel_str(8_030_000_000) == "8.03 B" and format_numel_str(1500) == "1.50 K" and format_numel_str(7) == "7"


def test_producer_eval_rollout_log_and_consumer_checkpoint(tmp_path):
    """Sharded prompt draws, periodic evaluation on held-out prompts, the rollout jsonl and consumer checkpoints."""
    import json

    from coati.distributed import merge_rollouts

    policy, sampler = _tiny(0, vocab_size=32), _tiny(0, vocab_size=32)
    log = tmp_path / "rollouts.jsonl"
    prod = Producer(ModelRolloutBackend(sampler, dict(max_new_tokens=4)), _prompt_loader(), num_generations=4,
                    rollout_log=str(log))
    cons = GRPOConsumer(policy, torch.optim.AdamW(policy.parameters(), lr=3e-3), _count_reward, num_generations=4)
    torch.manual_seed(0)
    h = launch_distributed(prod, cons, num_steps=4, sync_every=1, eval_dataloaders={"held_out": _prompt_loader()},
                           eval_interval=2, save_dir=str(tmp_path / "ckpt"), save_interval=4)
    assert "eval/held_out" in h[1] and "eval/held_out" in h[3] and "eval/held_out" not in h[0]
    assert 0.0 <= h[3]["eval/held_out"] <= 1.0
    lines = [json.loads(l) for l in log.read_text().splitlines()]
    assert len(lines) == 4 and len(lines[0]["response_ids"]) == 16 and lines[-1]["model_version"] == 3
    state = json.loads((tmp_path / "ckpt" / "step_4" / "state.json").read_text())
    assert state["version"] == 4 and (tmp_path / "ckpt" / "step_4" / "model.pt").exists()
    fresh = _tiny(1, vocab_size=32)
    fresh.load_state_dict(torch.load(tmp_path / "ckpt" / "step_4" / "model.pt"))
    for a, b in zip(fresh.parameters(), policy.parameters()):
        assert torch.equal(a, b)
    # two producers over one dataloader: disjoint, alternating batches
    items = [{"input_ids": [1, 10 + i, 11 + i]} for i in range(8)]
    mk = lambda: torch.utils.data.DataLoader(ListDataset(items), batch_size=2, collate_fn=DataCollatorForPromptDataset())
    p0 = Producer(ModelRolloutBackend(sampler, dict(max_new_tokens=3)), mk(), 2, producer_idx=0, num_producers=2)
    p1 = Producer(ModelRolloutBackend(sampler, dict(max_new_tokens=5)), mk(), 2, producer_idx=1, num_producers=2)
    r0, r1 = p0.rollout(), p1.rollout()
    assert r0["sequences"][0, 1].item() == 10 and r1["sequences"][0, 1].item() == 12          # batches 0 and 1
    r1["model_version"] = 3
    m = merge_rollouts([r0, None, r1], pad_token_id=0)
    assert m["sequences"].shape == (8, 3 + 5) and m["attention_mask"].shape[0] == 8 and m["model_version"] == 0
    assert m["producer"] == [0] * 4 + [1] * 4 and (m["sequences"][:4, 6:] == 0).all()
    out = cons.step(m)                                                                          # trains on the merged batch
    assert out["kept"] >= 0.0 and cons.version == 5


def _multi_producer_worker(rank, world_size, port):
    import torch.distributed as dist

    import colossalai_b200

    colossalai_b200.launch(rank, world_size, "127.0.0.1", port, backend="gloo", verbose=False)
    prod = cons = None
    if rank < 2:                                          # two producers, each with its shard of the prompts
        prod = Producer(ModelRolloutBackend(_tiny(0, vocab_size=32), dict(max_new_tokens=4)), _prompt_loader(),
                        num_generations=4, producer_idx=rank, num_producers=2)
    else:
        policy = _tiny(0, vocab_size=32)
        cons = GRPOConsumer(policy, torch.optim.AdamW(policy.parameters(), lr=3e-3), _count_reward, num_generations=4)
    torch.manual_seed(rank)
    h = launch_distributed(prod, cons, num_steps=3, sync_every=1, producer_ranks=(0, 1))
    if cons is not None:
        assert len(h) == 3 and cons.version == 3 and all("reward" in x for x in h)
        ref = torch.cat([p.detach().flatten() for p in cons.policy.parameters()])
    else:
        assert prod.model_version == 3
        ref = torch.cat([p.detach().flatten() for p in prod.backend.model.parameters()])
    got = [torch.empty_like(ref) for _ in range(world_size)]
    dist.all_gather(got, ref)
    assert torch.equal(got[0], got[2]) and torch.equal(got[1], got[2])      # both producers hold the consumer's weights
    dist.destroy_process_group()


@pytest.mark.dist
def test_two_producers_one_consumer_processes():
    from colossalai_b200.testing import spawn

    spawn(_multi_producer_worker, 3)


def test_colossal_eval_batched_pipeline_and_judge(tmp_path):
    """`EvalModel` (right-padded batches) agrees with the one-item-at-a-time functions; the two-phase pipeline writes
    answers and scores them offline; the LLM-as-judge path parses scores / battles from a stub judge."""
    import json

    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "applications", "ColossalEval"))
    from colossal_eval import (EvalModel, Evaluator, format_table, judge_battle, judge_scores, parse_battle, parse_score,
                               run_evaluation, run_inference, score_choices_by_loglikelihood)

    model = _tiny(3, vocab_size=300).eval()
    dec = lambda ids: bytes(max(0, min(255, i - 3)) for i in ids if i >= 3).decode(errors="replace")     # noqa: E731
    em = EvalModel(model, _tok, dec, batch_size=3, max_new_tokens=5, eos_token_id=2)
    # per-choice losses: batched == sequential (length-normalised log-likelihood, sign flipped)
    items = [{"instruction": f"Q{i}: pick", "choices": [" a", " bb", " ccc" * (i + 1)], "answer": i % 3, "category": f"c{i % 2}"}
             for i in range(5)]
    scored = em.score_choices(items)
    for it, s in zip(items, scored):
        ref = score_choices_by_loglikelihood(model, _tok, it["instruction"], it["choices"])
        torch.testing.assert_close(torch.tensor(s["choice_losses"]), -torch.tensor(ref), atol=1e-4, rtol=1e-4)
        assert s["output"] == max(range(3), key=ref.__getitem__)
    # generation: rows of different lengths in one right-padded batch == one at a time
    prompts = ["hi", "a much longer prompt than the others", "mid size"]
    single = Evaluator(model, _tok, dec, max_new_tokens=5)
    assert em.generate(prompts) == [single.generate(p) for p in prompts]
    # masked target loss == cross entropy over the target tokens only
    l = em.get_loss(["The quick", "x"], [" brown fox", " yz"])
    ids = torch.tensor([_tok("The quick") + _tok(" brown fox")])
    lp = torch.log_softmax(get_logits(model, ids)[0, :-1].float(), -1).gather(-1, ids[0, 1:, None]).squeeze(-1)
    assert abs(l[0] - float(-lp[len(_tok("The quick")) - 1:].mean())) < 1e-4
    # two-phase pipeline
    datasets = {"choice": items, "gen": [{"instruction": p, "target": "x", "category": "g"} for p in prompts],
                "lm": [{"instruction": "The quick", "target": " brown fox", "calculate_loss": True}]}
    paths = run_inference(em, datasets, str(tmp_path / "answers"))
    blob = json.loads(open(paths["choice"]).read())
    assert blob["num_items"] == 5 and [it["index"] for it in blob["items"]] == list(range(5))
    res = run_evaluation(str(tmp_path / "answers"), {"gen": ["exact_match", "f1"]}, str(tmp_path / "results.json"))
    acc = sum(int(s["output"] == s["answer"]) for s in scored) / 5
    assert abs(res["choice"]["micro_avg"]["accuracy"] - acc) < 1e-9 and set(res["choice"]) == {"c0", "c1", "macro_avg", "micro_avg"}
    assert abs(res["lm"]["micro_avg"]["loss"] - l[0]) < 1e-4 and "f1" in res["gen"]["g"]
    assert "perplexity" in format_table(res) and (tmp_path / "results.json").exists()
    # judge
    assert parse_score("fine.\nScore: 4/5") == 4 and parse_score("no verdict") is None
    assert parse_battle("B is vague. Winner: A") == "a" and parse_battle("Winner: Tie") == "tie"

    def stub(prompt):                                   # prefers long answers; always names the longer assistant
        if "[Assistant A]" in prompt:
            a = prompt.split("[Assistant A]\n")[1].split("\n[Assistant B]")[0]
            b = prompt.split("[Assistant B]\n")[1].split("\n\n")[0]
            return "Winner: " + ("A" if len(a) > len(b) else "B" if len(b) > len(a) else "tie")
        ans = prompt.split("[Assistant's answer]\n")[1].split("\n\n")[0]
        return f"ok. Score: {min(5, 1 + len(ans) // 4)}" if ans != "??" else "cannot say"

    rep = judge_scores(stub, [{"instruction": "q1", "output": "short", "category": "x"},
                              {"instruction": "q2", "output": "a considerably longer answer", "category": "y"},
                              {"instruction": "q3", "output": "??", "category": "y"}], criteria=("correctness",))
    assert rep["failed"] == 1 and rep["overall"]["correctness"] == (2 + 5) / 2 and rep["by_category"]["y"]["correctness"] == 5
    b = judge_battle(stub, ["q"] * 3, ["long answer here", "s", "same"], ["s", "long answer here", "same"])
    assert b["win_rate_a"] == b["win_rate_b"] == b["tie_rate"] == 1 / 3 and b["failed"] == 0


def test_colossalqa_universal_conversation(tmp_path):
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "applications", "ColossalQA"))
    from colossalqa import (LocalLLM, UniversalRetrievalConversation, classify_intent, detect_language, load_table,
                            split_chinese_text)

    assert detect_language("What is the warranty?") == "en" and detect_language("保修期是多久？") == "zh"
    zh = "极光笔记本的保修期为二十四个月。电池损耗仅在前十二个月内保修！退货须在收货后三十天内提出，且产品未使用、包装完好。"
    chunks = split_chinese_text(zh, 24)
    assert "".join(chunks) == zh and all(len(c) <= 24 for c in chunks) and chunks[0].endswith("。")
    (tmp_path / "t.csv").write_text("name,price,stock\nwidget,9.5,3\ngadget,20,\n")
    rows = load_table(tmp_path / "t.csv")
    assert rows[0]["text"] == "name: widget; price: 9.5; stock: 3" and rows[1]["text"] == "name: gadget; price: 20"

    seen = []

    def fake_model(prompt):                              # echoes the prompt like some engines do, then answers
        seen.append(prompt)
        if "intent:" in prompt:
            return prompt + " Refund."
        return prompt + (" 24 months\nquestion: something else" if "question:" in prompt else " 二十四个月\n\n多余")

    llm = LocalLLM(fake_model)
    conv = UniversalRetrievalConversation(llm, k=2, intents={"refund": "wants money back", "product": "asks about a product"},
                                          intent_replies={"refund": "Please open a refund ticket."})
    added = conv.add_documents(["The Aurora laptop has a warranty period of 24 months. Returns within 30 days.", zh], "kb")
    assert added["en"] >= 1 and added["zh"] >= 1
    assert classify_intent(lambda p: "Product.", "how long is the warranty", conv.intents) == "product"
    assert classify_intent(lambda p: "no idea", "??", conv.intents) == "other"
    ans, src, meta = conv.run("I want my money back")                 # intent routed: canned reply, no retrieval
    assert ans == "Please open a refund ticket." and src == [] and meta["intent"] == "refund"
    conv.intents = None
    ans, src, meta = conv.run("What is the warranty period of the laptop?")
    assert ans == "24 months" and meta["language"] == "en" and "warranty" in src[0]["text"] and "context:" in seen[-1]
    ans, src, meta = conv.run("笔记本的保修期是多久？")
    assert ans == "二十四个月" and meta["language"] == "zh" and "保修" in src[0]["text"] and "资料：" in seen[-1]
    n = llm.calls
    ans, src, _ = conv.run("Jupiter?")                                  # nothing relevant retrieved: refuse, no model call
    assert ans == "I do not know." and src == [] and llm.calls == n
    assert "warranty period" in conv.chains["en"].memory.render()
    conv.reset()
    assert conv.chains["en"].memory.render().strip() == ""
