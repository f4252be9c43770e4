"""Legacy Engine / Trainer / hooks (reference: tests/test_legacy/test_trainer)."""
import os

import torch
import torch.nn as nn
from torch.utils.data import DataLoader, TensorDataset

from colossalai_b200.legacy.engine import Engine
from colossalai_b200.legacy.trainer import Trainer
from colossalai_b200.legacy.trainer.hooks import LogMetricByEpochHook, LossHook, LRSchedulerHook, SaveCheckpointHook


def test_trainer_fit_with_hooks(tmp_path):
    torch.manual_seed(0)
    x = torch.randn(64, 8)
    y = (x.sum(-1, keepdim=True) > 0).float()
    loader = DataLoader(TensorDataset(x, y), batch_size=16)
    model = nn.Sequential(nn.Linear(8, 16), nn.Tanh(), nn.Linear(16, 1))
    opt = torch.optim.SGD(model.parameters(), lr=0.5)
    sched = torch.optim.lr_scheduler.StepLR(opt, step_size=1, gamma=0.9)
    engine = Engine(model, opt, nn.BCEWithLogitsLoss(), clip_grad_norm=1.0)
    log = LogMetricByEpochHook()
    trainer = Trainer(engine)
    trainer.fit(loader, epochs=4, test_dataloader=loader, hooks=[LossHook(), log, LRSchedulerHook(sched),
                                                                 SaveCheckpointHook(2, str(tmp_path))])
    train_losses = [v["train_loss"] for (_, mode, v) in log.history if mode == "train"]
    assert len(train_losses) == 4 and train_losses[-1] < train_losses[0]
    assert os.path.exists(tmp_path / "epoch_2.pt") and os.path.exists(tmp_path / "epoch_4.pt")
    assert abs(opt.param_groups[0]["lr"] - 0.5 * 0.9 ** 4) < 1e-9
