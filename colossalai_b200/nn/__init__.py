from . import init, lr_scheduler, optimizer
