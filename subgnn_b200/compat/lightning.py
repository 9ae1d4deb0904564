"""Minimal pytorch-lightning 0.7.1 protocol: what train_config.py:107-161 / train.py:326-447 ask of ``pl``.

Hook order of ``Trainer.fit(model)`` (0.7.1): prepare_data -> configure_optimizers -> train_dataloader / val_dataloader ->
per epoch [training_step -> model.backward(trainer, loss, optimizer, idx) -> clip_grad_norm_(gradient_clip_val) ->
optimizer.step / zero_grad] -> [validation_step ...] -> validation_epoch_end -> logger / ModelCheckpoint /
early_stop_callback.  ``Trainer.test(model)``: test_dataloader -> test_step ... -> test_epoch_end.

Fast path: a module that offers ``training_step_fused`` (subgnn_b200.SubGNN) gets each batch handed to it instead —
forward, loss, backward, clipping and Adam then run as one captured CUDA graph and the torch optimizer object is
not used.  Set ``Trainer(fused=False)`` (or SUBGNN_B200_AUTOGRAD_STEP=1) to force the hook-by-hook autograd path.
"""
import json
import os
import random
import re
import time

import numpy as np
import torch
import torch.nn as nn


def seed_everything(seed):
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    return seed


class LightningModule(nn.Module):
    """Base class of the 0.7.1 protocol: an nn.Module whose hooks the Trainer calls; ``hparams`` / ``device`` are plain
    attributes (the reference assigns both in __init__, SubGNN.py:99,102)."""
    trainer = None
    logger = None


class TensorBoardLogger:
    """save_dir/name/version directory layout of the real logger; scalars go to metrics.jsonl (tensorboard is not
    installed offline) — one JSON object per logged step."""

    def __init__(self, save_dir, name='default', version=None, **kw):
        self.save_dir, self.name = str(save_dir), name
        self.version = version if version is not None else 'version_0'

    @property
    def log_dir(self):
        v = self.version if isinstance(self.version, str) else 'version_%s' % self.version
        return os.path.join(self.save_dir, self.name, v)

    def log_metrics(self, metrics, step=None):
        os.makedirs(self.log_dir, exist_ok=True)
        row = {'step': step}
        for k, v in metrics.items():
            try:
                row[k] = float(v.detach()) if isinstance(v, torch.Tensor) else float(v)
            except (TypeError, ValueError):
                pass
        with open(os.path.join(self.log_dir, 'metrics.jsonl'), 'a') as f:
            f.write(json.dumps(row) + '\n')

    def log_hyperparams(self, params):
        pass

    def save(self):
        pass

    def finalize(self, status):
        pass


class AdvancedProfiler:
    def __init__(self, output_filename=None, **kw):
        self.output_filename = output_filename

    def describe(self):
        pass


class ModelCheckpoint:
    """filepath is a format template ('{epoch}-{val_micro_f1:.2f}-...'); the save_top_k best epochs by ``monitor`` are
    kept as '<epoch=E-val_micro_f1=0.52-...>.ckpt' holding {'epoch', 'global_step', 'state_dict', ...}."""

    def __init__(self, filepath=None, monitor='val_loss', verbose=False, save_top_k=1, save_weights_only=False, mode='auto', period=1, prefix=''):
        self.filepath, self.monitor, self.verbose, self.save_top_k, self.period, self.prefix = filepath, monitor, verbose, save_top_k, period, prefix
        if mode == 'auto':
            mode = 'min' if 'loss' in monitor else 'max'
        self.mode = mode
        self.best_k_models = {}                     # path -> score
        self.best = None

    def format_checkpoint_name(self, epoch, metrics):
        d, tmpl = os.path.split(self.filepath)
        m = dict(metrics)
        m['epoch'] = epoch
        if not tmpl:
            tmpl = '{epoch}'
        for tmp in re.findall(r'(\{.*?)[:\}]', tmpl):           # '{epoch' -> 'epoch={epoch' (0.7.1 format_checkpoint_name)
            name = tmp[1:]
            tmpl = tmpl.replace(tmp, name + '={' + name)
            m.setdefault(name, 0)
        m = {k: (float(v) if isinstance(v, (torch.Tensor, np.generic)) else v) for k, v in m.items()}
        return os.path.join(d, self.prefix + tmpl.format(**m) + '.ckpt')

    def _better(self, a, b):
        return a < b if self.mode == 'min' else a > b

    def on_validation_end(self, trainer, pl_module):
        metrics = trainer.callback_metrics
        epoch = trainer.current_epoch
        if self.save_top_k == 0 or (epoch + 1) % self.period or self.monitor not in metrics:
            return
        score = float(metrics[self.monitor])
        if self.save_top_k > 0 and len(self.best_k_models) >= self.save_top_k:
            worst = (max if self.mode == 'min' else min)(self.best_k_models, key=self.best_k_models.get)
            if not self._better(score, self.best_k_models[worst]):
                return
            self.best_k_models.pop(worst)
            if os.path.exists(worst):
                os.remove(worst)
        path = self.format_checkpoint_name(epoch, metrics)
        os.makedirs(os.path.dirname(path) or '.', exist_ok=True)
        trainer.save_checkpoint(path)
        self.best_k_models[path] = score
        self.best = (min if self.mode == 'min' else max)(self.best_k_models.values())
        if self.verbose:
            print('\nEpoch %05d: %s reached %.5f, saving model to %s' % (epoch, self.monitor, score, path))


class EarlyStopping:
    def __init__(self, monitor='val_loss', min_delta=0.0, patience=3, verbose=False, mode='auto'):
        self.monitor, self.min_delta, self.patience = monitor, min_delta, patience
        self.mode = ('min' if 'loss' in monitor else 'max') if mode == 'auto' else mode
        self.best, self.wait = None, 0

    def on_epoch_end(self, trainer, pl_module):
        cur = trainer.callback_metrics.get(self.monitor)
        if cur is None:
            return False
        cur = float(cur)
        if self.best is None or (cur < self.best - self.min_delta if self.mode == 'min' else cur > self.best + self.min_delta):
            self.best, self.wait = cur, 0
            return False
        self.wait += 1
        return self.wait >= self.patience


def _to_device(x, device):
    if isinstance(x, torch.Tensor):
        return x.to(device)
    if isinstance(x, dict):
        return {k: _to_device(v, device) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        return type(x)(_to_device(v, device) for v in x)
    return x


class Trainer:
    def __init__(self, max_epochs=1000, min_epochs=1, gpus=0, num_sanity_val_steps=0, progress_bar_refresh_rate=1, gradient_clip_val=0.0,
                 logger=True, checkpoint_callback=True, early_stop_callback=None, profiler=None, auto_lr_find=False, fused=None,
                 resume_from_checkpoint=None, **unused):
        self.max_epochs, self.min_epochs, self.gpus = max_epochs, min_epochs, gpus
        self.gradient_clip_val = float(gradient_clip_val or 0.0)
        self.progress_bar_refresh_rate = progress_bar_refresh_rate
        self.logger = logger if not isinstance(logger, bool) else None
        self.checkpoint_callback = checkpoint_callback if not isinstance(checkpoint_callback, bool) else None
        self.early_stop_callback = early_stop_callback if not isinstance(early_stop_callback, bool) else None
        self.profiler = profiler
        if auto_lr_find:
            print('[subgnn_b200.compat] auto_lr_find is accepted and ignored (0.7.1 runs its LR finder only on request of trainer.lr_find)')
        self.fused = fused if fused is not None else os.environ.get('SUBGNN_B200_AUTOGRAD_STEP', '0') != '1'
        self.current_epoch, self.global_step = 0, 0
        self.resume_from_checkpoint = resume_from_checkpoint
        self.callback_metrics = {}
        self.optimizers = []
        self.model = None
        self.epoch_times = []

    # ---- helpers ----------------------------------------------------------------------------------------------
    def _device(self, model):
        if self.gpus and torch.cuda.is_available():
            return torch.device('cuda')
        return torch.device('cpu')

    def get_model(self):
        return self.model

    def save_checkpoint(self, path):
        m = self.model
        ckpt = {'epoch': self.current_epoch + 1, 'global_step': self.global_step, 'state_dict': {k: v.detach().cpu() for k, v in m.state_dict().items()},
                # fused mode: the engine's own Adam moments and step count, in torch.optim.Adam's state_dict layout
                'optimizer_states': ([o.state_dict() for o in self.optimizers] if not self._fused_active
                                     else [m.engine.optimizer_state_dict()]),
                'checkpoint_callback_best': getattr(self.checkpoint_callback, 'best', None), 'hparams': dict(getattr(m, 'hparams', {}) or {})}
        torch.save(ckpt, path)

    def _log(self, metrics, step):
        if self.logger is not None and metrics:
            self.logger.log_metrics(metrics, step)

    # ---- fit ----------------------------------------------------------------------------------------------------
    def fit(self, model):
        self.model = model
        model.trainer, model.logger = self, self.logger
        dev = self._device(model)
        if dev.type == 'cuda' and not any(p.is_cuda for p in model.parameters()):
            model.to(dev)
        model.prepare_data()
        if dev.type == 'cuda':
            model.to(dev)                                     # parameters created by prepare_data (trainable cc tables)
        opt = model.configure_optimizers()
        self.optimizers = list(opt) if isinstance(opt, (list, tuple)) else [opt]
        opt = self.optimizers[0]
        self._fused_active = bool(self.fused and hasattr(model, 'training_step_fused'))
        if self._fused_active and getattr(model, 'engine', None) is not None:
            model.engine.grad_clip = self.gradient_clip_val  # Trainer(gradient_clip_val=...) is what Lightning clips with
        train_loader, val_loader = model.train_dataloader(), model.val_dataloader()
        first_epoch = 0
        if self.resume_from_checkpoint:            # Lightning 0.7.1 restore: weights, optimizer state, epoch / step counters
            ckpt = torch.load(self.resume_from_checkpoint, map_location='cpu', weights_only=False)
            model.load_state_dict(ckpt['state_dict'], strict=False)
            if ckpt.get('optimizer_states'):
                if self._fused_active:
                    model.engine.load_optimizer_state_dict(ckpt['optimizer_states'][0])
                else:
                    osd = {k: v for k, v in ckpt['optimizer_states'][0].items() if k in ('state', 'param_groups')}
                    opt.load_state_dict(osd)
            first_epoch, self.global_step = int(ckpt.get('epoch', 0)), int(ckpt.get('global_step', 0))
        stop = False
        for epoch in range(first_epoch, self.max_epochs):
            self.current_epoch = epoch
            t0 = time.time()
            model.train()
            run_loss, n_b = None, 0
            for bi, batch in enumerate(train_loader):
                if self._fused_active:
                    out = model.training_step_fused(batch)
                    loss = out['loss']
                else:
                    batch = _to_device(batch, dev)
                    out = model.training_step(batch, bi)
                    loss = out['loss']
                    model.backward(self, loss, opt, 0)
                    if self.gradient_clip_val > 0:
                        torch.nn.utils.clip_grad_norm_(model.parameters(), self.gradient_clip_val)
                    opt.step()
                    opt.zero_grad()
                run_loss = loss.detach().clone().reshape(()) if run_loss is None else run_loss + loss.detach().reshape(())
                n_b += 1
                self.global_step += 1
                if not self._fused_active and out.get('log'):
                    self._log(out['log'], self.global_step)
            logs = {'epoch': epoch}
            if n_b:
                logs['train_loss_epoch'] = float(run_loss) / n_b
            # ---- validation ----
            if val_loader is not None and len(val_loader) > 0:
                res = self._eval_loop(model, val_loader, dev, 'validation')
                logs.update(res.get('log', {}))
                self.callback_metrics.update({k: v for k, v in res.items() if k != 'log'})
                self.callback_metrics.update(res.get('log', {}))
            self.epoch_times.append(time.time() - t0)
            self._log(logs, self.global_step)
            if self.progress_bar_refresh_rate:
                shown = {k: round(float(v), 4) for k, v in logs.items() if k in ('train_loss_epoch', 'val_loss', 'val_micro_f1', 'val_acc', 'val_auroc')}
                print('epoch %d  %s  (%.2fs)' % (epoch, shown, self.epoch_times[-1]), flush=True)
            if self.checkpoint_callback is not None:
                self.checkpoint_callback.on_validation_end(self, model)
            if self.early_stop_callback is not None:
                stop = bool(self.early_stop_callback.on_epoch_end(self, model)) and epoch + 1 >= self.min_epochs
            if stop:
                break
        if self.logger is not None:
            self.logger.finalize('success')
        return 1

    def _eval_loop(self, model, loader, dev, kind):
        model.eval()
        step = model.validation_step if kind == 'validation' else model.test_step
        end = model.validation_epoch_end if kind == 'validation' else model.test_epoch_end
        outs = []
        with torch.no_grad():
            for bi, batch in enumerate(loader):
                outs.append(step(_to_device(batch, dev) if not self._fused_active else batch, bi))
        res = end(outs) or {}
        model.train()
        return res

    # ---- test ---------------------------------------------------------------------------------------------------
    def test(self, model=None):
        model = model if model is not None else self.model
        self.model = model
        model.trainer = self
        if not hasattr(self, '_fused_active'):
            self._fused_active = bool(self.fused and hasattr(model, 'training_step_fused'))
        dev = self._device(model)
        if getattr(model, 'engine', 1) is None or (dev.type == 'cuda' and not any(p.is_cuda for p in model.parameters())):
            model.prepare_data() if getattr(model, 'engine', 1) is None else model.to(dev)
        loader = model.test_dataloader()
        res = self._eval_loop(model, loader, dev, 'test')
        self.callback_metrics.update(res.get('log', {}))
        self._log(res.get('log', {}), self.global_step)
        return res
