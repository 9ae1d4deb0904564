"""Offline stand-ins for the reference's driver dependencies, so that train_config.py / train.py-style
drivers run unchanged where pytorch-lightning 0.7.1, optuna and commentjson are not installed:

    pytorch_lightning   Trainer (fit / test, 0.7.1 hook protocol), LightningModule, loggers.TensorBoardLogger,
                        callbacks.ModelCheckpoint, profiler.AdvancedProfiler                      -> compat/lightning.py
    optuna              create_study / Study.optimize / Trial.suggest_*, Random / Grid samplers (TPE falls back to random),
                        MedianPruner, integration.PyTorchLightningPruningCallback, sqlite storage  -> compat/optuna_lite.py
    commentjson         load / loads with // , # and /* */ comments                               -> compat/commentjson_lite.py

``install()`` registers a stand-in ONLY for packages that are not importable (or all of them with force=True).
Nothing here is on the hot path: the Trainer hands every training batch to ``SubGNN.training_step_fused`` (one captured
CUDA graph per step) when the module offers it.
"""
import importlib.util
import sys
import types


def _missing(name):
    try:
        return importlib.util.find_spec(name) is None
    except (ImportError, ValueError):
        return True


def install(force=False, only=None):
    """-> list of package names that were replaced by stand-ins."""
    from . import commentjson_lite, lightning, optuna_lite
    done = []

    def want(name):
        if only is not None and name not in only:
            return False
        if force:
            return True
        cur = sys.modules.get(name)
        if cur is not None:                                   # a bare placeholder module (no file behind it) is replaced, a real package is kept
            return not hasattr(cur, '__file__') and not getattr(cur, '__subgnn_compat__', False)
        return _missing(name)

    if want('pytorch_lightning'):
        pl = types.ModuleType('pytorch_lightning')
        pl.__subgnn_compat__ = True
        pl.Trainer, pl.LightningModule, pl.seed_everything = lightning.Trainer, lightning.LightningModule, lightning.seed_everything
        pl.__version__ = '0.7.1+subgnn_b200.compat'
        pl.loggers = types.ModuleType('pytorch_lightning.loggers')
        pl.loggers.TensorBoardLogger = lightning.TensorBoardLogger
        pl.callbacks = types.ModuleType('pytorch_lightning.callbacks')
        pl.callbacks.ModelCheckpoint, pl.callbacks.EarlyStopping = lightning.ModelCheckpoint, lightning.EarlyStopping
        pl.profiler = types.ModuleType('pytorch_lightning.profiler')
        pl.profiler.AdvancedProfiler = lightning.AdvancedProfiler
        for m in (pl, pl.loggers, pl.callbacks, pl.profiler):
            sys.modules[m.__name__] = m
        done.append('pytorch_lightning')
    if want('optuna'):
        op = types.ModuleType('optuna')
        op.__subgnn_compat__ = True
        for k in ('create_study', 'Study', 'Trial', 'TrialPruned', 'load_study'):
            setattr(op, k, getattr(optuna_lite, k))
        op.exceptions = types.ModuleType('optuna.exceptions')
        op.exceptions.TrialPruned = optuna_lite.TrialPruned
        op.samplers = types.ModuleType('optuna.samplers')
        op.samplers.RandomSampler, op.samplers.GridSampler, op.samplers.TPESampler = optuna_lite.RandomSampler, optuna_lite.GridSampler, optuna_lite.TPESampler
        op.pruners = types.ModuleType('optuna.pruners')
        op.pruners.MedianPruner, op.pruners.NopPruner = optuna_lite.MedianPruner, optuna_lite.NopPruner
        op.integration = types.ModuleType('optuna.integration')
        op.integration.PyTorchLightningPruningCallback = optuna_lite.PyTorchLightningPruningCallback
        for m in (op, op.exceptions, op.samplers, op.pruners, op.integration):
            sys.modules[m.__name__] = m
        done.append('optuna')
    if want('commentjson'):
        cj = types.ModuleType('commentjson')
        cj.__subgnn_compat__ = True
        cj.load, cj.loads, cj.dump, cj.dumps = commentjson_lite.load, commentjson_lite.loads, commentjson_lite.dump, commentjson_lite.dumps
        sys.modules['commentjson'] = cj
        done.append('commentjson')
    return done
