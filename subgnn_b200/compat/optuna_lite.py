"""The slice of optuna that train_config.py:40-88, 215-277 and train.py:447-497 use: a study that samples
hyper-parameters (random / grid; TPE requests fall back to random sampling, stated on first use), runs the objective
n_trials times, keeps every trial in a sqlite file (``storage='sqlite:///...'``; reloaded with load_if_exists), applies a
median pruner through the Lightning pruning callback, and reports best_params / best_value."""
import itertools
import json
import math
import os
import random
import sqlite3


class TrialPruned(Exception):
    pass


class RandomSampler:
    def __init__(self, seed=None):
        self.rng = random.Random(seed)

    def sample(self, study, trial, name, dist):
        kind = dist[0]
        if kind == 'cat':
            return self.rng.choice(list(dist[1]))
        _, low, high, log, step = dist
        if kind == 'float':
            if log:
                return math.exp(self.rng.uniform(math.log(low), math.log(high)))
            if step:
                return low + step * self.rng.randint(0, int(round((high - low) / step)))
            return self.rng.uniform(low, high)
        if log:
            return int(min(high, max(low, round(math.exp(self.rng.uniform(math.log(low), math.log(high)))))))
        return low + (step or 1) * self.rng.randint(0, (high - low) // (step or 1))


class TPESampler(RandomSampler):
    _warned = False

    def __init__(self, seed=None, **kw):
        super().__init__(seed)
        if not TPESampler._warned:
            print('[subgnn_b200.compat] optuna is not installed: TPESampler falls back to uniform random sampling')
            TPESampler._warned = True


class GridSampler(RandomSampler):
    def __init__(self, search_space, seed=None):
        super().__init__(seed)
        self.space = {k: list(v) for k, v in search_space.items()}
        self.grid = list(itertools.product(*self.space.values())) if self.space else [()]

    def sample(self, study, trial, name, dist):
        if name in self.space:
            point = self.grid[trial.number % len(self.grid)]
            return point[list(self.space).index(name)]
        return super().sample(study, trial, name, dist)


class NopPruner:
    def prune(self, study, trial):
        return False


class MedianPruner(NopPruner):
    def __init__(self, n_startup_trials=5, n_warmup_steps=0, interval_steps=1):
        self.n_startup_trials, self.n_warmup_steps = n_startup_trials, n_warmup_steps

    def prune(self, study, trial):
        if not trial.intermediate:
            return False
        step = max(trial.intermediate)
        done = [t for t in study.trials if t.state == 'COMPLETE' and t.intermediate]
        if len(done) < self.n_startup_trials or step < self.n_warmup_steps:
            return False
        sign = 1 if study.direction == 'minimize' else -1
        best_here = min(sign * v for v in trial.intermediate.values())
        others = sorted(min(sign * v for s, v in t.intermediate.items() if s <= step) for t in done if any(s <= step for s in t.intermediate))
        if not others:
            return False
        median = others[len(others) // 2] if len(others) % 2 else 0.5 * (others[len(others) // 2 - 1] + others[len(others) // 2])
        return best_here > median


class Trial:
    def __init__(self, study, number):
        self.study, self.number = study, number
        self.params, self.intermediate = {}, {}
        self.value, self.state = None, 'RUNNING'
        self.user_attrs = {}

    def _suggest(self, name, dist):
        if name not in self.params:                         # the same name suggested twice returns the same value
            self.params[name] = self.study.sampler.sample(self.study, self, name, dist)
        return self.params[name]

    def suggest_categorical(self, name, choices):
        return self._suggest(name, ('cat', choices))

    def suggest_float(self, name, low, high, step=None, log=False):
        return self._suggest(name, ('float', low, high, log, step))

    def suggest_uniform(self, name, low, high):
        return self.suggest_float(name, low, high)

    def suggest_loguniform(self, name, low, high):
        return self.suggest_float(name, low, high, log=True)

    def suggest_discrete_uniform(self, name, low, high, q):
        return self.suggest_float(name, low, high, step=q)

    def suggest_int(self, name, low, high, step=1, log=False):
        return self._suggest(name, ('int', low, high, log, step))

    def report(self, value, step):
        self.intermediate[int(step)] = float(value)

    def should_prune(self):
        return bool(self.study.pruner.prune(self.study, self))

    def set_user_attr(self, k, v):
        self.user_attrs[k] = v


class Study:
    def __init__(self, direction='minimize', sampler=None, pruner=None, storage=None, study_name=None):
        assert direction in ('minimize', 'maximize')
        self.direction, self.sampler, self.pruner = direction, sampler or RandomSampler(), pruner or NopPruner()
        self.study_name, self.storage = study_name, storage
        self.trials = []
        self._db = None
        if storage:
            assert storage.startswith('sqlite:///'), 'only sqlite storage is supported by the stand-in'
            self._db = storage[len('sqlite:///'):]
            os.makedirs(os.path.dirname(self._db) or '.', exist_ok=True)
            with sqlite3.connect(self._db) as con:
                con.execute('CREATE TABLE IF NOT EXISTS trials (study TEXT, number INTEGER, state TEXT, value REAL, params TEXT, intermediate TEXT)')

    def _load(self):
        if not self._db:
            return
        with sqlite3.connect(self._db) as con:
            for number, state, value, params, inter in con.execute('SELECT number, state, value, params, intermediate FROM trials WHERE study = ? ORDER BY number',
                                                                    (self.study_name,)):
                t = Trial(self, number)
                t.state, t.value, t.params = state, value, json.loads(params)
                t.intermediate = {int(k): v for k, v in json.loads(inter).items()}
                self.trials.append(t)

    def _store(self, t):
        if not self._db:
            return
        with sqlite3.connect(self._db) as con:
            con.execute('INSERT INTO trials VALUES (?, ?, ?, ?, ?, ?)', (self.study_name, t.number, t.state, t.value, json.dumps(t.params, default=str),
                                                                       json.dumps(t.intermediate)))

    def optimize(self, func, n_trials=None, n_jobs=1, timeout=None, catch=()):
        """trials run one after another (n_jobs is accepted: one GPU, one process)."""
        for _ in range(n_trials or 1):
            t = Trial(self, len(self.trials))
            self.trials.append(t)
            try:
                t.value = float(func(t))
                t.state = 'COMPLETE'
            except TrialPruned:
                t.state = 'PRUNED'
                t.value = t.intermediate[max(t.intermediate)] if t.intermediate else None
            except catch:
                t.state = 'FAIL'
            self._store(t)

    @property
    def best_trial(self):
        done = [t for t in self.trials if t.state == 'COMPLETE' and t.value is not None]
        if not done:
            raise ValueError('No trials are completed yet.')
        return (min if self.direction == 'minimize' else max)(done, key=lambda t: t.value)

    @property
    def best_params(self):
        return dict(self.best_trial.params)

    @property
    def best_value(self):
        return self.best_trial.value

    def __getstate__(self):                                  # joblib.dump(study, ...) (train_config.py:273)
        d = dict(self.__dict__)
        d['trials'] = [{'number': t.number, 'state': t.state, 'value': t.value, 'params': t.params, 'intermediate': t.intermediate} for t in self.trials]
        return d

    def __setstate__(self, d):
        rows = d.pop('trials')
        self.__dict__.update(d)
        self.trials = []
        for r in rows:
            t = Trial(self, r['number'])
            t.state, t.value, t.params, t.intermediate = r['state'], r['value'], r['params'], r['intermediate']
            self.trials.append(t)


def create_study(direction='minimize', sampler=None, pruner=None, storage=None, study_name=None, load_if_exists=False):
    s = Study(direction, sampler, pruner, storage, study_name)
    if load_if_exists:
        s._load()
    return s


def load_study(study_name, storage, sampler=None, pruner=None):
    s = Study('minimize', sampler, pruner, storage, study_name)
    s._load()
    return s


class PyTorchLightningPruningCallback:
    """early_stop_callback of the 0.7.1 Trainer: report the monitored metric every epoch, raise TrialPruned on request."""

    def __init__(self, trial, monitor):
        self.trial, self.monitor = trial, monitor

    def on_epoch_end(self, trainer, pl_module):
        cur = trainer.callback_metrics.get(self.monitor)
        if cur is None:
            return False
        self.trial.report(float(cur), step=trainer.current_epoch)
        if self.trial.should_prune():
            raise TrialPruned('Trial was pruned at epoch %d.' % trainer.current_epoch)
        return False
