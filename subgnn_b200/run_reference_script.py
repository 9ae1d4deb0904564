"""Run one of the reference's driver scripts (train_config.py, test.py, train.py) UNCHANGED on top of subgnn_b200:

    python -m subgnn_b200.run_reference_script --project-root /data/subgnn  /path/to/SubGNN/train_config.py -config_path cfg.json

What the launcher arranges before handing control to the script (runpy, run_name='__main__'):
  * ``import SubGNN`` / ``anchor_patch_samplers`` / ``gamma`` / ``subgraph_mpn`` / ``subgraph_utils`` resolve to the subgnn_b200
    modules of the same names (the script's own ``import SubGNN as md`` picks up the CUDA-backed LightningModule);
  * ``import config`` resolves to a module with PROJECT_ROOT (= --project-root, or $SUBGNN_PROJECT_ROOT) and PAD_VALUE (config.py:5-8);
  * pytorch_lightning / optuna / commentjson fall back to subgnn_b200.compat where they are not installed;
  * ``import train`` (test.py:4): the reference's train.py has one mis-indented docstring line (train.py:278) and does
    not compile; the launcher compiles it with that line's indentation normalised, in memory — the file is not modified.
"""
import argparse
import os
import runpy
import sys
import types
from pathlib import Path


def _load_train_module(path):
    src = Path(path).read_text().split('\n')
    for _ in range(8):
        try:
            code = compile('\n'.join(src), str(path), 'exec')
            break
        except IndentationError as e:
            fixed = False
            for ln in range(min(e.lineno, len(src)) - 1, -1, -1):
                ind = len(src[ln]) - len(src[ln].lstrip(' '))
                if src[ln].strip() and ind % 4:
                    src[ln] = ' ' * (ind - ind % 4) + src[ln].lstrip(' ')
                    fixed = True
                    break
            if not fixed:
                raise
    else:
        raise IndentationError('could not normalise the indentation of %s' % path)
    mod = types.ModuleType('train')
    mod.__file__ = str(path)
    sys.modules['train'] = mod
    exec(code, mod.__dict__)
    return mod


def setup(project_root, script=None):
    from . import compat
    replaced = compat.install()
    import subgnn_b200
    from subgnn_b200 import SubGNN as sg
    from subgnn_b200 import anchor_patch_samplers, gamma, subgraph_mpn
    root = Path(project_root)
    cfg = types.ModuleType('config')
    cfg.PROJECT_ROOT, cfg.PAD_VALUE = root, subgnn_b200.PAD_VALUE
    sys.modules['config'] = cfg
    sg.PROJECT_ROOT = root
    utils = types.ModuleType('subgraph_utils')
    utils.read_subgraphs, utils.calc_f1, utils.calc_accuracy = sg.read_subgraphs, sg.calc_f1, sg.calc_accuracy
    for name, mod in (('SubGNN', sg), ('anchor_patch_samplers', anchor_patch_samplers), ('gamma', gamma), ('subgraph_mpn', subgraph_mpn),
                      ('subgraph_utils', utils)):
        sys.modules[name] = mod
    if script is not None:
        train_py = Path(script).resolve().parent / 'train.py'
        if Path(script).name != 'train.py' and train_py.exists() and 'train' not in sys.modules:
            try:
                compile(train_py.read_text(), str(train_py), 'exec')
            except IndentationError:
                _load_train_module(train_py)
    return replaced


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument('--project-root', default=os.environ.get('SUBGNN_PROJECT_ROOT', '.'))
    ap.add_argument('script')
    ap.add_argument('script_args', nargs=argparse.REMAINDER)
    a = ap.parse_args(argv)
    replaced = setup(a.project_root, a.script)
    if replaced:
        print('[subgnn_b200] stand-ins active for: %s' % ', '.join(replaced))
    sys.argv = [a.script] + a.script_args
    if Path(a.script).name == 'train.py':
        mod = _load_train_module(a.script)
        return mod.main(mod.parse_arguments())
    runpy.run_path(a.script, run_name='__main__')


if __name__ == '__main__':
    main()
