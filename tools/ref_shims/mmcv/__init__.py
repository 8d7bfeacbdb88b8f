"""Minimal stand-in for the two mmcv names the reference imports (mmcv 1.6.1 is not installed and there is no network):
`mmcv.Config` / `mmcv.DictAction` (train.py:13, test.py) and `mmcv.cnn.normal_init` / `constant_init` (model/cdpn_resnet.py:6).
Only tools/train_step.py puts this directory on sys.path; nothing in lc_b200/ imports it."""
import yaml


class Config(dict):
    """Attribute-style nested dict, enough of mmcv.Config for the reference's cfg.x.y / cfg.get(...) accesses."""
    def __init__(self, d=None):
        super().__init__()
        for k, v in (d or {}).items():
            self[k] = Config(v) if isinstance(v, dict) else v

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v

    @staticmethod
    def fromfile(path):
        with open(path) as fh:
            return Config(yaml.safe_load(fh))


class DictAction:   # argparse action of mmcv; the harness does not use --opts
    pass
